#!/usr/bin/env python
"""BASELINE.json configs[4]: a 4096-frame stream sharded over the GPUs of one box from ONE process
(one handle + streams per GPU, host gather, no collective).   python examples/multi_gpu_stream.py [n_frames]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from openpose_plus_b200 import _capi as capi, synth  # noqa: E402
from openpose_plus_b200.engine import Engine  # noqa: E402
from openpose_plus_b200.sharding import process_stream_multi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
c, p = synth.render_batch(n, n_people=5, pool=16)
conf, paf = capi.pinned_empty(c.shape, np.float32), capi.pinned_empty(p.shape, np.float32)
conf[...] = c
paf[...] = p
for ndev in [d for d in (1, 2, 4, 8) if d <= torch.cuda.device_count()]:
    engines = [Engine(46, 54, max_batch=64, device=d) for d in range(ndev)]
    process_stream_multi(engines, conf[:256 * ndev], paf[:256 * ndev])
    t0 = time.perf_counter()
    humans, counts, flags = process_stream_multi(engines, conf, paf)
    dt = time.perf_counter() - t0
    print("%d GPU(s): %d frames in %.1f ms = %.0f frames/s (host maps in, skeletons out), %d skeletons" % (ndev, n, dt * 1e3, n / dt, int(counts.sum())))
    for e in engines:
        e.close()
