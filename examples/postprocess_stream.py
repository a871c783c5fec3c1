#!/usr/bin/env python
"""Streaming use of the post-processor (the role of the third stage of the reference's
examples/stream_detector.cpp:120-136): batches of feature maps in, skeletons out, several batches in
flight.  Synthetic maps stand in for the CNN.

    python examples/postprocess_stream.py [n_frames] [batch]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from openpose_plus_b200 import synth  # noqa: E402
from openpose_plus_b200.engine import Engine  # noqa: E402
from openpose_plus_b200.post_process import humans_from_records  # noqa: E402
from openpose_plus_b200.sharding import process_stream  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    conf, paf = synth.render_batch(n, n_people=5, pool=16)
    eng = Engine(46, 54, 368, 432, gauss_kernel_size=17, max_batch=batch)
    process_stream(eng, conf[:batch], paf[:batch])  # warm-up
    t0 = time.perf_counter()
    records, counts, flags = process_stream(eng, conf, paf)
    dt = time.perf_counter() - t0
    print("%d frames in %.1f ms: %.0f frames/s (pageable host maps), %d skeletons" % (n, dt * 1e3, n / dt, int(counts.sum())))
    for h in humans_from_records(records[0, :counts[0]], 368, 432):
        print(" ", h)


if __name__ == "__main__":
    main()
