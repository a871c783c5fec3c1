"""Stress: many mixed frames (typical, crowded, merged, noisy, near-threshold) through the pipelined engine,
every frame's humans compared with the CPU oracle.  python scripts/stress_parity.py [n_frames]"""
import os, sys, time
sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from openpose_plus_b200 import synth
from openpose_plus_b200 import _capi as capi
from openpose_plus_b200.engine import Engine
from openpose_plus_b200.sharding import process_stream
from oracle.oracle import Oracle, FLAG_UB_PEAK_INDEX
import helpers

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rng = np.random.default_rng(12345)

def make(i):
    kind = i % 8
    if kind in (0, 1, 2):
        return synth.render_frame(10000 + i, int(rng.integers(1, 9)), noise=1e-3 if kind == 2 else 0.0)
    if kind == 3:
        return synth.render_frame(10000 + i, int(rng.integers(20, 40)))
    if kind == 4:
        return synth.render_frame(10000 + i, int(rng.integers(20, 36)), drop_limbs=(12,))
    if kind == 5:
        c, p = synth.render_frame(10000 + i, 6)
        c[:18] *= np.float32(rng.uniform(0.05, 0.2))
        return c, p
    if kind == 6:
        c, p = synth.render_frame(10000 + i, 4)
        c[:18] = np.maximum(c[:18], np.float32(rng.uniform(0.03, 0.0499)))
        return c, p
    c = (rng.random((19, 46, 54), dtype=np.float32) ** 6).astype(np.float32)
    p = (rng.random((38, 46, 54), dtype=np.float32) * 2 - 1).astype(np.float32)
    return c, p

t0 = time.time()
frames = [make(i) for i in range(n)]
conf = np.stack([f[0] for f in frames]); paf = np.stack([f[1] for f in frames])
print("rendered %d frames in %.1fs" % (n, time.time() - t0), flush=True)
eng = Engine(46, 54, max_batch=64, max_peaks_per_part=512, max_cands_per_limb=8192, max_humans=1024)
for rep in range(3):
    humans, counts, flags = process_stream(eng, conf, paf)
print("gpu done; flags histogram:", {int(k): int(v) for k, v in zip(*np.unique(flags, return_counts=True))}, flush=True)

def check(i):
    o = Oracle(46, 54, 368, 432, 17).run(conf[i], paf[i])
    if flags[i] & capi.FLAG_OVERFLOW_MASK:
        return "overflow"
    if o["flags"] & FLAG_UB_PEAK_INDEX:
        return "ub"
    if counts[i] != o["n_humans"]:
        return "BAD count %d vs %d" % (counts[i], o["n_humans"])
    e = helpers.humans_equal(humans[i, :counts[i]], o["humans"])
    return "ok" if e is None else "BAD " + e

with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
    res = list(ex.map(check, range(n)))
bad = [(i, r) for i, r in enumerate(res) if r.startswith("BAD")]
print({k: res.count(k) for k in set(res)}, "in %.1fs" % (time.time() - t0))
print("BAD:", bad[:10])
sys.exit(1 if bad else 0)
