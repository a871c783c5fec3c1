#!/usr/bin/env python
"""CPU fuzz (run where /root/reference exists): the plain-C restatement (oracle/liborc.so) against the reference's own
src/paf.cpp (oracle/_ref/libopp_ref.so, strict build) on random geometries, kernel sizes, crowd sizes, missing limbs,
noise and quantised (tie-rich) maps.  Final human_t lists must be identical; frames where the reference itself reads
out of bounds (FLAG_UB_PEAK_INDEX) are counted and skipped.   python scripts/fuzz_oracle_vs_reference.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle, Reference, FLAG_UB_PEAK_INDEX  # noqa: E402
from openpose_plus_b200 import synth  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)


def same(a, b):
    if len(a) != len(b):
        return False
    ok = np.array_equal(a["score"].view(np.uint32), b["score"].view(np.uint32))
    ok &= np.array_equal(a["parts"]["has_value"] != 0, b["parts"]["has_value"] != 0)
    for f in ("x", "y", "score"):
        ok &= np.array_equal(np.ascontiguousarray(a["parts"][f]).view(np.uint32), np.ascontiguousarray(b["parts"][f]).view(np.uint32))
    return bool(ok)


t0, n, skipped, bad, merges, ties = time.time(), 0, 0, [], 0, 0
cache = {}
while time.time() - t0 < budget:
    fh, fw = [(46, 54), (23, 27), (30, 40), (12, 14)][int(rng.integers(4))]
    kind = int(rng.integers(6))
    scale = [8, 8, 4, 2, 1][int(rng.integers(5))]
    oh, ow = fh * scale, fw * scale
    if kind == 5:  # non-integer geometry
        oh, ow = int(fh * rng.uniform(1.0, 6.0)), int(fw * rng.uniform(1.0, 6.0))
    k = int(rng.choice([1, 3, 5, 7, 9, 13, 17, 25, 31]))
    if k <= 7 and (scale > 2 or kind in (3, 5)):
        continue  # small kernels on replicated maps are all plateaus: tens of thousands of tied peaks, minutes per frame
    if k // 2 >= min(oh, ow) - 1:
        continue
    seed = int(rng.integers(1 << 30))
    if kind in (0, 1, 5):
        conf, paf = synth.render_frame(seed, int(rng.integers(1, 12)), fh, fw, noise=1e-3 if kind == 1 else 0.0)
    elif kind == 2:
        drop = tuple(int(x) for x in rng.choice(19, size=int(rng.integers(0, 4)), replace=False))
        conf, paf = synth.render_frame(seed, int(rng.integers(15, 40)), fh, fw, drop_limbs=drop)
    elif kind == 3:
        conf, paf = synth.noise_frame(seed, min(fh, 14), min(fw, 16))
        fh, fw = conf.shape[1:]
        oh, ow = fh * 8, fw * 8
    else:  # quantised maps: many equal candidate scores -> std::sort tie order matters
        conf, paf = synth.render_frame(seed, int(rng.integers(10, 30)), fh, fw)
        conf = (np.round(conf * 8) / 8).astype(np.float32)
        paf = (np.round(paf * 4) / 4).astype(np.float32)
    if k // 2 >= min(oh, ow) - 1:
        continue
    key = (fh, fw, oh, ow, k)
    if key not in cache:
        if len(cache) > 24:
            cache.clear()
        cache[key] = (Oracle(fh, fw, oh, ow, k), Reference(fh, fw, oh, ow, k))
    orc, ref = cache[key]
    o = orc.run(conf, paf)
    n += 1
    merges += o["n_merges"]
    ties += sum(o["ties"])
    if o["flags"] & FLAG_UB_PEAK_INDEX:
        skipped += 1
        continue
    r = ref.run(conf, paf, cap=16384)
    if not same(o["humans"], r):
        bad.append((key, kind, seed, len(o["humans"]), len(r)))
        print("MISMATCH", bad[-1], flush=True)
print("frames %d, skipped (reference UB) %d, merges %d, limbs with tied scores %d, mismatches %d" % (n, skipped, merges, ties, len(bad)))
sys.exit(1 if bad else 0)
