#!/usr/bin/env python
"""GPU fuzz: the CUDA path against the CPU oracle on random geometries (integer and non-integer scales), kernel sizes,
both variants, crowd sizes, missing limbs, noise and quantised (tie-rich) maps, one- to five-frame batches, pinned and
pageable buffers, several capacity sets (the limb kernel takes different code paths with them), with and without the
up-sampled maps materialised.  Every stage is compared (tests/helpers.check_frame).
    python scripts/fuzz_gpu_vs_oracle.py [seconds] [seed]
With OPP_B200_LIB=openpose_plus_b200/libopp_b200_dbg.so (python -m openpose_plus_b200.build --debug) every shared-memory
array access of the kernels is bounds-checked as well and violations are reported (opp_debug_bounds_report)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.oracle import Oracle  # noqa: E402
from openpose_plus_b200 import synth, _capi as capi  # noqa: E402
from openpose_plus_b200.engine import Engine  # noqa: E402
import helpers  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0, n_frames, n_cfg, n_over, bad, oob = time.time(), 0, 0, 0, [], []
checked = None  # is this a bounds-checked build?
while time.time() - t0 < budget:
    fh, fw = [(46, 54), (23, 27), (30, 40), (12, 14)][int(rng.integers(4))]
    kind = int(rng.integers(6))
    scale = [8, 8, 8, 4, 2, 1][int(rng.integers(6))]
    oh, ow = fh * scale, fw * scale
    if kind == 5:
        oh, ow = int(fh * rng.uniform(1.0, 6.0)), int(fw * rng.uniform(1.0, 6.0))
    k = int(rng.choice([1, 3, 5, 7, 9, 13, 17, 19, 21, 25, 31, 33, 37]))
    if k <= 7 and (scale > 2 or kind in (3, 5)):
        continue
    if k // 2 >= min(oh, ow) - 1:
        continue
    variant = int(rng.integers(4) == 0)
    nb = int(rng.integers(1, 6))
    frames = []
    for _ in range(nb):
        seed = int(rng.integers(1 << 30))
        if kind in (0, 1, 5):
            c, p = synth.render_frame(seed, int(rng.integers(1, 12)), fh, fw, noise=1e-3 if kind == 1 else 0.0)
        elif kind == 2:
            drop = tuple(int(x) for x in rng.choice(19, size=int(rng.integers(0, 4)), replace=False))
            c, p = synth.render_frame(seed, int(rng.integers(15, 36)), fh, fw, drop_limbs=drop)
        elif kind == 3:
            c, p = synth.render_frame(seed, int(rng.integers(3, 10)), fh, fw, noise=0.02)
        else:
            c, p = synth.render_frame(seed, int(rng.integers(10, 30)), fh, fw)
            c, p = (np.round(c * 8) / 8).astype(np.float32), (np.round(p * 4) / 4).astype(np.float32)
        frames.append((c, p))
    conf, paf = np.stack([f[0] for f in frames]), np.stack([f[1] for f in frames])
    if rng.integers(2):
        hc, hp = capi.pinned_empty(conf.shape, np.float32), capi.pinned_empty(paf.shape, np.float32)
        hc[...], hp[...] = conf, paf
    else:
        hc, hp = conf, paf
    capP, capC, capH = [(512, 8192, 512), (512, 4096, 512), (128, 1024, 128), (256, 2048, 256)][int(rng.integers(4))]
    eng = Engine(fh, fw, oh, ow, k, max_batch=5, max_peaks_per_part=capP, max_cands_per_limb=capC, max_humans=capH, variant=variant)
    orc = Oracle(fh, fw, oh, ow, k, variant=variant)
    kw = {}
    if rng.integers(4) == 0:
        import torch
        kw = dict(conf_up=torch.empty((nb, 19, oh, ow), device="cuda"), paf_up=torch.empty((nb, 38, oh, ow), device="cuda"))
    what = "fuzz %s" % ((fh, fw, oh, ow, k, variant, kind, nb, capP, capC, capH, bool(kw)),)
    try:
        helpers.run_and_check(eng, orc, hc, hp, what, **kw)
        if kw:
            o = orc.run(conf[nb - 1], paf[nb - 1], maps=True)
            assert np.array_equal(kw["conf_up"][nb - 1].cpu().numpy(), o["conf_up"]) and np.array_equal(kw["paf_up"][nb - 1].cpu().numpy(), o["paf_up"]), what + ": up-sampled maps differ"
    except AssertionError as e:
        if "capacity overflow" in str(e):  # tie-rich maps can exceed the capacities chosen above: flagged, not a disagreement
            n_over += 1
        else:
            bad.append(str(e))
            print("MISMATCH", e, flush=True)
    if checked is not False:
        try:
            rec = eng.bounds_report()
            checked = True
            if rec[3]:
                oob.append((what, rec))
                print("OUT OF BOUNDS", what, "line %d index %d size %d (%d violations)" % rec, flush=True)
        except capi.OppError:
            checked = False
    eng.close()
    n_frames += nb
    n_cfg += 1
print("configurations %d, frames %d, batches over capacity (flagged) %d, mismatches %d, bounds-checked build: %s, out-of-bounds accesses in %d configurations"
      % (n_cfg, n_frames, n_over, len(bad), bool(checked), len(oob)))
sys.exit(1 if bad or oob else 0)
