"""Timeline of the limb kernel over a whole crowded batch: OPP_TRACE=1 OPP_TRACE_DUMP=gpurun_out/k3_stamps.bin python scripts/k3_batch_trace.py [people]
Prints, from the %globaltimer stamps of every (frame, limb) CTA: the kernel's span, CTA durations per phase, how many CTAs
ran at once, and the assembly tail."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from openpose_plus_b200 import synth
from openpose_plus_b200.engine import Engine
people = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
conf, paf = synth.render_batch(64, n_people=people, seed0=2000, pool=8)
dc, dp = torch.from_numpy(conf).to(dev), torch.from_numpy(paf).to(dev)
eng = Engine(46, 54, max_batch=64, max_humans=256, n_slots=1)
for i in range(3):
    eng.process(dc, dp)
t = np.fromfile(os.environ["OPP_TRACE_DUMP"], dtype=np.uint64).reshape(64, 19, 12).astype(np.int64)
t0 = t[:, :, 0].min()
us = lambda a: (a - t0) * 1e-3
start, staged, scored, sorted_, matched = (us(t[:, :, k]) for k in range(5))
print("CTA start: min %.1f  median %.1f  max %.1f us" % (start.min(), np.median(start), start.max()))
print("limb phases (us, median / p90 / max over 1216 CTAs): stage %s score %s sort %s match %s" % tuple(
    "%.1f/%.1f/%.1f" % (np.median(d), np.percentile(d, 90), d.max()) for d in (staged - start, scored - staged, sorted_ - scored, matched - sorted_)))
print("limb CTA total: median %.1f p90 %.1f max %.1f; last limb done at %.1f us" % (np.median(matched - start), np.percentile(matched - start, 90), (matched - start).max(), matched.max()))
asm = t[:, :, 9] > 0
ent, done = us(t[:, :, 5][asm]), us(t[:, :, 9][asm])
stg, tree, assembled = us(t[:, :, 6][asm]), us(t[:, :, 10][asm]), us(t[:, :, 7][asm])
print("assembly (64 frames): stage %.1f tree %.1f virtual %.1f out %.1f  total median %.1f max %.1f; last done at %.1f us" % (
    np.median(stg - ent), np.median(tree - stg), np.median(assembled - tree), np.median(done - assembled), np.median(done - ent), (done - ent).max(), done.max()))
ev = sorted([(s, 1) for s in start.ravel()] + [(e, -1) for e in matched.ravel()])
cur, best, area, last = 0, 0, 0.0, 0.0
for x, d in ev:
    area += cur * (x - last); last = x; cur += d; best = max(best, cur)
print("limb CTAs in flight: max %d, mean %.0f over %.1f us" % (best, area / max(last, 1e-9), last))
sm = t[:, :, 11] - 1
print("SMs used: %d; CTAs per SM min/max %d/%d" % (len(np.unique(sm)), np.bincount(sm.ravel()).min(), np.bincount(sm.ravel()).max()))
