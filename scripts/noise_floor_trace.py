"""Timeline of the limb kernel on configs[1] with a uniform noise floor in [0.03, 0.07] under the heat maps (bench.py dense_maps):
OPP_TRACE=1 OPP_TRACE_DUMP=gpurun_out/k3_nf.bin python scripts/noise_floor_trace.py"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from openpose_plus_b200.engine import Engine
conf, paf = bench.make_ring(bench.render_pool(floor=(0.03, 0.07)), 1)[0]
dev = torch.device("cuda", 0)
dc, dp = torch.from_numpy(conf).to(dev), torch.from_numpy(paf).to(dev)
eng = Engine(46, 54, max_batch=64, max_peaks_per_part=512, max_cands_per_limb=4096, max_humans=256, n_slots=1)
for i in range(3):
    h, c, f = eng.process(dc, dp)
print("humans", c[:4].tolist(), "flags", np.unique(f).tolist(), "counts", eng.debug_counts(0, 0).tolist(), "peaks", len(eng.debug_peaks(0, 0, cap=18 * 512)))
t = np.fromfile(os.environ["OPP_TRACE_DUMP"], dtype=np.uint64).reshape(64, 19, 12).astype(np.int64)
t0 = t[:, :, 0].min()
us = lambda a: (a - t0) * 1e-3
start, staged, scored, sorted_, matched = (us(t[:, :, k]) for k in range(5))
print("limb phases (us, median / max): stage %s score %s sort %s match %s" % tuple("%.1f/%.1f" % (np.median(d), d.max()) for d in (staged - start, scored - staged, sorted_ - scored, matched - sorted_)))
print("limb CTA total median %.1f max %.1f; last limb done at %.1f us" % (np.median(matched - start), (matched - start).max(), matched.max()))
asm = t[:, :, 9] > 0
print("assembly: median %.1f max %.1f us; last done at %.1f us" % (np.median(us(t[:, :, 9][asm]) - us(t[:, :, 5][asm])), (us(t[:, :, 9][asm]) - us(t[:, :, 5][asm])).max(), us(t[:, :, 9][asm]).max()))
