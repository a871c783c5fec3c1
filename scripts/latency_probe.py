import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
from openpose_plus_b200 import _capi as capi
from openpose_plus_b200.engine import Engine
ring = bench.make_inputs(1)
c, p = ring[0]
hc, hp = capi.pinned_empty((1,) + c.shape[1:], np.float32), capi.pinned_empty((1,) + p.shape[1:], np.float32)
hc[...] = c[:1]; hp[...] = p[:1]
eng = Engine(46, 54, 368, 432, 17, max_batch=64, n_slots=3)
out = (capi.pinned_empty((1, eng.max_humans), capi.HUMAN_DT), capi.pinned_empty((1,), np.int32), capi.pinned_empty((1,), np.int32))
tot = []
for i in range(330):
    t0 = time.perf_counter(); t = eng.submit(hc, hp, out=out); t1 = time.perf_counter(); eng.wait(t); t2 = time.perf_counter()
    if i >= 30: tot.append((t2 - t0) * 1e6)
    if i >= 326: print("submit %.1f us  wait %.1f us  total %.1f us  device %.1f us" % ((t1-t0)*1e6, (t2-t1)*1e6, (t2-t0)*1e6, eng.last_batch_ms(t)*1e3), file=sys.stderr)
print("p50 %.1f us  p10 %.1f us  p90 %.1f us over %d calls" % (np.percentile(tot, 50), np.percentile(tot, 10), np.percentile(tot, 90), len(tot)), file=sys.stderr)
