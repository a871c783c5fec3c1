import os, sys, time, statistics
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
from openpose_plus_b200 import _capi as capi
from openpose_plus_b200.engine import Engine
ring = bench.make_inputs(1)
c, p = np.ascontiguousarray(ring[0][0][:1]), np.ascontiguousarray(ring[0][1][:1])
eng = Engine(46, 54, 368, 432, 17, max_batch=1, n_slots=1)
out = (np.zeros((1, eng.max_humans), capi.HUMAN_DT), np.zeros(1, np.int32), np.zeros(1, np.int32))
lat = []
for i in range(300):
    t0 = time.perf_counter(); eng.process(c, p, out=out); lat.append((time.perf_counter() - t0) * 1e6)
print("pageable in/out: p50 %.1f us  p90 %.1f us" % (statistics.median(lat[50:]), sorted(lat[50:])[int(0.9 * 250)]))
hc, hp = capi.pinned_empty(c.shape, np.float32), capi.pinned_empty(p.shape, np.float32); hc[...] = c; hp[...] = p
pout = (capi.pinned_empty((1, eng.max_humans), capi.HUMAN_DT), capi.pinned_empty((1,), np.int32), capi.pinned_empty((1,), np.int32))
lat = []
for i in range(300):
    t0 = time.perf_counter(); eng.process(hc, hp, out=pout); lat.append((time.perf_counter() - t0) * 1e6)
print("pinned in/out:   p50 %.1f us  p90 %.1f us" % (statistics.median(lat[50:]), sorted(lat[50:])[int(0.9 * 250)]))
