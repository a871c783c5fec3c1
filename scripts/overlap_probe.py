"""Probe: how do the store-bound resize (K1) and the issue-bound peak kernel (K2) share the GPU?"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from openpose_plus_b200 import _capi as capi
from openpose_plus_b200.engine import Engine
ring = bench.make_inputs(4)
dev = torch.device('cuda', 0)
d_ring = [(torch.from_numpy(c).to(dev), torch.from_numpy(p).to(dev)) for c, p in ring]
eng = Engine(46, 54, 368, 432, 17, max_batch=64, n_slots=3)
up = [(torch.empty((64, 19, 368, 432), device=dev), torch.empty((64, 38, 368, 432), device=dev)) for _ in range(3)]
L, h = eng.L, eng.h
streams = [torch.cuda.Stream(device=dev) for _ in range(4)]

def k1(i, st):
    c, p = d_ring[i % 4]
    eng._check(L.opp_resize_pair_device(h, c.data_ptr(), p.data_ptr(), 64, up[i % 3][0].data_ptr(), up[i % 3][1].data_ptr(), 0, st.cuda_stream))

def k2(i, st):
    eng._check(L.opp_peaks_device(h, d_ring[i % 4][0].data_ptr(), None, 64, None, None, st.cuda_stream))

def run(label, plan, reps=30):
    # plan: list of (fn, stream index) launched round-robin per rep
    for _ in range(3):
        for fn, si in plan: fn(0, streams[si])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(reps):
        for fn, si in plan: fn(r, streams[si])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps * 1e3
    n1 = sum(1 for fn, _ in plan if fn is k1)
    print("%-40s %.3f ms per round  (K1 writes %.2f TB/s)" % (label, dt, n1 * 2.356 / dt))

run("K1 x1 one stream", [(k1, 0)])
run("K1 x3 one stream", [(k1, 0)] * 3)
run("K1 x3 three streams", [(k1, 0), (k1, 1), (k1, 2)])
run("K2 x1", [(k2, 3)])
run("K1 (s0) + K2 (s3)", [(k1, 0), (k2, 3)])
run("K1 x2 (s0,s1) + K2 (s3)", [(k1, 0), (k1, 1), (k2, 3)])
