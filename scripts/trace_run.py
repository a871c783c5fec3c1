import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from openpose_plus_b200.engine import Engine
ring = bench.make_inputs(4)
dev = torch.device('cuda', 0)
d_ring = [(torch.from_numpy(c).to(dev), torch.from_numpy(p).to(dev)) for c, p in ring]
S = 3
eng = Engine(46, 54, 368, 432, 17, max_batch=64, n_slots=S)
up = [(torch.empty((64, 19, 368, 432), device=dev), torch.empty((64, 38, 368, 432), device=dev)) for _ in range(S)]
infl = []
for k in range(18):
    if len(infl) == S: eng.wait(infl.pop(0))
    c, p = d_ring[k % 4]
    infl.append(eng.submit(c, p, conf_up=up[k % S][0], paf_up=up[k % S][1]))
for t in infl: eng.wait(t)
