"""Small end-to-end run for compute-sanitizer: typical, crowded, materialised, generic-path frames."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from openpose_plus_b200 import synth
from openpose_plus_b200.engine import Engine
conf, paf = synth.render_batch(3, n_people=5, seed0=1)
c2, p2 = synth.render_frame(200, 34, drop_limbs=(12,))
conf = np.concatenate([conf, c2[None]]); paf = np.concatenate([paf, p2[None]])
eng = Engine(46, 54, max_batch=4, max_humans=256)
h, c, f = eng.process(conf, paf)
print("skeleton-only", c.tolist(), f.tolist())
cu = torch.empty((4, 19, 368, 432), device="cuda"); pu = torch.empty((4, 38, 368, 432), device="cuda")
h, c, f = eng.process(conf, paf, conf_up=cu, paf_up=pu)
print("materialised", c.tolist(), float(cu.sum()))
h, c, f = eng.process(torch.from_numpy(conf).cuda(), torch.from_numpy(paf).cuda())
print("device in", c.tolist())
eng2 = Engine(46, 54, 300, 400, 25, max_batch=4)
h, c, f = eng2.process(conf, paf)
print("generic", c.tolist())
h, c, f = eng.process(conf[:1], paf[:1])
print("single", c.tolist())
torch.cuda.synchronize()
