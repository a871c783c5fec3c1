import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from openpose_plus_b200 import synth
from openpose_plus_b200.engine import Engine
dev = torch.device("cuda", 0)
conf, paf = synth.render_batch(64, n_people=32, seed0=2000, pool=8)
dc, dp = torch.from_numpy(conf).to(dev), torch.from_numpy(paf).to(dev)
eng = Engine(46, 54, max_batch=64, max_humans=256, n_slots=1)
for i in range(4):
    h, c, f = eng.process(dc, dp)
print("humans", c[:8].tolist(), "flags", np.unique(f).tolist(), file=sys.stderr)
