"""Per-phase %globaltimer stamps of the limb kernel on ONE crowded frame (OPP_TRACE=1 python scripts/crowded_trace.py)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from openpose_plus_b200 import synth
from openpose_plus_b200.engine import Engine
dev = torch.device("cuda", 0)
people = int(sys.argv[1]) if len(sys.argv) > 1 else 32
conf, paf = synth.render_batch(1, n_people=people, seed0=2000)
dc, dp = torch.from_numpy(conf).to(dev), torch.from_numpy(paf).to(dev)
eng = Engine(46, 54, max_batch=1, max_humans=256, n_slots=1)
for i in range(3):
    h, c, f = eng.process(dc, dp)
print("humans", c.tolist(), "flags", f.tolist(), "counts", eng.debug_counts(0, 0) if False else "", file=sys.stderr)
