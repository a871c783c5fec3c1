import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from openpose_plus_b200 import synth
from openpose_plus_b200.engine import Engine
dev = torch.device("cuda", 0)
conf, paf = synth.render_batch(32, n_people=5, seed0=2000, pool=8)
dc, dp = torch.from_numpy(conf).to(dev), torch.from_numpy(paf).to(dev)
for (oh, ow, k) in [(368, 432, 17), (368, 432, 25), (300, 400, 17), (184, 216, 9), (46 * 3, 54 * 3, 7)]:
    eng = Engine(46, 54, oh, ow, k, max_batch=32, max_peaks_per_part=256)
    def go(n):
        infl = []
        for i in range(n):
            if len(infl) == 3: eng.wait(infl.pop(0))
            infl.append(eng.submit(dc, dp))
        r = None
        for t in infl: r = eng.wait(t)
        return r
    r = go(3); torch.cuda.synchronize()
    eng._check(eng.L.opp_timer_start(eng.h)); go(20); ms = float(eng.L.opp_timer_stop(eng.h))
    print(json.dumps({"out": [oh, ow], "k": k, "frames_per_s": round(20 * 32 / (ms * 1e-3)), "humans0": int(r[1][0]), "flags": int(np.bitwise_or.reduce(r[2]))}))
    eng.close()
