"""Throughput of the paths outside the fast peak kernel: the Python variant (k = 25, zero border) and the C++ variant with k = 25,
both at x8 (replication-aware generic kernel), and a non-integer geometry (generic kernel on the materialised map)."""
import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from openpose_plus_b200 import synth, _capi as capi
from openpose_plus_b200.engine import Engine

def run(label, steps=30, batch=64, **kw):
    dev = torch.device("cuda", 0)
    conf, paf = synth.render_batch(batch, n_people=5, seed0=2000, pool=8)
    dc, dp = torch.from_numpy(conf).to(dev), torch.from_numpy(paf).to(dev)
    eng = Engine(46, 54, max_batch=batch, **kw)
    outs = [(capi.pinned_empty((batch, eng.max_humans), capi.HUMAN_DT), capi.pinned_empty((batch,), np.int32), capi.pinned_empty((batch,), np.int32)) for _ in range(3)]
    def go(n):
        infl = []
        for k in range(n):
            if len(infl) == 3: eng.wait(infl.pop(0))
            infl.append(eng.submit(dc, dp, out=outs[k % 3]))
        for t in infl: eng.wait(t)
    go(3); torch.cuda.synchronize()
    eng._check(eng.L.opp_timer_start(eng.h)); go(steps); ms = float(eng.L.opp_timer_stop(eng.h))
    print(json.dumps({"config": label, "frames_per_s": steps * batch / (ms * 1e-3), "ms_per_batch": ms / steps, "humans_frame0": int(outs[0][1][0])}))
    eng.close()

run("python variant k=25 x8", gauss_kernel_size=25, variant=capi.VARIANT_PYTHON)
run("cpp variant k=25 x8", gauss_kernel_size=25)
run("cpp variant k=17 x8 (fast kernel)", gauss_kernel_size=17)
run("cpp variant k=17 300x400 (non-integer scale)", out_h=300, out_w=400, gauss_kernel_size=17)
