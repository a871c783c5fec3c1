"""Throughput of the other BASELINE.json configurations (device-resident inputs, skeleton-only and materialised)."""
import os, sys, time, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from openpose_plus_b200 import synth
from openpose_plus_b200.engine import Engine

def run(label, fh, fw, people, batch, steps=60, materialize=False, **kw):
    dev = torch.device("cuda", 0)
    conf, paf = synth.render_batch(batch, n_people=people, feat_h=fh, feat_w=fw, seed0=2000, pool=8)
    dc, dp = torch.from_numpy(conf).to(dev), torch.from_numpy(paf).to(dev)
    eng = Engine(fh, fw, max_batch=batch, **kw)
    S = 3
    ups = [(torch.empty((batch, 19, 8 * fh, 8 * fw), device=dev), torch.empty((batch, 38, 8 * fh, 8 * fw), device=dev)) for _ in range(S)] if materialize else None
    def go(n):
        infl = []
        for k in range(n):
            if len(infl) == S: eng.wait(infl.pop(0))
            extra = dict(conf_up=ups[k % S][0], paf_up=ups[k % S][1]) if materialize else {}
            infl.append(eng.submit(dc, dp, **extra))
        res = None
        for t in infl: res = eng.wait(t)
        return res
    res = go(5)
    torch.cuda.synchronize()
    eng._check(eng.L.opp_timer_start(eng.h))
    go(steps)
    ms = float(eng.L.opp_timer_stop(eng.h))
    print(json.dumps({"config": label, "frames_per_s": steps * batch / (ms * 1e-3), "ms_per_batch": ms / steps, "batch": batch,
                      "humans_frame0": int(res[1][0]), "flags_any": int(np.bitwise_or.reduce(res[2]))}))
    eng.close()

mode = sys.argv[1] if len(sys.argv) > 1 else "all"
run("368x432 5 people b64 skeleton-only", 46, 54, 5, 64)
run("368x432 5 people b64 materialised", 46, 54, 5, 64, materialize=True)
run("736x864 12 people b32 skeleton-only", 92, 108, 12, 32)
run("736x864 12 people b32 materialised", 92, 108, 12, 32, materialize=True)
run("368x432 crowded 32 people b64 skeleton-only", 46, 54, 32, 64, max_humans=256)
run("368x432 crowded 32 people b64 materialised", 46, 54, 32, 64, materialize=True, max_humans=256)
