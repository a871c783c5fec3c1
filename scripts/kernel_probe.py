"""A few batches of ONE configuration in ONE mode, for ncu captures and quick A/B timing on the GPU box:

    python scripts/kernel_probe.py <config> <skel|store> [steps]

config: typical | crowded | hires | k25 | py25 | dense (typical + OPP_K2_NOSKIP=1: every block active) | x300 (300x400 output)
Prints frames/s over `steps` pipelined batches (CUDA events over all slot streams) and the peak kernel selected."""
import json, os, sys
sys.path.insert(0, os.getcwd())
cfg = sys.argv[1] if len(sys.argv) > 1 else "typical"
mode = sys.argv[2] if len(sys.argv) > 2 else "skel"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
if cfg == "dense":
    os.environ["OPP_K2_NOSKIP"] = "1"
import numpy as np, torch
from openpose_plus_b200 import synth, _capi as capi
from openpose_plus_b200.engine import Engine
fh, fw, oh, ow, people, batch, k, kw = 46, 54, 368, 432, 5, 64, 17, {}
if cfg == "crowded": people, kw = 32, dict(max_humans=256)
if cfg == "hires": fh, fw, oh, ow, people, batch = 92, 108, 736, 864, 12, 32
if cfg == "k25": k = 25
if cfg == "py25": k, kw = 25, dict(variant=capi.VARIANT_PYTHON)
if cfg == "x300": oh, ow = 300, 400
dev = torch.device("cuda", 0)
conf, paf = synth.render_batch(batch, n_people=people, feat_h=fh, feat_w=fw, seed0=2000, pool=8)
dc, dp = torch.from_numpy(conf).to(dev), torch.from_numpy(paf).to(dev)
eng = Engine(fh, fw, oh, ow, gauss_kernel_size=k, max_batch=batch, **kw)
S = 3
ups = [(torch.empty((batch, 19, oh, ow), device=dev), torch.empty((batch, 38, oh, ow), device=dev)) for _ in range(S)] if mode == "store" else None
outs = [(capi.pinned_empty((batch, eng.max_humans), capi.HUMAN_DT), capi.pinned_empty(batch, np.int32), capi.pinned_empty(batch, np.int32)) for _ in range(S)]
def go(n):
    infl, res = [], None
    for i in range(n):
        if len(infl) == S: eng.wait(infl.pop(0))
        extra = dict(conf_up=ups[i % S][0], paf_up=ups[i % S][1]) if ups else {}
        infl.append(eng.submit(dc, dp, out=outs[i % S], **extra))
    for t in infl: res = eng.wait(t)
    return res
res = go(4)
torch.cuda.synchronize()
eng._check(eng.L.opp_timer_start(eng.h))
go(steps)
ms = float(eng.L.opp_timer_stop(eng.h))
print(json.dumps({"config": cfg, "mode": mode, "peak_kernel": eng.peak_kernel(), "frames_per_s": round(steps * batch / (ms * 1e-3)), "ms_per_batch": round(ms / steps, 4),
                  "humans_frame0": int(res[1][0]), "flags_any": int(np.bitwise_or.reduce(res[2]))}))
