import os, sys, json, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from openpose_plus_b200 import synth, _capi as capi
from openpose_plus_b200.engine import Engine
dev = torch.device("cuda", 0)
conf, paf = synth.render_batch(32, n_people=5, seed0=2000, pool=8)
dc, dp = torch.from_numpy(conf).to(dev), torch.from_numpy(paf).to(dev)
eng = Engine(46, 54, max_batch=32)
cu = torch.empty((32, 368, 432, 19), device=dev); pu = torch.empty((32, 368, 432, 38), device=dev)
st = torch.cuda.Stream()
def k(): 
    eng._check(eng.L.opp_resize_pair_device(eng.h, dc.data_ptr(), dp.data_ptr(), 32, cu.data_ptr(), pu.data_ptr(), capi.LAYOUT_HWC, st.cuda_stream))
for _ in range(3): k()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(10): k()
e1.record(st); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("HWC resize pair, 32 frames: %.3f ms -> %.0f GB/s" % (ms, 32 * 36812880 / ms / 1e6))
# full pipeline with HWC maps
def go(n):
    for i in range(n):
        eng.process(dc, dp, conf_up=cu, paf_up=pu, up_layout=capi.LAYOUT_HWC)
go(3); torch.cuda.synchronize(); t0 = time.perf_counter(); go(20); torch.cuda.synchronize()
print("pipeline with HWC maps: %.0f frames/s" % (20 * 32 / (time.perf_counter() - t0)))
