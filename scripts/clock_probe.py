"""Clocks / power while K1 (HBM stores) and K2 (FP32 issue) run alone and together."""
import os, sys, time, subprocess, threading
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
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
def sample(label, plan, secs=1.5):
    rows = []
    proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown", "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
    th = threading.Thread(target=lambda: [rows.append(l.strip()) for l in proc.stdout], daemon=True); th.start()
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < secs:
        for r in range(20):
            for fn, si in plan: fn(n, streams[si])
            n += 1
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    proc.terminate()
    time.sleep(0.1)
    vals = [[x.strip() for x in r.split(",")] for r in rows[3:]]
    sm = sorted(float(v[0]) for v in vals); pw = sorted(float(v[2]) for v in vals)
    cap = sum(1 for v in vals if v[3].startswith("Active"))
    print("%-24s %.3f ms/round  sm_mhz med %.0f min %.0f  power med %.0f max %.0f W  sw_power_cap active in %d/%d samples" % (label, dt / n * 1e3, sm[len(sm)//2], sm[0], pw[len(pw)//2], pw[-1], cap, len(vals)))
sample("K1 alone", [(k1, 0)])
sample("K2 alone", [(k2, 3)])
sample("K1 + K2", [(k1, 0), (k2, 3)])
