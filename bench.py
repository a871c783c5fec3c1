#!/usr/bin/env python
"""Headline benchmark: post-process frames/sec @368x432 COCO-18 (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One process per GPU (torchrun for N > 1); frames shard across ranks with no data-path collective.

  value  frames/s of BASELINE.json configs[1] (batches of 64 synthetic 368x432 frames: resize + smooth/NMS/peaks + limb
         scoring + matching + assembly), feature maps already resident in HBM, CUDA-event time over all slot streams, max
         over ranks.  A step is 16 such batches (1024 frames) per GPU, weak scaling.  The up-sampled maps ARE
         materialised to HBM in this number (the reference keeps them as members, src/paf.cpp:74-75); "fused" reports
         the skeleton-only mode (the C++ paf_processor contract) beside it.
  e2e    BASELINE.json configs[4] through the product's streaming driver: a 4096-frame stream in pinned HOST memory,
         contiguous shards over the N ranks (sharding.process_stream: per-GPU slots / streams, H2D copies inside), the
         skeletons gathered on the host into one buffer on rank 0 (sharding.HostGather), wall clock from barrier to
         barrier over K passes, max over ranks.  A pass is the WHOLE stream on ALL ranks (strong scaling).  Beside it:
         the H2D ceiling of the same buffers measured in the same run on all ranks at once (plain cudaMemcpyAsync, no
         kernels) and the fraction of it the stream reaches.
  --impl reference   the reference's own unmodified src/paf.cpp (oracle/_ref, -O3 -ffast-math like its release build)
         on all host cores, rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FEAT_H, FEAT_W, STRIDE, KSIZE = 46, 54, 8, 17
OUT_H, OUT_W = FEAT_H * STRIDE, FEAT_W * STRIDE
BATCH = 64
STEP_BATCHES = 16           # a step = 16 batches of 64 = 1024 frames per GPU
STREAM = 4096               # BASELINE.json configs[4]
PEOPLE = 5
POOL = 16                   # distinct rendered frames the synthetic stream is drawn from
# algorithmic bytes per frame
K1_BYTES_PER_FRAME = 4 * 57 * (FEAT_H * FEAT_W + OUT_H * OUT_W)  # 36 812 880
IN_BYTES_PER_FRAME = 4 * 57 * FEAT_H * FEAT_W                     # 566 352: what crosses PCIe per frame
METRIC = "post-process frames/sec @368x432 COCO-18"
WORKLOAD = "batch 64 synthetic 368x432 maps (46x54 features, 19 heat + 38 PAF, %d people), gauss 17" % PEOPLE


def ncu_csv(name):
    """{metric: (value, unit)} of a committed `ncu --page raw --csv` capture (profiles/), or {}."""
    import csv
    try:
        rows = list(csv.reader(open(os.path.join(ROOT, "profiles", name))))
        return {h: (v, u) for h, v, u in zip(rows[0], rows[2], rows[1])}
    except Exception:
        return {}


def ncu_traffic(name):
    """dram read + write bytes per launch from the committed ncu capture of that kernel, or None."""
    d = ncu_csv(name)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        return sum(float(d[k][0].replace(",", "")) * scale[d[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    except Exception:
        return None


def ncu_value(name, key):
    try:
        return float(ncu_csv(name)[key][0].replace(",", ""))
    except Exception:
        return None


def first_profile(*names):
    for n in names:
        if os.path.exists(os.path.join(ROOT, "profiles", n)):
            return n
    return names[-1]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed regions (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, windows=()):
        """windows = [(t0, t1)] perf_counter intervals of the timed regions: samples inside them are counted separately."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, sm_in, mx, reasons = [], [], [], set()
        for t, r in self.rows:
            try:
                inside = any(a <= t <= b for a, b in windows)
                sm.append(float(r[1])), mx.append(float(r[2]))
                if inside:
                    sm_in.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active") and (inside or not windows):
                        reasons.add(name)
            except Exception:
                continue
        use = sm_in if sm_in else sm
        return {"sm_mhz": statistics.median(use) if use else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "samples_inside_timed_regions": len(sm_in)}


def render_pool(people=PEOPLE, fh=FEAT_H, fw=FEAT_W, n=POOL, seed0=1000, floor=None, salt=None):
    """`n` distinct rendered frames.  floor=(lo, hi) puts a uniform noise floor under the heat maps: with lo..hi around
    the 0.05 threshold every block of the peak kernel is active AND the noise itself yields ~7000 peaks per frame (a
    grouping stress: 2.9 M candidate pairs per frame).  salt=(base, spike) is the peak kernel's worst case on its own: a
    floor of `base` with one cell in every 3x3 at `spike` > threshold, so that every block must be computed while the
    smoothed maps stay below the threshold away from the joints (no extra peaks, grouping load unchanged)."""
    from openpose_plus_b200 import synth
    conf, paf = synth.render_batch(n, n_people=people, feat_h=fh, feat_w=fw, stride=STRIDE, seed0=seed0)
    if floor:
        rng = np.random.default_rng(seed0 + 77)
        conf = np.maximum(conf, rng.uniform(floor[0], floor[1], conf.shape).astype(np.float32))
    if salt:
        lattice = np.full((fh, fw), salt[0], np.float32)
        lattice[1::3, 1::3] = salt[1]
        conf = np.maximum(conf, lattice[None, None])
    return np.ascontiguousarray(conf), np.ascontiguousarray(paf)


def stream_index(n=STREAM):
    """Which pool frame stands at each position of the synthetic stream (every 64-frame batch differs from its neighbours)."""
    f = np.arange(n)
    return (f * 5 + f // BATCH) % POOL


def make_ring(pool, n_batches):
    conf, paf = pool
    ring = []
    for b in range(n_batches):
        idx = [(b * 5 + i) % POOL for i in range(BATCH)]
        ring.append((np.ascontiguousarray(conf[idx]), np.ascontiguousarray(paf[idx])))
    return ring


def make_inputs(n_batches):
    """`n_batches` distinct 64-frame batches of configs[1] (what the probe scripts under scripts/ feed the engine)."""
    return make_ring(render_pool(), n_batches)


def run_reference(args, rank):
    """The reference's own CPU implementation on all host cores (rank 0 only)."""
    if rank != 0:
        return
    from oracle.oracle import ref_available
    cores = os.cpu_count() or 1
    kind = "reference" if ref_available(fast=True) else "port"
    conf, paf = make_ring(render_pool(), 1)[0]
    geom = (FEAT_H, FEAT_W, OUT_H, OUT_W, KSIZE)
    if kind == "reference":
        from oracle.oracle import ReferencePool
        pool = ReferencePool(geom, min(cores, BATCH), fast=True)  # processors are built once, like a long-running caller would

        def step(n):
            return pool.run(conf[:n], paf[:n])[0]
    else:
        from oracle.oracle import Oracle
        orc = Oracle(*geom)

        def step(n):
            t0 = time.perf_counter()
            for i in range(n):
                orc.run(conf[i], paf[i], lazy=True)
            return time.perf_counter() - t0
        cores = 1
    # bounded sample: size a step so that steps+warmup finish in about two minutes
    t1 = step(min(cores, BATCH))
    per_frame_wall = t1 / min(cores, BATCH)
    budget = 120.0 / max(1, args.steps + args.warmup)
    n = int(max(min(cores, BATCH), min(BATCH, budget / max(per_frame_wall, 1e-6))))
    for _ in range(args.warmup):
        step(n)
    t = [step(n) for _ in range(args.steps)]
    total = sum(t)
    fps = n * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD + "; reference CPU paf_processor", "frames_per_step": n},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": min(cores, n), "kind": kind,
                         "sample": "%d frames per step on %d threads, one paf_processor per thread (use_gpu=false)" % (n, min(cores, n))},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def pin_rank_to_cpus(local, n_local, bus_id):
    """Per-rank CPU affinity, set BEFORE the pinned rings are allocated (first touch next to the GPU): the GPU's
    NUMA-local CPUs where the host exposes its topology, else an even split of the visible CPUs between the ranks."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
        how = None
        base = "/sys/bus/pci/devices/%s" % bus_id.lower()
        node = -1
        try:
            node = int(open(base + "/numa_node").read())
        except Exception:
            pass
        n_nodes = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]) if os.path.isdir("/sys/devices/system/node") else 1
        if node >= 0 and n_nodes > 1:
            lst = open(base + "/local_cpulist").read().strip()
            mine = set()
            for part in lst.split(","):
                a, _, b = part.partition("-")
                mine.update(range(int(a), int(b or a) + 1))
            mine &= set(cpus)
            if mine:
                os.sched_setaffinity(0, mine)
                how = "numa node %d (%s)" % (node, lst)
        if how is None and n_local > 1 and len(cpus) >= 2 * n_local:
            per = len(cpus) // n_local
            os.sched_setaffinity(0, set(cpus[local * per:(local + 1) * per]))
            how = "even split: cpus %d-%d (host exposes one NUMA node)" % (cpus[local * per], cpus[(local + 1) * per - 1])
        return how or "none (single rank)"
    except Exception as e:  # affinity is an optimisation, never a failure
        return "unavailable: %s" % e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--slots", type=int, default=3)
    ap.add_argument("--latency-iters", type=int, default=320, help="one-frame calls per p50 latency measurement (the first 20 are warm-up)")
    ap.add_argument("--input-memory", default="pinned", choices=["pinned", "wc"], help="host memory of the e2e input rings: pinned, or pinned write-combined")
    ap.add_argument("--no-affinity", action="store_true")
    ap.add_argument("--stream-skeleton-only", action="store_true", help="e2e stream without materialising the up-sampled maps")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from openpose_plus_b200 import _capi as capi
    from openpose_plus_b200.engine import Engine
    from openpose_plus_b200.sharding import HostGather, process_stream, shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    prop = torch.cuda.get_device_properties(local)
    bus = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id) if hasattr(prop, "pci_bus_id") else ""
    affinity = "off" if args.no_affinity else pin_rank_to_cpus(local, int(os.environ.get("LOCAL_WORLD_SIZE", world)), bus)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_max(*vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def all_min(*vals):
        return [-v for v in all_max(*[-v for v in vals])]

    S = args.slots
    pool = render_pool()
    RING = 8  # 8 x 36.2 MB of distinct feature maps = 290 MB > 126 MB L2
    ring = make_ring(pool, RING)
    d_ring = [(torch.from_numpy(c).to(dev), torch.from_numpy(p).to(dev)) for c, p in ring]
    eng = Engine(FEAT_H, FEAT_W, OUT_H, OUT_W, KSIZE, max_batch=BATCH, device=local, n_slots=S)
    up = [(torch.empty((BATCH, 19, OUT_H, OUT_W), device=dev), torch.empty((BATCH, 38, OUT_H, OUT_W), device=dev)) for _ in range(S)]
    outs = [(capi.pinned_empty((BATCH, eng.max_humans), capi.HUMAN_DT), capi.pinned_empty((BATCH,), np.int32), capi.pinned_empty((BATCH,), np.int32)) for _ in range(S)]

    # ---- this rank's shard of the 4096-frame host stream (configs[4]), pinned, filled after the affinity was set
    idx = stream_index()
    lo, hi = shard_range(STREAM, rank, world)
    wc = args.input_memory == "wc"
    h_conf = capi.pinned_empty((hi - lo, 19, FEAT_H, FEAT_W), np.float32, write_combined=wc)
    h_paf = capi.pinned_empty((hi - lo, 38, FEAT_H, FEAT_W), np.float32, write_combined=wc)
    for s in range(0, hi - lo, 256):
        e = min(s + 256, hi - lo)
        h_conf[s:e] = pool[0][idx[lo + s:lo + e]]
        h_paf[s:e] = pool[1][idx[lo + s:lo + e]]
    gather = HostGather("opp_bench_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getppid() if world > 1 else os.getpid()), STREAM, eng.max_humans, rank, world)

    # ---- parity before anything is timed, on EVERY rank: frames of this rank's own shard against the CPU oracle
    # (skeletons bit for bit; on rank 0 also the materialised maps), flags all-reduced
    from oracle.oracle import Oracle
    orc = Oracle(FEAT_H, FEAT_W, OUT_H, OUT_W, KSIZE)
    d_pool = (torch.from_numpy(pool[0]).to(dev), torch.from_numpy(pool[1]).to(dev))
    pool_h, pool_c, pool_f = eng.process(d_pool[0], d_pool[1], conf_up=up[0][0], paf_up=up[0][1])
    pool_h, pool_c = pool_h.copy(), pool_c.copy()
    ok = True
    for k in range(4):
        f = int(idx[lo + (k * 997) % (hi - lo)])
        o = orc.run(pool[0][f], pool[1][f], maps=(k == 0 and rank == 0))
        g = pool_h[f, :pool_c[f]]
        ok &= pool_c[f] == o["n_humans"] and all(
            np.array_equal(np.ascontiguousarray(g["parts"][key]).view(np.uint8), np.ascontiguousarray(o["humans"]["parts"][key]).view(np.uint8)) for key in ("x", "y", "score"))
        ok &= np.array_equal(g["score"].view(np.uint32), o["humans"]["score"].view(np.uint32))
        if k == 0 and rank == 0:
            ok &= bool(np.array_equal(up[0][0][f].cpu().numpy(), o["conf_up"])) and bool(np.array_equal(up[0][1][f].cpu().numpy(), o["paf_up"]))
    parity_ranks = world if all_min(1.0 if ok else 0.0)[0] > 0.5 else 0
    if not parity_ranks:
        raise SystemExit("bench.py: CUDA path disagrees with the oracle on rank %d (or another rank); refusing to report a number" % rank)

    def stream_pass(check=False):
        """One pass of configs[4]: every rank streams its shard from pinned host memory, rank 0 gets the whole result."""
        if rank:
            gather.wait_released()
        ups = None if args.stream_skeleton_only else up
        res = process_stream(eng, h_conf, h_paf, rank, world, gather=gather, shard_only=True, up_buffers=ups)
        wait_s = 0.0
        n_h = 0
        if rank == 0:
            (humans, counts, flags) = res
            n_h = int(counts.sum())         # the consumer touches the gathered result
            if check:                       # byte for byte against the results of the distinct frames processed alone on this GPU
                assert np.array_equal(counts, pool_c[idx]) and not (flags & capi.FLAG_OVERFLOW_MASK).any()
                for f in range(STREAM):
                    n = counts[f]
                    assert np.array_equal(humans[f, :n].view(np.uint8), pool_h[idx[f], :n].view(np.uint8)), "stream frame %d differs" % f
            gather.release()
        return n_h

    stream_ok = None
    barrier()
    n_h = stream_pass(check=True)           # ranks >= 1 are checked here on hardware: their frames must equal rank 0's own results
    single_gpu_ok = None
    if rank == 0:
        stream_ok = True
        if world > 1:                       # ... and against rank 0 running the WHOLE stream by itself (device-resident copy of it)
            hh = np.zeros((STREAM, eng.max_humans), capi.HUMAN_DT)
            cc, ff = np.zeros(STREAM, np.int32), np.zeros(STREAM, np.int32)
            d_idx = torch.from_numpy(idx).to(dev)
            infl = []
            for s in range(0, STREAM, BATCH):
                if len(infl) == S:
                    eng.wait(infl.pop(0))
                sel = d_idx[s:s + BATCH]
                infl.append(eng.submit(d_pool[0][sel], d_pool[1][sel], out=(hh[s:s + BATCH], cc[s:s + BATCH], ff[s:s + BATCH])))
            for t in infl:
                eng.wait(t)
            single_gpu_ok = bool(np.array_equal(cc, gather.counts)) and all(
                np.array_equal(hh[f, :cc[f]].view(np.uint8), gather.humans[f, :cc[f]].view(np.uint8)) for f in range(STREAM))
            if not single_gpu_ok:
                raise SystemExit("bench.py: gathered multi-GPU stream differs from the single-GPU result")
    barrier()

    def run_batches(n_batches, mode):
        """mode: 'materialize' (device maps, up-sampled maps written), 'fused' (device maps, skeletons only)."""
        inflight = []
        for k in range(n_batches):
            s = k % S
            if len(inflight) == S:
                eng.wait(inflight.pop(0))
            c, p = d_ring[k % RING]
            kw = {} if mode == "fused" else {"conf_up": up[s][0], "paf_up": up[s][1]}
            inflight.append(eng.submit(c, p, out=outs[s], **kw))
        for t in inflight:
            eng.wait(t)
        return n_batches * BATCH

    windows = []

    def timed(mode, n_steps):
        run_batches(args.warmup * STEP_BATCHES, mode)
        barrier()
        l0 = eng.launch_count()
        eng._check(eng.L.opp_timer_start(eng.h))
        t0 = time.perf_counter()
        frames = run_batches(n_steps * STEP_BATCHES, mode)
        dev_ms = float(eng.L.opp_timer_stop(eng.h))
        windows.append((t0, time.perf_counter()))
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        launches = eng.launch_count() - l0
        dev_ms, wall_ms = all_max(dev_ms, wall_ms)
        return frames, dev_ms, wall_ms, launches

    def timed_stream(n_passes):
        for _ in range(max(2, min(args.warmup, 5))):
            stream_pass()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_passes):
            stream_pass()
        t_local = time.perf_counter() - t0      # rank 0's includes waiting for every rank's results (the gather)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        windows.append((t0, time.perf_counter()))
        wall_ms, local_ms = all_max(wall_ms, t_local * 1e3)
        return wall_ms, local_ms

    def gather_wait_ms(n_passes):
        """How long rank 0 waits, after finishing its own shard, for the other ranks' results to be in place."""
        waits = []
        for _ in range(n_passes):
            if rank:
                gather.wait_released()
            ups = None if args.stream_skeleton_only else up
            hu, cn, fl = gather.local()
            inflight = []
            for s in range(0, hi - lo, BATCH):
                e = min(s + BATCH, hi - lo)
                if len(inflight) == S:
                    eng.wait(inflight.pop(0))
                kw = {} if ups is None else dict(conf_up=ups[(s // BATCH) % S][0], paf_up=ups[(s // BATCH) % S][1])
                inflight.append(eng.submit(h_conf[s:e], h_paf[s:e], out=(hu[s:e], cn[s:e], fl[s:e]), **kw))
            for t in inflight:
                eng.wait(t)
            gather.publish()
            res, w = gather.collect()
            if rank == 0:
                waits.append(w * 1e3)
                gather.release()
        return statistics.median(waits) if waits else 0.0

    def h2d_ceiling(iters):
        """Plain cudaMemcpyAsync of the same pinned batches into the same staging buffers on the same streams, all ranks
        at once, no kernels: what the host and PCIe deliver to N GPUs concurrently."""
        nb = (hi - lo) // BATCH
        ms = C.c_float(0)
        eng._check(eng.L.opp_bench_h2d(eng.h, h_conf.ctypes.data, h_paf.ctypes.data, BATCH, nb, min(iters, 8), C.byref(ms)))  # warm-up
        barrier()
        eng._check(eng.L.opp_bench_h2d(eng.h, h_conf.ctypes.data, h_paf.ctypes.data, BATCH, nb, iters, C.byref(ms)))
        barrier()
        (mx,) = all_max(ms.value)
        return world * iters * BATCH * IN_BYTES_PER_FRAME / (mx * 1e-3) / 1e9, ms.value

    import ctypes as C

    # ---- p50 latency, one frame, pinned host buffers, submit -> result (wall clock around the public call), measured
    # on the otherwise idle GPU before the throughput runs
    one = (capi.pinned_empty((1, 19, FEAT_H, FEAT_W), np.float32), capi.pinned_empty((1, 38, FEAT_H, FEAT_W), np.float32))
    one[0][...] = ring[0][0][:1]
    one[1][...] = ring[0][1][:1]

    def latency_p50():
        lat = []
        for i in range(max(args.latency_iters, 1)):
            t0 = time.perf_counter()
            eng.process(one[0], one[1], out=(outs[0][0][:1], outs[0][1][:1], outs[0][2][:1]))
            lat.append((time.perf_counter() - t0) * 1e3)
        return all_max(statistics.median(lat[20:] if len(lat) > 40 else lat))[0]

    # the same call from C (opp_bench_latency loops over opp_process inside the library): what a C++ paf_processor
    # caller sees, without the interpreter's share of every call
    def latency_p50_capi():
        b = capi.Batch()
        b.conf, b.paf, b.n_frames = one[0].ctypes.data, one[1].ctypes.data, 1
        b.in_mem, b.in_layout, b.out_mem = capi.MEM_HOST, capi.LAYOUT_CHW, capi.MEM_HOST
        b.humans, b.n_humans, b.frame_flags = outs[0][0].ctypes.data, outs[0][1].ctypes.data, outs[0][2].ctypes.data
        n_it = max(args.latency_iters, 1)
        us = np.zeros(n_it, np.float32)
        eng._check(eng.L.opp_bench_latency(eng.h, C.byref(b), n_it, us.ctypes.data))
        return float(np.median(us[20:] if n_it > 40 else us)) * 1e-3

    # pageable buffers in and out: what a caller of the reference's paf_processor contract passes (any host pointer)
    def latency_p50_pageable():
        c1, p1 = np.array(one[0]), np.array(one[1])
        out = (np.zeros((1, eng.max_humans), capi.HUMAN_DT), np.zeros(1, np.int32), np.zeros(1, np.int32))
        lat = []
        for i in range(max(args.latency_iters, 1)):
            t0 = time.perf_counter()
            eng.process(c1, p1, out=out)
            lat.append((time.perf_counter() - t0) * 1e3)
        return statistics.median(lat[20:] if len(lat) > 40 else lat)

    barrier()
    lat_p50 = latency_p50()
    lat_p50_capi = latency_p50_capi()
    lat_p50_pageable = latency_p50_pageable()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    frames, dev_ms, wall_ms, launches = timed("materialize", args.steps)
    f_frames, f_dev_ms, f_wall_ms, _ = timed("fused", args.steps)
    s_wall_ms, s_local_ms = timed_stream(args.steps)
    g_wait_ms = gather_wait_ms(5)
    ceiling_gbs, _ = h2d_ceiling(max(STEP_BATCHES * 4, 64))
    clocks = sampler.stop(windows) if rank == 0 else None

    # ---- dominant kernels alone, CUDA events on the launching stream
    def time_kernel(fn, iters):
        st = torch.cuda.Stream(device=dev)  # a real (non-NULL) stream: the launch and both events share it
        for _ in range(3):
            fn(0, st.cuda_stream)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for i in range(iters):
            ev[i][0].record(st)
            fn(i, st.cuda_stream)
            ev[i][1].record(st)
        torch.cuda.synchronize()
        return statistics.mean(a.elapsed_time(b) for a, b in ev)

    def kernel_fns(e, ringd, ups):
        def k1(i, stream):
            c, p = ringd[i % len(ringd)]
            e._check(e.L.opp_resize_pair_device(e.h, c.data_ptr(), p.data_ptr(), BATCH, ups[i % S][0].data_ptr(), ups[i % S][1].data_ptr(), capi.LAYOUT_CHW, stream))

        def k2(i, stream):
            e._check(e.L.opp_peaks_device(e.h, ringd[i % len(ringd)][0].data_ptr(), None, BATCH, None, None, stream))

        def k2_store(i, stream):  # what a materialising step launches: peaks + both up-sampled tensors in one kernel
            c, p = ringd[i % len(ringd)]
            e._check(e.L.opp_peaks_device(e.h, c.data_ptr(), p.data_ptr(), BATCH, ups[i % S][0].data_ptr(), ups[i % S][1].data_ptr(), stream))
        return k1, k2, k2_store

    iters = 100
    k1, k2, k2_store = kernel_fns(eng, d_ring, up)
    k1_ms = time_kernel(k1, iters)
    k2_ms = time_kernel(k2, iters)
    k2s_ms = time_kernel(k2_store, iters)
    peak, peak_src = measured_peaks()
    gbs = lambda ms: K1_BYTES_PER_FRAME * BATCH / (ms * 1e-3) / 1e9

    # ---- p50 latency again, right after the sustained runs (clocks and power state of a busy GPU)
    lat_p50_loaded = latency_p50()

    # ---- the other BASELINE.json configurations and the dense-map variants of configs[1], short runs (device-resident
    # maps), rank 0 at N=1 only
    other = None
    dense = None
    if rank == 0 and world == 1 and not args.no_other_configs:
        from openpose_plus_b200 import synth

        def other_config(fh, fw, people, batch, materialize, steps=40, inputs=None, env=None, out_hw=None, **kw):
            oh, ow = out_hw if out_hw else (8 * fh, 8 * fw)
            conf, paf = inputs if inputs is not None else synth.render_batch(batch, n_people=people, feat_h=fh, feat_w=fw, seed0=2000, pool=8)
            dc, dp = torch.from_numpy(conf).to(dev), torch.from_numpy(paf).to(dev)
            for k_, v_ in (env or {}).items():
                os.environ[k_] = v_
            try:
                e2 = Engine(fh, fw, oh, ow, max_batch=batch, device=local, n_slots=S, **kw)
            finally:
                for k_ in (env or {}):
                    del os.environ[k_]
            ups = [(torch.empty((batch, 19, oh, ow), device=dev), torch.empty((batch, 38, oh, ow), device=dev)) for _ in range(S)] if materialize else None
            o2 = [(capi.pinned_empty((batch, e2.max_humans), capi.HUMAN_DT), capi.pinned_empty((batch,), np.int32), capi.pinned_empty((batch,), np.int32)) for _ in range(S)]

            def go(n):
                infl = []
                for k in range(n):
                    if len(infl) == S:
                        e2.wait(infl.pop(0))
                    extra = dict(conf_up=ups[k % S][0], paf_up=ups[k % S][1]) if materialize else {}
                    infl.append(e2.submit(dc, dp, out=o2[k % S], **extra))
                for t in infl:
                    e2.wait(t)
            go(5)
            torch.cuda.synchronize()
            ms = None
            for _ in range(2):  # the better of two short runs: a busy host core shows up as a slow run, not as a fast one
                e2._check(e2.L.opp_timer_start(e2.h))
                go(steps)
                t = float(e2.L.opp_timer_stop(e2.h))
                ms = t if ms is None else min(ms, t)
            assert not (o2[0][2] & capi.FLAG_OVERFLOW_MASK).any(), "capacity overflow in a bench configuration"
            kernel_ms = None
            if materialize and batch == BATCH and fh == FEAT_H and not out_hw:   # the fused peak + resize kernel of this input alone
                try:
                    kernel_ms = time_kernel(kernel_fns(e2, [(dc, dp)], ups)[2], 50)
                except capi.OppError:   # configurations the integer-scale kernel does not cover have no stand-alone entry
                    kernel_ms = None
            e2.close()
            del ups
            torch.cuda.empty_cache()
            return steps * batch / (ms * 1e-3), kernel_ms

        def pair(fh, fw, people, batch, **kw):
            m, kms = other_config(fh, fw, people, batch, True, **kw)
            s_, _ = other_config(fh, fw, people, batch, False, **kw)
            oh, ow = kw.get("out_hw") or (8 * fh, 8 * fw)
            d = {"materialised": m, "skeleton_only": s_, "unit": "frames/s",
                 "materialised_hbm_frac": m * 4 * 57 * (fh * fw + oh * ow) / 1e9 / peak}
            if kms:
                d["fused_kernel_ms"] = kms
                d["fused_kernel_hbm_frac"] = gbs(kms) / peak
            return d

        other = {
            "736x864_b32_12people": pair(92, 108, 12, 32),
            "368x432_b64_crowded_32people": pair(46, 54, 32, 64, max_humans=256),
            # the semantics of the reference's Python graph (k = 25 CDF-derived kernel, zero border,
            # pafprocess-style grouping)
            "368x432_b64_python_variant_k25": pair(46, 54, PEOPLE, 64, gauss_kernel_size=25, variant=capi.VARIANT_PYTHON),
            "368x432_b64_cpp_k25": pair(46, 54, PEOPLE, 64, gauss_kernel_size=25),
            # a non-integer scale (46x54 -> 300x400): cv::resize's 2-tap area-mode form, the generic peak kernel
            "300x400_b64_non_integer_scale": pair(46, 54, PEOPLE, 64, out_hw=(300, 400)),
        }
        # configs[1] when the block skipping of the peak kernel finds nothing to skip: (a) a uniform noise floor in
        # [0.03, 0.07] under the rendered maps (what a CNN's heat maps look like away from the joints), (b) skipping off
        # [0.03, 0.07] under the rendered maps (also ~7000 noise peaks and 2.9 M candidate pairs per frame: the limb
        # kernel's load, which the reference's CPU path pays for as well), (a') a salted floor that activates every block
        # without adding peaks, (b) skipping switched off
        dense = {
            "noise_floor_0.03_0.07": pair(46, 54, PEOPLE, 64, inputs=make_ring(render_pool(floor=(0.03, 0.07)), 1)[0], steps=10,
                                          max_peaks_per_part=512, max_cands_per_limb=4096, max_humans=256),
            "salted_floor_all_blocks_active": pair(46, 54, PEOPLE, 64, inputs=make_ring(render_pool(salt=(0.02, 0.055)), 1)[0]),
            "OPP_K2_NOSKIP": pair(46, 54, PEOPLE, 64, inputs=ring[0], env={"OPP_K2_NOSKIP": "1"}),
        }

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.oracle import Reference, ref_available
        cores = os.cpu_count() or 1
        if ref_available(fast=True):
            def cpu_fps(geom, conf, paf, target_s=8.0):
                n = min(conf.shape[0], max(cores, 8))
                th = min(cores, n)
                Reference.time_frames(geom, conf[:th], paf[:th], 1, th, fast=True)
                rep = 1
                secs, _ = Reference.time_frames(geom, conf[:n], paf[:n], rep, th, fast=True)
                while secs < target_s and rep < 64:
                    rep *= 2
                    secs, _ = Reference.time_frames(geom, conf[:n], paf[:n], rep, th, fast=True)
                return n * rep / secs, th, "%d frames x %d passes on %d threads, %.1f s" % (n, rep, th, secs)

            geom = (FEAT_H, FEAT_W, OUT_H, OUT_W, KSIZE)
            fps, th, sample = cpu_fps(geom, ring[0][0], ring[0][1])
            # single thread, one paf_processor, one frame per call, stdout silenced: >= 200 frames
            ref1 = Reference(*geom, fast=True)
            ref1.latency_ms(ring[0][0][:2], ring[0][1][:2], 2)
            lat = ref1.latency_ms(ring[0][0], ring[0][1], 200)
            from openpose_plus_b200 import synth
            big = synth.render_batch(16, n_people=12, feat_h=92, feat_w=108, seed0=2000, pool=8)
            crowd = synth.render_batch(16, n_people=32, feat_h=46, feat_w=54, seed0=2000, pool=8)
            fps_big, _, s_big = cpu_fps((92, 108, 736, 864, KSIZE), big[0], big[1], 5.0)
            fps_crowd, _, s_crowd = cpu_fps(geom, crowd[0], crowd[1], 5.0)
            cpu_baseline = {"value": fps, "unit": "frames/s", "cores": th, "kind": "reference",
                            "sample": "reference src/paf.cpp (-O3 -ffast-math), one paf_processor per thread, use_gpu=false: " + sample,
                            "single_thread_ms_p50": float(np.median(lat)), "single_thread_frames": int(len(lat)),
                            "other_configs": {"736x864_12people": {"value": fps_big, "unit": "frames/s", "sample": s_big},
                                              "368x432_crowded_32people": {"value": fps_crowd, "unit": "frames/s", "sample": s_crowd}}}
        else:
            t0 = time.perf_counter()
            n = 0
            while time.perf_counter() - t0 < 10.0:
                orc.run(ring[0][0][n % BATCH], ring[0][1][n % BATCH], lazy=True)
                n += 1
            cpu_baseline = {"value": n / (time.perf_counter() - t0), "unit": "frames/s", "cores": 1, "kind": "port",
                            "sample": "%d frames, oracle/opp_oracle.c on 1 thread" % n}

    if rank == 0:
        N = world
        passes = args.steps
        e2e_fps = STREAM * passes / (s_wall_ms * 1e-3)
        mean_humans = float(np.mean(pool_c[idx]))
        d2h = int(STREAM * 8 + 292 * mean_humans * STREAM)  # counts + flags + the humans actually found, written over PCIe by the assembly kernels
        store_prof = first_profile("r2_k2store_ncu_raw.csv", "r1_final_k2store_ncu_raw.csv")
        k2_prof = first_profile("r2_k2_ncu_raw.csv", "r1_final_k2_ncu_raw.csv")
        line = {
            "metric": METRIC, "value": N * frames / (dev_ms * 1e-3), "unit": "frames/s", "n_gpus": N, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD + ": resize + smooth/NMS/peaks + limb scoring + matching + assembly, up-sampled maps materialised in HBM",
                       "frames_per_step_per_gpu": BATCH * STEP_BATCHES, "batch": BATCH, "slots_in_flight": S,
                       "l2": "inputs cycle through a ring of %d distinct batches (%.0f MB > 126 MB L2); each batch also writes %.0f MB of up-sampled maps" % (RING, RING * BATCH * IN_BYTES_PER_FRAME / 1e6, BATCH * 57 * OUT_H * OUT_W * 4 / 1e6),
                       "sharding": "frames sharded over ranks, no collective; value = N independent shards (weak), e2e = one 4096-frame stream over all ranks (strong) with the host gather inside the timed region",
                       "cpu_affinity": affinity, "input_memory": args.input_memory},
            "wall_ms_per_step": wall_ms / args.steps,
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": STREAM * IN_BYTES_PER_FRAME, "d2h_bytes_per_step": d2h,
                    "frames_per_step": STREAM, "ms_per_step": s_wall_ms / passes, "scaling": "strong",
                    "h2d_gbs": e2e_fps * IN_BYTES_PER_FRAME / 1e9, "h2d_ceiling_gbs": ceiling_gbs,
                    "frac_of_ceiling": e2e_fps * IN_BYTES_PER_FRAME / 1e9 / ceiling_gbs,
                    "note": "BASELINE configs[4]: a 4096-frame stream in pinned host memory through sharding.process_stream on every rank (cudaMemcpyAsync H2D per 64-frame batch, %s), skeletons written by the assembly kernels into one shared pinned host buffer (sharding.HostGather) that rank 0 reads; wall clock barrier to barrier over all passes, max over ranks; h2d_ceiling_gbs = the same batches copied by all ranks at once with no kernels" % ("skeletons only" if args.stream_skeleton_only else "up-sampled maps materialised on the device")},
            "stream_4096": {"frames_s": e2e_fps, "ms_per_pass": s_wall_ms / passes, "gather_ms": g_wait_ms, "ranks": N,
                            "gathered_equals_per_frame_results": stream_ok, "gathered_equals_single_gpu_stream": single_gpu_ok,
                            "humans_per_pass": n_h},
            "fused": {"value": N * f_frames / (f_dev_ms * 1e-3), "unit": "frames/s", "note": "skeletons only (C++ paf_processor contract): up-sampled maps never written to HBM"},
            "latency_ms_p50": lat_p50,
            "latency_ms_p50_after_load": lat_p50_loaded,
            "latency_ms_p50_capi": lat_p50_capi,
            "latency_ms_p50_pageable": lat_p50_pageable,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k2_peaks_fast<8,8,STORE> (resize of the 19+38 maps fused into smooth + NMS + peak list)",
                         "achieved": gbs(k2s_ms), "peak": peak, "unit": "GB/s", "frac": gbs(k2s_ms) / peak, "traffic": ncu_traffic(store_prof),
                         "traffic_source": "profiles/%s (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one 64-frame launch)" % store_prof,
                         "algorithmic_bytes_per_launch": K1_BYTES_PER_FRAME * BATCH, "peak_source": peak_src,
                         "ms_per_launch": k2s_ms, "algorithmic_bytes_per_frame": K1_BYTES_PER_FRAME,
                         "note": "algorithmic bytes = 4*57*(h*w + H*W): feature maps read once, up-sampled maps written once; the smoothed / pooled maps never touch HBM"},
            "roofline_k1": {"bound": "hbm", "kernel": "k1_replicate_chw<8> (stand-alone resize, used when the maps are requested without fusion)",
                            "achieved": gbs(k1_ms), "peak": peak, "unit": "GB/s", "frac": gbs(k1_ms) / peak, "ms_per_launch": k1_ms,
                            "algorithmic_bytes_per_frame": K1_BYTES_PER_FRAME},
            # not a roofline row: skeleton-only peak finding reads 12 MB of feature maps per launch and is bound by FP32
            # issue slots (the separable filter), so it is described by its time and its issue-slot utilisation
            "k2_skeleton_only": {"kernel": "k2_peaks_fast<8,8> (smooth + NMS + peak list, nothing written but peak keys)", "bound": "fp32 issue",
                                 "ms_per_launch": k2_ms, "frames_per_s_kernel_alone": BATCH / (k2_ms * 1e-3),
                                 "hbm_bytes_read_per_launch": 4 * 19 * FEAT_H * FEAT_W * BATCH,
                                 "hbm_gbs": 4 * 19 * FEAT_H * FEAT_W * BATCH / (k2_ms * 1e-3) / 1e9,
                                 "issue_slot_utilisation_pct": ncu_value(k2_prof, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                 "issue_source": "profiles/%s (smsp__issue_active.avg.pct_of_peak_sustained_active, ncu --set full of one launch)" % k2_prof},
            "clocks": clocks,
            "parity_checked": True,
            "parity_ranks_ok": parity_ranks,
        }
        if other:
            line["other_configs"] = other
        if dense:
            line["dense_maps"] = dense
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    barrier()
    gather.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
