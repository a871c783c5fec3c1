#!/usr/bin/env python
"""Turns the artefacts of a profiling gpurun (gpurun_out/) into the tracked summaries under profiles/.

    python scripts/make_profile_summary.py <tag>        e.g. r1_final

Expects gpurun_out/{bench.json, bench_ref.json, launches.csv, prof_k1.ncu-rep, prof_k2.ncu-rep,
prof_k2store.ncu-rep, prof_k3.ncu-rep}; writes profiles/<tag>_* and profiles/README.md."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1_final"
KERNELS = [("k1", "k1_replicate_chw<8> (stand-alone resize, 64 frames x 57 planes)"), ("k2store", "k2_peaks_fast<8,8,STORE> (peaks + fused resize)"),
           ("k2", "k2_peaks_fast<8,8> (peaks only)"), ("k3", "k3_limbs (scoring + matching + assembly)")]
for k, _ in KERNELS:
    with open(os.path.join(P, "%s_%s_ncu_raw.csv" % (tag, k)), "w") as f:
        subprocess.run(["ncu", "-i", os.path.join(G, "prof_%s.ncu-rep" % k), "--page", "raw", "--csv"], stdout=f, stderr=subprocess.DEVNULL, check=True)
shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, tag + "_launches.csv"))
shutil.copy(os.path.join(G, "bench.json"), os.path.join(P, tag + "_bench.json"))
shutil.copy(os.path.join(G, "bench_ref.json"), os.path.join(P, tag + "_bench_reference.json"))

out = ["# profiles/ — round 1 (B200, sm_100a)\n"]
out.append("All captures: `gpurun` on one B200, `ncu --clock-control none`, command `python bench.py --steps 3 --warmup 3 --no-cpu-baseline "
           "--no-other-configs --latency-iters 8` (the whole pass is `scripts/profile_run.sh`), taken only after the same command exited 0 without ncu. Timings under ncu are cold-cache and serialised: compare "
           "shares, not absolutes; the bench numbers are CUDA-event timings from a plain run. Regenerate with `scripts/make_profile_summary.py`.\n")
b = json.load(open(os.path.join(P, tag + "_bench.json")))
r = json.load(open(os.path.join(P, tag + "_bench_reference.json")))
out.append("## Bench line of this round (`%s_bench.json`, plain run)\n" % tag)
out.append("| quantity | value |\n|---|---|")
out.append("| value (materialised, device-resident maps) | %.0f frames/s (%.3f ms / 64-frame step) |" % (b["value"], b["ms_per_step"]))
out.append("| fused (skeleton-only, C++ contract) | %.0f frames/s |" % b["fused"]["value"])
out.append("| e2e (pinned host maps in, skeletons out) | %.0f frames/s (H2D %.1f MB/step: PCIe-bound) |" % (b["e2e"]["value"], b["e2e"]["h2d_bytes_per_step"] / 1e6))
out.append("| p50 latency, 1 frame, pinned host buffers, through the Python engine | %.4f ms on the idle GPU, %.4f ms right after the sustained runs |"
           % (b["latency_ms_p50"], b.get("latency_ms_p50_after_load", float("nan"))))
if "latency_ms_p50_capi" in b:
    out.append("| the same call looped from C (`opp_bench_latency`) | %.4f ms |" % b["latency_ms_p50_capi"])
if "latency_ms_p50_pageable" in b:
    out.append("| the same call with pageable buffers in and out (the reference's contract: any host pointer) | %.4f ms |" % b["latency_ms_p50_pageable"])
out.append("| roofline, dominant kernel of the materialised step (peaks + resize fused) | %.0f GB/s = %.3f of measured %.1f GB/s; ncu traffic %.3f GB vs %.3f GB algorithmic |"
           % (b["roofline"]["achieved"], b["roofline"]["frac"], b["roofline"]["peak"], (b["roofline"].get("traffic") or 0) / 1e9, b["roofline"]["algorithmic_bytes_per_launch"] / 1e9))
out.append("| stand-alone resize kernel | %.0f GB/s = %.3f |" % (b["roofline_k1"]["achieved"], b["roofline_k1"]["frac"]))
out.append("| peak kernel alone, SURVEY 8(d) algorithmic bytes (12.08 MB/frame) | %.0f GB/s = %.3f (exact block skipping: most of the map is provably below the threshold) |"
           % (b["roofline_k2"]["achieved"], b["roofline_k2"]["frac"]))
out.append("| CPU reference inside the b200 run (its own src/paf.cpp, -O3 -ffast-math) | %.0f frames/s on %d threads |" % (b["cpu_baseline"]["value"], b["cpu_baseline"]["cores"]))
out.append("| `--impl reference` arm | %.0f frames/s on %d threads |" % (r["value"], r["cpu_baseline"]["cores"]))
for k, v in b.get("other_configs", {}).items():
    out.append("| %s | materialised %.0f, skeleton-only %.0f frames/s |" % (k, v["materialised"], v["skeleton_only"]))
out.append("")
lines = [l for l in open(os.path.join(P, tag + "_launches.csv")) if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    agg[(row["Kernel Name"].replace("<unnamed>::", "")[:70], row.get("Grid Size", ""))].append(v)
out.append("## Launch list (`%s_launches.csv`, gpu__time_duration.sum)\n" % tag)
out.append("| kernel | grid | launches | mean us |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    out.append("| `%s` | %s | %d | %.1f |" % (k[0], k[1], len(v), sum(v) / len(v) / 1e3))
out.append("")
out.append("## `ncu --set full` captures (`%s_*_ncu_raw.csv` = `--page raw --csv` of the .ncu-rep)\n" % tag)
out.append("| kernel | duration us | dram read | dram write | issue active % | warps active % | regs | dyn smem |\n|---|---|---|---|---|---|---|---|")
for k, name in KERNELS:
    rows = list(csv.reader(open(os.path.join(P, "%s_%s_ncu_raw.csv" % (tag, k)))))
    d = {h: (v, u) for h, v, u in zip(rows[0], rows[2], rows[1])}
    g = lambda key: " ".join(d.get(key, ("?", "")))
    out.append("| %s | %s | %s | %s | %s | %s | %s | %s |" % (name, d["gpu__time_duration.sum"][0], g("dram__bytes_read.sum"), g("dram__bytes_write.sum"),
                                                           d["smsp__issue_active.avg.pct_of_peak_sustained_active"][0][:5], d["sm__warps_active.avg.pct_of_peak_sustained_active"][0][:5],
                                                           d["launch__registers_per_thread"][0], g("launch__shared_mem_per_block_dynamic")))
out.append("")
out.append("Earlier captures of this round are kept for the optimisation history: `r1_v0_*` (first working version: separate resize, one column per "
           "lane, 86.5 k frames/s materialised / 156 k skeleton-only) and `r1_v2_k2_ncu_raw.csv` (two columns per lane, shared tap products).\n")
out.append("SASS evidence (`cuobjdump -sass openpose_plus_b200/libopp_b200.so`): `UBLKCP.S.G` + `SYNCS.ARRIVE.TRANS64` / `SYNCS.PHASECHK.TRANS64.TRYWAIT` "
           "(TMA bulk load of the PAF tile in `k3_limbs`, mbarrier completion), `UBLKCP.G.S` (TMA bulk-store resize variant), `LDGSTS` (cp.async tile "
           "staging in `k2_peaks_fast`), `STG.E.EF.128` (streaming 16-byte stores of the up-sampled maps), `ACQBULK` / `PREEXIT` (`griddepcontrol.wait` / `.launch_dependents` of the "
           "programmatic dependent launches on the latency path), `MEMBAR.SC.SYS` (system-scope fence ahead of the pinned completion word). "
           "No tensor-core instructions: no stage is a contraction.\n")
out.append("compute-sanitizer is closed on this pool (the tool answers so); memory safety is covered by capacity checks in the kernels, the overflow "
           "flags and the stage-by-stage comparison with the CPU oracle in the GPU tests (`pytest -m gpu`).\n")
open(os.path.join(P, "README.md"), "w").write("\n".join(out))
print("\n".join(out)[:1500])
