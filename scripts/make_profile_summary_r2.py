#!/usr/bin/env python
"""Turns the artefacts of scripts/profile_run_r2.sh (gpurun_out/) into the tracked round-2 summaries under profiles/:
r2_bench.json, r2_bench_reference.json, r2_launches.csv, r2_<capture>_ncu_raw.csv, r2_probe_lines.jsonl and profiles/README_r2.md.

    python scripts/make_profile_summary_r2.py"""
import collections
import csv
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = "r2"
CAPS = [("k2store", "k2_peaks_fast<8,8,STORE> (peaks + fused resize), configs[1]"), ("k2", "k2_peaks_fast<8,8> (peaks only), configs[1]"),
        ("k1", "k1_replicate_chw<8> (stand-alone resize)"), ("k3", "k3_limbs, configs[1] (5 people)"),
        ("k2_k25", "k2_peaks_fast<8,12> (k = 25, cv::GaussianBlur border), peaks only"), ("k2store_py25", "k2_peaks_fast<8,12,STORE,zero border> (Python variant, k = 25)"),
        ("k2_crowded", "k2_peaks_fast<8,8>, 32 people per frame"), ("k3_crowded", "k3_limbs, 32 people per frame"),
        ("k2_dense", "k2_peaks_fast<8,8>, every block active (OPP_K2_NOSKIP=1)"), ("k2store_dense", "k2_peaks_fast<8,8,STORE>, every block active"),
        ("k2store_hires", "k2_peaks_fast<8,8,STORE>, 736x864, batch 32"), ("k2_x300", "k2_peaks_generic, 46x54 -> 300x400 (non-integer scale)"),
        ("k1_hwc", "k1_replicate_hwc<8> (channels-last maps, the Python PostProcessor contract)")]
have = []
for k, name in CAPS:
    src = os.path.join(G, "prof_%s_raw.csv" % k)
    if os.path.exists(src) and os.path.getsize(src) > 1000:
        shutil.copy(src, os.path.join(P, "%s_%s_ncu_raw.csv" % (tag, k)))
        have.append((k, name))
for a, b in (("launches.csv", tag + "_launches.csv"), ("bench.json", tag + "_bench.json"), ("bench_ref.json", tag + "_bench_reference.json"),
             ("probe_lines.jsonl", tag + "_probe_lines.jsonl")):
    if os.path.exists(os.path.join(G, a)):
        shutil.copy(os.path.join(G, a), os.path.join(P, b))

b = json.load(open(os.path.join(P, tag + "_bench.json")))
r = json.load(open(os.path.join(P, tag + "_bench_reference.json")))
out = ["# profiles/ - round 2 (B200, sm_100a)\n"]
out.append("All captures: `gpurun` on one B200, `ncu --set full --clock-control none`, each taken only after the same command exited 0 without ncu "
           "(`scripts/profile_run_r2.sh`, `scripts/ncu_cap.sh`). The bench's own kernels are captured from `python bench.py --steps 1 --warmup 3 "
           "--no-cpu-baseline --no-other-configs --latency-iters 8`, the other configurations from `scripts/kernel_probe.py <config> <skel|store> 4`. "
           "Timings under ncu are cold-cache and serialised: compare shares, not absolutes; the bench numbers are CUDA-event timings of plain runs. "
           "Round-1 files (`r1_*`, `README.md`) are kept for the history.\n")
out.append("## Bench line (`r2_bench.json`, plain run)\n")
out.append("| quantity | value |\n|---|---|")
out.append("| value (configs[1], up-sampled maps materialised, device-resident maps) | %.0f frames/s (%.3f ms per 1024-frame step) |" % (b["value"], b["ms_per_step"]))
out.append("| fused (skeleton-only, the C++ paf_processor contract) | %.0f frames/s |" % b["fused"]["value"])
e = b["e2e"]
out.append("| e2e (configs[4]: 4096-frame stream from pinned host memory, host gather inside) | %.0f frames/s; H2D %.1f GB/s of a measured ceiling of %.1f GB/s (%.3f) |"
           % (e["value"], e["h2d_gbs"], e["h2d_ceiling_gbs"], e["frac_of_ceiling"]))
out.append("| p50 latency, one frame, pinned buffers (Python engine / C loop / pageable buffers) | %.4f / %.4f / %.4f ms |"
           % (b["latency_ms_p50"], b.get("latency_ms_p50_capi", float("nan")), b.get("latency_ms_p50_pageable", float("nan"))))
rf = b["roofline"]
out.append("| roofline, dominant kernel (peaks + fused resize) | %.0f GB/s = %.3f of the measured %.1f GB/s; ncu dram traffic %.3f GB vs %.3f GB algorithmic per launch |"
           % (rf["achieved"], rf["frac"], rf["peak"], (rf.get("traffic") or 0) / 1e9, rf["algorithmic_bytes_per_launch"] / 1e9))
out.append("| stand-alone resize kernel | %.0f GB/s = %.3f |" % (b["roofline_k1"]["achieved"], b["roofline_k1"]["frac"]))
ks = b.get("k2_skeleton_only", {})
if ks:
    out.append("| peak kernel alone (skeleton-only) | %.4f ms per 64 frames, FP32-issue bound: issue slots %.1f %% busy (ncu), HBM read %.0f GB/s |"
               % (ks["ms_per_launch"], ks.get("issue_slot_utilisation_pct") or float("nan"), ks["hbm_gbs"]))
cb = b.get("cpu_baseline") or {}
if cb:
    out.append("| CPU reference in the same run (its own src/paf.cpp, -O3 -ffast-math) | %.0f frames/s on %d threads; one thread: p50 %.2f ms per frame |"
               % (cb["value"], cb["cores"], cb.get("single_thread_ms_p50", float("nan"))))
out.append("| `--impl reference` arm | %.0f frames/s on %d threads |" % (r["value"], r["cpu_baseline"]["cores"]))
for grp in ("other_configs", "dense_maps"):
    for k, v in (b.get(grp) or {}).items():
        extra = ", fused kernel %.3f ms = %.3f of HBM" % (v["fused_kernel_ms"], v["fused_kernel_hbm_frac"]) if "fused_kernel_ms" in v else ""
        out.append("| %s | materialised %.0f (%.3f of HBM), skeleton-only %.0f frames/s%s |" % (k, v["materialised"], v["materialised_hbm_frac"], v["skeleton_only"], extra))
out.append("")
lp = os.path.join(P, tag + "_launches.csv")
if os.path.exists(lp):
    lines = [l for l in open(lp) if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        agg[(row["Kernel Name"].replace("<unnamed>::", "")[:72], row.get("Grid Size", ""))].append(v)
    tot = sum(sum(v) for v in agg.values())
    out.append("## Launch list (`r2_launches.csv`, gpu__time_duration.sum, serialised under ncu)\n")
    out.append("| kernel | grid | launches | mean us | share of kernel time |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append("| `%s` | %s | %d | %.1f | %.1f %% |" % (k[0], k[1], len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    out.append("")
out.append("## `ncu --set full` captures (`r2_<name>_ncu_raw.csv` = `--page raw --csv` of the .ncu-rep)\n")
out.append("| capture | kernel / configuration | grid | duration us | dram read | dram write | issue active % | alu pipe % | fma pipe % | warps active % | regs | dyn smem |\n|---|---|---|---|---|---|---|---|---|---|---|---|")
for k, name in have:
    rows = list(csv.reader(open(os.path.join(P, "%s_%s_ncu_raw.csv" % (tag, k)))))
    d = {h: (v, u) for h, v, u in zip(rows[0], rows[2], rows[1])}
    g = lambda key: " ".join(d.get(key, ("?", "")))
    f5 = lambda key: d.get(key, ("?",))[0][:5]
    out.append("| `r2_%s` | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (
        k, name, d.get("Grid Size", ("?",))[0], d["gpu__time_duration.sum"][0], g("dram__bytes_read.sum"), g("dram__bytes_write.sum"),
        f5("smsp__issue_active.avg.pct_of_peak_sustained_active"), f5("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        f5("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), f5("sm__warps_active.avg.pct_of_peak_sustained_active"),
        d["launch__registers_per_thread"][0], g("launch__shared_mem_per_block_dynamic")))
out.append("")
pl = os.path.join(P, tag + "_probe_lines.jsonl")
if os.path.exists(pl):
    out.append("## Probe lines (`r2_probe_lines.jsonl`: `scripts/kernel_probe.py <config> <mode> 60`, pipelined batches, CUDA events)\n")
    out.append("| configuration | mode | peak kernel | frames/s | ms per batch |\n|---|---|---|---|---|")
    for l in open(pl):
        try:
            j = json.loads(l)
        except Exception:
            continue
        out.append("| %s | %s | %s | %d | %.4f |" % (j["config"], j["mode"], j["peak_kernel"], j["frames_per_s"], j["ms_per_batch"]))
    out.append("")
open(os.path.join(P, "README_r2.md"), "w").write("\n".join(out))
print("\n".join(out)[:3000])


def opcode_mix(path, top=8):
    """Share of the warp-stall SAMPLES per opcode (where the warps of the kernel spend their time) from a prof_*_sass.csv."""
    import re
    agg = collections.Counter()
    for row in csv.reader(open(path)):
        if len(row) < 4 or not row[0].startswith("0x"):
            continue
        nums = [c for c in row[2:] if re.fullmatch(r"\d+", c.strip())]
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", row[1])
        if nums and m:
            agg[m.group(2)] += int(nums[0])
    tot = sum(agg.values()) or 1
    return ", ".join("%s %.0f %%" % (k, 100.0 * v / tot) for k, v in agg.most_common(top))


mix = ["## Where the warps spend their time (stall samples per opcode; per-instruction tables kept as `r2_<name>_sass.csv` for four of them)\n", "| capture | opcodes by share of samples |\n|---|---|"]
for k, name in have:
    src = os.path.join(G, "prof_%s_sass.csv" % k)
    if os.path.exists(src):
        if k in ("k2store", "k2_dense", "k2_crowded", "k3_crowded"):  # the per-instruction tables of the kernels the text discusses (the rest: opcode shares only)
            shutil.copy(src, os.path.join(P, "%s_%s_sass.csv" % (tag, k)))
        mix.append("| `r2_%s` | %s |" % (k, opcode_mix(src)))
with open(os.path.join(P, "README_r2.md"), "a") as f:
    f.write("\n".join(mix) + "\n")
