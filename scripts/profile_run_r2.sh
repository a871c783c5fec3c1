#!/bin/bash
# Round-2 profiling pass on the GPU box (run through gpurun from the repo root).  Plain bench lines first, then the ncu
# launch list and one `--set full` capture per kernel - each only after the same command exited 0 without ncu.
# Artefacts land in gpurun_out/; scripts/make_profile_summary_r2.py turns them into profiles/r2_*.
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err || { echo "bench failed"; tail -5 gpurun_out/bench.err; exit 1; }
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs --latency-iters 8"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
cap() { # name, demangled-name regex, launches to skip, command...
    bash scripts/ncu_cap.sh "$@" | tail -1
}
# the bench's own kernels (configs[1]): fused peaks + resize, peaks alone, stand-alone resize, limbs
cap k2store 'k2_peaks_fast<\(int\)8, \(int\)8, \(bool\)1' 4 $CMD
cap k2 'k2_peaks_fast<\(int\)8, \(int\)8, \(bool\)0' 30 $CMD
cap k1 'k1_replicate_chw' 0 $CMD
cap k3 'k3_limbs' 30 $CMD
# the other configurations (scripts/kernel_probe.py: a few batches of one configuration in one mode)
cap k2_k25 'k2_peaks_fast<\(int\)8, \(int\)12, \(bool\)0, \(bool\)0' 6 python scripts/kernel_probe.py k25 skel 4
cap k2store_py25 'k2_peaks_fast<\(int\)8, \(int\)12, \(bool\)1, \(bool\)1' 6 python scripts/kernel_probe.py py25 store 4
cap k2_crowded 'k2_peaks_fast' 6 python scripts/kernel_probe.py crowded skel 4
cap k3_crowded 'k3_limbs' 6 python scripts/kernel_probe.py crowded skel 4
cap k2_dense 'k2_peaks_fast' 6 python scripts/kernel_probe.py dense skel 4
cap k2store_dense 'k2_peaks_fast' 6 python scripts/kernel_probe.py dense store 4
cap k2store_hires 'k2_peaks_fast' 6 python scripts/kernel_probe.py hires store 4
cap k2_x300 'k2_peaks_generic' 6 python scripts/kernel_probe.py x300 skel 4
cap k1_hwc 'k1_replicate_hwc' 2 python scripts/hwc_probe.py
for c in typical crowded hires k25 py25 dense x300; do for m in skel store; do python scripts/kernel_probe.py $c $m 60; done; done > gpurun_out/probe_lines.jsonl 2>&1
# experiments: limb CTAs of 128 threads, the limb kernel's shared-memory footprint, shorter peak tiles
for e in "OPP_K3_THREADS=128" "OPP_K3_THREADS=192" "OPP_K3_SMEM_MIN=56000" "OPP_K3_SMEM_MIN=75000" "OPP_K2_TH=12" "OPP_K2_TH=16"; do
  for c in typical crowded; do for m in skel store; do echo "$e $(env $e python scripts/kernel_probe.py $c $m 60)"; done; done
done > gpurun_out/experiments.txt 2>&1
tail -c 400 gpurun_out/bench.json
