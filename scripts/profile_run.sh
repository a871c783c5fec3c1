#!/bin/bash
# One profiling pass on the GPU box (run through gpurun from the repo root): plain bench lines first, then the
# ncu launch list and one `--set full` capture per kernel of the SAME short command, each only after the plain
# command exited 0.  Artefacts land in gpurun_out/; scripts/make_profile_summary.py <tag> turns them into profiles/.
set -u
mkdir -p gpurun_out
if [ "${ONLY_K2:-0}" != 1 ]; then
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err || { echo "bench failed"; tail -5 gpurun_out/bench.err; exit 1; }
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
fi
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --latency-iters 8"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; exit 1; }
[ "${ONLY_K2:-0}" != 1 ] && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
cap() { # name, demangled-name regex, launches of that kernel to skip first
    ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip ${3:-0} -c 1 -f -o gpurun_out/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
    ls -la gpurun_out/prof_$1.ncu-rep
}
[ "${ONLY_K2:-0}" != 1 ] && cap k2store 'k2_peaks_fast<\(int\)8, \(int\)8, \(bool\)1>'
cap k2 'k2_peaks_fast<\(int\)8, \(int\)8, \(bool\)0>' 24   # past the one-frame launches of the three latency loops (8 calls each): a 64-frame launch
if [ "${ONLY_K2:-0}" != 1 ]; then
cap k1 'k1_replicate_chw'
cap k3 'k3_limbs'
fi
tail -c 600 gpurun_out/bench.json
