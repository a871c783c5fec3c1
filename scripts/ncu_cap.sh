#!/bin/bash
# One `ncu --set full` capture of one kernel launch of a command (run on the GPU box through gpurun, from the repo root):
#   bash scripts/ncu_cap.sh <name> <demangled-name regex> <launches to skip> <command...>
# Writes gpurun_out/prof_<name>_raw.csv (the --page raw export read by make_profile_summary_r2.py) and prof_<name>_sass.csv;
# the .ncu-rep itself only with KEEP_REP=1.
name=$1; regex=$2; skip=$3; shift 3
mkdir -p gpurun_out
"$@" > gpurun_out/plain_$name.log 2>&1 || { echo "plain run of $name failed"; tail -5 gpurun_out/plain_$name.log; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" --launch-skip $skip -c 1 -f -o gpurun_out/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/prof_${name}_raw.csv 2>/dev/null
# per-instruction execution counts and stall samples (SASS view): what the opcode-mix tables in profiles/ are made from
ncu -i gpurun_out/prof_$name.ncu-rep --page source --csv --print-source sass 2>/dev/null | python -c "
import csv, sys
w = csv.writer(sys.stdout)
for row in csv.reader(sys.stdin):
    w.writerow(row[:8])" > gpurun_out/prof_${name}_sass.csv
# gpurun brings back at most 64 MiB of gpurun_out/: the 17 MB reports stay on the box unless asked for
[ "${KEEP_REP:-0}" = 1 ] || rm -f gpurun_out/prof_$name.ncu-rep
ls -la gpurun_out/prof_${name}_raw.csv gpurun_out/prof_${name}_sass.csv | awk '{print $5, $9}' | tr '\n' ' '; echo
