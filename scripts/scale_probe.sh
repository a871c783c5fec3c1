#!/bin/bash
# Multi-GPU probe of bench.py on one box (run under `gpurun --gpus N`): the driver's own launch line for N = 1, 2, 4(, 8)
# plus the H2D experiments of DESIGN section 7 at the largest N.  Output: gpurun_out/scale_*.json
#   bash scripts/scale_probe.sh <max_gpus> [steps]
MAXN=${1:-4}
STEPS=${2:-20}
OUT=gpurun_out
mkdir -p $OUT
(nvidia-smi topo -m; nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv; lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket"; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302" $d/class; then echo $d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist); fi; done) > $OUT/scale_host.txt 2>&1
run() { # name N extra-args...
  name=$1; n=$2; shift 2
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps $STEPS --warmup 5 --no-cpu-baseline --no-other-configs "$@" > $OUT/scale_$name.json 2> $OUT/scale_$name.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps $STEPS --warmup 5 --no-cpu-baseline --no-other-configs "$@" > $OUT/scale_$name.json 2> $OUT/scale_$name.err
  fi
  echo "$name rc=$? $(cut -c1-120 $OUT/scale_$name.json)"
}
for n in 1 2 4 8; do
  [ $n -le $MAXN ] && run n$n $n
done
run n${MAXN}_wc $MAXN --input-memory wc
run n${MAXN}_noaff $MAXN --no-affinity
OPP_H2D_SPLIT=1 run n${MAXN}_split $MAXN
OPP_INGEST_MAX=64 run n${MAXN}_smpull $MAXN
run n${MAXN}_skel $MAXN --stream-skeleton-only
