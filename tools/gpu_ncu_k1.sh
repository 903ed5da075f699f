#!/bin/bash
# ncu --set full of the push kernel for one workload: gpu_ncu_k1.sh <workload>
mkdir -p gpurun_out
WL=${1:-c5}
ncu --set full --clock-control none --import-source on -k regex:"k_push_deposit" -s 5 -c 1 -f -o gpurun_out/k1_${WL} \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_k1_${WL}.log 2>&1
ls -la gpurun_out/
