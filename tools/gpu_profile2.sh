#!/bin/bash
mkdir -p gpurun_out
WL=${1:-c4}
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${WL}.csv \
    python bench.py --workload $WL --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_${WL}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_inv_gemm|k_fwd_thomas|k_row_bounds" -s 9 -c 3 -f -o gpurun_out/solve_${WL} \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_solve_${WL}.log 2>&1
