#!/bin/bash
# ncu captures: launch list of a short bench run + full-set capture of the push kernel and the solver kernels.
mkdir -p gpurun_out
WL=${1:-c4}
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${WL}.csv \
    python bench.py --workload $WL --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_${WL}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_push_deposit -s 4 -c 2 -f -o gpurun_out/k1_${WL} \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_k1_${WL}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_inv_gemm|k_fwd_thomas|k_row_bounds|k_node_field" -s 8 -c 4 -f -o gpurun_out/solve_${WL} \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_solve_${WL}.log 2>&1
ls -la gpurun_out/ | grep -E "ncu-rep|launches"
