#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_fwd_dct|k_thomas_wide|k_thomas_expand|k_idct_fft_field" -s 6 -c 4 -f -o gpurun_out/solve_c5 \
    python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_solve_c5.log 2>&1
ls -la gpurun_out/
