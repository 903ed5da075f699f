#!/bin/bash
# ncu --set full of the push kernel and the radix-16 inverse on c5 (first two steps).
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_push_deposit|k_idct_r16_field" -s 3 -c 4 -f -o gpurun_out/full_c5_k1_r16 \
    python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_c5b.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
