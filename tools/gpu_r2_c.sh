#!/bin/bash
# Round-2 third check: full GPU tests + ncu --set full captures of the step's kernels (shard-sized c4, c4, c1).
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))s" | tee -a $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
Q="--no-e2e --no-cpu-baseline --min-time 0 --graph off --steps 3 --warmup 3"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_inv_field_bulk|k_fwd_thomas|k_push_deposit" -s 11 -c 3 -f -o $O/full_c4shard \
    python bench.py --workload c4 --total 12500000 $Q > $O/ncu_full_c4shard.log 2>&1; echo "ncu c4shard rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_push_deposit" -s 4 -c 1 -f -o $O/full_c4_k1 \
    python bench.py --workload c4 $Q > $O/ncu_full_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_inv_field_bulk|k_fwd_thomas|k_push_deposit" -s 16 -c 4 -f -o $O/full_c1 \
    python bench.py --workload c1 $Q > $O/ncu_full_c1.log 2>&1; echo "ncu c1 rc=$?"
ls -la $O/*.ncu-rep
echo "total t=$(( $(date +%s)-S ))s"
