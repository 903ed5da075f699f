#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench lines. Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 600 python bench.py --workload c2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1; echo "bench c2 rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c4.log 2>&1; echo "bench c4 rc=$?"
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -1 gpurun_out/bench_c2.log; tail -1 gpurun_out/bench_c4.log
