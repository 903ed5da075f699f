#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench lines. Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x -s --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
run() { # name, args...
  n=$1; shift
  timeout 900 python bench.py "$@" > gpurun_out/bench_$n.log 2>&1; echo "bench $n rc=$?"
  tail -1 gpurun_out/bench_$n.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  value %.3e  ms/step %.4f  k1 %.4f ms  frac %s  solve %.4f ms  e2e %s' % (d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'], d['phases_ms_per_step']['solve_node_field'], d['e2e'] and '%.3e' % d['e2e']['value']))
except Exception as e: print('  parse fail', e)
"
}
run c2 --workload c2 --steps 200 --warmup 5 --no-cpu-baseline
run c4_default --steps 100 --warmup 3 --no-cpu-baseline --no-e2e
run c3 --workload c3 --steps 100 --warmup 3 --no-cpu-baseline --no-e2e
run c2_graph --workload c2 --steps 200 --warmup 5 --no-cpu-baseline --no-e2e --graph
run c3_graph --workload c3 --steps 100 --warmup 3 --no-cpu-baseline --no-e2e --graph
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c4.csv \
    python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_c4.log 2>&1
