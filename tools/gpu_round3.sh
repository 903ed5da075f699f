#!/bin/bash
# One gpurun call: GPU parity tests, smoke, the driver's default bench line, c5, launch lists. Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))s" | tee -a gpurun_out/pytest_gpu.log
tail -14 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
run() { # name, args...
  n=$1; shift
  S=$(date +%s)
  timeout 900 python bench.py "$@" > gpurun_out/bench_$n.log 2>&1; echo "bench $n rc=$? t=$(( $(date +%s)-S ))s"
  tail -1 gpurun_out/bench_$n.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  value %.3e  ms/step %.4f  k1 %.4f ms  frac %s  solve %.4f ms  e2e %s cpu %s' % (d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'], d['phases_ms_per_step']['solve_node_field'], d['e2e'] and '%.3e' % d['e2e']['value'], d.get('cpu_baseline') and '%.3e' % d['cpu_baseline']['value']))
except Exception as e: print('  parse fail', e)
"
}
run default
run c5 --workload c5 --steps 50 --warmup 3 --no-cpu-baseline --no-e2e
run c2 --workload c2 --steps 200 --warmup 5 --no-cpu-baseline --no-e2e
run c3 --workload c3 --steps 100 --warmup 3 --no-cpu-baseline --no-e2e
for WL in c4 c5; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${WL}.csv \
    python bench.py --workload $WL --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_${WL}.log 2>&1
done
