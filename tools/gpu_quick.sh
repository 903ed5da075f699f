#!/bin/bash
# quick GPU check: selected tests + selected bench workloads. usage: gpu_quick.sh "<pytest -k expr>" "<workloads>"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "$1" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_quick.log
for WL in $2; do
  timeout 600 python bench.py --workload $WL --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_q_$WL.log 2>&1; echo "bench $WL rc=$?"
  tail -1 gpurun_out/bench_q_$WL.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  value %.3e  ms/step %.4f  k1 %.4f ms  frac %s  solve %.4f ms' % (d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'], d['phases_ms_per_step']['solve_node_field']))
except Exception as e: print('  parse fail', e)
"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_q_${WL}.csv \
    python bench.py --workload $WL --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_q_${WL}.log 2>&1
done
