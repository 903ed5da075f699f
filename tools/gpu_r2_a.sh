#!/bin/bash
# Round-2 first check: GPU tests of the row-limited solve + new parity cases, bench lines of every named config, A/B of the whole-grid solve.
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/gpu.txt 2>&1
free -g >> $O/gpu.txt; nproc >> $O/gpu.txt
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))s" | tee -a $O/pytest_gpu.log
tail -14 $O/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
tail -2 $O/smoke.log
show() {
  tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  value %.3e  ms/step %.4f  k1 %.4f ms  frac %.3f  solve %.4f ms  e2e %s e2ed %s cpu %s sorts %s batches %s' % (d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'] or 0, d['phases_ms_per_step']['solve_node_field'], d['e2e'] and '%.3e' % d['e2e']['value'], d.get('e2e_from_density') and '%.3e' % d['e2e_from_density']['value'], d.get('cpu_baseline') and '%.3e' % d['cpu_baseline']['value'], d['tuning'].get('sorts_in_run_rank0'), d['timing']['batches']))
except Exception as e: print('  parse fail', e)
"
}
run() { # name, args...
  local name=$1; shift
  local T0=$(date +%s)
  timeout 400 python bench.py "$@" > $O/bench_$name.log 2>&1; echo "bench $name rc=$? t=$(( $(date +%s)-T0 ))s"; show $O/bench_$name.log
}
run c4 --workload c4
run c1 --workload c1 --min-time 0.3
run c2 --workload c2 --min-time 0.3
run c3 --workload c3 --min-time 0.3
run c5 --workload c5 --steps 200 --min-time 0.3 --no-e2e
run c4shard --workload c4 --total 12500000 --no-e2e --no-cpu-baseline --min-time 0.3
PTP_FULL_SOLVE=1 run c4shard_full --workload c4 --total 12500000 --no-e2e --no-cpu-baseline --min-time 0.3
PTP_FULL_SOLVE=1 run c5_full --workload c5 --steps 200 --min-time 0.3 --no-e2e
PTP_FULL_SOLVE=1 run c1_full --workload c1 --min-time 0.3 --no-e2e --no-cpu-baseline
run c1_graph --workload c1 --min-time 0.3 --no-e2e --no-cpu-baseline --graph on
run c2_graph --workload c2 --min-time 0.3 --no-e2e --no-cpu-baseline --graph on
run c4e --workload c4 --electrons --steps 100 --min-time 0.3 --no-e2e --no-cpu-baseline
run c5e --workload c5 --electrons --steps 100 --min-time 0.3 --no-e2e --no-cpu-baseline
for WL in c4 c5 c1; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_${WL}.csv \
    python bench.py --workload $WL --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --min-time 0 > $O/ncu_bench_${WL}.log 2>&1
done
echo "total t=$(( $(date +%s)-S ))s"
