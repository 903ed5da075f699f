#!/bin/bash
# Quick A/B of the compacted cluster solve kernel.
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
S=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "populated_rows or lockstep or free_running_175 or other_grids or graph_replay" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))s" | tee -a $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
show() {
  tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  value %.3e  ms/step %.4f  k1 %.4f ms  frac %.3f  solve %.4f ms  graph %s batches %s launches %s' % (d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'] or 0, d['phases_ms_per_step']['solve_node_field'], d['timing'].get('graph_replay'), d['timing']['batches'], d['gpu_launches']))
except Exception as e: print('  parse fail', e)
"
}
run() { # name, args...
  local name=$1; shift
  local T0=$(date +%s)
  timeout 400 python bench.py "$@" > $O/bench_$name.log 2>&1; echo "bench $name rc=$? t=$(( $(date +%s)-T0 ))s"; show $O/bench_$name.log
}
Q="--no-e2e --no-cpu-baseline --min-time 0.3"
run c4shard --workload c4 --total 12500000 $Q
PTP_CLUSTER_SOLVE=0 run c4shard_nocl --workload c4 --total 12500000 $Q
run c1 --workload c1 $Q
PTP_CLUSTER_SOLVE=0 run c1_nocl --workload c1 $Q
run c2 --workload c2 $Q
PTP_CLUSTER_SOLVE=0 run c2_nocl --workload c2 $Q
run c4shard_graph --workload c4 --total 12500000 $Q --graph on
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c4shard.csv \
    python bench.py --workload c4 --total 12500000 --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --min-time 0 --graph off > $O/ncu_bench.log 2>&1
python tools/ncu_summary.py $O/launches_c4shard.csv | grep -E "cluster|push_deposit<512, 4, 1"
echo "total t=$(( $(date +%s)-S ))s"
