#!/bin/bash
# Round-2 fourth check: the cluster solve kernel. Full GPU tests, A/B against the two-kernel path, launch lists.
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))s" | tee -a $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
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
run c4 --workload c4 $Q
run c1 --workload c1 $Q
run c2 --workload c2 $Q
run c3 --workload c3 $Q
run c4shard --workload c4 --total 12500000 $Q
PTP_CLUSTER_SOLVE=0 run c4shard_nocl --workload c4 --total 12500000 $Q
PTP_CLUSTER_SOLVE=0 run c1_nocl --workload c1 $Q
run c1_nograph --workload c1 $Q --graph off
run c4shard_graph --workload c4 --total 12500000 $Q --graph on
for WL in c4 c1; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_${WL}.csv \
    python bench.py --workload $WL --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --min-time 0 --graph off > $O/ncu_bench_${WL}.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_solve_cluster" -s 4 -c 1 -f -o $O/full_cluster \
    python bench.py --workload c4 --total 12500000 --no-e2e --no-cpu-baseline --min-time 0 --graph off --steps 3 --warmup 3 > $O/ncu_full_cluster.log 2>&1; echo "ncu cluster rc=$?"
echo "total t=$(( $(date +%s)-S ))s"
