#!/bin/bash
# Single-GPU validation of the multi-species push launch + species-parallel cluster solve; full GPU tests.
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
S=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 --durations=4 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))s" | tee -a $O/pytest_gpu.log
tail -9 $O/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
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
  timeout 300 python bench.py "$@" > $O/bench_$name.log 2>&1; echo "bench $name rc=$? t=$(( $(date +%s)-T0 ))s"; show $O/bench_$name.log
}
Q="--no-e2e --no-cpu-baseline --min-time 0.3"
run c1 --workload c1 $Q
PTP_MULTI_PUSH=0 run c1_nomulti --workload c1 $Q
run c3 --workload c3 $Q
PTP_MULTI_PUSH=0 run c3_nomulti --workload c3 $Q
run c3_graph --workload c3 $Q --graph on
run c2 --workload c2 $Q
run c4 --workload c4 $Q
run c4_graph --workload c4 $Q --graph on
run c5 --workload c5 --steps 200 $Q
run c5_graph --workload c5 --steps 200 $Q --graph on
echo "total t=$(( $(date +%s)-S ))s"
