#!/bin/bash
# Round-2 second check: programmatic dependent launch chain, per-parity step graphs, bulk-async inverse; A/B switches; sanitizer.
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))s" | tee -a $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
show() {
  tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  value %.3e  ms/step %.4f  k1 %.4f ms  frac %.3f  solve %.4f ms  e2e %s e2ed %s cpu %s graph %s batches %s' % (d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'] or 0, d['phases_ms_per_step']['solve_node_field'], d['e2e'] and '%.3e' % d['e2e']['value'], d.get('e2e_from_density') and '%.3e' % d['e2e_from_density']['value'], d.get('cpu_baseline') and '%.3e' % d['cpu_baseline']['value'], d['timing'].get('graph_replay'), d['timing']['batches']))
except Exception as e: print('  parse fail', e)
"
}
run() { # name, args...
  local name=$1; shift
  local T0=$(date +%s)
  timeout 400 python bench.py "$@" > $O/bench_$name.log 2>&1; echo "bench $name rc=$? t=$(( $(date +%s)-T0 ))s"; show $O/bench_$name.log
}
Q="--no-e2e --no-cpu-baseline --min-time 0.3"
run c4 --workload c4
run c1 --workload c1 --min-time 0.3
run c2 --workload c2 $Q
run c3 --workload c3 $Q
run c5 --workload c5 --steps 200 $Q
run c4shard --workload c4 --total 12500000 $Q
run c4shard_nograph --workload c4 --total 12500000 $Q --graph off
PTP_PDL=0 run c4shard_nograph_nopdl --workload c4 --total 12500000 $Q --graph off
PTP_PDL=0 run c4shard_nopdl --workload c4 --total 12500000 $Q
PTP_INV_BULK=0 run c4shard_nobulk --workload c4 --total 12500000 $Q --graph off
PTP_PDL=0 run c1_nopdl --workload c1 $Q
run c1_nograph --workload c1 $Q --graph off
PTP_PDL=0 run c1_nograph_nopdl --workload c1 $Q --graph off
PTP_PDL=0 run c5_nopdl --workload c5 --steps 200 $Q
for WL in c4 c1; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_${WL}.csv \
    python bench.py --workload $WL --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --min-time 0 --graph off > $O/ncu_bench_${WL}.log 2>&1
done
T0=$(date +%s)
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > $O/sanitize_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$? t=$(( $(date +%s)-T0 ))s"; tail -3 $O/sanitize_memcheck_smoke.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > $O/sanitize_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$? t=$(( $(date +%s)-T0 ))s"; tail -3 $O/sanitize_racecheck_smoke.log
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python __graft_entry__.py smoke > $O/sanitize_synccheck_smoke.log 2>&1; echo "synccheck smoke rc=$? t=$(( $(date +%s)-T0 ))s"; tail -3 $O/sanitize_synccheck_smoke.log
echo "total t=$(( $(date +%s)-S ))s"
