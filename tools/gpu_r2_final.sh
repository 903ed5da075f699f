#!/bin/bash
# Round-end style check of the final tree on one GPU: all GPU tests, smoke, the driver's default bench line, every named configuration,
# the reference arm, launch lists and the ncu capture of K1 that profiles/k1_traffic.json is made from.
mkdir -p gpurun_out/r2final
O=gpurun_out/r2final
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/gpu.txt 2>&1
S=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 --durations=6 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))s" | tee -a $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log; tail -2 $O/smoke.log
show() {
  tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  value %.3e  ms/step %.4f  k1 %.4f ms  frac %.3f  solve %.4f ms  e2e %s e2ed %s cpu %s graph %s batches %s traffic %s' % (d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'] or 0, d['phases_ms_per_step']['solve_node_field'], d['e2e'] and '%.3e' % d['e2e']['value'], d.get('e2e_from_density') and '%.3e' % d['e2e_from_density']['value'], d.get('cpu_baseline') and '%.3e' % d['cpu_baseline']['value'], d['timing'].get('graph_replay'), d['timing']['batches'], d['roofline'].get('traffic')))
except Exception as e: print('  parse fail', e)
"
}
run() { # name, args...
  local name=$1; shift
  local T0=$(date +%s)
  timeout 400 python bench.py "$@" > $O/bench_$name.log 2>&1; echo "bench $name rc=$? t=$(( $(date +%s)-T0 ))s"; show $O/bench_$name.log
}
run default --steps 20 --warmup 5
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.log 2>&1; echo "reference arm rc=$?"; tail -1 $O/bench_reference.log | cut -c1-200
run c1 --workload c1
run c2 --workload c2
run c3 --workload c3
run c5 --workload c5 --steps 200
run c4_fixed --workload c4 --deposit fixed --no-e2e --no-cpu-baseline
for WL in c4 c5 c1; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_${WL}.csv \
    python bench.py --workload $WL --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --min-time 0 --graph off > $O/ncu_bench_${WL}.log 2>&1
done
for WL in c4 c5; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_push_deposit" -s 4 -c 1 -f -o $O/full_${WL}_k1 \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --min-time 0 --graph off > $O/ncu_full_${WL}.log 2>&1; echo "ncu full $WL rc=$?"
done
ls -la $O/*.ncu-rep
echo "total t=$(( $(date +%s)-S ))s"
