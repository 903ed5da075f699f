#!/bin/bash
# Short gpurun call: rewritten sort (parity), c5 over 400 steps: planner slack x re-sort threshold, per-step K1 curves.
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 400 -k "adaptive or large_load or driver_d or losses or edge or fine_grid_full_size_properties-1" --durations=5 > gpurun_out/pytest_sort2.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))s"
tail -8 gpurun_out/pytest_sort2.log
show() {
  tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  value %.3e  ms/step %.4f  k1 %.4f ms  frac %.3f  solve %.4f ms  sorts %s' % (d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'], d['phases_ms_per_step']['solve_node_field'], d['tuning'].get('sorts_in_run_rank0')))
except Exception as e: print('  parse fail', e)
"
}
run() { n=$1; shift; timeout 300 env PTP_STEP_TIMES_FILE=gpurun_out/steps_$n.csv "$@" > gpurun_out/bench_$n.log 2>&1; echo "bench $n rc=$?"; show gpurun_out/bench_$n.log; }
B="python bench.py --workload c5 --steps 400 --warmup 3 --no-cpu-baseline --no-e2e"
run s5_nosort    PTP_PLAN_SLACK=5  $B --sort-interval 0
run s5_f005      PTP_PLAN_SLACK=5  PTP_SORT_FAR_FRACTION=0.005 PTP_SORT_CHECK_STEPS=16 $B
run s14_f005     PTP_PLAN_SLACK=14 PTP_SORT_FAR_FRACTION=0.005 PTP_SORT_CHECK_STEPS=16 $B
run s22_f005     PTP_PLAN_SLACK=22 PTP_SORT_FAR_FRACTION=0.005 PTP_SORT_CHECK_STEPS=16 $B
run s22_f001     PTP_PLAN_SLACK=22 PTP_SORT_FAR_FRACTION=0.001 PTP_SORT_CHECK_STEPS=16 $B
run s14_i32      PTP_PLAN_SLACK=14 $B --sort-interval 32
run c4           python bench.py --workload c4 --steps 200 --warmup 3 --no-cpu-baseline --no-e2e
echo "total t=$(( $(date +%s)-S ))s"
