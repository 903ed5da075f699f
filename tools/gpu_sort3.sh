#!/bin/bash
# Short gpurun call: default policy (wide-row slack + adaptive re-sort) on c5 over 400 steps, sort kernel times, other configs.
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 400 -k "adaptive or large_load or losses or edge or lockstep or free_running or graph or other_grids" --durations=4 > gpurun_out/pytest_sort3.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))s"
tail -7 gpurun_out/pytest_sort3.log
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
B="python bench.py --steps 400 --warmup 3 --no-cpu-baseline --no-e2e"
run d_c5       $B --workload c5
run d_c5_i32   $B --workload c5 --sort-interval 32
run d_c4       $B --workload c4 --steps 200
run d_c3       $B --workload c3
run d_c2       $B --workload c2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c5_sort.csv \
    python bench.py --workload c5 --steps 6 --warmup 3 --sort-interval 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c5_sort.log 2>&1
grep -E "k_sort|k_tile_bounds" gpurun_out/launches_c5_sort.csv | awk -F'","' '{print $5, $NF}' | cut -c1-60,200- | head -12
echo "total t=$(( $(date +%s)-S ))s"
