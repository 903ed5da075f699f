#!/bin/bash
# Short gpurun call: adaptive re-sort + wide field window (parity, c5 over 400 steps under three policies), c4 unaffected.
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 400 -k "adaptive or lockstep or large_load or fine_grid or free_running or graph or losses or edge" --durations=5 > gpurun_out/pytest_sort.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))s"
tail -12 gpurun_out/pytest_sort.log
show() {
  tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  value %.3e  ms/step %.4f  k1 %.4f ms  frac %.3f  solve %.4f ms  sorts %s' % (d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'], d['phases_ms_per_step']['solve_node_field'], d['tuning'].get('sorts_in_run_rank0')))
except Exception as e: print('  parse fail', e)
"
}
run() { n=$1; shift; timeout 300 env "$@" > gpurun_out/bench_$n.log 2>&1; echo "bench $n rc=$?"; show gpurun_out/bench_$n.log; }
run c5_nosort   python bench.py --workload c5 --steps 400 --warmup 3 --no-cpu-baseline --no-e2e --sort-interval 0
run c5_adapt02  python bench.py --workload c5 --steps 400 --warmup 3 --no-cpu-baseline --no-e2e
run c5_adapt005 PTP_SORT_FAR_FRACTION=0.005 python bench.py --workload c5 --steps 400 --warmup 3 --no-cpu-baseline --no-e2e
run c5_first50  python bench.py --workload c5 --steps 50 --warmup 3 --no-cpu-baseline --no-e2e
run c4          python bench.py --workload c4 --steps 200 --warmup 3 --no-cpu-baseline --no-e2e
run c3          python bench.py --workload c3 --steps 200 --warmup 3 --no-cpu-baseline --no-e2e
echo "total t=$(( $(date +%s)-S ))s"
