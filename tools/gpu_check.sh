#!/bin/bash
# Short sanity call: GPU parity subset + smoke + quick lines of every workload.
mkdir -p gpurun_out
S=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "lockstep or large_load or adaptive or free_running or losses or edge or fine_grid_full_size_properties-1 or driver_d or 256-24-2-1" > gpurun_out/pytest_check.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))s"
tail -3 gpurun_out/pytest_check.log
timeout 100 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
show() {
  tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  value %.3e  ms/step %.4f  k1 %.4f ms  frac %.3f  solve %.4f ms  sorts %s' % (d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'], d['phases_ms_per_step']['solve_node_field'], d['tuning'].get('sorts_in_run_rank0')))
except Exception as e: print('  parse fail', e)
"
}
for WL in c4 c5 c3 c2; do
  timeout 200 python bench.py --workload $WL --steps 200 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_chk_$WL.log 2>&1; echo "bench $WL rc=$?"; show gpurun_out/bench_chk_$WL.log
done
echo "total t=$(( $(date +%s)-S ))s"
