#!/bin/bash
# Round-end style check in one gpurun call: all GPU tests, smoke, the driver's default bench line, c5, launch lists, ncu --set full.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
S=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))s" | tee -a gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
show() {
  tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  value %.3e  ms/step %.4f  k1 %.4f ms  frac %.3f  solve %.4f ms  e2e %s cpu %s sorts %s' % (d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'], d['phases_ms_per_step']['solve_node_field'], d['e2e'] and '%.3e' % d['e2e']['value'], d.get('cpu_baseline') and '%.3e' % d['cpu_baseline']['value'], d['tuning'].get('sorts_in_run_rank0')))
except Exception as e: print('  parse fail', e)
"
}
T0=$(date +%s); timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench default rc=$? t=$(( $(date +%s)-T0 ))s"; show gpurun_out/bench_default.log
timeout 300 env PTP_STEP_TIMES_FILE=gpurun_out/steps_c5.csv python bench.py --workload c5 --steps 400 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c5.log 2>&1; echo "bench c5 rc=$?"; show gpurun_out/bench_c5.log
for WL in c4 c5; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${WL}.csv \
    python bench.py --workload $WL --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_${WL}.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_push_deposit<512, 4, 1|k_idct_r16_field<1>|k_fwd_dct|k_thomas_wide" -s 8 -c 4 -f -o gpurun_out/full_c5 \
    python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_c5.log 2>&1; echo "ncu full c5 rc=$?"
ls -la gpurun_out/*.ncu-rep
echo "total t=$(( $(date +%s)-S ))s"
