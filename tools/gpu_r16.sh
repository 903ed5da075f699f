#!/bin/bash
# One short gpurun call: the radix-16 inverse kernel (parity + A/B timing on c5), the push kernel after the row-range change (c4, c5).
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 400 -k "4096 or radix16 or fine_grid or other_grids or lockstep" --durations=5 > gpurun_out/pytest_r16.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))s"
tail -12 gpurun_out/pytest_r16.log
show() {
  tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  value %.3e  ms/step %.4f  k1 %.4f ms  frac %s  solve %.4f ms' % (d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'], d['phases_ms_per_step']['solve_node_field']))
except Exception as e: print('  parse fail', e)
"
}
for R in 0 1; do
  PTP_FFT_R16=$R timeout 300 python bench.py --workload c5 --steps 100 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c5_r16_$R.log 2>&1; echo "bench c5 r16=$R rc=$?"
  show gpurun_out/bench_c5_r16_$R.log
done
timeout 300 python bench.py --workload c4 --steps 100 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c4_q.log 2>&1; echo "bench c4 rc=$?"
show gpurun_out/bench_c4_q.log
PTP_FFT_R16=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_c5_r16.csv \
    python bench.py --workload c5 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c5_r16.log 2>&1
grep -E "k_idct|k_push_deposit<512, 4, 1" gpurun_out/launches_c5_r16.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -8
echo "total t=$(( $(date +%s)-S ))s"
