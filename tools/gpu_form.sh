#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
PTP_FFT_FORM_ROWS=1 timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 90 -k "(4096_node_rows and 1-1) or fine_grid_full_size_properties-1" > gpurun_out/pytest_form.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))s"; tail -2 gpurun_out/pytest_form.log
for F in 0 1; do
PTP_FFT_FORM_ROWS=$F timeout 60 python bench.py --workload c5 --steps 100 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_form_$F.log 2>&1; echo "form=$F rc=$?"
tail -1 gpurun_out/bench_form_$F.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('  ms/step %.4f k1 %.4f solve %.4f launches %d' % (d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['phases_ms_per_step']['solve_node_field'], d['gpu_launches']))"
done
echo "total t=$(( $(date +%s)-S ))s"
