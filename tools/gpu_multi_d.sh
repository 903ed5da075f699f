#!/bin/bash
# 2-GPU check of the current tree: sharded-step parity (both exchange modes), c4 and c5 lines with the adaptive re-sort active.
N=${1:-2}
mkdir -p gpurun_out
S=$(date +%s)
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 300 -k "test_sharded_step_matches_single_gpu and 2-" > gpurun_out/pytest_multi_d.log 2>&1; echo "pytest multi rc=$? t=$(( $(date +%s)-S ))s"; tail -3 gpurun_out/pytest_multi_d.log
summ() { tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  n=%d value %.3e  ms/step %.4f  k1 %.4f ms frac %.3f  exch %.4f solve %.4f ms sorts %s [%s]' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'], d['phases_ms_per_step']['allreduce'], d['phases_ms_per_step']['solve_node_field'], d['tuning'].get('sorts_in_run_rank0'), d['config']['exchange']))
except Exception as e: print('  parse fail', e)
"; }
P=29715
for WL in c5 c4; do
  P=$((P+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload $WL --steps 300 --warmup 5 --no-e2e > gpurun_out/scaled_${WL}_$N.log 2>&1
  echo "$WL n=$N rc=$?"; summ gpurun_out/scaled_${WL}_$N.log
done
P=$((P+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --impl reference --steps 3 --warmup 1 > gpurun_out/scaled_ref_$N.log 2>&1
echo "reference arm n=$N rc=$?"; tail -1 gpurun_out/scaled_ref_$N.log | cut -c1-300
echo "total t=$(( $(date +%s)-S ))s"
