#!/bin/bash
# usage: gpu_multi_c.sh N : NCCL-mode parity at 2 ranks + c5 scaling lines
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 600 -k "2-nccl" > gpurun_out/pytest_multi_c.log 2>&1; echo "pytest multi rc=$?"; tail -2 gpurun_out/pytest_multi_c.log
summ() { tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  n=%d value %.3e  ms/step %.4f  k1 %.4f ms frac %.3f  exch %.4f solve %.4f ms  e2e %s  [%s]' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'], d['phases_ms_per_step']['allreduce'], d['phases_ms_per_step']['solve_node_field'], d['e2e'] and '%.3e' % d['e2e']['value'], d['config']['exchange']))
except Exception as e: print('  parse fail', e)
"; }
P=29615
for WL in c5 c4; do
  P=$((P+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload $WL --steps 100 --warmup 5 --no-e2e > gpurun_out/scalec_${WL}_$N.log 2>&1
  echo "$WL n=$N rc=$?"; summ gpurun_out/scalec_${WL}_$N.log
done
P=$((P+1))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload c4 --allreduce nccl --steps 100 --warmup 5 --no-e2e > gpurun_out/scalec_c4_nccl_$N.log 2>&1
echo "c4 nccl n=$N rc=$?"; summ gpurun_out/scalec_c4_nccl_$N.log
