#!/bin/bash
# usage: gpu_multi_b.sh N   (run under gpurun --gpus N): multi-GPU parity tests + scaling lines for c4 and c5
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest multi rc=$?"; tail -2 gpurun_out/pytest_multi_$N.log
summ() { tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  n=%d value %.3e  ms/step %.4f  k1 %.4f ms frac %.3f  exch %.4f solve %.4f ms  e2e %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'], d['phases_ms_per_step']['allreduce'], d['phases_ms_per_step']['solve_node_field'], d['e2e'] and '%.3e' % d['e2e']['value']))
except Exception as e: print('  parse fail', e)
"; }
P=29515
for WL in c4 c5; do
 for n in 1 2 4 8; do
  [ $n -gt $N ] && continue
  P=$((P+1))
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --workload $WL --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/scale_${WL}_$n.log 2>&1
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --workload $WL --steps 100 --warmup 5 > gpurun_out/scale_${WL}_$n.log 2>&1
  fi
  echo "$WL n=$n rc=$?"; summ gpurun_out/scale_${WL}_$n.log
 done
done
P=$((P+1))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload c5 --steps 100 --warmup 5 --allreduce nccl --no-e2e > gpurun_out/scale_c5_${N}_nccl.log 2>&1
echo "c5 nccl n=$N rc=$?"; summ gpurun_out/scale_c5_${N}_nccl.log
P=$((P+1))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload c4 --steps 100 --warmup 5 --graph --no-e2e > gpurun_out/scale_c4_${N}_graph.log 2>&1
echo "c4 graph n=$N rc=$?"; summ gpurun_out/scale_c4_${N}_graph.log
