#!/bin/bash
# Multi-GPU call (gpurun --gpus N): parity tests incl. the sharded step, then the scaling bench lines.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu_multi.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_multi.log
summ() { tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  n=%d value %.3e  ms/step %.4f  k1 %.4f ms frac %.3f  allreduce %.4f solve %.4f ms  e2e %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'], d['phases_ms_per_step']['allreduce'], d['phases_ms_per_step']['solve_node_field'], d['e2e'] and '%.3e' % d['e2e']['value']))
except Exception as e: print('  parse fail', e)
"; }
timeout 900 python bench.py --gpus 1 --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/scale_1.log 2>&1; echo "bench 1 rc=$?"; summ gpurun_out/scale_1.log
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 100 --warmup 3 > gpurun_out/scale_$n.log 2>&1; echo "bench $n rc=$?"; summ gpurun_out/scale_$n.log
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 100 --warmup 3 --allreduce nccl --no-e2e > gpurun_out/scale_${n}_nccl.log 2>&1; echo "bench $n nccl rc=$?"; summ gpurun_out/scale_${n}_nccl.log
  fi
done
