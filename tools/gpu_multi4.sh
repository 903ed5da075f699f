#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 600 -k "$N-" > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest multi (default) rc=$?"; tail -2 gpurun_out/pytest_multi_$N.log
PTP_PEER_FENCE=1 timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 600 -k "$N-peer" > gpurun_out/pytest_multi_${N}_fence.log 2>&1; echo "pytest multi (with fence) rc=$?"; tail -2 gpurun_out/pytest_multi_${N}_fence.log
summ() { tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  n=%d value %.3e  ms/step %.4f  k1 %.4f ms frac %.3f  exch %.4f solve %.4f ms  e2e %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'], d['phases_ms_per_step']['allreduce'], d['phases_ms_per_step']['solve_node_field'], d['e2e'] and '%.3e' % d['e2e']['value']))
except Exception as e: print('  parse fail', e)
"; }
for mode in nccl peer; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 200 --warmup 5 --allreduce $mode --no-e2e > gpurun_out/scale${N}_$mode.log 2>&1; echo "$mode rc=$?"; summ gpurun_out/scale${N}_$mode.log
done
