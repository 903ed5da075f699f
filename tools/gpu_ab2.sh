#!/bin/bash
mkdir -p gpurun_out
summ() { tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  n=%d value %.3e  ms/step %.4f  k1 %.4f ms frac %.3f  exch %.4f solve %.4f ms ' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'], d['phases_ms_per_step']['allreduce'], d['phases_ms_per_step']['solve_node_field']), d.get('phases_ms_per_step_per_rank[whole,push,exchange,solve]'))
except Exception as e: print('  parse fail', e)
"; }
for rep in 1; do for mode in peer peernofence nccl; do
  if [ $mode = peernofence ]; then export PTP_PEER_NOFENCE=1; else unset PTP_PEER_NOFENCE; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 5 --allreduce ${mode%nofence} --no-e2e > gpurun_out/ab_${mode}_$rep.log 2>&1; echo "$mode $rep rc=$?"; summ gpurun_out/ab_${mode}_$rep.log
done; done
