#!/bin/bash
# A/B runs prepared at the end of round 1 (never executed): the switches that exist but have not been timed.
#   usage: gpurun --timeout 600 -- 'bash tools/gpu_next_ab.sh'            (1 GPU part)
#          gpurun --gpus 2 --timeout 600 -- 'bash tools/gpu_next_ab.sh 2' (exchange part)
N=${1:-1}
mkdir -p gpurun_out
show() { tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  n=%d value %.3e  ms/step %.4f  k1 %.4f ms  frac %s  exch %.4f  solve %.4f ms  sorts %s [%s]' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'], d['phases_ms_per_step']['allreduce'], d['phases_ms_per_step']['solve_node_field'], d['tuning'].get('sorts_in_run_rank0'), d['config']['exchange']))
except Exception as e: print('  parse fail', e)
"; }
if [ "$N" = "1" ]; then
  # the 1 M-ring free-running parity test at its full five periods (only one period was run in round 1)
  timeout 300 python -m pytest tests/test_zz_gpu_free_running_1m.py -m gpu -q -x --timeout 250 > gpurun_out/ab_pytest_1m.log 2>&1; echo "1M free-running rc=$?"; tail -2 gpurun_out/ab_pytest_1m.log
  # merged bin updates in the push kernel: parity first, then c4 / c5 / c3 with and without
  PTP_MERGE_BINS=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 250 -k "lockstep or large_load or free_running or losses or fixed_point" > gpurun_out/ab_pytest_merge.log 2>&1; echo "merge parity rc=$?"; tail -2 gpurun_out/ab_pytest_merge.log
  for WL in c4 c5 c3; do for M in 0 1; do
    PTP_MERGE_BINS=$M timeout 200 python bench.py --workload $WL --steps 200 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab_merge_${WL}_$M.log 2>&1; echo "$WL merge=$M rc=$?"; show gpurun_out/ab_merge_${WL}_$M.log
  done; done
  # graph replay of the step on small configurations
  for WL in c2 c3; do for G in "" "--graph"; do
    timeout 200 python bench.py --workload $WL --steps 400 --warmup 5 --no-cpu-baseline --no-e2e $G > gpurun_out/ab_graph_${WL}_${G:-stream}.log 2>&1; echo "$WL ${G:-stream} rc=$?"; show gpurun_out/ab_graph_${WL}_${G:-stream}.log
  done; done
else
  P=29900
  run() { n=$1; shift; P=$((P+1)); timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 300 --warmup 5 --no-e2e "$@" > gpurun_out/ab_multi_$n.log 2>&1; echo "$n rc=$?"; show gpurun_out/ab_multi_$n.log; }
  # exchange on the large grid: NCCL all-reduce (default there) against the peer-memory push now that misses are rare
  run c5_nccl --workload c5 --allreduce nccl
  run c5_peer --workload c5 --allreduce peer
  # graph replay of step pairs in peer-memory mode
  run c4_stream --workload c4
  run c4_graph --workload c4 --graph
fi
