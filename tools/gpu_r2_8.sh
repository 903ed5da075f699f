#!/bin/bash
# 8-GPU evidence run: sharded-step parity at 4 and 8 ranks (all exchanges), c4 and c5 at 4 and 8 GPUs, exchange A/B at 8.
N=${1:-8}
mkdir -p gpurun_out/r2m8
O=gpurun_out/r2m8
S=$(date +%s)
nvidia-smi -L > $O/gpus.txt
export PTP_TEST_LAUNCH_TIMEOUT=240
export PTP_TEST_MULTI_LOG=$PWD/$O/multi_parity.jsonl
rm -f $PTP_TEST_MULTI_LOG
timeout 520 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 -rs -k "sharded and (4 or 8)" > $O/pytest_multi.log 2>&1; echo "pytest multi rc=$? t=$(( $(date +%s)-S ))s" | tee -a $O/pytest_multi.log
tail -6 $O/pytest_multi.log
summ() { tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    ph=d['phases_ms_per_step']
    print('  n=%d value %.3e  ms/step %.4f  k1 %.4f frac %.3f  exch %.4f solve %.4f  e2e %s parity %s graph %s [%s]' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'] or 0, ph['allreduce'], ph['solve_node_field'], d['e2e'] and '%.3e' % d['e2e']['value'], d.get('parity') and d['parity']['ok'], d['timing'].get('graph_replay'), d['config']['exchange'][:40]))
except Exception as e: print('  parse fail', e)
"; }
P=29615
run() { # name nranks args...
  local name=$1; local n=$2; shift; shift
  P=$((P+1))
  local T0=$(date +%s)
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n "$@" > $O/bench_${name}_$n.log 2>&1
  echo "bench $name n=$n rc=$? t=$(( $(date +%s)-T0 ))s"; summ $O/bench_${name}_$n.log
}
Q="--no-e2e --no-cpu-baseline --min-time 0.3 --no-verify"
run c4 8 --workload c4
run c4 4 --workload c4
run c5 8 --workload c5 --steps 100 --no-e2e --no-cpu-baseline --min-time 0.3
run c5 4 --workload c5 --steps 100 $Q
run c4_peer 8 --workload c4 $Q --allreduce peer
run c4_nccl 8 --workload c4 $Q --allreduce nccl
run c4_nograph 8 --workload c4 $Q --graph off
run c3 8 --workload c3 $Q
echo "total t=$(( $(date +%s)-S ))s"
