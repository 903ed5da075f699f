#!/bin/bash
# Multi-GPU check (N = number of GPUs on the box): sharded-step parity at every rank count up to N, exchange modes A/B, c4 / c5 lines.
N=${1:-2}
mkdir -p gpurun_out/r2m$N
O=gpurun_out/r2m$N
S=$(date +%s)
nvidia-smi -L > $O/gpus.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 900 -rs > $O/pytest_multi.log 2>&1; echo "pytest multi rc=$? t=$(( $(date +%s)-S ))s" | tee -a $O/pytest_multi.log
tail -8 $O/pytest_multi.log
summ() { tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    ph=d['phases_ms_per_step']
    print('  n=%d value %.3e  ms/step %.4f  k1 %.4f frac %.3f  exch %.4f solve %.4f  e2e %s parity %s graph %s [%s]' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'] or 0, ph['allreduce'], ph['solve_node_field'], d['e2e'] and '%.3e' % d['e2e']['value'], d.get('parity') and d['parity']['ok'], d['timing'].get('graph_replay'), d['config']['exchange'][:40]))
except Exception as e: print('  parse fail', e)
"; }
P=29815
run() { # name nranks args...
  local name=$1; local n=$2; shift; shift
  P=$((P+1))
  local T0=$(date +%s)
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n "$@" > $O/bench_${name}_$n.log 2>&1
  echo "bench $name n=$n rc=$? t=$(( $(date +%s)-T0 ))s"; summ $O/bench_${name}_$n.log
}
Q="--no-e2e --no-cpu-baseline --min-time 0.3"
for n in 2 4 8; do
  [ $n -le $N ] || continue
  run c4 $n --workload c4
  run c4_peer $n --workload c4 $Q --allreduce peer --no-verify
  run c4_nccl $n --workload c4 $Q --allreduce nccl --no-verify
  run c4_nograph $n --workload c4 $Q --graph off --no-verify
  PTP_CLUSTER_SOLVE=0 run c4_nocl $n --workload c4 $Q --no-verify
  run c5 $n --workload c5 --steps 100 $Q
  run c5_gather $n --workload c5 --steps 100 $Q --allreduce gather --no-verify
  run c3 $n --workload c3 $Q --no-verify
done
run ref $N --impl reference --steps 3 --warmup 1
echo "total t=$(( $(date +%s)-S ))s"
