#!/bin/bash
# 2-GPU check of the exchange kernel (loads issued ahead of their use).
mkdir -p gpurun_out/r2dbg4
O=gpurun_out/r2dbg4
S=$(date +%s)
export PTP_TEST_LAUNCH_TIMEOUT=200
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 280 -k "sharded and 2" > $O/pytest_multi.log 2>&1; echo "pytest multi rc=$? t=$(( $(date +%s)-S ))s"; tail -3 $O/pytest_multi.log
P=29715
run() { # name nranks args...
  local name=$1; local n=$2; shift; shift
  P=$((P+1))
  local T0=$(date +%s)
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n "$@" > $O/bench_${name}_$n.log 2>&1
  echo "bench $name n=$n rc=$? t=$(( $(date +%s)-T0 ))s"; tail -1 $O/bench_${name}_$n.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); ph=d['phases_ms_per_step']
print('  value %.3e ms/step %.4f k1 %.4f exch %.4f solve %.4f parity %s' % (d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], ph['allreduce'], ph['solve_node_field'], d.get('parity') and d['parity']['ok']))"
}
Q="--no-e2e --no-cpu-baseline --min-time 0.2"
run c4 2 --workload c4 $Q
run c5 2 --workload c5 --steps 100 $Q --no-verify
echo "total t=$(( $(date +%s)-S ))s"
