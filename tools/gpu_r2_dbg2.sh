#!/bin/bash
# 2-GPU debug: which multi-GPU cases fail or hang (strict timeouts, full logs).
mkdir -p gpurun_out/r2dbg
O=gpurun_out/r2dbg
S=$(date +%s)
export PTP_TEST_LAUNCH_TIMEOUT=150
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 -rs -k "2- or power" > $O/pytest_multi.log 2>&1; echo "pytest multi rc=$? t=$(( $(date +%s)-S ))s" | tee -a $O/pytest_multi.log
tail -5 $O/pytest_multi.log
P=29915
run() { # name nranks args...
  local name=$1; local n=$2; shift; shift
  P=$((P+1))
  local T0=$(date +%s)
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n "$@" > $O/bench_${name}_$n.log 2>&1
  echo "bench $name n=$n rc=$? t=$(( $(date +%s)-T0 ))s"; tail -1 $O/bench_${name}_$n.log | cut -c1-400
}
Q="--no-e2e --no-cpu-baseline --min-time 0.2 --no-verify"
run c5_gather 2 --workload c5 --steps 100 $Q
run c5_gather_nograph 2 --workload c5 --steps 100 $Q --graph off
run c5_nccl 2 --workload c5 --steps 100 $Q --allreduce nccl
run c4 2 --workload c4 $Q
run c3 2 --workload c3 $Q
echo "total t=$(( $(date +%s)-S ))s"
