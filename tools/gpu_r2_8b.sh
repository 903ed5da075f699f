#!/bin/bash
# 8-GPU scaling lines of the final exchange kernel: c4 at 8 and 4 (default line incl. parity / e2e / cpu baseline), c5 at 8 and 4, fused mode at 8.
mkdir -p gpurun_out/r2m8b
O=gpurun_out/r2m8b
S=$(date +%s)
summ() { tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    ph=d['phases_ms_per_step']
    print('  n=%d value %.3e  ms/step %.4f  k1 %.4f frac %.3f  exch %.4f solve %.4f  e2e %s parity %s graph %s [%s]' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'] or 0, ph['allreduce'], ph['solve_node_field'], d['e2e'] and '%.3e' % d['e2e']['value'], d.get('parity') and d['parity']['ok'], d['timing'].get('graph_replay'), d['config']['exchange'][:40]))
except Exception as e: print('  parse fail', e)
"; }
P=29515
run() { # name nranks args...
  local name=$1; local n=$2; shift; shift
  P=$((P+1))
  local T0=$(date +%s)
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n "$@" > $O/bench_${name}_$n.log 2>&1
  echo "bench $name n=$n rc=$? t=$(( $(date +%s)-T0 ))s"; summ $O/bench_${name}_$n.log
}
Q="--no-e2e --no-cpu-baseline --min-time 0.3"
run c4 8 --workload c4 --steps 20 --warmup 5
run c4 4 --workload c4 --steps 20 --warmup 5
run c5 8 --workload c5 --steps 100 $Q
run c5 4 --workload c5 --steps 100 $Q --no-verify
run c4_peer 8 --workload c4 $Q --allreduce peer
echo "total t=$(( $(date +%s)-S ))s"
