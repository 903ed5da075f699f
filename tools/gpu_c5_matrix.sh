#!/bin/bash
mkdir -p gpurun_out
run() { n=$1; wl=$2; shift; shift
  timeout 600 python bench.py --workload $wl --steps 30 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/bench_m_$n.log 2>&1
  tail -1 gpurun_out/bench_m_$n.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('$n: ms/step %.4f  k1 %.4f ms  frac %.3f  solve %.4f ms' % (d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'], d['phases_ms_per_step']['solve_node_field']))
except Exception as e: print('$n parse fail', e)
"
}
run c5_base c5
run c5_t256w88 c5 --threads 256 --window 88
run c4_base c4
PTP_NO_L2_PERSIST=1 run c4_nopersist c4
