#!/bin/bash
# single-GPU A/B of library builds / tunings on the c4 workload: lines "name k1_ms frac ms/step"
mkdir -p gpurun_out
summ() { tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  value %.3e  ms/step %.4f  k1 %.4f ms frac %.3f  solve %.4f ms' % (d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'], d['phases_ms_per_step']['solve_node_field']))
except Exception as e: print('  parse fail', e)
"; }
run() { n=$1; lib=$2; shift 2; if [ -n "$lib" ]; then export PTP_LIB=$PWD/$lib; else unset PTP_LIB; fi
  timeout 600 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ab1_$n.log 2>&1; echo "$n rc=$?"; summ gpurun_out/ab1_$n.log; }
run pf2 ""
run pf0 build/pf0/libptp.so
run pf4 build/pf4/libptp.so
run pf2_t256r8 "" --threads 256 --window 64 --rings 8
run pf2_r8 "" --rings 8
run pf2_fixed56 "" --deposit fixed --window 56
