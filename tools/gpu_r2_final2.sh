#!/bin/bash
# (historical: PTP_SCATTER_FORM selected between variants of the hot form that existed when this ran; one form is shipped, the switch is gone)
# Final single-GPU check of the round-2 tree after the hot-species work: all GPU tests (both hot forms are parametrised inside),
# smoke, the driver's default bench line + reference arm, every named configuration, the hot-species workloads with both ways
# of grouping a warp's rings (PTP_SCATTER_FORM), one full ncu capture of each hot form.
mkdir -p gpurun_out/r2final2
O=gpurun_out/r2final2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/gpu.txt 2>&1
S=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))s" | tee -a $O/pytest_gpu.log
tail -14 $O/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log; tail -2 $O/smoke.log
show() {
  tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  value %.3e  ms/step %.4f  k1 %.4f ms  frac %.3f  solve %.4f ms  e2e %s e2ed %s cpu %s graph %s batches %s traffic %s sorts %s hot %s' % (d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'] or 0, d['phases_ms_per_step']['solve_node_field'], d['e2e'] and '%.3e' % d['e2e']['value'], d.get('e2e_from_density') and '%.3e' % d['e2e_from_density']['value'], d.get('cpu_baseline') and '%.3e' % d['cpu_baseline']['value'], d['timing'].get('graph_replay'), d['timing']['batches'], d['roofline'].get('traffic'), d['tuning']['sorts_in_run_rank0'], d['tuning']['hot_form_in_use_rank0']))
except Exception as e: print('  parse fail', e)
"
}
run() { # name, args...
  local name=$1; shift
  local T0=$(date +%s)
  timeout 300 python bench.py "$@" > $O/bench_$name.log 2>&1; echo "bench $name rc=$? t=$(( $(date +%s)-T0 ))s"; show $O/bench_$name.log
}
run default --steps 20 --warmup 5
Q="--no-e2e --no-cpu-baseline --min-time 0.3"
run c5e_auto_form1 --workload c5 --electrons --steps 100 $Q
PTP_SCATTER_FORM=2 run c5e_auto_form2 --workload c5 --electrons --steps 100 $Q
PTP_SCATTER_FORM=2 run c5e_hot_form2 --workload c5 --electrons --steps 100 $Q --hot on
PTP_SCATTER_FORM=2 run c4e_hot_form2 --workload c4 --electrons $Q --hot on
PTP_SCATTER_FORM=2 run c5p_hot_form2 --workload c5 --steps 100 $Q --hot on
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.log 2>&1; echo "reference arm rc=$?"; tail -1 $O/bench_reference.log | cut -c1-200
run c5 --workload c5 --steps 200
run c3 --workload c3
run c1 --workload c1
run c2 --workload c2
echo "benches done t=$(( $(date +%s)-S ))s"
PTP_SCATTER_FORM=2 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_push_deposit" -s 30 -c 1 -f -o $O/full_c5e_k1_form2 \
    python bench.py --workload c5 --electrons --hot on --steps 12 --warmup 20 --no-e2e --no-cpu-baseline --min-time 0 --graph off > $O/ncu_full_c5e_form2.log 2>&1; echo "ncu full c5e form2 rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_push_deposit" -s 30 -c 1 -f -o $O/full_c5e_k1_form1 \
    python bench.py --workload c5 --electrons --hot on --steps 12 --warmup 20 --no-e2e --no-cpu-baseline --min-time 0 --graph off > $O/ncu_full_c5e_form1.log 2>&1; echo "ncu full c5e form1 rc=$?"
ls -la $O/*.ncu-rep
echo "total t=$(( $(date +%s)-S ))s"
