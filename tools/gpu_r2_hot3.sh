#!/bin/bash
# (historical: PTP_SCATTER_FORM selected between variants of the hot form that existed when this ran; one form is shipped, the switch is gone)
# Hot species, third way of sharing a warp's bins (warp sort + segmented sums, PTP_SCATTER_FORM=3): parity tests of all forms,
# timings on electrons (fine grid, default grid) and on ordered antiprotons, one full ncu capture.
mkdir -p gpurun_out/r2hot3
O=gpurun_out/r2hot3
S=$(date +%s)
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 --durations=5 -k "hot_species_form or losses_match" > $O/pytest_hot.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))s" | tee -a $O/pytest_hot.log
tail -6 $O/pytest_hot.log
show() {
  tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('  value %.3e  ms/step %.4f  k1 %.4f ms  frac %.3f  solve %.4f ms  sorts %s hot %s batches %s batch_ms %s' % (d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'] or 0, d['phases_ms_per_step']['solve_node_field'], d['tuning']['sorts_in_run_rank0'], d['tuning']['hot_form_in_use_rank0'], d['timing']['batches'], d['timing']['batch_ms']))
except Exception as e: print('  parse fail', e)
"
}
run() { # name, args...
  local name=$1; shift
  local T0=$(date +%s)
  timeout 200 python bench.py "$@" > $O/bench_$name.log 2>&1; echo "bench $name rc=$? t=$(( $(date +%s)-T0 ))s"; show $O/bench_$name.log
}
Q="--no-e2e --no-cpu-baseline --min-time 0.3"
export PTP_SCATTER_FORM=3
run c5e_form3 --workload c5 --electrons --steps 100 $Q --hot on
run c5e_auto --workload c5 --electrons --steps 100 $Q
run c4e_hot --workload c4 --electrons $Q --hot on
run c5p_hot --workload c5 --steps 100 $Q --hot on
echo "benches done t=$(( $(date +%s)-S ))s"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_push_deposit" -s 30 -c 1 -f -o $O/full_c5e_k1_form3 \
    python bench.py --workload c5 --electrons --hot on --steps 12 --warmup 20 --no-e2e --no-cpu-baseline --min-time 0 --graph off > $O/ncu_full_c5e.log 2>&1; echo "ncu full c5e rc=$?"
ls -la $O/*.ncu-rep
echo "total t=$(( $(date +%s)-S ))s"
