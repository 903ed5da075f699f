#!/bin/bash
# (historical: PTP_SCATTER_FORM selected between variants of the hot form that existed when this ran; one form is shipped, the switch is gone)
# Hot species (electrons on the fine grid): parity tests of the per-warp-bin form of K1, then timings of the three ways to
# push them - thread-private bins + re-sorts, per-warp bins with match.any groups (form 1), per-warp bins with tags (form 2) -
# launch list and one full ncu capture of the hot form.
mkdir -p gpurun_out/r2hot
O=gpurun_out/r2hot
S=$(date +%s)
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 --durations=5 -k "hot or losses_match or adaptive_resort" > $O/pytest_hot.log 2>&1; echo "pytest form1 rc=$? t=$(( $(date +%s)-S ))s" | tee -a $O/pytest_hot.log
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
run c5e_form1 --workload c5 --electrons --steps 100 $Q --hot on
PTP_SCATTER_FORM=2 run c5e_form2 --workload c5 --electrons --steps 100 $Q --hot on
# the faster of the two forms for everything that follows
BEST=$(python - <<PY
import json
def v(f):
    try:
        return json.loads(open(f).read().strip().splitlines()[-1])["value"]
    except Exception:
        return 0.0
print(2 if v("$O/bench_c5e_form2.log") > 1.03 * v("$O/bench_c5e_form1.log") else 1)
PY
)
echo "faster form: $BEST"
export PTP_SCATTER_FORM=$BEST
run c5e_auto --workload c5 --electrons --steps 100 $Q
run c5e_sorts --workload c5 --electrons --steps 100 $Q --hot off
run c4e_hot --workload c4 --electrons $Q --hot on
run c4e_window --workload c4 --electrons $Q --hot off
run c5p_hot --workload c5 --steps 100 $Q --hot on
echo "benches done t=$(( $(date +%s)-S ))s"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c5e.csv \
    python bench.py --workload c5 --electrons --hot on --steps 12 --warmup 20 --no-e2e --no-cpu-baseline --min-time 0 --graph off > $O/ncu_bench_c5e.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_push_deposit" -s 30 -c 1 -f -o $O/full_c5e_k1 \
    python bench.py --workload c5 --electrons --hot on --steps 12 --warmup 20 --no-e2e --no-cpu-baseline --min-time 0 --graph off > $O/ncu_full_c5e.log 2>&1; echo "ncu full c5e rc=$?"
ls -la $O/*.ncu-rep
echo "total t=$(( $(date +%s)-S ))s"
