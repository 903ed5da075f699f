#!/bin/bash
# 2-GPU c5 line with the baseline-relative adaptive re-sort.
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --workload c5 --steps 300 --warmup 5 --no-e2e > gpurun_out/scalee_c5_2.log 2>&1
echo "c5 n=2 rc=$?"; tail -1 gpurun_out/scalee_c5_2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('  n=%d value %.3e  ms/step %.4f  k1 %.4f ms frac %.3f  exch %.4f solve %.4f ms sorts %s [%s]' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['frac'], d['phases_ms_per_step']['allreduce'], d['phases_ms_per_step']['solve_node_field'], d['tuning'].get('sorts_in_run_rank0'), d['config']['exchange']))
"
