#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_scenario.py > gpurun_out/sanitize_scenario.log 2>&1; echo "memcheck scenario rc=$? t=$(( $(date +%s)-S ))s"
tail -6 gpurun_out/sanitize_scenario.log
timeout 120 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitize_smoke.log 2>&1; echo "memcheck smoke rc=$? t=$(( $(date +%s)-S ))s"
tail -3 gpurun_out/sanitize_smoke.log
