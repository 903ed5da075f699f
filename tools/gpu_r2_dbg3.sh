#!/bin/bash
mkdir -p gpurun_out/r2dbg3
O=gpurun_out/r2dbg3
timeout 300 python -m pytest tests/test_gpu_drivers.py -m gpu -q -k "history_row_order" > $O/plain.log 2>&1; echo "plain rc=$?"; tail -3 $O/plain.log
PTP_TEST_WRAP="compute-sanitizer --tool memcheck --print-limit 5" timeout 500 python -m pytest tests/test_gpu_drivers.py -m gpu -q -s -k "history_row_order" > $O/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "=========|Invalid|illegal|at 0x|by thread|k_" $O/memcheck.log | head -40
PTP_GRAPH=0 timeout 300 python -m pytest tests/test_gpu_drivers.py -m gpu -q -k "history_row_order" > $O/nograph.log 2>&1; echo "nograph rc=$?"; tail -3 $O/nograph.log
PTP_CLUSTER_SOLVE=0 timeout 300 python -m pytest tests/test_gpu_drivers.py -m gpu -q -k "history_row_order" > $O/nocluster.log 2>&1; echo "nocluster rc=$?"; tail -3 $O/nocluster.log
PTP_PDL=0 timeout 300 python -m pytest tests/test_gpu_drivers.py -m gpu -q -k "history_row_order" > $O/nopdl.log 2>&1; echo "nopdl rc=$?"; tail -3 $O/nopdl.log
