#!/bin/bash
# compute-sanitizer memcheck over the small paths: smoke (default grid, 2 species) and the large-grid solver on a small grid
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitize_smoke.log 2>&1; echo "memcheck smoke rc=$?"
tail -4 gpurun_out/sanitize_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "other_grids and (256-24-2-0 or 64-300 or 48-420 or 8-70)" > gpurun_out/sanitize_grids.log 2>&1; echo "memcheck grids rc=$?"
tail -4 gpurun_out/sanitize_grids.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "device_loader and 4000 or electrode_programme or edge_cases or losses_match" > gpurun_out/sanitize_misc.log 2>&1; echo "memcheck misc rc=$?"
tail -4 gpurun_out/sanitize_misc.log
