#!/bin/bash
# Race and bounds check of the kernel text on host threads: every emulation harness rebuilt with ThreadSanitizer (missing
# __syncthreads / __syncwarp between shared-memory accesses show up as data races) and with AddressSanitizer (out-of-bounds
# global or shared accesses), run on the case files the pytest run left behind (pytest --basetemp or /tmp/pytest-of-$USER).
# usage: tests/emu/sanitize_all.sh <dir with test_*/case.bin>      (run `pytest tests/test_kernel_models.py` first)
set -e
cd "$(dirname "$0")/../.."
CASES=${1:-/tmp/pytest-of-$(id -un)/pytest-current}
bash tests/emu/emu_push.sh; bash tests/emu/emu_wide.sh; bash tests/emu/emu_solve.sh; bash tests/emu/emu_cluster.sh; bash tests/emu/emu_place.sh; bash tests/emu/emu_sort.sh > /dev/null; bash tests/emu/emu_r16.sh > /dev/null; bash tests/emu/emu_rng.sh > /dev/null
rc=0
for SAN in thread address; do
  for H in push wide solve cluster place sort r16 rng; do
    g++ -std=c++20 -O1 -g -pthread -fsanitize=$SAN -ffp-contract=off -frounding-math -Ibuild/emu -Itests/emu -o build/emu/emu_${H}_$SAN tests/emu/emu_$H.cpp
  done
  for H in sort r16 rng; do
    if ! TSAN_OPTIONS="exitcode=66" ASAN_OPTIONS="exitcode=66:detect_leaks=0" build/emu/emu_${H}_$SAN > build/emu/san_${H}_$SAN.log 2>&1; then echo "$SAN $H: FAILED"; tail -20 build/emu/san_${H}_$SAN.log; rc=1; else echo "$SAN $H: clean"; fi
  done
  for d in $CASES/test_cluster_solve_kernel_text[0-9]; do
    [ -f "$d/case.bin" ] || continue
    case $(basename $d) in *0) ARGS="12 2";; *1) ARGS="20 1";; *) ARGS="4 3";; esac
    if ! TSAN_OPTIONS="exitcode=66" ASAN_OPTIONS="exitcode=66:detect_leaks=0" build/emu/emu_cluster_$SAN "$d/case.bin" build/emu/san_out.bin $ARGS > build/emu/san_cluster_$SAN.log 2>&1; then echo "$SAN cluster $(basename $d): FAILED"; tail -20 build/emu/san_cluster_$SAN.log; rc=1; else echo "$SAN cluster $(basename $d): clean"; fi
  done
  for H in push wide solve place; do
    case $H in place) pat="test_loader_text_on_host_threa[0-9]";; push) pat="test_push_kernel_text_on_host_[0-9]*";; wide) pat="test_large_grid_solver_text_on[0-9]";; solve) pat="test_default_grid_solver_text_[0-9]";; esac
    EXTRA=""
    [ $H = push ] && EXTRA="$CASES/test_hot_form_*[0-9]"          # the hot form's own cases (longest segment, widest window)
    for d in $CASES/$pat $EXTRA; do
      [ -f "$d/case.bin" ] || continue
      if ! TSAN_OPTIONS="exitcode=66" ASAN_OPTIONS="exitcode=66:detect_leaks=0" build/emu/emu_${H}_$SAN "$d/case.bin" build/emu/san_out.bin > build/emu/san_${H}_$SAN.log 2>&1; then echo "$SAN $H $(basename $d): FAILED"; tail -20 build/emu/san_${H}_$SAN.log; rc=1; else echo "$SAN $H $(basename $d): clean"; fi
    done
  done
done
exit $rc
