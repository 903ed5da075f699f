#!/bin/bash
# Builds the host emulation of the default-grid solver (see tests/emu/emu_solve.cpp): build/emu/emu_solve
set -e
cd "$(dirname "$0")/../.."
mkdir -p build/emu
awk '/\[solve-begin\]/{f=1;next} /\[solve-end\]/{f=0} f' pic-trapped-plasma_b200/csrc/ptp_solve.cu \
  | awk '/^__device__ __forceinline__ void (cp_async(8|16)|mbar_init|mbar_fence_init|mbar_expect_tx|mbar_wait|bulk_g2s|fence_proxy_async)\(/{skip=1} skip{ if ($0 ~ /^}/) skip=0; next } {print}' \
  | sed -e 's/asm volatile("cp.async.commit_group;\\n" ::);/;/' \
        -e 's/asm volatile("cp.async.wait_group [^;]*;\\n" ::[^;]*);/;/' \
        -e 's/extern __shared__ double sB\[\];/double* sB = reinterpret_cast<double*>(g_smem);/' \
        -e 's/extern __shared__ double sm\[\];/double* sm = reinterpret_cast<double*>(g_smem);/' \
        -e 's/extern __shared__ __align__(16) double sm\[\];/double* sm = reinterpret_cast<double*>(g_smem);/' > build/emu/solve_snippet.inc
if grep -q "asm volatile" build/emu/solve_snippet.inc; then echo "emu_solve.sh: inline PTX left in the snippet"; grep -n "asm volatile" build/emu/solve_snippet.inc; exit 1; fi
awk '/\[tables-begin\]/{f=1;next} /\[tables-end\]/{f=0} f' pic-trapped-plasma_b200/csrc/ptp_solve.cu > build/emu/tables_snippet.inc
g++ -std=c++20 -O1 -pthread -ffp-contract=off -Ibuild/emu -Itests/emu -o build/emu/emu_solve tests/emu/emu_solve.cpp
