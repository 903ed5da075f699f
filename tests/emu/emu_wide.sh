#!/bin/bash
# Builds the host emulation of the large-grid solver (see tests/emu/emu_wide.cpp): build/emu/emu_wide
set -e
cd "$(dirname "$0")/../.."
mkdir -p build/emu
awk '/\[wide-begin\]/{f=1;next} /\[r16-end\]/{f=0} f' pic-trapped-plasma_b200/csrc/ptp_solve_wide.cu \
  | sed -e 's/extern __shared__ double2 fbw\[\];.*$/ /' \
        -e 's/extern __shared__ __align__(16) double smw\[\];/double* smw = reinterpret_cast<double*>(g_smem);/' > build/emu/wide_snippet.inc
awk '/\[tables-begin\]/{f=1;next} /\[tables-end\]/{f=0} f' pic-trapped-plasma_b200/csrc/ptp_solve.cu > build/emu/tables_snippet.inc
g++ -std=c++20 -O1 -pthread -ffp-contract=off -Ibuild/emu -Itests/emu -o build/emu/emu_wide tests/emu/emu_wide.cpp
