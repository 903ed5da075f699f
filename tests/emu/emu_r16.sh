#!/bin/bash
# Host emulation of the radix-16 inverse-DCT kernel (see tests/emu/emu_r16.cpp).
set -e
cd "$(dirname "$0")/../.."
mkdir -p build/emu
awk '/\[r16-begin\]/{f=1;next} /\[r16-end\]/{f=0} f' pic-trapped-plasma_b200/csrc/ptp_solve_wide.cu \
  | sed -e 's/extern __shared__ double2 fbw\[\];/ /' > build/emu/r16_snippet.inc
g++ -std=c++20 -O2 -pthread -ffp-contract=off -Ibuild/emu -Itests/emu -o build/emu/emu_r16 tests/emu/emu_r16.cpp
build/emu/emu_r16
