#!/bin/bash
# Host emulation of the sort kernels (see tests/emu/emu_sort.cpp).
set -e
cd "$(dirname "$0")/../.."
mkdir -p build/emu
awk '/\[emu-begin\]/{f=1;next} /\[emu-end\]/{f=0} f' pic-trapped-plasma_b200/csrc/ptp_particles.cu \
  | sed -e 's/extern __shared__ unsigned int hist\[\];.*$/unsigned int* hist = reinterpret_cast<unsigned int*>(g_smem);/' \
        -e 's/extern __shared__ unsigned int sh\[\];.*$/unsigned int* sh = reinterpret_cast<unsigned int*>(g_smem);/' > build/emu/sort_snippet.inc
g++ -std=c++20 -O1 -pthread -ffp-contract=off -Ibuild/emu -Itests/emu -o build/emu/emu_sort tests/emu/emu_sort.cpp
build/emu/emu_sort
