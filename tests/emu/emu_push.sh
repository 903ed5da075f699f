#!/bin/bash
# Builds the host emulation of the push kernel (see tests/emu/emu_push.cpp): build/emu/emu_push
set -e
cd "$(dirname "$0")/../.."
mkdir -p build/emu
awk '/\[emu-begin\]/{f=1;next} /\[emu-end\]/{f=0} f' pic-trapped-plasma_b200/csrc/ptp_push.cu \
  | sed -e 's/extern __shared__ __align__(16) unsigned char smem\[\];/unsigned char* smem = g_smem;/' > build/emu/push_snippet.inc
g++ -std=c++20 -O1 -pthread -ffp-contract=off -frounding-math -Ibuild/emu -Itests/emu -o build/emu/emu_push tests/emu/emu_push.cpp
