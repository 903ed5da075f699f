#!/bin/bash
# Host emulation of the loader's deviate-stream kernels (see tests/emu/emu_rng.cpp).
set -e
cd "$(dirname "$0")/../.."
mkdir -p build/emu
awk '/\[emu-begin\]/{f=1;next} /\[emu-end\]/{f=0} f' pic-trapped-plasma_b200/csrc/ptp_load.cu > build/emu/rng_snippet.inc
g++ -std=c++20 -O1 -pthread -ffp-contract=off -Ibuild/emu -Itests/emu -o build/emu/emu_rng tests/emu/emu_rng.cpp
build/emu/emu_rng
