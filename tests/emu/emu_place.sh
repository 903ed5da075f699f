#!/bin/bash
# Builds the host emulation of the device-side loader (see tests/emu/emu_place.cpp): build/emu/emu_place
set -e
cd "$(dirname "$0")/../.."
mkdir -p build/emu
awk '/\[place-begin\]/{f=1;next} /\[place-end\]/{f=0} f' pic-trapped-plasma_b200/csrc/ptp_load.cu > build/emu/place_snippet.inc
awk '/\[counts-begin\]/{f=1;next} /\[counts-end\]/{f=0} f' pic-trapped-plasma_b200/csrc/ptp_load.cu > build/emu/counts_snippet.inc
g++ -std=c++20 -O1 -pthread -ffp-contract=off -Ibuild/emu -Itests/emu -o build/emu/emu_place tests/emu/emu_place.cpp
