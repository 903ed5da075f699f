#!/bin/bash
# Builds the host emulation of the cluster solve kernel (see tests/emu/emu_cluster.cpp): build/emu/emu_cluster
set -e
cd "$(dirname "$0")/../.."
mkdir -p build/emu
awk '/\[cluster-begin\]/{f=1;next} /\[cluster-end\]/{f=0} f' pic-trapped-plasma_b200/csrc/ptp_solve_cluster.cu \
  | awk '/^__device__ __forceinline__ void cs_(cp8|commit|wait_all)\(/{skip=1} skip{ if ($0 ~ /}/) skip=0; next } {print}' \
  | sed -e 's/cg::cluster_group cluster = cg::this_cluster();/EmuCluster cluster;/' \
        -e 's/extern __shared__ __align__(16) double smc\[\];/double* smc = reinterpret_cast<double*>(g_smem);/' > build/emu/cluster_snippet.inc
if grep -q "asm volatile" build/emu/cluster_snippet.inc; then echo "emu_cluster.sh: inline PTX left in the snippet"; grep -n "asm volatile" build/emu/cluster_snippet.inc; exit 1; fi
awk '/\[tables-begin\]/{f=1;next} /\[tables-end\]/{f=0} f' pic-trapped-plasma_b200/csrc/ptp_solve.cu > build/emu/tables_snippet.inc
g++ -std=c++20 -O1 -pthread -ffp-contract=off -Ibuild/emu -Itests/emu -o build/emu/emu_cluster tests/emu/emu_cluster.cpp
