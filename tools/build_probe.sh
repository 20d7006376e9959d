#!/bin/bash
# Builds the tcgen05/TMA hardware probe (developer tool).  Run it with:
#   gpurun -- 'timeout 90 ./build/probe_umma'
set -e
cd "$(dirname "$0")/.."
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -lineinfo -std=c++17 -o build/probe_umma tools/probe_umma.cu
