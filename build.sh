#!/bin/bash
# Builds libr2dm_b200.so (hand-written sm_100a kernels + C ABI) in-tree.
set -e
cd "$(dirname "$0")"
SRC=r2dm_b200/csrc
OUT=r2dm_b200/libr2dm_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
mkdir -p build/obj
pids=()
for f in conv_umma elementwise attention attention_umma model; do
  nvcc $FLAGS ${PTXAS_V:+-Xptxas -v} -c $SRC/$f.cu -o build/obj/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o $OUT build/obj/conv_umma.o build/obj/elementwise.o build/obj/attention.o build/obj/attention_umma.o build/obj/model.o -cudart static
echo "built $OUT"
