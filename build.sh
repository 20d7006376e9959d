#!/bin/bash
# Builds libr2dm_b200.so (hand-written sm_100a kernels + C ABI) in-tree.
set -e
cd "$(dirname "$0")"
SRC=r2dm_b200/csrc
OUT=${OUT:-r2dm_b200/libr2dm_b200.so}   # developer A/B builds: OUT=... EXTRA_FLAGS="-D..." OBJ=build/alt ./build.sh
OBJ=${OBJ:-build/obj}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $EXTRA_FLAGS"
mkdir -p $OBJ
pids=()
for f in conv_umma conv_chain elementwise attention attention_umma render pointnet model; do
  nvcc $FLAGS ${PTXAS_V:+-Xptxas -v} -c $SRC/$f.cu -o $OBJ/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o $OUT $OBJ/conv_umma.o $OBJ/conv_chain.o $OBJ/elementwise.o $OBJ/attention.o $OBJ/attention_umma.o $OBJ/render.o $OBJ/pointnet.o $OBJ/model.o -cudart static
echo "built $OUT"
