#!/bin/bash
# Builds libsovits_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=../libsovits_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
mkdir -p build
for f in kernels_f32 kernels_tc kernels_convn kernels_resblock kernels_flow kernels_attn kernels_prefix api; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ kernels.h -nt build/$f.o ] || [ tc_common.cuh -nt build/$f.o ] || [ ../../include/sovits_b200.h -nt build/$f.o ]; then
    $NVCC $FLAGS ${EXTRA_NVCC_FLAGS} -c $f.cu -o build/$f.o
  fi
done
# link next to the target and rename: a snapshot taken while we build never sees a half-written library
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT.tmp build/kernels_f32.o build/kernels_tc.o build/kernels_convn.o build/kernels_resblock.o build/kernels_flow.o build/kernels_attn.o build/kernels_prefix.o build/api.o -lcudart
mv -f $OUT.tmp $OUT
echo "built $OUT"
