#!/bin/bash
# Build librfk.so in-tree for sm_100a (the only target). Usage: csrc/build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
OUT=../librfk.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}

# --use_fast_math would change expf/division accuracy in the fp32 validation kernels: keep it off.
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
mkdir -p build
pids=()
SRCS="rfk_api rfk_gemm rfk_gemm_epi0 rfk_gemm_epi1 rfk_gemm_epi2 rfk_gemm_epi3 rfk_gemm_epi4 rfk_gemm_conv rfk_elementwise rfk_favor rfk_favor_tc rfk_favor_tm rfk_favor_col rfk_embed rfk_heads"
for f in $SRCS; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ rfk_common.cuh -nt build/$f.o ] || [ rfk_gemm_device.cuh -nt build/$f.o ] || [ rfk_favor_device.cuh -nt build/$f.o ] || [ ../../include/rfk.h -nt build/$f.o ]; then
    $NVCC $FLAGS "$@" -c $f.cu -o build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $(for f in $SRCS; do echo build/$f.o; done) -lcudart
echo "built $(realpath $OUT)"
