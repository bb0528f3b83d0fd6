#!/bin/bash
# Build an experimental copy of librfk.so with extra nvcc flags, next to the shipped one:
#   tools/build_variant.sh NAME -DRFK_EPI3_WARPS=16 -DRFK_EPI3_RING=2
# -> rosettafold-pytorch_b200/librfk_NAME.so (git-ignored; select it with RFK_LIB_PATH=... in tools/ scripts).
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/rosettafold-pytorch_b200/csrc
BLD=$SRC/build_$NAME
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
mkdir -p $BLD
cd $SRC
SRCS=$(grep '^SRCS=' build.sh | sed 's/SRCS=//; s/"//g')
pids=()
for f in $SRCS; do
  $NVCC $FLAGS "$@" -c $f.cu -o $BLD/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/rosettafold-pytorch_b200/librfk_$NAME.so $(for f in $SRCS; do echo $BLD/$f.o; done) -lcudart
echo "built librfk_$NAME.so"
