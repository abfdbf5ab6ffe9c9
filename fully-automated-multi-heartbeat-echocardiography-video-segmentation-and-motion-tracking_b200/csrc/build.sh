#!/bin/sh
# Build libclasfv_b200.so for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="$EXTRA_NVCC_FLAGS -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
mkdir -p build
for f in api conv_simt conv_umma decoder decoder_umma fusion; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ internal.h -nt build/$f.o ] || [ umma_ptx.cuh -nt build/$f.o ] || [ ../../include/clasfv_b200.h -nt build/$f.o ]; then
    rm -f build/$f.o
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $f.cu -o build/$f.o &
  fi
done
wait
for f in api conv_simt conv_umma decoder decoder_umma fusion; do
  [ -f build/$f.o ] || { echo "build failed: $f.cu" >&2; exit 1; }      # a failed background compile must not link a stale object
done
$NVCC -shared -o libclasfv_b200.so build/api.o build/conv_simt.o build/conv_umma.o build/decoder.o build/decoder_umma.o build/fusion.o
echo "built $(pwd)/libclasfv_b200.so"
