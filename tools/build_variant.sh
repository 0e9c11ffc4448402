#!/bin/bash
# tools/build_variant.sh TAG "EXTRA_NVCC_FLAGS" -- builds flexpart_b200/csrc/build/var/libfpb_TAG.so
# (the production fast kernels with extra flags) for A/B runs: FPB_ENGINE_LIB=... python bench.py
set -e
cd "$(dirname "$0")/../flexpart_b200/csrc"
TAG=$1; EXTRA=$2
mkdir -p build/var
ARCH="-gencode arch=compute_100a,code=sm_100a"
NVF="$ARCH -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 177"
nvcc $NVF -DFPB_STRICT=0 --prec-div=false --prec-sqrt=false -ftz=true $EXTRA -c fpb_kernels.cu -o build/var/k_$TAG.o
nvcc $NVF $EXTRA -c fpb_sort.cu -o build/var/s_$TAG.o
nvcc $ARCH -shared -o build/var/libfpb_$TAG.so build/var/k_$TAG.o build/fpb_kernels_strict.o build/fpb_scatter.o build/var/s_$TAG.o build/fpb_output.o build/fpb_api.o -lcudart
echo build/var/libfpb_$TAG.so
