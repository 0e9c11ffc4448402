#!/bin/bash
# tools/build_conv_variant.sh TAG "EXTRA_NVCC_FLAGS" -- builds flexpart_b200/libfpb_TAG.so with fpb_convect.cu compiled
# with extra flags (everything else from the regular build) for A/B runs: FPB_ENGINE_LIB=... python tools/convmix_profile.py
set -e
cd "$(dirname "$0")/../flexpart_b200/csrc"
TAG=$1; EXTRA=$2
mkdir -p build/var
ARCH="-gencode arch=compute_100a,code=sm_100a"
NVF="$ARCH -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 177"
nvcc $NVF --fmad=false $EXTRA -c fpb_convect.cu -o build/var/c_$TAG.o
OTHERS=$(ls build/*.o | grep -v fpb_convect.o)
nvcc $ARCH -shared -o ../libfpb_$TAG.so build/var/c_$TAG.o $OTHERS -lcudart -ldl -lpthread
echo flexpart_b200/libfpb_$TAG.so
