#!/bin/bash
# Builds an experimental variant of the library (RGB only, for speed): tools/build_variant.sh <name> <extra nvcc -D flags...>
set -e
NAME=$1; shift
cd "$(dirname "$0")/../pbrt-r3_b200/csrc"
mkdir -p ../variants
NVFLAGS="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -fmad=false -prec-div=true -prec-sqrt=true -Xcompiler -fPIC,-ffp-contract=off --expt-relaxed-constexpr -DPT_NCS=3 $@"
/usr/local/cuda/bin/nvcc $NVFLAGS -c cuda/device_api.cu -o /tmp/dev_$NAME.o 2> /tmp/dev_$NAME.log || (cat /tmp/dev_$NAME.log; false)
/usr/local/cuda/bin/nvcc -shared -o ../variants/lib_$NAME.so host/front.o host/api_common.o host/bvh_build.o host/scene_accel.o /tmp/dev_$NAME.o -lcudart_static -ldl -lpthread -lrt 2>/dev/null
cp -r ../data ../variants/ 2>/dev/null || true
echo built ../variants/lib_$NAME.so
