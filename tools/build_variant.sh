#!/bin/bash
# Builds an experimental variant of the library (RGB only, for speed): tools/build_variant.sh <name> <extra nvcc -D flags...>
# SRC=<dir> selects another checkout of pbrt-r3_b200/csrc/cuda (e.g. an export of HEAD for a baseline);
# SHADE_FLAGS="..." replaces the division / sqrt flags of the shade translation unit (default: exact, like the product).
set -e
NAME=$1; shift
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC=${SRC:-$ROOT/pbrt-r3_b200/csrc/cuda}
SHADE_FLAGS=${SHADE_FLAGS:--fmad=false -prec-div=true -prec-sqrt=true}
cd "$ROOT/pbrt-r3_b200/csrc"
mkdir -p ../variants
COMMON="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-ffp-contract=off --expt-relaxed-constexpr -Xptxas -v -DPT_NCS=3 $@"
/usr/local/cuda/bin/nvcc $COMMON -fmad=false -prec-div=true -prec-sqrt=true -c $SRC/device_api.cu -o /tmp/dev_$NAME.o 2> /tmp/dev_$NAME.log || (cat /tmp/dev_$NAME.log; false)
OBJS=/tmp/dev_$NAME.o
if [ -f $SRC/shade_tu.cu ]; then
  /usr/local/cuda/bin/nvcc $COMMON $SHADE_FLAGS -c $SRC/shade_tu.cu -o /tmp/shade_$NAME.o 2> /tmp/shade_$NAME.log || (cat /tmp/shade_$NAME.log; false)
  OBJS="$OBJS /tmp/shade_$NAME.o"
fi
/usr/local/cuda/bin/nvcc -shared -o ../variants/lib_$NAME.so host/front.o host/api_common.o host/bvh_build.o host/scene_accel.o $OBJS -lcudart_static -ldl -lpthread -lrt -lz 2>/dev/null
cp -r ../data ../variants/ 2>/dev/null || true
grep -A2 "Function properties for _ZN3ptd7k_traceILi3ELb0ELb0" /tmp/dev_$NAME.log | grep -E "Used|spill" | tr '\n' ' '; echo
echo built ../variants/lib_$NAME.so
