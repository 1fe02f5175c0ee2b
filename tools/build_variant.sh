#!/bin/bash
# Build a variant of libdynamask_sm100.so with extra -D flags for dm_roi_align.cu into tools/bin/.
# usage: tools/build_variant.sh <name> "<extra nvcc flags>"   (select at run time with DYNAMASK_LIB=tools/bin/lib_<name>.so)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
OBJ=$ROOT/dynamask_b200/csrc/build
mkdir -p $ROOT/tools/bin
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I$ROOT/include -I$ROOT/dynamask_b200/csrc $@ \
  -c $ROOT/dynamask_b200/csrc/dm_roi_align.cu -o /tmp/ra_$NAME.o
OTHERS=$(ls $OBJ/*.o | grep -v dm_roi_align.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/tools/bin/lib_$NAME.so /tmp/ra_$NAME.o $OTHERS -lcudart
echo built tools/bin/lib_$NAME.so
