#!/bin/bash
# tools/build_variant.sh NAME FILE.cu "-DFLAG=..." : build gpurun_out/variants/NAME.so with one translation unit recompiled with extra flags
set -e
cd "$(dirname "$0")/.."
NAME=$1; FILE=$2; FLAGS=$3
OUT=_variants; mkdir -p $OUT
OBJ=torchebm_b200/csrc/_build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I include $FLAGS -c torchebm_b200/csrc/$FILE -o $OUT/$NAME.o
OTHERS=$(ls $OBJ/*.o | grep -v "$(basename $FILE .cu).o")
nvcc -shared -o $OUT/$NAME.so $OUT/$NAME.o $OTHERS -gencode arch=compute_100a,code=sm_100a -lcudart
rm $OUT/$NAME.o
echo built $OUT/$NAME.so
