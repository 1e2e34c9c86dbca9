#!/bin/bash
# Builds variants of the library that differ in $SRC.cu compile-time switches (probe only).
# usage: build_variants.sh <source without .cu> name:"-DFLAG=.. -DFLAG=.." ...
set -e
SRC=$1; shift
cd "$(dirname "$0")/../../daliid_b200/csrc"
make -s
OUT=../../tests/probes/_variants; mkdir -p $OUT; find $OUT -name "lib_*.so" -delete
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -cudart static --threads 0"
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  $NV $flags -c $SRC.cu -o $OUT/${SRC}_$name.o &
done
wait
for spec in "$@"; do
  name=${spec%%:*}
  objs=$(ls _obj/*.o | grep -v "_obj/$SRC.o")
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $OUT/lib_$name.so $objs $OUT/${SRC}_$name.o -lpthread
  rm $OUT/${SRC}_$name.o
  echo built $name
done
