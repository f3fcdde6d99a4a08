#!/bin/bash
# experiment helper: build library variants with extra -D flags into gpurun_variants/<name>.so
#   scripts/build_variants.sh name1 "-DX=1 -DY=2" name2 "-DX=2" ...
set -e
cd "$(dirname "$0")/../myrenderer_b200/csrc"
mkdir -p ../../gpurun_variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -Xcompiler -fPIC"
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  ( nvcc $FLAGS $defs -c triangulate.cu -o /tmp/tri_$name.o && nvcc $ARCH -shared -o ../../gpurun_variants/$name.so api.o terrain.o /tmp/tri_$name.o synth.o -lcudart && echo built $name ) &
done
wait
