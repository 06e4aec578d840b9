#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc flags...]  -> gpurun_variants/<name>.so (A/B builds for tools/sweep.sh)
name=$1; shift
mkdir -p gpurun_variants build/var_$name
for f in diffvg_b200/csrc/*.cu; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -DDVG_FMA_QUINTIC \
    -Xcompiler -fPIC "$@" -c -o build/var_$name/$(basename $f .cu).o $f &
done
wait
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o gpurun_variants/$name.so build/var_$name/*.o
