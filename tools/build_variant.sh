#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc flags...]  -> gpurun_variants/<name>.so (A/B builds for tools/sweep.sh)
name=$1; shift
mkdir -p gpurun_variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -DDVG_FMA_QUINTIC \
  -Xcompiler -fPIC -shared "$@" -o gpurun_variants/$name.so diffvg_b200/csrc/*.cu
