#!/bin/bash
# usage (on the GPU box): tools/gpu_prof.sh [regex] [count] [skip] -- one `ncu --set full` capture of the render kernels
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${1:-k_wave}" -s ${3:-3} -c ${2:-3} \
    -f -o gpurun_out/prof_render python bench.py --steps 1 --warmup 3 --no-cpu-baseline --quick > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/ncu_full.log
tail -3 gpurun_out/ncu_full.log
