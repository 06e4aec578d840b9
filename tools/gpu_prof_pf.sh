#!/bin/bash
# usage (on the GPU box): tools/gpu_prof_pf.sh [asset] -- one `ncu --set full` capture of the prefiltered path's kernels (flower.svg 2048^2 2x2)
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_render_pf|k_pf_backward|k_wave_classify_px|k_wave_solve_fill" -s 12 -c 6 \
    -f -o gpurun_out/prof_pf python tools/config_kernels.py ${1:-flower} 2 1 > gpurun_out/ncu_pf.log 2>&1
echo "ncu rc=$?" >> gpurun_out/ncu_pf.log
tail -3 gpurun_out/ncu_pf.log
