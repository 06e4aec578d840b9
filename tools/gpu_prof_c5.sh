#!/bin/bash
# usage (on the GPU box): tools/gpu_prof_c5.sh -- one `ncu --set full` capture of a step of the C5 batch scene (512 scenes x 16 strokes, 64^2, 2x2 spp)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_wave|k_boundary" -s 32 -c 16 \
    -f -o gpurun_out/prof_c5 python tools/measure_configs.py --c5-only > gpurun_out/ncu_c5.log 2>&1
echo "ncu rc=$?" >> gpurun_out/ncu_c5.log
grep -v "==PROF==" gpurun_out/ncu_c5.log | tail -4
