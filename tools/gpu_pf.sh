#!/bin/bash
# usage (on the GPU box): tools/gpu_pf.sh  -- prefiltered path: parity tests, then per-kernel times with the winding pre-pass and inline
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_modes_gpu.py tests/test_svg_assets_gpu.py -m gpu -x -q 2>&1 | tail -8
echo "== flower 2048^2 2x2 prefilter, winding pre-pass"; timeout 300 python tools/config_kernels.py flower 2 1 2>&1 | head -14
echo "== same, inline"; DVG_PF_INLINE=1 timeout 300 python tools/config_kernels.py flower 2 1 2>&1 | head -6
timeout 600 python tools/measure_configs.py 2>&1 | grep -E "^C[24]" 
