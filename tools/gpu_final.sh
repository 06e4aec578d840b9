#!/bin/bash
# usage (on the GPU box): tools/gpu_final.sh  -- the full record of a round in one call: GPU parity suite, tools/gpu_artifacts.sh
# (smoke, bench line, reference arm, ncu launch list, ncu --set full of a step), the other BASELINE configurations with the
# reference CPU figure, and one ncu --set full capture of the prefiltered path's kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
tools/gpu_artifacts.sh
timeout 900 python tools/measure_configs.py --cpu > gpurun_out/configs.txt 2> gpurun_out/configs.err; cat gpurun_out/configs.txt
tools/gpu_prof_pf.sh
