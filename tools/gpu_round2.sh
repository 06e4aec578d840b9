#!/bin/bash
# usage (on the GPU box): tools/gpu_round2.sh  -- parity tests, then the A/B sweep of gpurun_variants/*.so
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
tools/sweep.sh > gpurun_out/sweep.txt 2>&1; cat gpurun_out/sweep.txt
