#!/bin/bash
# usage (on the GPU box): tools/gpu_round.sh  -- parity tests (every failure reported), quick bench, other configurations
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --quick > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err || tail -5 gpurun_out/bench_quick.err
python tools/bench_brief.py gpurun_out/bench_quick.json
timeout 600 python tools/measure_configs.py > gpurun_out/configs.txt 2>&1; cat gpurun_out/configs.txt
