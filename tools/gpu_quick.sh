#!/bin/bash
# usage (on the GPU box): tools/gpu_quick.sh [fused]  -- parity tests + a short bench of the in-tree library
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err || tail -5 gpurun_out/bench_quick.err
python tools/bench_brief.py gpurun_out/bench_quick.json
if [ "$1" = "fused" ]; then
  DVG_FUSED=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err || tail -5 gpurun_out/bench_fused.err
  python tools/bench_brief.py gpurun_out/bench_fused.json
fi
