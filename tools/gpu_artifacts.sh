#!/bin/bash
# usage (on the GPU box): tools/gpu_artifacts.sh  -- the measurement record of a round: full bench line (with the CPU baseline and the
# reference CUDA build), the reference arm, ncu launch list of the same command, one `ncu --set full` capture of a whole step
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_wave|k_boundary" -s 45 -c 15 \
    -f -o gpurun_out/prof_render python bench.py --steps 1 --warmup 3 --no-cpu-baseline --quick > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/smoke.log; python tools/bench_brief.py gpurun_out/bench.json; tail -2 gpurun_out/bench.err; cat gpurun_out/bench_reference.json | cut -c1-400
