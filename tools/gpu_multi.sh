#!/bin/bash
# usage (on a GPU box with N GPUs): tools/gpu_multi.sh N  -- NCCL check of the row-sharded render, then the driver's own bench command at N GPUs
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > gpurun_out/mg.log 2>&1; echo "mg rc=$?" >> gpurun_out/mg.log
grep -E "OK|FAIL|rc=|Error" gpurun_out/mg.log | tail -12
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?" >> gpurun_out/bench_${N}gpu.err
tail -3 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
for s in d.get('strong') or []:
    print({k: s.get(k) for k in ('workload', 'ms_per_step_1gpu', 'ms_per_step', 'efficiency', 'error')})
PY
