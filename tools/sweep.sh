#!/bin/bash
# usage: tools/sweep.sh  -- runs bench.py once per library variant under gpurun_variants/
for so in gpurun_variants/*.so; do
  DVG_B200_LIB=$PWD/$so python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /tmp/sweep.json 2>/tmp/sweep.err || { echo "$so FAILED"; tail -3 /tmp/sweep.err; continue; }
  python - "$so" <<'PY'
import json,sys
d=json.load(open('/tmp/sweep.json')); k=d['kernels']
print('%-34s step %.2f ms | fwd %.2f int %.2f edge %.2f' % (sys.argv[1], d['ms_per_step'], k['k_render<false>']['ms_per_step'], k['k_render<true>']['ms_per_step'], k['k_edge']['ms_per_step']))
PY
done
