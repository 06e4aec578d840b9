#!/bin/bash
# usage: tools/sweep.sh  -- runs bench.py once per library variant under gpurun_variants/ (device-timed value + per-kernel times)
mkdir -p gpurun_out
for so in gpurun_variants/*.so; do
  DVG_B200_LIB=$PWD/$so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --quick > /tmp/sweep.json 2>/tmp/sweep.err || { echo "$so FAILED"; tail -3 /tmp/sweep.err; continue; }
  echo "== $so"; python tools/bench_brief.py /tmp/sweep.json | head -2 | python -c "import sys; [print(l[:420]) for l in sys.stdin]"
done
