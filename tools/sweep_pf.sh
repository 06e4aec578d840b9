#!/bin/bash
# usage: tools/sweep_pf.sh  -- the prefiltered configurations once per library variant under gpurun_variants/
for so in gpurun_variants/*.so; do
  echo "== $so"; DVG_B200_LIB=$PWD/$so timeout 300 python tools/measure_configs.py 2>&1 | grep -E "^C4"
  DVG_B200_LIB=$PWD/$so timeout 300 python tools/config_kernels.py flower 2 1 2>&1 | sed -n 2,5p
done
