#!/bin/bash
# one ncu --set full capture of a kernel out of tools/configs_bench.py
# usage: tools/gpu_ncu_kernel.sh <tag> <kernel-regex> <configs --only list> [extra configs_bench args]
TAG=$1; KRE=$2; ONLY=$3; EXTRA=${4:-}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 1 -c 1 -o gpurun_out/${TAG} -f \
  python tools/configs_bench.py --only $ONLY --reps 1 $EXTRA > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log | cut -c1-200
