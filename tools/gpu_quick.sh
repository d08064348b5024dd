#!/bin/bash
# quick GPU check: the gpu tests (optionally filtered with -k) and chosen configs_bench cases
# usage: tools/gpu_quick.sh <tag> [pytest -k expr | ""] [configs --only list | ""] [extra configs_bench args]
TAG=${1:-quick}; KEXPR=${2:-}; ONLY=${3:-}; EXTRA=${4:-}
mkdir -p gpurun_out
{
echo "== pytest gpu"
if [ -n "$KEXPR" ]; then timeout 1500 python -m pytest tests -m gpu -x -q -k "$KEXPR" 2>&1 | tail -40
else timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40; fi
if [ -n "$ONLY" ]; then echo "== configs"; timeout 900 python tools/configs_bench.py --only "$ONLY" $EXTRA 2>&1 | tail -30; fi
} > gpurun_out/${TAG}.log 2>&1
tail -40 gpurun_out/${TAG}.log | cut -c1-900
