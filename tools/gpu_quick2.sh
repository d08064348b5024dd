#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest gpu (all)"; timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== kbench"; timeout 600 python tools/kbench.py --reps 5 --variants "pair,pair/THB_PAIR_PREFETCH=1,pair"
timeout 600 python tools/kbench.py --reps 5 --win-ms 40 --n-mel 0 --variants "pair/THB_MEL_DIRECT=0,pair"
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/rq_bench.json 2> gpurun_out/rq_bench.err; tail -c 400 gpurun_out/rq_bench.err; python tools/design_table.py gpurun_out/rq_bench.json | head -8
} > gpurun_out/rq.log 2>&1
tail -40 gpurun_out/rq.log
