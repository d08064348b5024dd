#!/bin/bash
# round 2, call A: full GPU test suite, A/B of the frame-pair kernel changes (prefetch, lean mel), first full bench line
mkdir -p gpurun_out
{
echo "== pytest gpu"; timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== kbench C3-shape (32 ch x 150 s)"
timeout 600 python tools/kbench.py --reps 5 --variants "pair/THB_PAIR_PREFETCH=0/THB_MEL4=0,pair/THB_PAIR_PREFETCH=1/THB_MEL4=0,pair/THB_PAIR_PREFETCH=0,pair,pair/THB_PAIR_PREFETCH=0/THB_MEL4=0,pair"
echo "== kbench default mel (347), hop 480"
timeout 600 python tools/kbench.py --reps 5 --win-ms 40 --n-mel 0 --variants "pair/THB_PAIR_PREFETCH=0,pair"
echo "== kbench C2 shape hop 256"
timeout 600 python tools/kbench.py --reps 5 --t-overlap 8 --channels 8 --variants "pair/THB_PAIR_PREFETCH=0,pair"
echo "== kbench linear"
timeout 600 python tools/kbench.py --reps 5 --scale linear --variants "pair/THB_PAIR_PREFETCH=0,pair"
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.err; cut -c1-3000 gpurun_out/r2a_bench.json
} > gpurun_out/r2a.log 2>&1
tail -80 gpurun_out/r2a.log
