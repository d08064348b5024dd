#!/bin/bash
# round 2, call B: GPU tests (incl. the n_fft 1024 / 512 warp kernel), A/B of it against the previous kernels, bench line
mkdir -p gpurun_out
{
echo "== pytest gpu"; timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -25
echo "== kbench 16 kHz default (n_fft 1024)"
timeout 600 python tools/kbench.py --reps 5 --sr 16000 --win-ms 40 --n-mel 0 --seconds 600 --variants "big,warp,generic"
echo "== kbench 8 kHz default (n_fft 512)"
timeout 600 python tools/kbench.py --reps 5 --sr 8000 --win-ms 40 --n-mel 0 --seconds 600 --variants "generic,warp"
echo "== kbench 22.05 kHz default (n_fft 1024, odd hop)"
timeout 600 python tools/kbench.py --reps 5 --sr 22050 --win-ms 40 --n-mel 0 --seconds 600 --variants "big,warp"
echo "== kbench 16 kHz linear"
timeout 600 python tools/kbench.py --reps 5 --sr 16000 --win-ms 40 --scale linear --seconds 600 --variants "big,warp"
echo "== kbench C3-shape"
timeout 600 python tools/kbench.py --reps 5 --variants "pair/THB_MEL4=0,pair"
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 1500 gpurun_out/r2b_bench.err; cut -c1-1200 gpurun_out/r2b_bench.json
} > gpurun_out/r2b.log 2>&1
tail -100 gpurun_out/r2b.log
