#!/bin/bash
# round 2, call G: tests; band-major vs bin-major mel schedule on every kernel family; identity tiles; bench line
mkdir -p gpurun_out
{
echo "== pytest gpu (all)"; timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== C3' default 48k (347 bands)"; timeout 600 python tools/kbench.py --reps 5 --win-ms 40 --n-mel 0 --variants "pair/THB_MEL_DIRECT=0,pair/THB_MEL_DIRECT=1,pair"
echo "== C3 mel 128"; timeout 600 python tools/kbench.py --reps 5 --variants "pair/THB_MEL_DIRECT=0,pair/THB_MEL_DIRECT=1,pair"
echo "== 44.1k default"; timeout 600 python tools/kbench.py --reps 5 --sr 44100 --win-ms 40 --n-mel 0 --variants "pair/THB_MEL_DIRECT=0,pair/THB_MEL_DIRECT=1,pair"
echo "== 16k default"; timeout 600 python tools/kbench.py --reps 5 --sr 16000 --win-ms 40 --n-mel 0 --seconds 600 --variants "warp/THB_MEL_DIRECT=0,warp/THB_MEL_DIRECT=1,warp"
echo "== 8k default"; timeout 600 python tools/kbench.py --reps 5 --sr 8000 --win-ms 40 --n-mel 0 --seconds 600 --variants "warp/THB_MEL_DIRECT=0,warp/THB_MEL_DIRECT=1,warp"
echo "== C2 default mel hop 256"; timeout 600 python tools/kbench.py --reps 5 --t-overlap 8 --n-mel 0 --channels 8 --variants "pair/THB_MEL_DIRECT=0,pair/THB_MEL_DIRECT=1"
echo "== f2 tiles bench"; timeout 600 python tools/configs_bench.py --only F2 --reps 3 2>&1 | tail -3 | cut -c1-330
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 600 gpurun_out/r2g_bench.err; python tools/design_table.py gpurun_out/r2g_bench.json
} > gpurun_out/r2g.log 2>&1
tail -80 gpurun_out/r2g.log
