#!/bin/bash
# round 2, call C: tests again (C++ concurrency proof, 2x bars), L1-prefetch A/B, ncu of the frame-pair + warp kernels
mkdir -p gpurun_out
{
echo "== pytest gpu"; timeout 1700 python -m pytest tests -m gpu -q -s -k "cpp_host or seams or genuinely or real_audio" 2>&1 | grep -v "^$" | tail -25
echo "== pytest gpu (all)"; timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== kbench C3-shape: L1 prefetch of the next pair"
timeout 600 python tools/kbench.py --reps 5 --variants "pair,pair/THB_PAIR_PL=1,pair,pair/THB_PAIR_PL=1"
timeout 600 python tools/kbench.py --reps 3 --channels 128 --seconds 600 --variants "pair,pair/THB_PAIR_PL=1"
echo "== ncu full: pair kernel (default) at bench scale 0.25"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stft2048_pair -s 1 -c 1 -o gpurun_out/r2c_pair -f python bench.py --scale 0.25 --steps 1 --warmup 1 --no-e2e --no-cpu --no-strong --no-configs > gpurun_out/r2c_ncu_pair.log 2>&1; tail -c 300 gpurun_out/r2c_ncu_pair.log
echo "== ncu full: warp kernel, 16 kHz default"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stft_warp_kernel -s 1 -c 1 -o gpurun_out/r2c_warp16 -f python tools/kbench.py --reps 1 --sr 16000 --win-ms 40 --n-mel 0 --seconds 300 --variants warp > gpurun_out/r2c_ncu_warp.log 2>&1; tail -c 300 gpurun_out/r2c_ncu_warp.log
} > gpurun_out/r2c.log 2>&1
tail -60 gpurun_out/r2c.log
