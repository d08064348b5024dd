#!/bin/bash
# round 2, call D: tests; warps-per-SM experiment on the frame-pair kernel; exact-reciprocal image kernel A/B + ncu
mkdir -p gpurun_out
{
echo "== pytest gpu (all)"; timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== kbench C3-shape: warps per SM"
timeout 600 python tools/kbench.py --reps 5 --variants "pair:12,pair:14,pair:16,pair:10,pair:12"
timeout 600 python tools/kbench.py --reps 3 --channels 128 --seconds 600 --variants "pair:12,pair:16"
echo "== image kernel: tile mode 2 (fdiv) vs 3 (reciprocal + 2 FMA corrections), C3 size"
for m in 2 3 2 3; do THB_IMG_TILE=$m timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-strong --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); r=d['roofline']; print('THB_IMG_TILE=$m spec_to_img %.4f ms  %.0f GB/s  step %.3f ms'%(r['spec_to_img_avg_ms'], r['spec_to_img_gbs'], d['ms_per_step']))"; done
echo "== ncu full: spec_to_img"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spec_to_img_tile -s 1 -c 1 -o gpurun_out/r2d_img -f python bench.py --scale 0.25 --steps 1 --warmup 1 --no-e2e --no-cpu --no-strong --no-configs > gpurun_out/r2d_ncu_img.log 2>&1; tail -c 200 gpurun_out/r2d_ncu_img.log
echo "== ncu full: warp kernel packed, 16 kHz default"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stft_warp_kernelI6float2 -s 1 -c 1 -o gpurun_out/r2d_warp16 -f python tools/kbench.py --reps 1 --sr 16000 --win-ms 40 --n-mel 0 --seconds 300 --variants warp > gpurun_out/r2d_ncu_warp.log 2>&1; tail -c 300 gpurun_out/r2d_ncu_warp.log
} > gpurun_out/r2d.log 2>&1
tail -60 gpurun_out/r2d.log
