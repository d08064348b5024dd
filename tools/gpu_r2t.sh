#!/bin/bash
# round 2, call T (1 GPU): file-edge frames with their loads in batches of sixteen (scalar n_fft 2048 / 1024 / 512 kernels)
mkdir -p gpurun_out
{
nvidia-smi -L | head -1
echo "== pytest gpu (all)"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "== small steps"; timeout 300 python tools/smallstep.py 2>&1 | tail -4
echo "== C3 quarter (edges of 32 channels serial behind the packed kernel)"
timeout 300 python tools/kbench.py --channels 32 --seconds 150 --reps 5 --variants pair,pair 2>&1 | tail -2
echo "== 16 kHz / 8 kHz default"
for sr in 16000 8000; do timeout 300 python tools/kbench.py --channels 32 --seconds 600 --sr $sr --win-ms 40 --n-mel 0 --reps 5 --variants auto 2>&1 | tail -1; done
} > gpurun_out/r2t.log 2>&1
tail -30 gpurun_out/r2t.log
