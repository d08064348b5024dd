#!/bin/bash
# round 2, call O (1 GPU): band-major mel walk two rounds at a time (warp kernel at 8 - 24 kHz, pair kernel at the 48 / 44.1 kHz
# default bank), programmatic dependent launch + cached quantiser descriptors on small steps, ncu of the packed warp kernel
mkdir -p gpurun_out
{
nvidia-smi -L | head -1
echo "== pytest gpu (all)"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "== small steps"
THB_PDL=0 timeout 300 python tools/smallstep.py 2>&1 | tail -4
THB_PDL=1 timeout 300 python tools/smallstep.py 2>&1 | tail -4
THB_PDL=0 timeout 300 python tools/smallstep.py 2>&1 | tail -4
THB_PDL=1 timeout 300 python tools/smallstep.py 2>&1 | tail -4
echo "== default banks (band-major schedule, two rounds per walk)"
for sr in 16000 8000 22050 24000 48000 44100; do
  timeout 300 python tools/kbench.py --channels 32 --seconds 600 --sr $sr --win-ms 40 --n-mel 0 --reps 5 --variants auto,auto 2>&1 | tail -3
done
echo "== C3 (bin-major, unchanged code path)"
timeout 300 python tools/kbench.py --channels 32 --seconds 150 --reps 5 --variants pair,pair 2>&1 | tail -2
echo "== ncu full: packed warp kernel at 16 kHz and 8 kHz"
for sr in 16000 8000; do
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'stft_warp_kernel<float2' -s 1 -c 1 -o gpurun_out/r02o_warp_$sr -f \
    python tools/kbench.py --channels 32 --seconds 150 --sr $sr --win-ms 40 --n-mel 0 --reps 1 --variants auto > gpurun_out/r02o_ncu_$sr.log 2>&1
  tail -2 gpurun_out/r02o_ncu_$sr.log | cut -c1-160
done
} > gpurun_out/r2o.log 2>&1
tail -70 gpurun_out/r2o.log
