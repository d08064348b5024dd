#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on the shared-memory kernels) over a small subset of the GPU tests
mkdir -p gpurun_out
{
echo "== memcheck"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "spectrogram_tile or apply_gain or odd_hop or plan_cache or (large_fft and 4096) or (large_fft and 8192 and Linear) or (spec_parity and (22k05 or 192k or 4096 or C3 or C4-mel))" 2>&1 | tail -12
echo "== racecheck"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "(spec_parity and (22k05 or 96k or C4-mel or C3-mel))" 2>&1 | tail -12
} > gpurun_out/sanitize.log 2>&1
tail -30 gpurun_out/sanitize.log | cut -c1-220
