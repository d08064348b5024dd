#!/bin/bash
# compute-sanitizer memcheck over the WHOLE -m gpu suite (the full-size property tests run at reduced sizes under the
# tool: THB_TEST_SMALL=1), racecheck over the kernels with shared-memory choreography.  Log -> gpurun_out/sanitize_<tag>.log
TAG=${1:-r02}
mkdir -p gpurun_out
export THB_TEST_SMALL=1 THB_NO_TIMING=1
{
echo "== memcheck: tests -m gpu"
timeout 3000 compute-sanitizer --tool memcheck --leak-check no --error-exitcode 9 --target-processes all \
  python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | grep -v "^=========\s*$" | tail -40
echo "memcheck exit: ${PIPESTATUS[0]}"
echo "== racecheck: frame-pair / warp / large-FFT kernels, tiles"
timeout 2400 compute-sanitizer --tool racecheck --error-exitcode 9 \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "spec_parity or odd_hop or large_fft or genuinely or spectrogram_tile or i16" 2>&1 | tail -25
echo "racecheck exit: ${PIPESTATUS[0]}"
} > gpurun_out/sanitize_${TAG}.log 2>&1
tail -12 gpurun_out/sanitize_${TAG}.log
