#!/bin/bash
# round 2, call P (1 GPU): the band-major mel walk, two rounds at a time WITHOUT padding (joint loop + tail of the longer
# round), on the default banks; the C++ host mirror's wall-clock test with its output
mkdir -p gpurun_out
{
nvidia-smi -L | head -1
echo "== pytest gpu (all)"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "== C++ host mirror"; timeout 300 python -m pytest tests/test_cpp_host.py -m gpu -q -s 2>&1 | grep -i "tile readers\|passed\|failed\|EXPECT" | head
echo "== default banks (band-major schedule, two rounds per walk, no padding)"
for sr in 48000 44100 16000 8000 22050 24000; do
  timeout 300 python tools/kbench.py --channels 32 --seconds 600 --sr $sr --win-ms 40 --n-mel 0 --reps 5 --variants auto,auto 2>&1 | tail -3
done
timeout 300 python tools/kbench.py --channels 1 --seconds 3600 --t-overlap 8 --n-mel 0 --reps 5 --variants auto,auto 2>&1 | tail -3
echo "== C3 (bin-major, unchanged code path)"
timeout 300 python tools/kbench.py --channels 32 --seconds 150 --reps 5 --variants pair,pair 2>&1 | tail -2
} > gpurun_out/r2p.log 2>&1
tail -60 gpurun_out/r2p.log
