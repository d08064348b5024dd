#!/bin/bash
# round 2, call U (1 GPU): small jobs cut into frame-range descriptors so that every SM gets a work item
mkdir -p gpurun_out
{
nvidia-smi -L | head -1
echo "== pytest gpu (all)"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
for v in 1; do
  echo "== THB_SPLIT_SMALL=$v"
  THB_SPLIT_SMALL=$v timeout 300 python tools/smallstep.py 2>&1 | tail -3
  THB_SPLIT_SMALL=$v timeout 300 python tools/configs_bench.py --only C1 --reps 20 2>&1 | grep -E '^\{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d['config'], 'kernel scopes %.1f us, the call queued %.1f us' % (1e3 * d['stft_ms'], 1e3 * d['spec_batch_ms']))
"
  for secs in 5 20 60; do
    THB_SPLIT_SMALL=$v timeout 300 python tools/kbench.py --channels 2 --seconds $secs --reps 20 --variants auto 2>&1 | tail -2 | tr '\n' ' '; echo
    THB_SPLIT_SMALL=$v timeout 300 python tools/kbench.py --channels 2 --seconds $secs --sr 16000 --win-ms 40 --n-mel 0 --reps 20 --variants auto 2>&1 | tail -2 | tr '\n' ' '; echo
  done
done
echo "== C3 quarter (no split: 4 700 tiles)"; timeout 300 python tools/kbench.py --channels 32 --seconds 150 --reps 5 --variants pair 2>&1 | tail -1
} > gpurun_out/r2u.log 2>&1
tail -40 gpurun_out/r2u.log
