#!/bin/bash
# round 2, call M (1 GPU): level-0 tile kernel variants (colormap through L1 or shared memory, row groups per CTA; the
# resample kernels no longer launched for a level-0 batch), the warp kernel with / without the loads of the rows that
# hold no window tap, and one ncu --set full capture of the warp kernel at 16 kHz and 8 kHz
mkdir -p gpurun_out
{
nvidia-smi -L | head -1
echo "== pytest gpu (all)"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "== f2 level-0 tiles"
for v in "l1 1" "smem 1" "smem 2" "smem 4" "smem 8" "l1 4"; do
  set -- $v
  echo "-- THB_TILE_CM=$1 THB_TILE_GROUPS=$2"
  THB_TILE_CM=$1 THB_TILE_GROUPS=$2 timeout 300 python tools/configs_bench.py --only F2 --reps 5 2>&1 | grep -E '^\{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   level', d['level_x'], d['level_y'], 'tiles', d['tiles'], '%.3f ms' % d['tile_kernels_ms'], '%.0f GB/s' % d['algorithmic_GBps'], '%.3f of HBM' % d['hbm_frac'])
"
done
echo "== warp kernel: rows without a window tap loaded (1) or not (0)"
for sr in 16000 8000 22050 24000; do
  timeout 300 python tools/kbench.py --channels 32 --seconds 600 --sr $sr --win-ms 40 --n-mel 0 --reps 5 --variants warp/THB_WARP_ALLROWS=1,warp/THB_WARP_ALLROWS=0,warp/THB_WARP_ALLROWS=1,warp/THB_WARP_ALLROWS=0 2>&1 | tail -5
done
timeout 300 python tools/kbench.py --channels 32 --seconds 600 --sr 16000 --win-ms 40 --scale linear --reps 5 --variants warp/THB_WARP_ALLROWS=1,warp/THB_WARP_ALLROWS=0 2>&1 | tail -3
echo "== ncu full: warp kernel at 16 kHz and 8 kHz"
for sr in 16000 8000; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_warp_kernel -s 1 -c 1 -o gpurun_out/r02m_warp_$sr -f \
    python tools/kbench.py --channels 32 --seconds 150 --sr $sr --win-ms 40 --n-mel 0 --reps 1 --variants warp > gpurun_out/r02m_ncu_$sr.log 2>&1
  tail -2 gpurun_out/r02m_ncu_$sr.log | cut -c1-160
done
echo "== ncu full: level-0 tile kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_identity -s 1 -c 1 -o gpurun_out/r02m_tile_id -f \
  python tools/configs_bench.py --only F2 --reps 1 > gpurun_out/r02m_ncu_tile.log 2>&1
tail -2 gpurun_out/r02m_ncu_tile.log | cut -c1-160
} > gpurun_out/r2m.log 2>&1
tail -80 gpurun_out/r2m.log
