mkdir -p gpurun_out
{
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -x -q -k "large_fft or spec_parity or golden or c4" 2>&1 | tail -3
echo "== G"; timeout 600 python tools/configs_bench.py --only G96,C4M --c4-seconds 900 2>&1 | tail -2 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'][:50], round(d['stft_ms'], 3), 'ms', round(d['fp32_tflops'], 2), 'TF', round(d['audio_hours_per_s'], 1), 'h/s')"
} > gpurun_out/c4.log 2>&1
cat gpurun_out/c4.log
