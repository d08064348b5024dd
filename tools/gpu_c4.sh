mkdir -p gpurun_out
{
echo "== tests (default)"; timeout 900 python -m pytest tests -m gpu -x -q -k "large_fft or 16384 or golden or big" 2>&1 | tail -2
for cfg in "512 1"; do set -- $cfg
echo "== C4 threads=$1 wsmem=$2"; THB_BIG_THREADS=$1 THB_BIG_WSMEM=$2 timeout 600 python tools/configs_bench.py --only C4L,C4M --c4-seconds 900 2>&1 | tail -2 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'][:22], round(d['stft_ms'], 3), 'ms', round(d['fp32_tflops'], 2), 'TF')"
done
} > gpurun_out/c4.log 2>&1
cat gpurun_out/c4.log
