#!/bin/bash
# round 2, call I (2 GPUs): the NVLink peer-memory range exchange: 2-rank equality test, bench N = 2 with and without it
mkdir -p gpurun_out
{
echo "== pytest multi + seams"; timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_seams.py -m gpu -q -s 2>&1 | tail -8
for PX in 1 0; do
echo "== bench N=2 THB_PEER_EXCHANGE=$PX"
THB_PEER_EXCHANGE=$PX timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29502+PX)) bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2i_bench_n2_px$PX.log 2>&1; grep -E '^\{' gpurun_out/r2i_bench_n2_px$PX.log > gpurun_out/r2i_bench_n2_px$PX.json; grep -v '^{' gpurun_out/r2i_bench_n2_px$PX.log | grep -i "error\|Traceback\|assert" | tail -5
python - <<PY
import json
d=json.loads(open('gpurun_out/r2i_bench_n2_px$PX.json').read().strip().split('\n')[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus')}), d['strong']['global_range_exchange'], d['roofline']['minmax_allreduce_avg_ms'])
for k in ('c3','c2'):
    j=d['strong'][k]; print(k, round(j['ms_per_step'],4), round(j['efficiency'],4), j['minmax_allreduce_ms'], j['check'])
PY
done
} > gpurun_out/r2i.log 2>&1
tail -40 gpurun_out/r2i.log
