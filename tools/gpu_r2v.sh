#!/bin/bash
# round 2, call V (8 GPUs, the round's final build): the 2-GPU equality test and bench at N = 8 (weak value + strong scaling)
mkdir -p gpurun_out
{
nvidia-smi -L | wc -l
echo "== multi-gpu equality test"; timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2
N=8
echo "== bench N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e --no-configs > gpurun_out/r2v_bench_n$N.log 2>&1
grep -E '^\{' gpurun_out/r2v_bench_n$N.log > gpurun_out/r2v_bench_n$N.json; grep -v '^{' gpurun_out/r2v_bench_n$N.log | grep -i "error\|Traceback\|assert" | tail -5
python tools/design_table.py gpurun_out/r2v_bench_n$N.json | grep -A4 "strong scaling"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2v_bench_n$N.json').read().strip().split('\n')[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus')}))
for k in ('c3','c2'):
    j=d['strong'][k]; print(k, j['ms_per_step'], j['efficiency'], j['limiter'], j['check'])
PY
} > gpurun_out/r2v.log 2>&1
tail -20 gpurun_out/r2v.log
