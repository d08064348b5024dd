#!/bin/bash
# round 2, call N (8 GPUs, last build of the round: edge frames on a side stream): bench at N = 8 and N = 4 as the driver launches it
# (weak value, strong scaling of the fixed C3 batch and of C2's one file by frame range with the shard check, e2e)
mkdir -p gpurun_out
{
nvidia-smi -L | head -8; nproc; free -g | head -2
echo "== multi-gpu equality test"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2
for N in 8 4; do
  echo "== bench N=$N"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/r2n_bench_n$N.log 2>&1
  grep -E '^\{' gpurun_out/r2n_bench_n$N.log > gpurun_out/r2n_bench_n$N.json; grep -v '^{' gpurun_out/r2n_bench_n$N.log | grep -i "error\|Traceback\|assert" | tail -5
  python tools/design_table.py gpurun_out/r2n_bench_n$N.json | grep -A4 "strong scaling"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2n_bench_n$N.json').read().strip().split('\n')[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus')}))
e=d['e2e']; print('e2e', e['value'], e['ms_per_step'], json.dumps(e['host_feed_probe']), e['host_limit_ms_per_step'], e['host_affinity'])
print('e2e_i16', d['e2e_i16']['value'], d['e2e_i16']['ms_per_step'])
for k in ('c3','c2'):
    j=d['strong'][k]; print(k, j['ms_per_step'], j['efficiency'], j['limiter'], j['check'])
PY
done
} > gpurun_out/r2n.log 2>&1
tail -60 gpurun_out/r2n.log
