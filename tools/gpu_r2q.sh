#!/bin/bash
# round 2, call Q (8 GPUs, the round's last build: programmatic dependent launch + cached quantiser descriptors): weak value and
# strong scaling at N = 8 and N = 4 (no e2e / cpu / configs legs: those are in call N and the final 1-GPU cycle)
mkdir -p gpurun_out
{
nvidia-smi -L | wc -l
for N in 8 4; do
  echo "== bench N=$N"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e --no-configs > gpurun_out/r2q_bench_n$N.log 2>&1
  grep -E '^\{' gpurun_out/r2q_bench_n$N.log > gpurun_out/r2q_bench_n$N.json; grep -v '^{' gpurun_out/r2q_bench_n$N.log | grep -i "error\|Traceback\|assert" | tail -5
  python tools/design_table.py gpurun_out/r2q_bench_n$N.json | grep -A4 "strong scaling"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2q_bench_n$N.json').read().strip().split('\n')[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus')}))
for k in ('c3','c2'):
    j=d['strong'][k]; print(k, j['ms_per_step'], j['efficiency'], j['limiter'], j['check'], j['stft_kernel_ms'], j['scalar_launches_ms'], j['minmax_allreduce_ms'], j['spec_to_img_ms'])
PY
done
} > gpurun_out/r2q.log 2>&1
tail -40 gpurun_out/r2q.log
