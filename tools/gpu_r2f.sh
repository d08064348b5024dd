#!/bin/bash
# round 2, call F (2 GPUs): bench at N = 2 (strong scaling + shard check), e2e phases with / without the L2 prefetch, tile kernels
mkdir -p gpurun_out
{
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2f_bench_n2.log 2>&1; grep -E '^\{' gpurun_out/r2f_bench_n2.log > gpurun_out/r2f_bench_n2.json; grep -v '^{' gpurun_out/r2f_bench_n2.log | tail -12; python tools/design_table.py gpurun_out/r2f_bench_n2.json | tail -6
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench_n2.json').read().strip().split('\n')[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus')}), json.dumps(d['e2e'])[:1200])
print(json.dumps(d['strong'])[:3500])
PY
echo "== e2e phases, default"; timeout 600 python tools/e2e_phases.py 2>&1 | tail -6
echo "== e2e phases, THB_PAIR_PREFETCH=0"; THB_PAIR_PREFETCH=0 timeout 600 python tools/e2e_phases.py 2>&1 | tail -6
echo "== tests touching tiles + images"; timeout 900 python -m pytest tests -m gpu -q -k "tile or golden or img or image or smoke" 2>&1 | tail -4
echo "== f2 tiles bench"; timeout 600 python tools/configs_bench.py --only F2 --reps 3 2>&1 | tail -4
} > gpurun_out/r2f.log 2>&1
tail -70 gpurun_out/r2f.log
