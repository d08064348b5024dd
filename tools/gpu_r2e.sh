#!/bin/bash
# round 2, call E (2 GPUs): tests incl. the 2-rank NCCL equality test, bench at N = 1 and N = 2 (strong scaling + shard check)
mkdir -p gpurun_out
{
nvidia-smi -L
echo "== pytest gpu (all)"; timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== bench N=1"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; tail -c 600 gpurun_out/r2e_bench_n1.err; python tools/design_table.py gpurun_out/r2e_bench_n1.json
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2e_bench_n2.log 2>&1; grep -E '^\{' gpurun_out/r2e_bench_n2.log > gpurun_out/r2e_bench_n2.json; tail -c 1500 gpurun_out/r2e_bench_n2.log | grep -v '^{' | tail -8; python tools/design_table.py gpurun_out/r2e_bench_n2.json | tail -12
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e_bench_n2.json').read().strip().split('\n')[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus')}), json.dumps(d['e2e'])[:900])
print(json.dumps(d['strong'])[:3000])
PY
} > gpurun_out/r2e.log 2>&1
tail -70 gpurun_out/r2e.log
