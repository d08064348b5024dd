#!/bin/bash
# round 2, call K (2 GPUs): dynamic tile hand-out + file-edge frames on a side stream: tests, A/B, bench N = 1 / 2
mkdir -p gpurun_out
{
echo "== pytest gpu (all)"; timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== kbench C3 shape"; timeout 600 python tools/kbench.py --reps 5 --variants "pair,pair/THB_EDGE_SIDE=0,pair"
timeout 600 python tools/kbench.py --reps 3 --channels 128 --seconds 600 --variants "pair,pair/THB_EDGE_SIDE=0"
echo "== kbench C2 shape, 1/8 of the hour (one rank's share at N = 8)"; timeout 600 python tools/kbench.py --reps 20 --t-overlap 8 --channels 1 --seconds 450 --variants "pair,pair/THB_EDGE_SIDE=0,pair"
echo "== bench N=1"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; tail -c 300 gpurun_out/r2k_bench_n1.err; python tools/design_table.py gpurun_out/r2k_bench_n1.json | sed -n '3p;20,24p'
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2k_bench_n2.log 2>&1; grep -E '^\{' gpurun_out/r2k_bench_n2.log > gpurun_out/r2k_bench_n2.json; grep -v '^{' gpurun_out/r2k_bench_n2.log | grep -i "error\|Traceback\|assert" | tail -5
python tools/design_table.py gpurun_out/r2k_bench_n2.json | grep -A4 "strong scaling"
} > gpurun_out/r2k.log 2>&1
tail -50 gpurun_out/r2k.log
