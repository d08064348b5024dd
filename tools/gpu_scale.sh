#!/bin/bash
# multi-GPU check on an N-GPU box: the NCCL tests, then bench.py at N = 2, 4, 8 as the driver launches it
# usage: tools/gpu_scale.sh <tag> "<list of N>"
TAG=${1:-scale}; NS=${2:-"2 4 8"}
mkdir -p gpurun_out
{
echo "== nvidia-smi"; nvidia-smi -L
echo "== pytest multi"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5
for N in $NS; do
  echo "== bench N=$N"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 5 --warmup 3 --no-cpu 2>&1 | grep -E '^\{|Error|error' | tail -3
done
echo "== reference arm under torchrun N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | grep -E '^\{|Error|error' | tail -3
} > gpurun_out/${TAG}.log 2>&1
cut -c1-400 gpurun_out/${TAG}.log | tail -30
