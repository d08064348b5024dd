#!/bin/bash
# one measure cycle on the GPU box: parity tests, full bench, ncu launch list + full capture of the STFT kernel
# usage: tools/gpu_cycle.sh <tag> [kernel-regex] [pytest-args]
TAG=${1:-cycle}; KRE=${2:-stft}; PYARGS=${3:-}
mkdir -p gpurun_out
{
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q $PYARGS 2>&1 | tail -30
echo "== bench full"; timeout 900 python bench.py --steps 5 --warmup 3
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --scale 0.25 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_launch_bench.log 2>&1; tail -c 300 gpurun_out/${TAG}_ncu_launch_bench.log
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 1 -c 1 -o gpurun_out/${TAG}_stft -f python bench.py --scale 0.25 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_full_bench.log 2>&1; tail -c 300 gpurun_out/${TAG}_ncu_full_bench.log
echo "== ncu dram traffic at full size"; timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:$KRE -s 1 -c 1 --csv --log-file gpurun_out/${TAG}_traffic.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_traffic_bench.log 2>&1; tail -4 gpurun_out/${TAG}_traffic.csv
} > gpurun_out/${TAG}.log 2>&1
tail -5 gpurun_out/${TAG}.log | cut -c1-600
