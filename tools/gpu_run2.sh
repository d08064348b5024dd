#!/bin/bash
# second GPU pass: parity tests with the warp-per-frame kernel, bench, ncu launch list + full capture
mkdir -p gpurun_out
{
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
echo "== bench full"; timeout 900 python bench.py --steps 5 --warmup 3
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01_launches.csv python bench.py --scale 0.25 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch_bench.log 2>&1; tail -2 gpurun_out/ncu_launch_bench.log
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:stft2048 -s 1 -c 1 -o gpurun_out/r01_stft2048_mel -f python bench.py --scale 0.25 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_bench.log 2>&1; tail -2 gpurun_out/ncu_full_bench.log
} > gpurun_out/run2.log 2>&1
tail -5 gpurun_out/run2.log
