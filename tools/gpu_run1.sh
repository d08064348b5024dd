#!/bin/bash
# first GPU pass: environment, smoke, parity tests, microbench, small + full bench
mkdir -p gpurun_out
{
echo "== env"; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm,power.limit --format=csv; nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|Core|Thread" 
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "== sanitizer (smoke under memcheck)"; timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40
echo "== microbench"; timeout 120 ./tools/microbench/fp32_pipes
echo "== bench small"; timeout 600 python bench.py --scale 0.1 --steps 3 --warmup 1
echo "== bench full"; timeout 900 python bench.py --steps 5 --warmup 3
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1
} > gpurun_out/run1.log 2>&1
tail -5 gpurun_out/run1.log
