#!/bin/bash
# round 2, final 1-GPU cycle: tests, smoke, the bench line, ncu launch list of the bench command, ncu --set full of the
# frame-pair and image kernels, DRAM traffic at full size, compute-sanitizer over the whole suite
mkdir -p gpurun_out
{
echo "== pytest gpu (all)"; timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (final line)"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_final_bench_line.json 2> gpurun_out/r02_final_bench.err; tail -c 300 gpurun_out/r02_final_bench.err; python tools/design_table.py gpurun_out/r02_final_bench_line.json
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 1 --warmup 1 | cut -c1-400
echo "== ncu launch list of the bench command (scale 0.25)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --scale 0.25 --steps 2 --warmup 1 --no-e2e --no-cpu --no-strong --no-configs > gpurun_out/r02_final_ncu_launch_bench.log 2>&1; tail -c 200 gpurun_out/r02_final_ncu_launch_bench.log
echo "== ncu full: pair kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stft2048_pair -s 1 -c 1 -o gpurun_out/r02_final_pair -f python bench.py --scale 0.25 --steps 1 --warmup 1 --no-e2e --no-cpu --no-strong --no-configs > gpurun_out/r02_final_ncu_pair.log 2>&1; tail -c 150 gpurun_out/r02_final_ncu_pair.log
echo "== ncu full: image kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spec_to_img_tile -s 1 -c 1 -o gpurun_out/r02_final_img -f python bench.py --scale 0.25 --steps 1 --warmup 1 --no-e2e --no-cpu --no-strong --no-configs > gpurun_out/r02_final_ncu_img.log 2>&1; tail -c 150 gpurun_out/r02_final_ncu_img.log
echo "== ncu dram traffic at full size"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"stft2048_pair|spec_to_img_tile" -s 2 -c 2 --csv --log-file gpurun_out/r02_final_traffic.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-strong --no-configs > gpurun_out/r02_final_ncu_traffic.log 2>&1; tail -6 gpurun_out/r02_final_traffic.csv | cut -c1-300
echo "== sanitizer"; bash tools/gpu_sanitize.sh r02
} > gpurun_out/r02_final.log 2>&1
tail -90 gpurun_out/r02_final.log
