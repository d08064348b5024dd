#!/bin/bash
# round 2, call R (1 GPU): the tests added after the final cycle, and ncu --set full of the two small kernels that
# weigh on a 250 us step: the scalar kernel on nine file-edge frames, the quantiser on an eighth of C2
mkdir -p gpurun_out
{
nvidia-smi -L | head -1
echo "== pytest gpu (all)"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "== small steps"; timeout 300 python tools/smallstep.py 2>&1 | tail -4
echo "== ncu: edge kernel (scalar 2048, file edges of one channel)"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'stft2048_kernel<true, false>' -s 4 -c 1 -o gpurun_out/r02r_edges -f python tools/smallstep.py 3 > gpurun_out/r02r_ncu_edges.log 2>&1; tail -2 gpurun_out/r02r_ncu_edges.log | cut -c1-160
echo "== ncu: quantiser on 84 374 x 128"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spec_to_img_tile -s 4 -c 1 -o gpurun_out/r02r_img_small -f python tools/smallstep.py 3 > gpurun_out/r02r_ncu_img.log 2>&1; tail -2 gpurun_out/r02r_ncu_img.log | cut -c1-160
echo "== ncu: packed kernel on 84 374 frames"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft2048_pair -s 4 -c 1 -o gpurun_out/r02r_pair_small -f python tools/smallstep.py 3 > gpurun_out/r02r_ncu_pair.log 2>&1; tail -2 gpurun_out/r02r_ncu_pair.log | cut -c1-160
} > gpurun_out/r2r.log 2>&1
tail -30 gpurun_out/r2r.log
