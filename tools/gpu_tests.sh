#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -s 2>&1 | tail -80 > gpurun_out/tests.log
tail -3 gpurun_out/tests.log
