#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_round2.py -m gpu -q -s 2>&1 | grep -v "^  \|^$" | tail -90 > gpurun_out/pytest_gpu.log
tail -45 gpurun_out/pytest_gpu.log
