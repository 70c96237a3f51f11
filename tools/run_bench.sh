#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -s -k "calibrator or prototxt or npp or arbitrary or head_tensor" 2>&1 | grep -v "^$" | tail -40
