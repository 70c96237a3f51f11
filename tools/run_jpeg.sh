#!/bin/bash
mkdir -p gpurun_out
nproc
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -s -k "jpeg" 2>&1 | grep -v "^$" | tail -30 | tee gpurun_out/jpeg_test.log
for t in 1 2 4 8 16; do
  RF_JPEG_THREADS=$t timeout 300 python tools/raw_path_rate.py --jpeg-only 2>&1 | tail -1 | tee gpurun_out/jpeg_rate_t$t.json
done
