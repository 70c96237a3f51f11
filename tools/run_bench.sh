#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -s -k "letterbox or head_tensor or arbitrary or pinned or views" 2>&1 | grep -v "^$" | tail -30
python tools/raw_path_rate.py 2>&1 | tail -2 | tee gpurun_out/raw_path_rate.json
