#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/check_pytest.log 2>&1
tail -6 gpurun_out/check_pytest.log
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -s -k "head_tensor or all_images or config4" 2>&1 | grep -E "mean|matched|config 4|passed|failed" | tee gpurun_out/check_parity.log
python tools/mask_sweep.py --masks ${1:-9999} --batches 8,32 --streams 6,1 2>&1 | tee gpurun_out/check_sweep.log
