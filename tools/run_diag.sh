#!/bin/bash
mkdir -p gpurun_out
python tools/mask_sweep.py --masks "${1:-9999,448}" --batches "${2:-8}" --streams "${3:-4,6,8}" 2>&1 | tee gpurun_out/mask_sweep.log
