#!/bin/bash
mkdir -p gpurun_out
python tools/mask_sweep.py --masks "${1:-9999,10511,10495,10492,10483,10031}" --batches "${2:-8,32}" --streams "${3:-4,1}" 2>&1 | tee gpurun_out/mask_sweep.log
