#!/bin/bash
# A/B of library variants (RF_B200_LIB): step time of the throughput plan
mkdir -p gpurun_out
: > gpurun_out/ab_lib.log
for lib in retinaface_b200/librf_b200.so "$@"; do
  echo "== $lib" | tee -a gpurun_out/ab_lib.log
  RF_B200_LIB=$PWD/$lib python tools/mask_sweep.py --masks 9999 --batches 8,32 --streams 6,1 2>&1 | tee -a gpurun_out/ab_lib.log
done
