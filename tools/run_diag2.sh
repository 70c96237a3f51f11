#!/bin/bash
# A/B: programmatic dependent launch and graph lanes in throughput mode
mkdir -p gpurun_out
: > gpurun_out/ab_pdl.log
for cfg in "" "RF_NO_PDL=1" "RF_ONE_LANE=1" "RF_NO_PDL=1 RF_ONE_LANE=1"; do
  echo "== $cfg" | tee -a gpurun_out/ab_pdl.log
  env $cfg python tools/mask_sweep.py --masks "${1:-9999}" --batches "${2:-8,32}" --streams "${3:-6,1}" 2>&1 | tee -a gpurun_out/ab_pdl.log
done
