#!/bin/bash
# runs every umma_probe test in its own process (gpurun_out/probe.log)
mkdir -p gpurun_out
: > gpurun_out/probe.log
for t in "tma128" "tma64" "tma32" "s2" "conv128 0" "conv128 1" "conv64 0" "conv64 1" "conv32 0" "conv32 1" "dw 0" "dw128 0" "ts 0" "ts128 0"; do
  echo "=== $t" >> gpurun_out/probe.log
  timeout 30 ./tools/umma_probe $t >> gpurun_out/probe.log 2>&1
  echo "rc=$?" >> gpurun_out/probe.log
done
cat gpurun_out/probe.log
