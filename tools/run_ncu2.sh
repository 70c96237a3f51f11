#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${1:-k_tc_dwpw_2d} -s ${2:-6} -c ${3:-2} -f -o gpurun_out/full_k python tools/ncu_step.py --workload mnet25_fp16_b8_448 --warm 3 --steps 1 > gpurun_out/ncu_e.log 2>&1
tail -2 gpurun_out/ncu_e.log; ls -la gpurun_out/full_k.ncu-rep
