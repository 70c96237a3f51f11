#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,launch__shared_mem_per_block_dynamic,launch__shared_mem_per_block_static,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__occupancy_limit_warps,launch__registers_per_thread,launch__block_size,launch__grid_size
timeout 600 ncu --metrics $M --clock-control none -s 87 -c 29 --csv --log-file gpurun_out/occ.csv python tools/ncu_step.py --workload ${1:-mnet25_fp16_b8_448} --warm 3 --steps 1 > gpurun_out/occ.log 2>&1
tail -2 gpurun_out/occ.log
