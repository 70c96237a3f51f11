#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__cycles_active.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum,launch__registers_per_thread,launch__waves_per_multiprocessor,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
timeout 600 ncu --metrics $M --clock-control none -s 87 -c 29 --csv --log-file gpurun_out/launches_throughput.csv python tools/ncu_step.py --workload mnet25_fp16_b8_448 --warm 3 --steps 1 > gpurun_out/ncu_a.log 2>&1
tail -2 gpurun_out/ncu_a.log
