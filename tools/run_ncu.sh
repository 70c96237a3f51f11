#!/bin/bash
# Round-2 ncu evidence: launch lists of one step (throughput plan, latency plan) + full captures of the dominant kernels.
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__cycles_active.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum,launch__registers_per_thread,launch__waves_per_multiprocessor,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
WL=${1:-mnet25_fp16_b8_448}
# preprocess-free blocking call = 1 launch list per step; skip the 3 warm steps
L4=$(python tools/ncu_step.py --workload $WL --warm 0 --steps 1 | grep "launches per step" | awk '{print $4}')
L1=$(python tools/ncu_step.py --workload $WL --warm 0 --steps 1 --streams 1 | grep "launches per step" | awk '{print $4}')
echo "launches: throughput plan $L4, latency plan $L1"
timeout 600 ncu --metrics $M --clock-control none -s $((3 * L4)) -c $L4 --csv --log-file gpurun_out/launches_throughput.csv python tools/ncu_step.py --workload $WL --warm 3 --steps 1 > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -s $((3 * L1)) -c $L1 --csv --log-file gpurun_out/launches_latency.csv python tools/ncu_step.py --workload $WL --warm 3 --steps 1 --streams 1 > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stem_tc -s 3 -c 1 -f -o gpurun_out/full_stem python tools/ncu_step.py --workload $WL --warm 3 --steps 1 > gpurun_out/ncu_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_chain -s $((3 * ${2:-5})) -c ${2:-5} -f -o gpurun_out/full_tile python tools/ncu_step.py --workload $WL --warm 3 --steps 1 --streams 1 > gpurun_out/ncu_d.log 2>&1
tail -3 gpurun_out/ncu_*.log
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_*.csv
