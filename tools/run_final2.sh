#!/bin/bash
mkdir -p gpurun_out
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02_final_1gpu.json 2> gpurun_out/bench_r02_final_1gpu.err ) > gpurun_out/final_bench_time.log 2>&1
python tools/mask_sweep.py --masks 9999 --batches 8 --streams 6,8 > gpurun_out/mask_sweep_s8.log 2>&1
M=gpu__time_duration.sum,sm__cycles_active.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum,launch__registers_per_thread,launch__waves_per_multiprocessor,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
timeout 600 ncu --metrics $M --clock-control none -s 87 -c 29 --csv --log-file gpurun_out/launches_throughput.csv python tools/ncu_step.py --workload mnet25_fp16_b8_448 --warm 3 --steps 1 > gpurun_out/ncu_a.log 2>&1
L1=$(python tools/ncu_step.py --workload mnet25_fp16_b8_448 --warm 0 --steps 1 --streams 1 | grep "launches per step" | awk '{print $4}')
timeout 600 ncu --metrics $M --clock-control none -s $((3 * L1)) -c $L1 --csv --log-file gpurun_out/launches_latency.csv python tools/ncu_step.py --workload mnet25_fp16_b8_448 --warm 3 --steps 1 --streams 1 > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/final_bench_time.log; cat gpurun_out/mask_sweep_s8.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02_final_1gpu.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "images_per_s", "gpu_launches")}, d["clocks"])
print("e2e", {k: d["e2e"][k] for k in ("images_per_s", "ms_per_step")}, "blocking", d["e2e"]["blocking"]["ms_per_step"]); print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"])
for k, v in d.get("configs", {}).items():
    print(k, {x: v.get(x) for x in ("ms_per_step", "images_per_s", "error")}, "e2e", (v.get("e2e") or {}).get("images_per_s"))
PY
