#!/bin/bash
# Round-end check on one GPU: every GPU test, smoke, the driver's bench commands (timed), the mask sweep of DESIGN.md section 3
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/final_pytest.log 2>&1
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final_smoke.log 2>&1
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02_final_1gpu.json 2> gpurun_out/bench_r02_final_1gpu.err ) > gpurun_out/final_bench_time.log 2>&1
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02_final_ref.json 2> gpurun_out/bench_r02_final_ref.err ) > gpurun_out/final_ref_time.log 2>&1
python tools/mask_sweep.py --masks 9999,448,480,482,511 --batches 1,8,32 --streams 1,6 > gpurun_out/mask_sweep_final.log 2>&1
tail -4 gpurun_out/final_pytest.log; tail -4 gpurun_out/final_smoke.log; tail -4 gpurun_out/final_bench_time.log; tail -4 gpurun_out/final_ref_time.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02_final_1gpu.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "images_per_s", "gpu_launches", "clocks")})
print("e2e", d["e2e"]); print("roofline", d["roofline"]); print("cpu", d["cpu_baseline"])
for k, v in d.get("configs", {}).items():
    print(k, {x: v.get(x) for x in ("ms_per_step", "images_per_s", "error")}, "e2e", (v.get("e2e") or {}).get("images_per_s"))
PY
cat gpurun_out/mask_sweep_final.log
