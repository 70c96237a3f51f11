#!/bin/bash
# single-GPU: round-2 tests that failed + default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -s -k "head_tensor or all_images" 2>&1 | grep -v "^$" | tail -25
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/bench_r02_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r02_1gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02_1gpu.json"))
print({k: d[k] for k in ("value", "ms_per_step", "images_per_s", "launches_per_step")})
print("e2e", {k: d["e2e"][k] for k in ("value", "images_per_s", "ms_per_step")}, "blocking", d["e2e"].get("blocking"))
print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["kernel_ms"])
for k, v in d["configs"].items():
    print(k, {x: v.get(x) for x in ("ms_per_step", "images_per_s", "error")}, "e2e", (v.get("e2e") or {}).get("images_per_s"))
print("cpu", {k: d["cpu_baseline"][k] for k in ("value", "cores", "images_per_s")}, d["cpu_baseline"].get("one_core"), d["cpu_baseline"].get("single_image"))
print("clocks", d.get("clocks"))
PY
