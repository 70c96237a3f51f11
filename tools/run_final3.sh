#!/bin/bash
# what the driver runs at round end, on one GPU: GPU tests, smoke, bench with its defaults, the reference arm
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) 2>&1 | tail -8
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -5
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | tail -4
( time python bench.py --impl reference > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err ) 2>&1 | tail -4
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "images_per_s", "gpu_launches", "dtype")})
print("e2e", {k: d["e2e"][k] for k in ("value", "images_per_s", "ms_per_step")}, "config", d["config"].get("execution_contexts"))
r = json.loads(open("gpurun_out/bench_default_ref.json").read().strip().splitlines()[-1])
print("ref", {k: r.get(k) for k in ("impl", "value", "unit", "ms_per_step")})
PY
