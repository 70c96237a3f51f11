#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/comm_check.py > gpurun_out/comm_check_${N}.log 2>&1; echo "comm_check rc=$?"; grep -E "COMM CHECK|mismatch|Error|error" gpurun_out/comm_check_${N}.log | head -10
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/bench_r02_${N}gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r02_${N}gpu.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_r02_${N}gpu.json"))
print({k: d[k] for k in ("n_gpus", "value", "ms_per_step", "images_per_s")}, "e2e", {k: d["e2e"][k] for k in ("value", "images_per_s", "ms_per_step")})
for k, v in d["configs"].items():
    print(k, {x: v.get(x) for x in ("ms_per_step", "images_per_s", "error")}, "e2e", (v.get("e2e") or {}).get("images_per_s"))
PY
