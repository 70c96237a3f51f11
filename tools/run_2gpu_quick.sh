#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "comm" 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/comm_check.py > gpurun_out/comm_check_${N}.log 2>&1; echo "comm_check rc=$?"; grep -E "COMM CHECK|mismatch|Error|error" gpurun_out/comm_check_${N}.log | head -10
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-extra-configs > gpurun_out/bench_r02_${N}gpu_q.json 2> gpurun_out/bench_r02_${N}gpu_q.err; echo "bench rc=$?"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra-configs > gpurun_out/bench_r02_1gpu_q.json 2> gpurun_out/bench_r02_1gpu_q.err
python - <<PY
import json
for f in ("gpurun_out/bench_r02_${N}gpu_q.json", "gpurun_out/bench_r02_1gpu_q.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print({k: d[k] for k in ("n_gpus", "value", "ms_per_step", "images_per_s")}, "e2e", {k: d["e2e"][k] for k in ("images_per_s", "ms_per_step")})
PY
