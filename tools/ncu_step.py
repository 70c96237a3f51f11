#!/usr/bin/env python
"""Small driver for ncu captures: W warm steps then K steps of the device-resident hot path
(same engine / workload names as bench.py).  Usage under gpurun:
  ncu --metrics gpu__time_duration.sum --clock-control none -s <W*launches> -c <K*launches> --csv \
      --log-file gpurun_out/launches.csv python tools/ncu_step.py --workload mnet25_fp16_b8_448 --warm 3 --steps 2
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default=bench.DEFAULT_WORKLOAD)
    ap.add_argument("--warm", type=int, default=3)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--streams", type=int, default=0, help="execution contexts (1 = the latency plan: tile chains)")
    args = ap.parse_args()
    wl = bench.WORKLOADS[args.workload]
    from retinaface_b200 import RF_PREC_FP16, RF_PREC_FP32, Engine
    from retinaface_b200.capi import RF_FLAG_NO_GRAPH
    eng = Engine(os.path.join(bench.GOLD, "weights", wl["model"] + ".caffemodel"), wl["h"], wl["w"],
                 precision=RF_PREC_FP16 if wl["precision"] == "fp16" else RF_PREC_FP32, max_batch=wl["batch"], max_faces=128,
                 flags=RF_FLAG_NO_GRAPH if args.no_graph else 0, streams=args.streams)
    batch = bench.make_batches(wl, 1, 0)[0]
    eng.pinned_input()[:] = batch
    print("launches per step:", eng.launches_per_batch(wl["batch"]))
    for _ in range(args.warm + args.steps):
        eng.detect_pinned(wl["batch"], bench.SCORE_THR, bench.NMS_THR, bench.np.empty((wl["batch"], eng.max_faces, 15), bench.np.float32),
                          bench.np.zeros(wl["batch"], bench.np.int32))
    eng.close()


if __name__ == "__main__":
    main()
