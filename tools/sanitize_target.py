#!/usr/bin/env python
"""Short end-to-end scenario for compute-sanitizer (memcheck): throughput plan, chain plan, INT8, letter-box (both resize
definitions).  (The fused exchange needs one process per rank: tests/comm_worker.py, tools/comm_check.py.)  compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    import cv2
    from oracle.inputs import letterbox_bgr_u8
    from retinaface_b200 import RF_PREC_FP16, RF_PREC_INT8, Engine
    from retinaface_b200.capi import RF_FLAG_NPP_RESIZE
    img = cv2.imread(os.path.join(GOLD, "data", "img.jpg"))
    inp = letterbox_bgr_u8(img, 448, 448)
    batch = [np.roll(inp, 8 * i, axis=1) for i in range(3)]
    cm = os.path.join(GOLD, "weights", "mnet25.caffemodel")
    for streams in (0, 1):
        eng = Engine(cm, 448, 448, precision=RF_PREC_FP16, max_batch=3, streams=streams, max_image=img.shape[:2])
        f = eng.detect_batch(batch, 0.9, 0.4)
        g = eng.detect_batch([img, img[:600, :900].copy()], 0.9, 0.4)
        t = [eng.submit(batch, 0.9, 0.4) for _ in range(3)]
        for k in t:
            eng.collect(k)
        print("fp16 streams", streams, eng.launches_per_batch(3), [len(x) for x in f], [len(x) for x in g])
        eng.close()
    eng = Engine(cm, 288, 416, precision=RF_PREC_FP16, max_batch=2, streams=1, flags=RF_FLAG_NPP_RESIZE, max_image=img.shape[:2])
    print("288x416 npp", [len(x) for x in eng.detect_batch([img, img[::2, ::2].copy()], 0.8, 0.4)])
    eng.close()
    cm2 = os.path.join(GOLD, "weights", "mnet-deconv-0517.caffemodel")
    eng = Engine(cm2, 448, 448, precision=RF_PREC_INT8, max_batch=3, int8_table=os.path.join(GOLD, "weights", "mnet-deconv-0517.table.int8"))
    print("int8", [len(x) for x in eng.detect_batch(batch, 0.9, 0.4)])
    eng.close()


if __name__ == "__main__":
    main()
