#!/bin/bash
mkdir -p gpurun_out
RF_B200_LIB=$PWD/tools/librf_b200_trace.so RF_TILE_MASK=${1:-255} python - > gpurun_out/trace.log 2>&1 <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, cv2
from oracle.inputs import letterbox_bgr_u8
from retinaface_b200 import RF_PREC_FP16, Engine
from retinaface_b200.capi import RF_FLAG_NO_GRAPH
img = cv2.imread("tests/golden/data/img.jpg")
inp = letterbox_bgr_u8(img, 448, 448)
batch = [np.roll(inp, 8 * i, axis=1) for i in range(8)]
eng = Engine("tests/golden/weights/mnet25.caffemodel", 448, 448, precision=RF_PREC_FP16, max_batch=8, flags=RF_FLAG_NO_GRAPH)
for _ in range(4):
    f = eng.detect_batch(batch, 0.9, 0.4)
print([len(x) for x in f])
eng.close()
PY
grep TRACE gpurun_out/trace.log | tail -13
