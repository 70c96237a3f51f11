#!/usr/bin/env python
"""Generates the committed golden fixtures FROM THE REFERENCE ITSELF, run in this container.

Run once here (needs /root/reference):   python tests/golden/make_golden.py

What it freezes (the reference has no tests / golden vectors of its own):
  * data/img.jpg, weights/*.caffemodel, weights/*.table.int8 -- the reference's DATA artefacts
    (its only image fixture, the trained weights and the INT8 calibration cache the hot path
    consumes).  Binary/data files, byte-identical copies; no reference source code is copied.
  * heads_<model>_448.npz -- the 9 head blobs for data/img.jpg letterboxed to 448x448, from
    cv2.dnn executing the reference's OWN prototxt (input-dim line rewritten) + caffemodel.
  * dets_<model>_<HxW>.npz -- detections from the reference's OWN compiled post-process
    (oracle/_ref: RetinaFace::postProcess, thr 0.9 / NMS 0.4 as in main.cpp:43) on those heads,
    plus head checksums for the large input.
"""
import hashlib
import os
import re
import shutil
import sys
import tempfile

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import topology  # noqa: E402
from oracle.inputs import letterbox_bgr_u8  # noqa: E402
from oracle.mnet_numpy import preprocess_bgr_u8  # noqa: E402
from oracle.postproc import ReferencePostproc, build  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def ref_net(model: str, h: int, w: int):
    txt = open(f"{REF}/model/{model}.prototxt").read()
    txt, n = re.subn(r"shape: \{ dim: 1 dim: 3 dim: \d+ dim: \d+ \}", f"shape: {{ dim: 1 dim: 3 dim: {h} dim: {w} }}", txt)
    assert n == 1
    d = tempfile.mkdtemp()
    p = os.path.join(d, "ref.prototxt")
    open(p, "w").write(txt)
    return cv2.dnn.readNetFromCaffe(p, f"{REF}/model/{model}.caffemodel")


def main():
    build(force=True)
    os.makedirs(f"{OUT}/data", exist_ok=True)
    os.makedirs(f"{OUT}/weights", exist_ok=True)
    shutil.copyfile(f"{REF}/data/img.jpg", f"{OUT}/data/img.jpg")
    for f in ("mnet25.caffemodel", "mnet-deconv-0517.caffemodel", "mnet-deconv-0517.table.int8"):
        shutil.copyfile(f"{REF}/model/{f}", f"{OUT}/weights/{f}")
    img = cv2.imread(f"{OUT}/data/img.jpg")
    assert img.shape == (886, 1280, 3)
    for model in ("mnet-deconv-0517", "mnet25"):
        for (h, w) in ((448, 448), (896, 1280)):
            inp = letterbox_bgr_u8(img, h, w)
            net = ref_net(model, h, w)
            net.setInput(preprocess_bgr_u8(inp))
            heads = [np.ascontiguousarray(b[0]) for b in net.forward(topology.OUTPUT_BLOBS)]
            ref = ReferencePostproc(h, w)
            out = {"input_sha256": np.frombuffer(hashlib.sha256(inp.tobytes()).digest(), dtype=np.uint8)}
            for thr in (0.9, 0.5, 0.02):
                out[f"faces_thr{thr}"] = ref.postprocess(heads, thr)
            ref.close()
            out["head_sums"] = np.array([b.astype(np.float64).sum() for b in heads])
            out["head_abs_sums"] = np.array([np.abs(b.astype(np.float64)).sum() for b in heads])
            np.savez_compressed(f"{OUT}/dets_{model}_{h}x{w}.npz", **out)
            if (h, w) == (448, 448):
                np.savez_compressed(f"{OUT}/heads_{model}_448.npz", **{n: b for n, b in zip(topology.OUTPUT_BLOBS, heads)})
            print(model, h, w, {k: v.shape for k, v in out.items() if k.startswith("faces")})
            print("  top face thr0.9:", out["faces_thr0.9"][0][:5])


if __name__ == "__main__":
    main()
