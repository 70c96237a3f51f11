"""Python mirror of the reference's detector class surface, over the C ABI.

Mirrors ``class RetinaFace`` (retinaface/RetinaFace.h:63-70): same constructor arguments
(model directory, network name "net3", nms threshold 0.4), ``detect(img, threshold, scales)``
and ``detectBatchImages(imgs, threshold)``.  The reference's methods return ``void`` and drop
their results (RetinaFace.cpp:665,726,747); here they return the ``FaceDetectInfo`` list that the
reference computes and discards.  The C++ twin of this file is ``retinaface_b200/host/RetinaFace.h``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

from .capi import RF_PREC_FP16, Engine


@dataclass
class FaceDetectInfo:  # RetinaFace.h:37-42
    score: float
    rect: tuple   # anchor_box x1,y1,x2,y2
    pts_x: tuple  # FacePts.x[5]
    pts_y: tuple  # FacePts.y[5]

    @staticmethod
    def from_row(r: np.ndarray) -> "FaceDetectInfo":
        return FaceDetectInfo(float(r[0]), tuple(map(float, r[1:5])), tuple(map(float, r[5:10])), tuple(map(float, r[10:15])))


class RetinaFace:
    MODEL_FILE = "mnet-deconv-0517.caffemodel"  # the file the reference always loads, RetinaFace.cpp:276

    def __init__(self, model: str, network: str = "net3", nms: float = 0.4, *, net_w: int = 448, net_h: int = 448,
                 max_batch: int = 8, precision: int = RF_PREC_FP16, device: int = 0, max_faces: int = 256,
                 model_file: str = None, max_image=(3072, 4096)):
        if network != "net3":
            # RetinaFace.cpp:211-242 lists other names, but only the fmc==3 "net3" anchors are configured (:245-271)
            raise ValueError(f"network setting error {network}: only 'net3' is configured")
        self.nms_threshold = nms
        path = os.path.join(model, model_file or self.MODEL_FILE)
        self.engine = Engine(path, net_h, net_w, precision=precision, max_batch=max_batch, max_faces=max_faces,
                             device=device, max_image=max_image)

    def detect(self, img: np.ndarray, threshold: float = 0.5, scales: float = 1.0) -> List[FaceDetectInfo]:
        if img is None or img.size == 0:  # RetinaFace.cpp:578-580
            return []
        return self.detectBatchImages([img], threshold)[0]

    def detectBatchImages(self, imgs: Sequence[np.ndarray], threshold: float = 0.5) -> List[List[FaceDetectInfo]]:
        rows = self.engine.detect_batch(list(imgs), threshold, self.nms_threshold)
        return [[FaceDetectInfo.from_row(r) for r in per] for per in rows]

    @staticmethod
    def map_back_scale(img_w: int, img_h: int, net_w: int, net_h: int) -> float:
        """scale of RetinaFace.cpp:587-591: multiply coordinates by it to return to image pixels (:732-738)."""
        return max(img_w / net_w, img_h / net_h, 1.0)
