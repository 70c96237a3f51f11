"""Input preparation restating the reference's OpenCV preprocess branch, plus the seeded
synthetic inputs of SURVEY.md section 8d.  Test infrastructure -- see ``oracle/__init__.py``.
"""
from __future__ import annotations

import numpy as np


def letterbox_bgr_u8(img: np.ndarray, net_h: int, net_w: int) -> np.ndarray:
    """retinaface/RetinaFace.cpp:587-624 (non-NPP branch): isotropic shrink by
    1/max(cols/W, rows/H, 1) with cv::resize (INTER_LINEAR default), then zero-pad bottom /
    right to net_h x net_w (copyMakeBorder BORDER_CONSTANT 0).  Never up-scales."""
    import cv2
    rows, cols = img.shape[:2]
    sw = np.float32(1.0 * cols / net_w)
    sh = np.float32(1.0 * rows / net_h)
    scale = sw if sw > sh else sh
    scale = scale if scale > 1.0 else np.float32(1.0)
    if scale > 1:
        f = float(np.float32(1) / scale)
        res = cv2.resize(img, None, fx=f, fy=f)
    else:
        res = img
    out = np.zeros((net_h, net_w, 3), dtype=np.uint8)
    h = min(res.shape[0], net_h)
    w = min(res.shape[1], net_w)
    out[:h, :w] = res[:h, :w]
    return out


def s_real_batch(base: np.ndarray, batch: int) -> np.ndarray:
    """S-real: element i = np.roll(base, 8*i, axis=1): same faces, distinct content."""
    return np.stack([np.roll(base, 8 * i, axis=1) for i in range(batch)])


def s_noise_batch(batch: int, net_h: int, net_w: int, seed: int = 0) -> np.ndarray:
    """S-noise: uniform random u8, ~0 detections."""
    return np.random.default_rng(seed).integers(0, 256, (batch, net_h, net_w, 3), dtype=np.uint8)
