#!/usr/bin/env python
"""Dumps NPP super-sampling letter-boxes of a few shapes (gpurun_out/npp_*.npz) so the exact semantics of the reference's NPP
branch can be fitted offline.  Development tool; needs a GPU."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2
from oracle.npp import npp_letterbox
img = cv2.imread(os.path.join(ROOT, "tests", "golden", "data", "img.jpg"))
rng = np.random.default_rng(5)
cases = {"photo_448": (img, 448, 448), "photo_320": (img, 320, 320), "crop_700x500": (img[100:600, 200:900], 448, 448), "noise_97x61": (rng.integers(0, 256, (61, 97, 3), dtype=np.uint8), 32, 32),
         "noise_1000x333": (rng.integers(0, 256, (333, 1000, 3), dtype=np.uint8), 448, 448), "ramp_200x100": (np.tile(np.arange(200, dtype=np.uint8)[None, :, None], (100, 1, 3)), 64, 64),
         "small_100x80": (rng.integers(0, 256, (80, 100, 3), dtype=np.uint8), 448, 448), "noise_449x449": (rng.integers(0, 256, (449, 449, 3), dtype=np.uint8), 448, 448)}
out = {}
for k, (im, nh, nw) in cases.items():
    im = np.ascontiguousarray(im)
    out[k + "_src"] = im
    out[k + "_npp"] = npp_letterbox(im, nh, nw)
    nz = np.argwhere(out[k + "_npp"].any(axis=2))
    print(k, im.shape, "->", (nh, nw), "nonzero extent", nz.max(axis=0) + 1 if len(nz) else None, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "npp_dump.npz"), **out)
