"""ctypes loader of oracle/npp_oracle.cu: the reference's NPP letter-box (``nppiResizeSqrPixel_8u_C3R``, NPPI_INTER_SUPER) run by
NPP itself on a GPU box.  Test infrastructure -- see ``oracle/__init__.py``."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnpp_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "npp_oracle.cu")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["nvcc", "-O2", "-shared", "-Xcompiler", "-fPIC", "-o", _SO, src, "-lnppig", "-lnppc"], stdout=subprocess.DEVNULL)
    return _SO


def npp_letterbox(img: np.ndarray, net_h: int, net_w: int) -> np.ndarray:
    """u8 BGR HWC image -> the reference's NPP letter-box into net_h x net_w (needs a GPU)."""
    lib = C.CDLL(build())
    img = np.ascontiguousarray(img, dtype=np.uint8)
    out = np.empty((net_h, net_w, 3), dtype=np.uint8)
    rc = lib.npp_letterbox(img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], out.ctypes.data_as(C.c_void_p), net_w, net_h)
    if rc != 0:
        raise RuntimeError(f"nppiResizeSqrPixel_8u_C3R failed: {rc}")
    return out
