"""retinaface_b200 -- B200-native (sm_100a) RetinaFace mnet25 detect path.

The product is ``librf_b200.so`` (CUDA kernels behind the C ABI of ``include/rf_b200.h``); this
package is the thin Python host side used by the tests and ``bench.py``: a ctypes binding
(``capi``) and ``RetinaFace``, a mirror of the reference's C++ class surface
(``retinaface/RetinaFace.h:63-70``).  There is no CPU fallback anywhere in this package.
"""
from .capi import (RF_PREC_FP16, RF_PREC_FP32, RF_PREC_INT8, RfError, Engine, lib_path, load_library)  # noqa: F401
from .detector import FaceDetectInfo, RetinaFace  # noqa: F401

__all__ = ["Engine", "RetinaFace", "FaceDetectInfo", "RfError", "load_library", "lib_path",
           "RF_PREC_FP32", "RF_PREC_FP16", "RF_PREC_INT8"]
