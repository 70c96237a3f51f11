"""FP32 numpy restatement of the reference network forward (Caffe semantics).

Test infrastructure -- see ``oracle/__init__.py``.

Restates what the reference obtains from ``Net_->Forward()`` (``retinaface/RetinaFace.cpp:988``)
or ``context->enqueue`` (``retinaface/tensorrt/trtretinafacenet.cpp:60``) on the graph of
``model/mnet-deconv-0517.prototxt``: the arithmetic lives in BVLC Caffe / TensorRT, which
are NOT in /root/reference (un-vendored, unpinned), so the layer
definitions below follow BVLC Caffe's published layer semantics:

* Convolution: cross-correlation, zero padding, groups                  (conv_layer.cpp)
* BatchNorm(use_global_stats): (x - mean/sf) / sqrt(var/sf + eps)       (batch_norm_layer.cpp)
* Scale(bias_term): x * gamma + beta                                    (scale_layer.cpp)
* Deconvolution (grouped, k4 s2 p1): transposed conv, zero border       (deconv_layer.cpp)
* Crop axis 2 offset 0: keep top-left h x w                             (crop_layer.cpp)
* Softmax over axis 1 of the (N,2,2h,w) view, max-subtracted            (softmax_layer.cpp)

Pinned against cv2.dnn executing the reference's own prototxt + caffemodel
(``tests/golden/make_golden.py`` -> ``tests/test_oracle_forward.py``).
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import numpy as np

from . import topology
from .caffemodel import load_caffemodel


def preprocess_bgr_u8(img_hwc_bgr: np.ndarray) -> np.ndarray:
    """u8 HWC BGR (already network-sized) -> f32 1x3xHxW RGB, raw 0..255.

    retinaface/RetinaFace.cpp:626-645 (convertTo CV_32FC3, cvtColor BGR2RGB, split);
    pixel_means = 0, pixel_stds = 1 for "net3" (RetinaFace.h:85-87).
    """
    x = img_hwc_bgr[..., ::-1].astype(np.float32)
    return np.ascontiguousarray(x.transpose(2, 0, 1))[None]


def _conv2d(x, w, b, s, p, g):
    n, cin, h, wd = x.shape
    cout, cig, k, _ = w.shape
    oh = (h + 2 * p - k) // s + 1
    ow = (wd + 2 * p - k) // s + 1
    xp = np.pad(x, ((0, 0), (0, 0), (p, p), (p, p))) if p else x
    if g == 1:
        if k == 1 and s == 1:
            y = np.einsum("oc,nchw->nohw", w[:, :, 0, 0], xp, optimize=True)
        else:
            # im2col: (n, cin*k*k, oh*ow)
            cols = np.empty((n, cin, k, k, oh, ow), dtype=np.float32)
            for dy in range(k):
                for dx in range(k):
                    cols[:, :, dy, dx] = xp[:, :, dy:dy + s * oh:s, dx:dx + s * ow:s]
            y = np.matmul(w.reshape(cout, -1)[None], cols.reshape(n, cin * k * k, oh * ow))
            y = y.reshape(n, cout, oh, ow)
    else:
        assert g == cin == cout and cig == 1
        y = np.zeros((n, cout, oh, ow), dtype=np.float32)
        for dy in range(k):
            for dx in range(k):
                y += xp[:, :, dy:dy + s * oh:s, dx:dx + s * ow:s] * w[None, :, 0, dy, dx, None, None]
    if b is not None:
        y = y + b[None, :, None, None]
    return y.astype(np.float32, copy=False)


def _deconv_dw_k4s2p1(x, w):
    """Grouped (depthwise) transposed conv, k=4 s=2 p=1: out[y] = sum_i in[i] * w[y - 2i + 1]."""
    n, c, h, wd = x.shape
    full = np.zeros((n, c, 2 * h + 2, 2 * wd + 2), dtype=np.float32)  # un-cropped (pad 0) output
    for ky in range(4):
        for kx in range(4):
            full[:, :, ky:ky + 2 * h:2, kx:kx + 2 * wd:2] += x * w[None, :, 0, ky, kx, None, None]
    return full[:, :, 1:1 + 2 * h, 1:1 + 2 * wd]  # pad=1 crops one ring


class MnetOracle:
    """Runs the graph of ``topology.ops()`` with the weights of one caffemodel."""

    def __init__(self, caffemodel_path: str):
        self.layers = load_caffemodel(caffemodel_path)

    def _blobs(self, name):
        return self.layers[name]["blobs"]

    def forward(self, data_nchw_f32: np.ndarray, want: Optional[Iterable[str]] = None) -> Dict[str, np.ndarray]:
        want = list(want) if want is not None else list(topology.OUTPUT_BLOBS)
        t: Dict[str, np.ndarray] = {"data": np.asarray(data_nchw_f32, dtype=np.float32)}
        for op in topology.ops():
            kind = op["op"]
            if kind == "conv":
                bl = self._blobs(op["name"])
                w = bl[0].reshape(op["cout"], op["cin"] // op["g"], op["k"], op["k"])
                b = bl[1].reshape(-1) if op["bias"] else None
                y = _conv2d(t[op["src"]], w, b, op["s"], op["p"], op["g"])
                t[op["name"]] = y
                if op["bn"]:
                    mean, var, sf = (a.reshape(-1) for a in self._blobs(op["bn"]))
                    scale = np.float32(0.0) if sf[0] == 0 else np.float32(1.0) / sf[0]
                    gamma, beta = (a.reshape(-1) for a in self._blobs(op["bn"] + "_scale"))
                    m = (mean * scale)[None, :, None, None]
                    v = (var * scale)[None, :, None, None]
                    y = (y - m) / np.sqrt(v + np.float32(op["eps"]))
                    y = y * gamma[None, :, None, None] + beta[None, :, None, None]
                    y = y.astype(np.float32)
                    t[op["bn"]] = y
                if op["relu"]:
                    y = np.maximum(y, np.float32(0))
                    t[op["relu"]] = y
            elif kind == "deconv":
                w = self._blobs(op["name"])[0].reshape(op["c"], 1, 4, 4)
                t[op["name"]] = _deconv_dw_k4s2p1(t[op["src"]], w)
            elif kind == "crop":
                h, wd = t[op["like"]].shape[2:]
                t[op["name"]] = t[op["src"]][:, :, :h, :wd]
            elif kind == "add":
                t[op["name"]] = t[op["a"]] + t[op["b"]]
            elif kind == "concat":
                t[op["name"]] = np.concatenate([t[s] for s in op["srcs"]], axis=1)
            elif kind == "relu":
                t[op["name"]] = np.maximum(t[op["src"]], np.float32(0))
            elif kind == "cls_softmax":
                s = op["stride"]
                x = t[op["src"]]
                n, c, h, wd = x.shape
                v = x.reshape(n, 2, (c // 2) * h, wd)
                v = v - v.max(axis=1, keepdims=True)
                e = np.exp(v)
                pr = (e / e.sum(axis=1, keepdims=True)).astype(np.float32)
                t[f"face_rpn_cls_prob_reshape_stride{s}"] = pr.reshape(n, c, h, wd)
        return {k: t[k] for k in want}


def folded_params(caffemodel_path: str) -> Dict[str, dict]:
    """BN+Scale(+bias) folded per-conv (w', b') in FP32: the numbers the CUDA engine must hold.

    w' = w * gamma / sqrt(var/sf + eps);  b' = (bias - mean/sf) * gamma / sqrt(var/sf+eps) + beta
    """
    layers = load_caffemodel(caffemodel_path)
    out = {}
    for op in topology.ops():
        if op["op"] == "conv":
            bl = layers[op["name"]]["blobs"]
            w = bl[0].reshape(op["cout"], op["cin"] // op["g"], op["k"], op["k"]).astype(np.float64)
            b = bl[1].reshape(-1).astype(np.float64) if op["bias"] else np.zeros(op["cout"])
            if op["bn"]:
                mean, var, sf = (a.reshape(-1).astype(np.float64) for a in layers[op["bn"]]["blobs"])
                sc = 0.0 if sf[0] == 0 else 1.0 / sf[0]
                gamma, beta = (a.reshape(-1).astype(np.float64) for a in layers[op["bn"] + "_scale"]["blobs"])
                k = gamma / np.sqrt(var * sc + op["eps"])
                w = w * k[:, None, None, None]
                b = (b - mean * sc) * k + beta
            out[op["name"]] = dict(w=w.astype(np.float32), b=b.astype(np.float32))
        elif op["op"] == "deconv":
            out[op["name"]] = dict(w=layers[op["name"]]["blobs"][0].reshape(op["c"], 1, 4, 4))
    return out
