"""mnet25 / mnet-deconv-0517 layer graph, restated from the reference prototxt.

Test infrastructure -- see ``oracle/__init__.py``.

Source of truth restated here: ``/root/reference/model/mnet-deconv-0517.prototxt``
(2302 lines; ``model/mnet25.prototxt`` has the identical layer list).
Line numbers in comments refer to that file.  ``tests/test_oracle_forward.py`` checks this
restatement against the reference prototxt itself (cv2.dnn on both, bit-for-bit equal
outputs) whenever /root/reference is present.

``ops()`` yields a flat, ordered list of dict ops; ``to_prototxt()`` renders them as a
Caffe prototxt so that cv2.dnn can execute the graph on machines where /root/reference
does not exist (the GPU box) -- this is what the CPU baseline in ``bench.py`` runs.
"""
from __future__ import annotations

from typing import Dict, List

BACKBONE_EPS = 1e-5   # prototxt:35 (all 27 backbone BatchNorm layers)
HEAD_EPS = 2e-5       # prototxt:1222 (all 20 rf_* BatchNorm layers)

# (cout, kind, stride) for mobilenet0_conv{i}_fwd, i = 0..26   (prototxt:11-1192)
_BACKBONE = [
    (8, "full", 2),
    (8, "dw", 1), (16, "pw", 1),
    (16, "dw", 2), (32, "pw", 1),
    (32, "dw", 1), (32, "pw", 1),
    (32, "dw", 2), (64, "pw", 1),
    (64, "dw", 1), (64, "pw", 1),      # relu10 = C1 (stride 8)   prototxt:488
    (64, "dw", 2), (128, "pw", 1),
    (128, "dw", 1), (128, "pw", 1),
    (128, "dw", 1), (128, "pw", 1),
    (128, "dw", 1), (128, "pw", 1),
    (128, "dw", 1), (128, "pw", 1),
    (128, "dw", 1), (128, "pw", 1),    # relu22 = C2 (stride 16)  prototxt:1016
    (128, "dw", 2), (256, "pw", 1),
    (256, "dw", 1), (256, "pw", 1),    # relu26 = C3 (stride 32)  prototxt:1192
]

OUTPUT_BLOBS = [  # order of retinaface/tensorrt/trtretinafacenet.cpp:23-31
    "face_rpn_cls_prob_reshape_stride32", "face_rpn_bbox_pred_stride32", "face_rpn_landmark_pred_stride32",
    "face_rpn_cls_prob_reshape_stride16", "face_rpn_bbox_pred_stride16", "face_rpn_landmark_pred_stride16",
    "face_rpn_cls_prob_reshape_stride8", "face_rpn_bbox_pred_stride8", "face_rpn_landmark_pred_stride8",
]


def _conv(name, src, cin, cout, k, s, p, g, bias, bn=None, eps=None, relu=None):
    return dict(op="conv", name=name, src=src, cin=cin, cout=cout, k=k, s=s, p=p, g=g,
                bias=bias, bn=bn, eps=eps, relu=relu)


def _head_conv(name, src, cin, cout, k, relu):
    # rf_* convs: bias_term true, BN eps 2e-5, optional ReLU   (prototxt:1199-1238 pattern)
    return _conv(name, src, cin, cout, k, 1, 1 if k == 3 else 0, 1, True,
                 bn=name + "_bn", eps=HEAD_EPS, relu=(name + "_relu") if relu else None)


def _ssh(level: str, src: str, stride: int) -> List[dict]:
    """SSH context head + the three 1x1 predictors of one FPN level (prototxt:1239-1512)."""
    p = f"rf_{level}_det"
    o = [
        _head_conv(f"{p}_conv1", src, 64, 32, 3, False),
        _head_conv(f"{p}_context_conv1", src, 64, 16, 3, True),
        _head_conv(f"{p}_context_conv2", f"{p}_context_conv1_relu", 16, 16, 3, False),
        _head_conv(f"{p}_context_conv3_1", f"{p}_context_conv1_relu", 16, 16, 3, True),
        _head_conv(f"{p}_context_conv3_2", f"{p}_context_conv3_1_relu", 16, 16, 3, False),
        dict(op="concat", name=f"{p}_concat",
             srcs=[f"{p}_conv1_bn", f"{p}_context_conv2_bn", f"{p}_context_conv3_2_bn"]),
        dict(op="relu", name=f"{p}_concat_relu", src=f"{p}_concat"),
        _conv(f"face_rpn_cls_score_stride{stride}", f"{p}_concat_relu", 64, 4, 1, 1, 0, 1, True),
        dict(op="cls_softmax", stride=stride, src=f"face_rpn_cls_score_stride{stride}"),
        _conv(f"face_rpn_bbox_pred_stride{stride}", f"{p}_concat_relu", 64, 8, 1, 1, 0, 1, True),
        _conv(f"face_rpn_landmark_pred_stride{stride}", f"{p}_concat_relu", 64, 20, 1, 1, 0, 1, True),
    ]
    return o


def ops() -> List[dict]:
    o: List[dict] = []
    src, cin = "data", 3
    for i, (cout, kind, s) in enumerate(_BACKBONE):
        name = f"mobilenet0_conv{i}_fwd"
        k, p, g = (1, 0, 1) if kind == "pw" else (3, 1, cin if kind == "dw" else 1)
        o.append(_conv(name, src, cin, cout, k, s, p, g, False,
                       bn=f"mobilenet0_batchnorm{i}_fwd", eps=BACKBONE_EPS,
                       relu=f"mobilenet0_relu{i}_fwd"))
        src, cin = f"mobilenet0_relu{i}_fwd", cout
    # FPN top level (prototxt:1199-1238) and its SSH head (stride 32)
    o.append(_head_conv("rf_c3_lateral", "mobilenet0_relu26_fwd", 256, 64, 1, True))
    o += _ssh("c3", "rf_c3_lateral_relu", 32)
    # stride 16 (prototxt:1513-1906)
    o.append(_head_conv("rf_c2_lateral", "mobilenet0_relu22_fwd", 128, 64, 1, True))
    o.append(dict(op="deconv", name="rf_c3_upsampling", src="rf_c3_lateral_relu", c=64))
    o.append(dict(op="crop", name="crop0", src="rf_c3_upsampling", like="rf_c2_lateral_relu"))
    o.append(dict(op="add", name="_plus0", a="rf_c2_lateral_relu", b="crop0"))
    o.append(_head_conv("rf_c2_aggr", "_plus0", 64, 64, 3, True))
    o += _ssh("c2", "rf_c2_aggr_relu", 16)
    # stride 8 (prototxt:1908-2302)
    o.append(_head_conv("rf_c1_red_conv", "mobilenet0_relu10_fwd", 64, 64, 1, True))
    o.append(dict(op="deconv", name="rf_c2_upsampling", src="rf_c2_aggr_relu", c=64))
    o.append(dict(op="crop", name="crop1", src="rf_c2_upsampling", like="rf_c1_red_conv_relu"))
    o.append(dict(op="add", name="_plus1", a="rf_c1_red_conv_relu", b="crop1"))
    o.append(_head_conv("rf_c1_aggr", "_plus1", 64, 64, 3, True))
    o += _ssh("c1", "rf_c1_aggr_relu", 8)
    return o


def conv_macs(h: int, w: int) -> Dict[str, int]:
    """MACs per image by class for an HxW input (multiples of 32)."""
    shape = {"data": (h, w)}
    out = {"full3x3": 0, "pw": 0, "dw": 0, "deconv": 0}
    for op in ops():
        if op["op"] == "conv":
            ih, iw = shape[op["src"]]
            oh = (ih + 2 * op["p"] - op["k"]) // op["s"] + 1
            ow = (iw + 2 * op["p"] - op["k"]) // op["s"] + 1
            macs = oh * ow * op["cout"] * (op["cin"] // op["g"]) * op["k"] ** 2
            cls = "dw" if op["g"] > 1 else ("pw" if op["k"] == 1 else "full3x3")
            out[cls] += macs
            for t in (op["name"], op["bn"], op["relu"]):
                if t:
                    shape[t] = (oh, ow)
        elif op["op"] == "deconv":
            ih, iw = shape[op["src"]]
            shape[op["name"]] = (2 * ih, 2 * iw)
            out["deconv"] += 4 * ih * 4 * iw * op["c"]  # 16 taps per input px... = 4 per output px
        elif op["op"] == "crop":
            shape[op["name"]] = shape[op["like"]]
        elif op["op"] == "add":
            shape[op["name"]] = shape[op["a"]]
        elif op["op"] == "concat":
            shape[op["name"]] = shape[op["srcs"][0]]
        elif op["op"] == "relu":
            shape[op["name"]] = shape[op["src"]]
    out["total"] = sum(out.values())
    return out


def to_prototxt(h: int, w: int, n: int = 1) -> str:
    """Render the graph as Caffe prototxt text for an n x 3 x h x w input."""
    L: List[str] = ['name: "rf_b200_oracle_mnet"']
    L.append('layer { name: "data" type: "Input" top: "data" input_param { shape: '
             f'{{ dim: {n} dim: 3 dim: {h} dim: {w} }} }} }}')
    for op in ops():
        if op["op"] == "conv":
            L.append(
                f'layer {{ name: "{op["name"]}" type: "Convolution" bottom: "{op["src"]}" top: "{op["name"]}" '
                f'convolution_param {{ num_output: {op["cout"]} kernel_size: {op["k"]} pad: {op["p"]} '
                f'group: {op["g"]} stride: {op["s"]} bias_term: {"true" if op["bias"] else "false"} }} }}')
            top = op["name"]
            if op["bn"]:
                L.append(
                    f'layer {{ name: "{op["bn"]}" type: "BatchNorm" bottom: "{top}" top: "{op["bn"]}" '
                    f'batch_norm_param {{ use_global_stats: true eps: {op["eps"]} }} }}')
                L.append(
                    f'layer {{ name: "{op["bn"]}_scale" type: "Scale" bottom: "{op["bn"]}" top: "{op["bn"]}" '
                    f'scale_param {{ bias_term: true }} }}')
                top = op["bn"]
            if op["relu"]:
                L.append(f'layer {{ name: "{op["relu"]}" type: "ReLU" bottom: "{top}" top: "{op["relu"]}" }}')
        elif op["op"] == "deconv":
            L.append(
                f'layer {{ name: "{op["name"]}" type: "Deconvolution" bottom: "{op["src"]}" top: "{op["name"]}" '
                f'convolution_param {{ kernel_size: 4 stride: 2 pad: 1 num_output: {op["c"]} group: {op["c"]} '
                f'bias_term: false }} }}')
        elif op["op"] == "crop":
            L.append(
                f'layer {{ name: "{op["name"]}" type: "Crop" bottom: "{op["src"]}" bottom: "{op["like"]}" '
                f'top: "{op["name"]}" crop_param {{ axis: 2 offset: 0 offset: 0 }} }}')
        elif op["op"] == "add":
            L.append(
                f'layer {{ name: "{op["name"]}" type: "Eltwise" bottom: "{op["a"]}" bottom: "{op["b"]}" '
                f'top: "{op["name"]}" eltwise_param {{ operation: SUM }} }}')
        elif op["op"] == "concat":
            bots = " ".join(f'bottom: "{s}"' for s in op["srcs"])
            L.append(f'layer {{ name: "{op["name"]}" type: "Concat" {bots} top: "{op["name"]}" }}')
        elif op["op"] == "relu":
            L.append(f'layer {{ name: "{op["name"]}" type: "ReLU" bottom: "{op["src"]}" top: "{op["name"]}" }}')
        elif op["op"] == "cls_softmax":
            s = op["stride"]
            L.append(
                f'layer {{ name: "face_rpn_cls_score_reshape_stride{s}" type: "Reshape" bottom: "{op["src"]}" '
                f'top: "face_rpn_cls_score_reshape_stride{s}" reshape_param {{ shape {{ dim: 0 dim: 2 dim: -1 dim: 0 }} }} }}')
            L.append(
                f'layer {{ name: "face_rpn_cls_prob_stride{s}" type: "Softmax" '
                f'bottom: "face_rpn_cls_score_reshape_stride{s}" top: "face_rpn_cls_prob_stride{s}" }}')
            L.append(
                f'layer {{ name: "face_rpn_cls_prob_reshape_stride{s}" type: "Reshape" '
                f'bottom: "face_rpn_cls_prob_stride{s}" top: "face_rpn_cls_prob_reshape_stride{s}" '
                f'reshape_param {{ shape {{ dim: 0 dim: 4 dim: -1 dim: 0 }} }} }}')
    return "\n".join(L) + "\n"
