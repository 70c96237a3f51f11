"""CPU: pins the plain-C post-process restatement (oracle/postproc.c) against the reference's own
compiled code (oracle/_ref = /root/reference/retinaface/RetinaFace.cpp built unmodified behind a
fake engine) and against the committed golden detections."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.postproc import STRIDES, PostprocOracle, ReferencePostproc, synth_heads

needs_ref = pytest.mark.skipif(not ReferencePostproc.available(), reason="oracle/_ref not built (no /root/reference)")


@pytest.fixture(scope="module")
def oracle():
    return PostprocOracle()


def test_base_anchors_values(oracle):
    # values printed by the reference's own generate_anchors_fpn
    exp = {32: [[-248, -248, 263, 263], [-120, -120, 135, 135]], 16: [[-56, -56, 71, 71], [-24, -24, 39, 39]],
           8: [[-8, -8, 23, 23], [0, 0, 15, 15]]}
    for s in STRIDES:
        assert oracle.base_anchors(s).tolist() == exp[s]


@needs_ref
@pytest.mark.parametrize("hw", [(448, 448), (896, 1280), (320, 320)])
def test_anchors_match_reference(oracle, hw):
    ref = ReferencePostproc(*hw)
    try:
        for s in STRIDES:
            base = oracle.base_anchors(s)
            assert np.array_equal(base, ref.base_anchors(s))
            plane = ref.anchor_plane(s)  # anchors_plane: index k*H*W + ih*W + iw
            h, w = hw[0] // s, hw[1] // s
            ys, xs = np.mgrid[0:h, 0:w]
            for k in range(2):
                mine = np.stack([base[k, 0] + xs * s, base[k, 1] + ys * s, base[k, 2] + xs * s, base[k, 3] + ys * s], -1)
                assert np.array_equal(mine.reshape(-1, 4).astype(np.float32), plane[k * h * w:(k + 1) * h * w])
    finally:
        ref.close()


@needs_ref
@pytest.mark.parametrize("hw", [(448, 448), (896, 1280)])
@pytest.mark.parametrize("ncand", [0, 1, 7, 64, 1024, 4000])
def test_postprocess_bit_exact_vs_reference(oracle, hw, ncand):
    heads = synth_heads(hw[0], hw[1], ncand, seed=ncand + 11)
    ref = ReferencePostproc(*hw)
    try:
        for thr in (0.9, 0.5):
            mine = oracle.postprocess(heads, hw[0], hw[1], thr, 0.4)   # reference postProcess hard-codes NMS 0.4
            theirs = ref.postprocess(heads, thr)
            assert np.array_equal(mine["faces"], theirs)
            if thr == 0.9:
                assert len(mine["cand"]) == ncand
        # RetinaFace::nms with other thresholds, on the same candidates
        cands = oracle.postprocess(heads, hw[0], hw[1], 0.9, 0.4)["cand"]
        for nt in (0.0, 0.3, 0.7, 1.0):
            a, _ = oracle.nms(cands, nt)
            assert np.array_equal(a, ref.nms(cands, nt))
    finally:
        ref.close()


def test_edge_cases(oracle):
    h = w = 64
    heads = synth_heads(h, w, 0)
    r = oracle.postprocess(heads, h, w, 0.9, 0.4)
    assert len(r["faces"]) == 0 and len(r["cand"]) == 0
    # every anchor is a candidate
    heads = synth_heads(h, w, 10_000)
    r = oracle.postprocess(heads, h, w, 0.9, 0.4)
    assert len(r["cand"]) == 2 * (2 * 2 + 4 * 4 + 8 * 8)
    # strict threshold: conf == thr is dropped (RetinaFace.cpp:693 `conf <= threshold`)
    heads = synth_heads(h, w, 0)
    heads[0][2, 0, 0] = np.float32(0.9)
    assert len(oracle.postprocess(heads, h, w, np.float32(0.9), 0.4)["cand"]) == 0
    heads[0][2, 0, 0] = np.nextafter(np.float32(0.9), np.float32(1))
    assert len(oracle.postprocess(heads, h, w, np.float32(0.9), 0.4)["cand"]) == 1
    # ties: equal scores keep emission order
    heads = synth_heads(h, w, 0)
    heads[6][2, 0, 0] = 0.95   # stride 8, anchor 0, j 0
    heads[0][2, 1, 1] = 0.95   # stride 32, anchor 0, j 3 -> emitted first
    r = oracle.postprocess(heads, h, w, 0.9, 1.0)
    assert r["idx"].tolist() == sorted(r["idx"].tolist()) and len(r["idx"]) == 2


@pytest.mark.parametrize("model", ["mnet-deconv-0517", "mnet25"])
def test_golden_detections_from_golden_heads(oracle, model):
    """heads frozen from cv2.dnn(reference prototxt+caffemodel) -> oracle == detections frozen from oracle/_ref."""
    heads_npz = np.load(os.path.join(GOLDEN, f"heads_{model}_448.npz"))
    from oracle.topology import OUTPUT_BLOBS
    heads = [heads_npz[n] for n in OUTPUT_BLOBS]
    dets = np.load(os.path.join(GOLDEN, f"dets_{model}_448x448.npz"))
    for thr in (0.9, 0.5, 0.02):
        r = oracle.postprocess(heads, 448, 448, thr, 0.4)
        assert np.array_equal(r["faces"], dets[f"faces_thr{thr}"])
    assert dets["faces_thr0.9"].shape == (5, 15)   # 27 candidates -> 5 faces
