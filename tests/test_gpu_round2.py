"""GPU (-m gpu): round-2 parity -- the persistent tile chains against the round-1 per-layer kernels and the FP32 engine,
BASELINE.json configs[3] exactly as stated, FP16 / INT8 detections of EVERY image of a batch against FP32 detections of the
same images (match rate printed), the distribution of the FP16 head-tensor error, and the multi-GPU exchange fused into
the NMS kernel (two handles of one process standing in for two ranks).  Everything goes through the C ABI.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, caffemodel
from oracle import topology
from oracle.inputs import letterbox_bgr_u8, s_noise_batch, s_real_batch

pytestmark = pytest.mark.gpu


def _engine(model, h, w, prec, **kw):
    from retinaface_b200 import Engine
    return Engine(caffemodel(model), h, w, precision=prec, **kw)


def _iou(a, b):
    x1, y1, x2, y2 = max(a[0], b[0]), max(a[1], b[1]), min(a[2], b[2]), min(a[3], b[3])
    iw, ih = max(0.0, x2 - x1 + 1), max(0.0, y2 - y1 + 1)
    inter = iw * ih
    return inter / ((a[2] - a[0] + 1) * (a[3] - a[1] + 1) + (b[2] - b[0] + 1) * (b[3] - b[1] + 1) - inter)


def _match(mine, ref, iou_min=0.7):
    """Greedy one-to-one matching of two face lists by IoU.  Returns (pairs, unmatched_mine, unmatched_ref)."""
    pairs, used = [], set()
    for i, m in enumerate(mine):
        best, bj = 0.0, -1
        for j, r in enumerate(ref):
            if j in used:
                continue
            v = _iou(m[1:5], r[1:5])
            if v > best:
                best, bj = v, j
        if bj >= 0 and best >= iou_min:
            used.add(bj)
            pairs.append((i, bj))
    return pairs, len(mine) - len(pairs), len(ref) - len(pairs)


@pytest.mark.parametrize("hw,nb", [((448, 448), 5), ((896, 1280), 2), ((288, 416), 3)])
@pytest.mark.parametrize("mask", [511, 490, 255, 127, 63])
def test_tile_chains_equal_round1_kernels(mask, hw, nb, golden_image, monkeypatch):
    """The tile-chain plan (RF_TILE_MASK: 511 every chain; 490 the default selection; 255 stand-alone NMS; 127 stand-alone predictors
    + NMS; 63 round-1 SSH)
    against one round-1 kernel per layer (RF_FLAG_LEGACY_TC): every tensor both plans materialise within 2e-2 of its max
    (depthwise weights are FP16 diagonal tiles in the chains, FP32 in the round-1 stencil), head blobs within 1e-2, and the
    SAME faces (anchor indices) within 0.25 px / 5e-3 score -- on the photo, on noise and on shifted copies."""
    from retinaface_b200 import RF_PREC_FP16
    from retinaface_b200.capi import RF_FLAG_LEGACY_TC, RfError
    h, w = hw
    inp = letterbox_bgr_u8(golden_image, h, w)
    batch = np.stack([inp, s_noise_batch(1, h, w, seed=1)[0]] + [np.roll(inp, 24 * k, axis=1) for k in range(1, nb - 1)])
    monkeypatch.setenv("RF_TILE_MASK", str(mask))
    new = _engine("mnet25", h, w, RF_PREC_FP16, max_batch=nb)
    old = _engine("mnet25", h, w, RF_PREC_FP16, max_batch=nb, flags=RF_FLAG_LEGACY_TC)
    try:
        new.debug_keep_all()
        old.debug_keep_all()
        hn, ho = new.forward_heads(batch), old.forward_heads(batch)
        common = 0
        for name in ["mobilenet0_relu2_fwd", "mobilenet0_relu6_fwd", "mobilenet0_relu10_fwd", "rf_c1_red_conv_relu", "mobilenet0_relu16_fwd",
                     "mobilenet0_relu22_fwd", "rf_c2_lateral_relu", "mobilenet0_relu24_fwd", "mobilenet0_relu26_fwd", "rf_c3_lateral_relu",
                     "rf_c3_det_concat_relu", "rf_c2_aggr_relu", "rf_c2_det_concat_relu", "rf_c1_aggr_relu", "rf_c1_det_concat_relu"]:
            try:
                a, b = new.debug_tensor(name, nb), old.debug_tensor(name, nb)
            except RfError:
                continue
            common += 1
            e = float(np.abs(a - b).max() / (np.abs(b).max() or 1.0))
            assert e < 2e-2, (name, e)
        assert common >= 8
        for k in range(9):
            assert np.abs(hn[k] - ho[k]).max() < 1e-2, k
        fn, idn = new.detect_batch(list(batch), 0.9, 0.4, want_index=True)
        fo, ido = old.detect_batch(list(batch), 0.9, 0.4, want_index=True)
        for i in range(nb):
            assert idn[i].tolist() == ido[i].tolist(), i
            if len(fn[i]):
                assert np.abs(fn[i][:, 1:] - fo[i][:, 1:]).max() < 0.25 and np.abs(fn[i][:, 0] - fo[i][:, 0]).max() < 5e-3, i
        assert len(fn[0]) >= 4 and len(fn[1]) == 0
        # graph replay + self-cleaning last-block counters: three more runs, bit-identical
        for _ in range(3):
            again = new.detect_batch(list(batch), 0.9, 0.4)
            for i in range(nb):
                assert again[i].shape == fn[i].shape and np.array_equal(again[i], fn[i])
        # smaller batches through the same engine (other tile counts per launch)
        one = new.detect_batch([batch[0]], 0.9, 0.4)
        assert np.array_equal(one[0], fn[0])
    finally:
        new.close()
        old.close()


@pytest.mark.parametrize("mask", [511, 482])
@pytest.mark.parametrize("model", ["mnet25", "mnet-deconv-0517"])
def test_tile_plans_against_golden_fp32(mask, model, golden_image, monkeypatch):
    """The tile-chain plans held to the same bars as the round-1 FP16 engine (tests/test_gpu_parity.py::test_fp16_forward_and_detect):
    head blobs vs the golden FP32 heads (cls_prob 5e-3, deltas 2e-2 over ALL anchors), detections on the golden photo vs the
    golden FP32 detections (scores 1e-3, coordinates 0.1 px: north_star's FP16 tolerance), and -- decode + NMS running inside
    the SSH chain -- its own heads through the oracle post-process == its own detections, selection bit-exact."""
    from oracle.postproc import PostprocOracle
    from retinaface_b200 import RF_PREC_FP16
    monkeypatch.setenv("RF_TILE_MASK", str(mask))
    eng = _engine(model, 448, 448, RF_PREC_FP16, max_batch=8)
    try:
        inp = letterbox_bgr_u8(golden_image, 448, 448)
        batch = s_real_batch(inp, 8)
        heads = eng.forward_heads(batch)
        gold = np.load(os.path.join(GOLDEN, f"heads_{model}_448.npz"))
        for k, name in enumerate(topology.OUTPUT_BLOBS):
            err = np.abs(heads[k][0] - gold[name]).max()
            assert err < (5e-3 if "cls_prob" in name else 2e-2), (name, err)
        dets = np.load(os.path.join(GOLDEN, f"dets_{model}_448x448.npz"))["faces_thr0.9"]
        faces, idx = eng.detect_batch(list(batch), 0.9, 0.4, want_index=True)
        assert faces[0].shape == dets.shape
        assert np.abs(faces[0][:, 0] - dets[:, 0]).max() < 1e-3
        assert np.abs(faces[0][:, 1:] - dets[:, 1:]).max() < 0.1
        post = PostprocOracle()
        for i in range(8):
            ref = post.postprocess([x[i] for x in heads], 448, 448, 0.9, 0.4)
            assert idx[i].tolist() == ref["idx"].tolist(), i
            assert np.array_equal(faces[i][:, 0], ref["faces"][:, 0]) and np.array_equal(faces[i][:, 5:], ref["faces"][:, 5:]), i
            assert np.allclose(faces[i][:, 1:5], ref["faces"][:, 1:5], rtol=4e-6, atol=1e-4), i
    finally:
        eng.close()


def test_config4_mnet25_fp16_b8_1280x896(golden_image):
    """BASELINE.json configs[3] exactly: mnet25, FP16, batch 8, 1280x896 (47,040 anchors / image).  Image 0 against the golden
    FP32 detections: scores <= 1e-3, boxes and landmarks <= 0.1 px (north_star's FP16 bar); ALL 8 images against the FP32
    engine's detections of the same images: every face matched, scores <= 2e-3, coordinates <= 0.2 px."""
    from retinaface_b200 import RF_PREC_FP16, RF_PREC_FP32
    h, w = 896, 1280
    inp = letterbox_bgr_u8(golden_image, h, w)
    batch = s_real_batch(inp, 8)
    e16 = _engine("mnet25", h, w, RF_PREC_FP16, max_batch=8)
    e32 = _engine("mnet25", h, w, RF_PREC_FP32, max_batch=8)
    try:
        f16 = e16.detect_batch(list(batch), 0.9, 0.4)
        f32 = e32.detect_batch(list(batch), 0.9, 0.4)
        gold = np.load(os.path.join(GOLDEN, f"dets_mnet25_{h}x{w}.npz"))["faces_thr0.9"]
        assert f16[0].shape == gold.shape
        assert np.abs(f16[0][:, 0] - gold[:, 0]).max() < 1e-3
        assert np.abs(f16[0][:, 1:] - gold[:, 1:]).max() < 0.1
        total = matched = 0
        for i in range(8):
            pairs, um, ur = _match(f16[i], f32[i])
            total += len(f32[i])
            matched += len(pairs)
            for a, b in pairs:
                assert abs(f16[i][a, 0] - f32[i][b, 0]) < 2e-3, (i, f16[i][a, :5], f32[i][b, :5])
                assert np.abs(f16[i][a, 1:] - f32[i][b, 1:]).max() < 0.2, (i, f16[i][a, :5], f32[i][b, :5])
            assert len(f32[i]) >= 5
        print(f"config 4: {matched}/{total} FP32 faces matched by the FP16 engine over 8 images")
        assert matched == total
    finally:
        e16.close()
        e32.close()


@pytest.mark.parametrize("prec", ["fp16", "int8"])
def test_all_images_against_fp32_detections(prec, golden_image):
    """FP16 / INT8 detections of EVERY image -- 8 S-real images (the photo rolled by 8 i pixels), 8 S-noise images and 8 images
    of a second photo-derived family (mirrored + vertically shifted) -- against the FP32 engine's detections of the same
    images (the FP32 engine is held to the oracle at 2e-3 px elsewhere).  Reported: match rate over all FP32 faces; gated:
    FP16 all faces matched with scores <= 2e-3 / coordinates <= 0.2 px; INT8 (mnet-deconv-0517 + the reference's table)
    match rate >= 0.9 with scores <= 0.05 / coordinates <= 6 px -- the calibration's own tolerance (observed on a B200: 77 / 79
    matched, 1 extra face, worst score difference 0.023, worst coordinate 4.6 px on a 200-px face)."""
    from retinaface_b200 import RF_PREC_FP16, RF_PREC_FP32, RF_PREC_INT8, Engine
    model = "mnet25" if prec == "fp16" else "mnet-deconv-0517"
    table = os.path.join(GOLDEN, "weights", model + ".table.int8")
    inp = letterbox_bgr_u8(golden_image, 448, 448)
    fam2 = np.ascontiguousarray(inp[:, ::-1])
    batches = [s_real_batch(inp, 8), s_noise_batch(8, 448, 448, seed=0), np.stack([np.roll(np.roll(fam2, 16 * i, axis=1), 4 * i, axis=0) for i in range(8)])]
    eng = Engine(caffemodel(model), 448, 448, precision=RF_PREC_FP16 if prec == "fp16" else RF_PREC_INT8, max_batch=8,
                 int8_table=table if prec == "int8" else None)
    ref = _engine(model, 448, 448, RF_PREC_FP32, max_batch=8)
    tol_s, tol_px = (2e-3, 0.2) if prec == "fp16" else (0.05, 6.0)
    try:
        total = matched = extra = 0
        worst_s = worst_px = 0.0
        for batch in batches:
            mine = eng.detect_batch(list(batch), 0.9, 0.4)
            gold = ref.detect_batch(list(batch), 0.9, 0.4)
            for i in range(len(batch)):
                pairs, um, ur = _match(mine[i], gold[i])
                total += len(gold[i]); matched += len(pairs); extra += um
                for a, b in pairs:
                    worst_s = max(worst_s, abs(float(mine[i][a, 0] - gold[i][b, 0])))
                    worst_px = max(worst_px, float(np.abs(mine[i][a, 1:] - gold[i][b, 1:]).max()))
        rate = matched / max(total, 1)
        print(f"{prec}: {matched}/{total} FP32 faces matched (rate {rate:.3f}), {extra} extra faces, worst score diff {worst_s:.2e}, worst coordinate diff {worst_px:.3f} px")
        assert total >= 60
        assert worst_s < tol_s and worst_px < tol_px
        if prec == "fp16":
            assert matched == total and extra == 0
        else:
            assert rate >= 0.9 and extra <= 0.1 * total
    finally:
        eng.close()
        ref.close()


def test_fp16_head_tensor_error_distribution(golden_image):
    """How far the FP16 engine's 9 head blobs are from the golden FP32 ones, over ALL anchors of the golden photo (printed: mean,
    p99, p99.9, max per blob).  north_star asks 1e-3 for FP16; what a B200 run shows (profiles/README.md): cls_prob max 3e-5 at
    stride 32, 1.8e-3 at stride 16, ~3e-3 at stride 8; regression deltas max 2e-3 .. 1e-2.  The floor is not the predictor
    arithmetic (FP32-grade: hi + lo FP16 weights, FP32 accumulate) but FP16 STORAGE of ~30 layers of activations in front of it
    (11-bit mantissa on values up to ~30: 1e-2 absolute per tensor), which a probability near 0.5 sees at 1/4 of the logit
    error.  Gates: cls_prob p99 <= 1e-3 and max <= 5e-3, and <= 1e-3 wherever P(face) > 0.9 -- the anchors that become
    detections at the reference's threshold (main.cpp:43), whose scores test_fp16_forward_and_detect holds to 1e-3; deltas p99 <= 5e-3,
    max <= 2e-2."""
    from retinaface_b200 import RF_PREC_FP16
    eng = _engine("mnet25", 448, 448, RF_PREC_FP16, max_batch=1)
    try:
        inp = letterbox_bgr_u8(golden_image, 448, 448)
        heads = eng.forward_heads(inp[None])
        gold = np.load(os.path.join(GOLDEN, "heads_mnet25_448.npz"))
        bad = []
        for k, name in enumerate(topology.OUTPUT_BLOBS):
            err = np.abs(heads[k][0] - gold[name]).ravel()
            p99, mx = float(np.quantile(err, 0.99)), float(err.max())
            print(f"{name:40s} mean {err.mean():.2e}  p99 {p99:.2e}  p99.9 {np.quantile(err, 0.999):.2e}  max {mx:.2e}")
            if "cls_prob" in name:
                ok = p99 < 1e-3 and mx < 5e-3
                face = gold[name][2:] > 0.9
                if face.any():
                    ok &= float(np.abs(heads[k][0][2:] - gold[name][2:])[face].max()) < 1e-3
            else:
                ok = p99 < 5e-3 and mx < 2e-2
            if not ok:
                bad.append((name, p99, mx))
        assert not bad, bad
    finally:
        eng.close()


def test_comm_allgather_fused_into_nms_two_ranks(tmp_path):
    """The multi-GPU exchange (csrc/comm.cu) with two RANKS = two processes, each with its own handle, sharing ONE GPU (CUDA IPC
    windows, blobs exchanged through files -- tests/comm_worker.py): every rank's rf_submit_batch_allgather /
    rf_collect_batch_allgather returns the faces of BOTH ranks (rank r's image i at row r * max_batch + i), bit-equal to what each
    rank detects locally, for several steps in flight (ring slots, sequence numbers); a plain local call still works afterwards.
    (One process per rank as in production: two handles of one process could deadlock on hardware-queue aliasing, rank B's
    kernels queued behind rank A's spinning wait kernel.  tools/comm_check.py is the torchrun / multi-GPU version.)"""
    import subprocess
    import sys
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "comm_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(tmp_path), "0"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = []
    for p_ in procs:
        try:
            out, _ = p_.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            p_.kill()
            out, _ = p_.communicate()
        outs.append(out)
    for r, (p_, out) in enumerate(zip(procs, outs)):
        assert p_.returncode == 0, f"rank {r}:\n{out[-3000:]}"
        assert "0 mismatches" in out, out[-1000:]


def test_letterbox_npp_super_sampling_semantics(golden_image):
    """RF_FLAG_NPP_RESIZE: the reference's USE_NPP letter-box (resizeconvertion.cu:279-316) against NPP itself (oracle/npp_oracle.cu
    calls nppiResizeSqrPixel_8u_C3R exactly as the reference does) on 8 shapes: same extent, every byte within 1 LSB, at most
    0.5 % of the bytes off by that 1 LSB (NPP accumulates in single precision), byte-identical on the shapes with integral
    coverage.  And what the choice of branch does to the detections of the golden photo (printed)."""
    from oracle.npp import npp_letterbox
    from retinaface_b200 import RF_PREC_FP16, Engine
    from retinaface_b200.capi import RF_FLAG_NPP_RESIZE
    rng = np.random.default_rng(5)
    cases = [(golden_image, 448, 448), (golden_image, 320, 320), (golden_image[100:600, 200:900], 448, 448),
             (rng.integers(0, 256, (181, 297, 3), dtype=np.uint8), 96, 160), (rng.integers(0, 256, (333, 1000, 3), dtype=np.uint8), 448, 448),
             (np.tile(np.arange(200, dtype=np.uint8).repeat(2)[None, :, None], (200, 1, 3)), 96, 160), (rng.integers(0, 256, (80, 100, 3), dtype=np.uint8), 448, 448),
             (rng.integers(0, 256, (449, 449, 3), dtype=np.uint8), 448, 448)]
    exact = 0
    for img, nh, nw in cases:
        img = np.ascontiguousarray(img)
        eng = Engine(caffemodel("mnet25"), nh, nw, precision=RF_PREC_FP16, max_batch=1, max_image=(max(img.shape[0], nh), max(img.shape[1], nw)), flags=RF_FLAG_NPP_RESIZE)
        try:
            mine = eng.preprocess(img)
        finally:
            eng.close()
        ref = npp_letterbox(img, nh, nw)
        diff = np.abs(mine.astype(int) - ref.astype(int))
        frac = np.count_nonzero(diff) / diff.size
        print(f"{img.shape[1]}x{img.shape[0]} -> {nw}x{nh}: max diff {diff.max()}, {np.count_nonzero(diff)} of {diff.size} bytes differ ({frac:.4%})")
        assert diff.max() <= 1 and frac <= 5e-3, (img.shape, nh, nw, diff.max(), frac)
        assert np.array_equal(mine.any(axis=2), ref.any(axis=2)) or frac < 5e-3
        exact += int(diff.max() == 0)
    assert exact >= 4
    # detections of the photo under either branch
    e_lin = Engine(caffemodel("mnet25"), 448, 448, precision=RF_PREC_FP16, max_batch=1, max_image=golden_image.shape[:2])
    e_npp = Engine(caffemodel("mnet25"), 448, 448, precision=RF_PREC_FP16, max_batch=1, max_image=golden_image.shape[:2], flags=RF_FLAG_NPP_RESIZE)
    try:
        a = e_lin.detect_batch([golden_image], 0.9, 0.4)[0]
        b = e_npp.detect_batch([golden_image], 0.9, 0.4)[0]
        pairs, ua, ub = _match(a, b)
        ds = max((abs(float(a[i, 0] - b[j, 0])) for i, j in pairs), default=0.0)
        dp = max((float(np.abs(a[i, 1:5] - b[j, 1:5]).max()) for i, j in pairs), default=0.0)
        print(f"golden photo, OpenCV-bilinear vs NPP-super-sampling letter-box: {len(a)} vs {len(b)} faces, {len(pairs)} matched, score diff <= {ds:.4f}, box diff <= {dp:.2f} px")
        assert len(pairs) >= 4
    finally:
        e_lin.close()
        e_npp.close()


def test_arbitrary_size_batch_is_one_letterbox_launch(golden_image):
    """rf_detect_batch on a batch of differently sized caller images (pinned or pageable): uploaded into per-image raw buffers and
    letter-boxed by one launch -- results equal those of the same images letter-boxed one by one through rf_preprocess."""
    from retinaface_b200 import RF_PREC_FP16, Engine
    imgs = [golden_image, np.ascontiguousarray(golden_image[:600, :900]), np.ascontiguousarray(golden_image[100:, 300:]), np.ascontiguousarray(golden_image[::2, ::2])]
    eng = Engine(caffemodel("mnet25"), 448, 448, precision=RF_PREC_FP16, max_batch=4, max_image=golden_image.shape[:2])
    try:
        together = eng.detect_batch(imgs, 0.9, 0.4)
        for i, im in enumerate(imgs):
            alone = eng.detect_batch([eng.preprocess(im)], 0.9, 0.4)[0]
            assert together[i].shape == alone.shape and np.array_equal(together[i], alone), i
        assert len(together[0]) >= 4
    finally:
        eng.close()


def test_engine_from_prototxt_with_cache(golden_image, tmp_path):
    """rf_create through the model front end: the prototxt supplies the network size (the reference reads it from prototxt line 7), the
    folded model is cached (miss, then hit) and the detections equal those of the built-in graph at the same size."""
    from retinaface_b200 import RF_PREC_FP16, Engine
    proto = tmp_path / "net.prototxt"
    proto.write_text(topology.to_prototxt(416, 288, 1))          # mnet25.prototxt's own size (H 416, W 288)
    cache = str(tmp_path / "net.rfcache")
    img = letterbox_bgr_u8(golden_image, 416, 288)
    ref = _engine("mnet25", 416, 288, RF_PREC_FP16, max_batch=1)
    want = ref.detect_batch([img], 0.5, 0.4)[0]
    ref.close()
    for expect in (1, 2):
        eng = Engine(caffemodel("mnet25"), 0, 0, precision=RF_PREC_FP16, max_batch=1, prototxt=str(proto), cache=cache, network="net3")
        try:
            assert (eng.net_h, eng.net_w) == (416, 288)
            assert eng.lib.rf_cache_status(eng.h) == expect
            got = eng.detect_batch([img], 0.5, 0.4)[0]
            assert got.shape == want.shape and np.array_equal(got, want)
        finally:
            eng.close()
    assert len(want) >= 1


def test_calibrator_under_the_references_conditions(golden_image, tmp_path):
    """rf_calibrate_int8 pinned against the one calibration output the reference holds (model/mnet-deconv-0517.table.int8), under the
    conditions its tool used: a 320 x 320 network fed UNRESIZED pixels, one image per batch (INT8-Calibration-Tool/
    CalibrationTableImpl.cpp:5-9, 28-34: BGR->RGB float, no resize, num_per_batch = 1).  The reference's calibration images are not
    shipped; here 96 unresized 320 x 320 crops of the one shipped photo (a 4 x 6 grid of positions x 4 scales-free offsets).  Printed:
    the distribution of (our scale / TensorRT's scale) over the tensors both tables name -- on a B200: 43 tensors, min 0.57, p10 0.67,
    median 0.94, p90 1.08, max 1.44 (TensorRT's calibrator and its images are closed / not shipped; the method family and the
    conditions are the same).  Gates (round 1: median within 0.5 .. 2): median in 0.8 .. 1.15, 80 % of the tensors within 0.65 .. 1.3,
    every tensor within 0.45 .. 1.8; and the table is good enough to run with: an INT8 engine built from it finds the FP32 faces."""
    from oracle.mnet_int8 import read_table
    from retinaface_b200 import RF_PREC_FP32, RF_PREC_INT8, Engine
    model = "mnet-deconv-0517"
    H, W = golden_image.shape[:2]
    crops = []
    for oy in range(0, H - 320 + 1, 80):
        for ox in range(0, W - 320 + 1, 64):
            crops.append(golden_image[oy:oy + 320, ox:ox + 320])
    crops = np.ascontiguousarray(np.stack(crops[:96]))
    assert crops.shape[0] >= 90
    table = str(tmp_path / "recalibrated.table.int8")
    fp32 = Engine(caffemodel(model), 320, 320, precision=RF_PREC_FP32, max_batch=1)       # num_per_batch = 1
    try:
        fp32.calibrate_int8(crops, table)
    finally:
        fp32.close()
    mine, shipped = read_table(table), read_table(os.path.join(GOLDEN, "weights", model + ".table.int8"))
    names = [k for k in mine if k in shipped and k != "data"]
    ratios = np.array([mine[k] / shipped[k] for k in names])
    q = np.quantile(ratios, [0.0, 0.1, 0.5, 0.9, 1.0])
    print(f"calibrator vs the shipped TensorRT table over {len(names)} tensors: ratio min {q[0]:.3f}  p10 {q[1]:.3f}  median {q[2]:.3f}  p90 {q[3]:.3f}  max {q[4]:.3f}")
    worst = sorted(zip(ratios, names))
    print("  lowest:", [(n, round(float(r), 3)) for r, n in worst[:3]], " highest:", [(n, round(float(r), 3)) for r, n in worst[-3:]])
    assert len(names) >= 40
    assert 0.8 < q[2] < 1.15, q
    assert ((ratios > 0.65) & (ratios < 1.3)).mean() >= 0.8, q
    assert q[0] > 0.45 and q[4] < 1.8, q
    inp = letterbox_bgr_u8(golden_image, 448, 448)
    eng = Engine(caffemodel(model), 448, 448, precision=RF_PREC_INT8, max_batch=1, int8_table=table)
    try:
        faces = eng.detect_batch([inp], 0.9, 0.4)[0]
        gold = np.load(os.path.join(GOLDEN, f"dets_{model}_448x448.npz"))["faces_thr0.9"]
        pairs, _, missing = _match(faces, gold)
        assert missing == 0, (len(faces), len(gold))
    finally:
        eng.close()


def test_jpeg_ingest_decodes_on_the_gpu(golden_image):
    """rf_detect_jpeg_batch (f1 ingest, compressed half; main.cpp:18-26 decodes with cv::imread on the host): JPEG bitstreams ->
    nvJPEG decode in device memory -> batched letter-box -> detect.  Checked against the host path on cv2.imdecode pixels of the
    same streams: the decoders (libjpeg-turbo vs nvJPEG) are allowed their IDCT / chroma-upsampling differences -- decoded bytes
    within a small mean error -- and the detections must be the same faces."""
    import cv2
    from retinaface_b200 import RF_PREC_FP16, Engine
    from retinaface_b200.capi import RfError
    raw = open(os.path.join(GOLDEN, "data", "img.jpg"), "rb").read()
    streams = [raw]
    for im, q, extra in [(golden_image[:, ::-1], 92, []), (golden_image[::2, ::2], 85, []), (golden_image[100:548, 300:748], 95, []),
                         (golden_image, 90, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])]:
        ok, enc = cv2.imencode(".jpg", np.ascontiguousarray(im), [cv2.IMWRITE_JPEG_QUALITY, q] + extra)
        assert ok
        streams.append(enc.tobytes())
    eng = Engine(caffemodel("mnet25"), 448, 448, precision=RF_PREC_FP16, max_batch=8, max_image=golden_image.shape[:2])
    try:
        try:
            faces, sizes = eng.detect_jpeg(streams, 0.9, 0.4)
        except RfError as e:
            if e.status == -7 or "libnvjpeg" in str(e):
                pytest.skip("libnvjpeg not present on this box")
            raise
        print("nvJPEG back end:", eng.jpeg_backend())
        host_imgs = [cv2.imdecode(np.frombuffer(s, np.uint8), cv2.IMREAD_COLOR) for s in streams]
        ref = eng.detect_batch(host_imgs, 0.9, 0.4)
        for i, (s, im) in enumerate(zip(streams, host_imgs)):
            assert sizes[i] == (im.shape[1], im.shape[0]), (i, sizes[i], im.shape)
            dec = eng.decode_jpeg(s)
            assert dec.shape == im.shape
            d = np.abs(dec.astype(np.int32) - im.astype(np.int32))
            print(f"stream {i} ({im.shape[1]}x{im.shape[0]}, {len(s)} bytes): decoded bytes vs cv2.imdecode mean |diff| {d.mean():.3f}, max {d.max()}, "
                  f"faces {len(faces[i])} vs {len(ref[i])}")
            assert d.mean() < 1.5, (i, d.mean())
            pairs, extra_mine, missed = _match(faces[i], ref[i])
            assert extra_mine == 0 and missed == 0, (i, len(faces[i]), len(ref[i]))
            for a, b in pairs:
                assert abs(faces[i][a][0] - ref[i][b][0]) < 0.03 and np.abs(faces[i][a][1:] - ref[i][b][1:]).max() < 2.0
        assert len(faces[0]) >= 4
        # a stream that is not a JPEG is an error, not a crash
        with pytest.raises(RfError):
            eng.detect_jpeg([b"not a jpeg at all" * 10], 0.9, 0.4)
        # and the pixel path still works afterwards
        assert len(eng.detect_batch([host_imgs[0]], 0.9, 0.4)[0]) == len(ref[0])
    finally:
        eng.close()
