#!/usr/bin/env python
"""Bring-up of the tile-chain plan on a B200 (development tool): every RF_TILE_MASK subset in its own process (a trap in
one kernel cannot poison the next case), compared with the round-1 per-layer kernels (RF_FLAG_LEGACY_TC) and the FP32
numpy oracle.  Usage: python tools/tile_bringup.py [--masks 1,2,...] [--hw 448x448] [--batch 3]; child: --child MASK."""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def child(mask, h, w, nb, model):
    os.environ["RF_TILE_MASK"] = str(mask)
    import cv2
    from oracle.inputs import letterbox_bgr_u8, s_noise_batch
    from retinaface_b200 import RF_PREC_FP16, Engine
    from retinaface_b200.capi import RF_FLAG_LEGACY_TC, RfError, plan_describe
    cm = os.path.join(GOLD, "weights", model + ".caffemodel")
    img = cv2.imread(os.path.join(GOLD, "data", "img.jpg"))
    inp = letterbox_bgr_u8(img, h, w)
    imgs = [inp, s_noise_batch(1, h, w, seed=1)[0], np.roll(inp, 40, axis=1)]
    batch = np.stack([imgs[i % 3] if i < 3 else np.roll(inp, 8 * i, axis=1) for i in range(nb)])
    print(plan_describe(cm, h, w, max_batch=nb), flush=True)
    new = Engine(cm, h, w, precision=RF_PREC_FP16, max_batch=nb)
    old = Engine(cm, h, w, precision=RF_PREC_FP16, max_batch=nb, flags=RF_FLAG_LEGACY_TC)
    new.debug_keep_all()
    old.debug_keep_all()
    t0 = time.time()
    hn = new.forward_heads(batch)
    print(f"forward_heads new: {time.time() - t0:.3f}s", flush=True)
    ho = old.forward_heads(batch)
    res = {"mask": mask, "tensors": {}, "heads": []}
    names = ["mobilenet0_relu2_fwd", "mobilenet0_relu6_fwd", "mobilenet0_relu10_fwd", "rf_c1_red_conv_relu", "mobilenet0_relu16_fwd", "mobilenet0_relu22_fwd",
             "rf_c2_lateral_relu", "mobilenet0_relu24_fwd", "mobilenet0_relu26_fwd", "rf_c3_lateral_relu", "rf_c3_det_concat_relu", "rf_c2_aggr_relu",
             "rf_c2_det_concat_relu", "rf_c1_aggr_relu", "rf_c1_det_concat_relu"]
    ok = True
    for name in names:
        try:
            a = new.debug_tensor(name, nb)
            b = old.debug_tensor(name, nb)
        except RfError:
            continue
        scale = float(np.abs(b).max()) or 1.0
        e = float(np.abs(a - b).max() / scale)
        where = np.unravel_index(np.argmax(np.abs(a - b)), a.shape)
        res["tensors"][name] = e
        flag = "" if e < 2e-2 else "   <-- BAD at " + str(tuple(int(x) for x in where))
        print(f"  {name:34s} max|new-old|/max = {e:.5f}{flag}", flush=True)
        ok &= e < 2e-2
    for k in range(9):
        e = float(np.abs(hn[k] - ho[k]).max())
        res["heads"].append(e)
        ok &= e < 3e-2
    print("  heads max abs diff:", " ".join(f"{e:.4f}" for e in res["heads"]), flush=True)
    fn, idn = new.detect_batch(list(batch), 0.9, 0.4, want_index=True)
    fo, ido = old.detect_batch(list(batch), 0.9, 0.4, want_index=True)
    for i in range(nb):
        same = len(fn[i]) == len(fo[i]) and list(idn[i]) == list(ido[i])
        d = float(np.abs(fn[i] - fo[i]).max()) if same and len(fn[i]) else 0.0
        print(f"  image {i}: {len(fn[i])} faces (legacy {len(fo[i])}), same anchors {same}, max diff {d:.4f}", flush=True)
        ok &= same and d < 0.5
    # repeat: graph replay + self-cleaning counters
    fn2 = new.detect_batch(list(batch), 0.9, 0.4)
    for i in range(nb):
        ok &= len(fn2[i]) == len(fn[i]) and (len(fn[i]) == 0 or np.array_equal(fn2[i], fn[i]))
    print("  second run identical:", all(len(fn2[i]) == len(fn[i]) and (len(fn[i]) == 0 or np.array_equal(fn2[i], fn[i])) for i in range(nb)), flush=True)
    prof = new.profile_layers(nb, iters=20)
    for p in prof:
        print(f"  {p['name']:44s} {p['ms'] * 1e3:8.2f} us", flush=True)
    print(f"  sum of kernels {sum(p['ms'] for p in prof) * 1e3:.1f} us; launches {new.launches_per_batch(nb)}", flush=True)
    fn3 = new.detect_batch(list(batch), 0.9, 0.4)
    ok &= all(len(fn3[i]) == len(fn[i]) and (len(fn[i]) == 0 or np.array_equal(fn3[i], fn[i])) for i in range(nb))
    print("RESULT", "PASS" if ok else "FAIL", json.dumps(res), flush=True)
    new.close()
    old.close()
    return 0 if ok else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", type=int, default=None)
    ap.add_argument("--masks", default="511,490,0")
    ap.add_argument("--hw", default="448x448")
    ap.add_argument("--batch", type=int, default=3)
    ap.add_argument("--model", default="mnet25")
    ap.add_argument("--timeout", type=int, default=120)
    args = ap.parse_args()
    h, w = (int(x) for x in args.hw.split("x"))
    if args.child is not None:
        sys.exit(child(args.child, h, w, args.batch, args.model))
    summary = {}
    for m in [int(x) for x in args.masks.split(",")]:
        print(f"===== RF_TILE_MASK={m} ({h}x{w}, batch {args.batch}) =====", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(m), "--hw", args.hw, "--batch", str(args.batch), "--model", args.model],
                               capture_output=True, text=True, timeout=args.timeout)
            out = r.stdout + r.stderr[-3000:]
            rc = r.returncode
        except subprocess.TimeoutExpired as e:
            out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            out += "\nTIMEOUT"
            rc = -9
        print(out, flush=True)
        summary[m] = rc
    print("SUMMARY", summary, flush=True)


if __name__ == "__main__":
    main()
