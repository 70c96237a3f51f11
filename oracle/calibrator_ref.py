"""numpy restatement of the entropy-calibration threshold search used by ``rf_calibrate_int8``.

Test infrastructure -- see ``oracle/__init__.py``.  The reference's calibrator is TensorRT's closed-source
``IInt8EntropyCalibrator2`` (``INT8-Calibration-Tool/calibrationtable.h:142-215`` only feeds it batches), so what
is restated is the published procedure (NVIDIA GTC 2017, "8-bit inference with TensorRT") in the open formulation
also used by MXNet's quantisation tools -- the same formulation ``retinaface_b200/csrc/calibrate.cu`` implements.
"""
import numpy as np


def kl_threshold_bins(hist, levels=128):
    """Threshold (in bins) of the entropy calibration: bin 0 := bin 1; for i in [levels, bins]: P = first i bins with the
    outliers folded into bin i-1, Q = bins assigned uniformly to `levels` groups (k -> floor(k*levels/i)), every non-empty
    bin getting its group's mean over non-empty bins; normalise; KL(P||Q) (skip i when Q = 0 < P); last argmin."""
    bins = np.asarray(hist, dtype=np.float64).copy()
    n = bins.size
    if n > 1:
        bins[0] = bins[1]
    if bins.sum() == 0:
        return float(n)
    best, best_i = np.inf, n
    csum = np.concatenate([[0.0], np.cumsum(bins)])
    for i in range(levels, n + 1):
        b = bins[:i]
        grp = (np.arange(i) * levels) // i
        nz = b != 0
        s = np.bincount(grp, weights=b, minlength=levels)
        c = np.bincount(grp, weights=nz.astype(np.float64), minlength=levels)
        avg = np.where(c > 0, s / np.maximum(c, 1), 0.0)
        q = np.where(nz, avg[grp], 0.0)
        p = b.copy()
        p[i - 1] += csum[n] - csum[i]
        if q.sum() == 0:
            continue
        q = q / q.sum()
        p = p / p.sum()
        m = p > 0
        if np.any(q[m] == 0):
            continue
        kl = float(np.sum(p[m] * np.log(p[m] / q[m])))
        if kl <= best:
            best, best_i = kl, i
    return float(best_i)
