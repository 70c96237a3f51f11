// calibrate.cuh -- INT8 entropy calibrator.
//
// Replaces the reference's offline INT8-Calibration-Tool (INT8-Calibration-Tool/calibrationtable.cpp:399-583),
// whose actual algorithm lives inside closed-source TensorRT (IInt8EntropyCalibrator2): this is the published
// entropy-calibration procedure (NVIDIA, "8-bit inference with TensorRT", GTC 2017) on GPU-built histograms:
//   1. absmax of every activation tensor over the calibration set          (k_absmax)
//   2. 2048-bin histogram of |x| over [0, absmax]                          (k_hist)
//   3. per tensor, the clipping threshold T in {128..2048 bins} minimising KL(P || Q), P = clipped reference
//      distribution (outliers folded into the last bin), Q = P quantised to 128 levels      (host, kl_threshold)
//   4. scale = T / 127, written in the reference's own cache format, "TRT-5102-EntropyCalibration2" + one
//      "<caffe top name>: <big-endian float32 hex>" line per tensor, so that the table
//      is consumable both by rf_create(RF_PREC_INT8) and by the reference's Int8EntropyCalibrator2 reader
//      (retinaface/tensorrt/trtnetbase.cpp:31-44).
#pragma once
#include <string>
#include <vector>

#include "common.cuh"

namespace rf {

constexpr int CALIB_BINS = 2048;

// |x| max of a float / half tensor, accumulated into *out (float bits, atomicMax on non-negative floats)
template <typename T>
void launch_absmax(const T *x, size_t n, float *out, cudaStream_t s);
// histogram of |x| with bin = min(bins-1, floor(|x| * inv_width)), accumulated into hist[CALIB_BINS]
template <typename T>
void launch_hist(const T *x, size_t n, float inv_width, unsigned *hist, cudaStream_t s);

// Threshold (in units of bin width, i.e. the real threshold is result * absmax / CALIB_BINS) minimising the KL
// divergence; returns CALIB_BINS when the histogram is empty.
double kl_threshold_bins(const unsigned *hist, int bins = CALIB_BINS, int levels = 128);

bool write_int8_table(const std::string &path, const std::vector<std::pair<std::string, float>> &scales, std::string &err);

}  // namespace rf
