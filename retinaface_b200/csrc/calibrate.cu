// calibrate.cu -- see calibrate.cuh.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>

#include "calibrate.cuh"

namespace rf {

namespace {

template <typename T>
__global__ void __launch_bounds__(256) k_absmax(const T *__restrict__ x, size_t n, float *__restrict__ out) {
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(to_f(x[i])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned *>(out), __float_as_uint(m));   // non-negative floats order as uints
}

template <typename T>
__global__ void __launch_bounds__(256) k_hist(const T *__restrict__ x, size_t n, float inv_width, unsigned *__restrict__ hist) {
    __shared__ unsigned sh[CALIB_BINS];
    for (int i = threadIdx.x; i < CALIB_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int b = (int)(fabsf(to_f(x[i])) * inv_width);
        atomicAdd(&sh[min(b, CALIB_BINS - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CALIB_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

}  // namespace

template <typename T>
void launch_absmax(const T *x, size_t n, float *out, cudaStream_t s) {
    unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 8);
    k_absmax<T><<<grid ? grid : 1, 256, 0, s>>>(x, n, out);
}
template <typename T>
void launch_hist(const T *x, size_t n, float inv_width, unsigned *hist, cudaStream_t s) {
    unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 4);
    k_hist<T><<<grid ? grid : 1, 256, 0, s>>>(x, n, inv_width, hist);
}
template void launch_absmax<float>(const float *, size_t, float *, cudaStream_t);
template void launch_absmax<__half>(const __half *, size_t, float *, cudaStream_t);
template void launch_hist<float>(const float *, size_t, float, unsigned *, cudaStream_t);
template void launch_hist<__half>(const __half *, size_t, float, unsigned *, cudaStream_t);

// Entropy calibration threshold search, in the formulation NVIDIA publishes with its own open-source calibrator
// (pytorch-quantization, calib/histogram.py `_compute_amax_entropy`, which mirrors TensorRT's entropy calibration):
// the zero-spike bin 0 is replaced by bin 1; for every candidate number of bins i in [levels, bins]: P = the first i
// bins with the outliers folded into bin i-1; Q = the first i bins assigned uniformly to `levels` groups (bin k ->
// group floor(k * levels / i)), each NON-EMPTY bin receiving its group's mean over the non-empty bins; both
// normalised; KL(P || Q) (infinite where Q = 0 < P); the LAST argmin wins.  Returns the threshold in bins.
double kl_threshold_bins(const unsigned *hist_in, int bins, int levels) {
    std::vector<double> h(hist_in, hist_in + bins);
    if (bins > 1) h[0] = h[1];
    double total = 0;
    for (int i = 0; i < bins; i++) total += h[i];
    if (total == 0) return bins;
    std::vector<double> suffix(bins + 1, 0.0);
    for (int i = bins - 1; i >= 0; i--) suffix[i] = suffix[i + 1] + h[i];
    std::vector<double> gsum(levels), q(bins);
    std::vector<int> gcnt(levels);
    double best = 1e300;
    int best_i = bins;
    for (int i = levels; i <= bins; i++) {
        std::fill(gsum.begin(), gsum.end(), 0.0);
        std::fill(gcnt.begin(), gcnt.end(), 0);
        for (int k = 0; k < i; k++) {
            const int g = (int)(((long)k * levels) / i);
            gsum[g] += h[k];
            gcnt[g] += h[k] != 0;
        }
        double qs = 0;
        for (int k = 0; k < i; k++) {
            const int g = (int)(((long)k * levels) / i);
            q[k] = h[k] != 0 ? gsum[g] / gcnt[g] : 0.0;
            qs += q[k];
        }
        if (qs == 0) continue;
        const double ps = suffix[0];             // P sums to the whole histogram (outliers folded in)
        double kl = 0;
        bool inf = false;
        for (int k = 0; k < i && !inf; k++) {
            const double pk = (k == i - 1 ? h[k] + suffix[i] : h[k]) / ps;
            if (pk <= 0) continue;
            const double qk = q[k] / qs;
            if (qk <= 0) { inf = true; break; }
            kl += pk * std::log(pk / qk);
        }
        if (inf) continue;
        if (kl <= best) { best = kl; best_i = i; }
    }
    return best_i;
}

bool write_int8_table(const std::string &path, const std::vector<std::pair<std::string, float>> &scales, std::string &err) {
    std::ofstream f(path);
    if (!f) { err = "cannot write calibration table '" + path + "'"; return false; }
    f << "TRT-5102-EntropyCalibration2\n";
    for (auto &kv : scales) {
        uint32_t bits;
        memcpy(&bits, &kv.second, 4);
        char hex[16];
        snprintf(hex, sizeof hex, "%08x", bits);
        f << kv.first << ": " << hex << "\n";
    }
    return (bool)f;
}

}  // namespace rf
