// preprocess.cu -- see preprocess.cuh.
#include <cmath>

#include "preprocess.cuh"

namespace rf {

void letterbox_geometry(int w, int h, int net_w, int net_h, int *dw, int *dh, double *scale) {
    // RetinaFace.cpp:587-591: float sw = 1.0*cols/inputW, sh = 1.0*rows/inputH; scale = max(sw, sh, 1)
    float sw = (float)(1.0 * w / net_w), sh = (float)(1.0 * h / net_h);
    float sc = sw > sh ? sw : sh;
    sc = sc > 1.0f ? sc : 1.0f;
    if (sc > 1.0f) {
        double f = (double)(1 / sc);              // cv::resize(..., fx = 1/scale (float), fy = same)
        *dw = (int)std::nearbyint(w * f);         // saturate_cast<int>(double): round half to even
        *dh = (int)std::nearbyint(h * f);
        *scale = 1.0 / f;                         // resize.cpp: scale_x = 1. / inv_scale_x
    } else {
        *dw = w; *dh = h; *scale = 1.0;
    }
    if (*dw > net_w) *dw = net_w;                 // copyMakeBorder would assert; clamp instead
    if (*dh > net_h) *dh = net_h;
}

namespace {

struct Tap { int s0, s1, a0, a1; };

// OpenCV resizeGeneric_ linear coefficient of one destination coordinate (resize.cpp, INTER_LINEAR).
// Horizontal taps zero the fraction at the borders (xmin/xmax handling); vertical taps keep the
// fraction and clamp the source ROW indices instead (resizeGeneric_Invoker: sy = clip(sy0 + k, 0, h)),
// which differs by one count after the truncating vertical pass.
template <bool HORIZONTAL>
__device__ __forceinline__ Tap tap_of(int d, int sn, double scale) {
    float fx = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(fx);
    fx -= (float)s;
    if (HORIZONTAL) {
        if (s < 0) { fx = 0.f; s = 0; }
        if (s >= sn - 1) { fx = 0.f; s = sn - 1; }
    }
    Tap t;
    t.s0 = min(max(s, 0), sn - 1);
    t.s1 = min(max(s + 1, 0), sn - 1);
    t.a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, fx), 2048.f));   // saturate_cast<short>(cbuf*INTER_RESIZE_COEF_SCALE)
    t.a1 = __float2int_rn(__fmul_rn(fx, 2048.f));
    return t;
}

__global__ void __launch_bounds__(256) k_letterbox(const uint8_t *__restrict__ src, int sw, int sh,
                                                   uint8_t *__restrict__ dst, int net_w, int net_h, int dw, int dh,
                                                   double scale, int identity, int flip) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= net_w) return;
    uint8_t *o = dst + ((size_t)y * net_w + x) * 3;
    if (x >= dw || y >= dh) { o[0] = 0; o[1] = 0; o[2] = 0; return; }
    // flip: the view is the letter-box of the horizontally mirrored image -- every source column index is mirrored, the
    // taps are those of the mirrored image (== cv::resize(cv::flip(img, 1)) bit for bit)
    if (identity) {
        const uint8_t *p = src + ((size_t)y * sw + (flip ? sw - 1 - x : x)) * 3;
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
        return;
    }
    Tap tx = tap_of<true>(x, sw, scale);
    const Tap ty = tap_of<false>(y, sh, scale);
    if (flip) { tx.s0 = sw - 1 - tx.s0; tx.s1 = sw - 1 - tx.s1; }
    const uint8_t *r0 = src + (size_t)ty.s0 * sw * 3, *r1 = src + (size_t)ty.s1 * sw * 3;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        int h0 = (int)r0[tx.s0 * 3 + c] * tx.a0 + (int)r0[tx.s1 * 3 + c] * tx.a1;   // HResizeLinear
        int h1 = (int)r1[tx.s0 * 3 + c] * tx.a0 + (int)r1[tx.s1 * 3 + c] * tx.a1;
        int v = (((ty.a0 * (h0 >> 4)) >> 16) + ((ty.a1 * (h1 >> 4)) >> 16) + 2) >> 2;  // VResizeLinear 8u
        o[c] = (uint8_t)min(max(v, 0), 255);
    }
}

}  // namespace

void launch_letterbox(const uint8_t *src, int w, int h, uint8_t *dst, int net_w, int net_h, cudaStream_t s) {
    launch_letterbox_view(src, w, h, dst, net_w, net_h, net_w, net_h, 0, s);
}

float launch_letterbox_view(const uint8_t *src, int w, int h, uint8_t *dst, int net_w, int net_h, int box_w, int box_h, int flip,
                            cudaStream_t s) {
    int dw, dh;
    double scale;
    letterbox_geometry(w, h, box_w, box_h, &dw, &dh, &scale);
    dim3 grid((net_w + 255) / 256, net_h);
    k_letterbox<<<grid, 256, 0, s>>>(src, w, h, dst, net_w, net_h, dw, dh, scale, scale == 1.0 ? 1 : 0, flip);
    // the reference's own float `scale` (RetinaFace.cpp:587-591), the factor its map-back multiplies by (:732-738)
    const float sw = (float)(1.0 * w / box_w), sh = (float)(1.0 * h / box_h);
    const float sc = sw > sh ? sw : sh;
    return sc > 1.0f ? sc : 1.0f;
}

}  // namespace rf
