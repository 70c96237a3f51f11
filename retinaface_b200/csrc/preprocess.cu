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

// One launch letter-boxes up to LB_MAX_IMAGES images: blockIdx.z = image, blockIdx.y = output row, a thread = 4 consecutive
// output pixels = 12 bytes = three aligned 32-bit stores (net_w is a multiple of 32, so rows start 4-byte aligned).
struct LbImage {
    const uint8_t *src;      // packed rows, w x h x 3 u8 BGR
    uint8_t *dst;            // net_h x net_w x 3
    int sw, sh, dw, dh;      // source size, size of the resized image inside the output (top-left), rest = 0
    double scale;            // source pixels per output pixel
    int identity, flip, area;
};
struct LbBatch { LbImage img[LB_MAX_IMAGES]; };

// NPP's NPPI_INTER_SUPER as measured against nppiResizeSqrPixel_8u_C3R on a B200 (tools/npp_dump.py, oracle/npp_oracle.cu):
// output pixel (x, y) = the coverage-weighted mean of the source rectangle [x / f, (x + 1) / f) x [y / f, (y + 1) / f), the
// resized extent is ceil(w f) x ceil(h f), source samples beyond the image count as ZERO (the last row / column is darker, not
// renormalised), round half up.  Matches NPP byte for byte on 5 of 8 probe shapes and within 1 LSB on < 0.5 % of the bytes of
// the others (NPP's own arithmetic is single precision).
__device__ __forceinline__ void area_pixel(const LbImage &im, int x, int y, int out[3]) {
    const double inv = im.scale;
    const double ax = x * inv, bx = (x + 1) * inv, ay = y * inv, by = (y + 1) * inv;
    const int x0 = (int)floor(ax), x1 = min((int)ceil(bx - 1e-12), im.sw), y0 = (int)floor(ay), y1 = min((int)ceil(by - 1e-12), im.sh);
    double acc[3] = {0.0, 0.0, 0.0};
    for (int sy = y0; sy < y1; sy++) {
        const double wy = fmin((double)(sy + 1), by) - fmax((double)sy, ay);
        const uint8_t *row = im.src + (size_t)sy * im.sw * 3;
        for (int sx = x0; sx < x1; sx++) {
            const double w = wy * (fmin((double)(sx + 1), bx) - fmax((double)sx, ax));
            const uint8_t *p = row + (im.flip ? im.sw - 1 - sx : sx) * 3;
            acc[0] += w * p[0]; acc[1] += w * p[1]; acc[2] += w * p[2];
        }
    }
    const double norm = 1.0 / (inv * inv);
#pragma unroll
    for (int c = 0; c < 3; c++) out[c] = min(max((int)floor(acc[c] * norm + 0.5), 0), 255);
}

__device__ __forceinline__ void linear_pixel(const LbImage &im, int x, const Tap &ty, int out[3]) {
    Tap tx = tap_of<true>(x, im.sw, im.scale);
    if (im.flip) { tx.s0 = im.sw - 1 - tx.s0; tx.s1 = im.sw - 1 - tx.s1; }
    const uint8_t *r0 = im.src + (size_t)ty.s0 * im.sw * 3, *r1 = im.src + (size_t)ty.s1 * im.sw * 3;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        int h0 = (int)r0[tx.s0 * 3 + c] * tx.a0 + (int)r0[tx.s1 * 3 + c] * tx.a1;   // HResizeLinear
        int h1 = (int)r1[tx.s0 * 3 + c] * tx.a0 + (int)r1[tx.s1 * 3 + c] * tx.a1;
        int v = (((ty.a0 * (h0 >> 4)) >> 16) + ((ty.a1 * (h1 >> 4)) >> 16) + 2) >> 2;  // VResizeLinear 8u
        out[c] = min(max(v, 0), 255);
    }
}

__global__ void __launch_bounds__(128) k_letterbox_batch(const __grid_constant__ LbBatch B, int net_w, int net_h) {
    const LbImage &im = B.img[blockIdx.z];
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    if (x4 >= net_w) return;
    unsigned char px[12];
    const bool row_in = y < im.dh;
    Tap ty{};
    if (row_in && !im.identity && !im.area) ty = tap_of<false>(y, im.sh, im.scale);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = x4 + k;
        int v[3] = {0, 0, 0};
        if (row_in && x < im.dw) {
            // flip: the view is the letter-box of the horizontally mirrored image -- every source column index is mirrored, the
            // taps are those of the mirrored image (== cv::resize(cv::flip(img, 1)) bit for bit)
            if (im.identity) {
                const uint8_t *p = im.src + ((size_t)y * im.sw + (im.flip ? im.sw - 1 - x : x)) * 3;
                v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
            } else if (im.area) {
                area_pixel(im, x, y, v);
            } else {
                linear_pixel(im, x, ty, v);
            }
        }
        px[3 * k] = (unsigned char)v[0]; px[3 * k + 1] = (unsigned char)v[1]; px[3 * k + 2] = (unsigned char)v[2];
    }
    uint32_t *o = reinterpret_cast<uint32_t *>(im.dst + ((size_t)y * net_w + x4) * 3);
    o[0] = px[0] | (px[1] << 8) | (px[2] << 16) | ((uint32_t)px[3] << 24);
    o[1] = px[4] | (px[5] << 8) | (px[6] << 16) | ((uint32_t)px[7] << 24);
    o[2] = px[8] | (px[9] << 8) | (px[10] << 16) | ((uint32_t)px[11] << 24);
}

}  // namespace

void letterbox_geometry_npp(int w, int h, int net_w, int net_h, int *dw, int *dh, double *scale) {
    // resizeconvertion.cu:298-303: factor = min(dstW / srcW, dstH / srcH), clamped to <= 1 (never up-scaled)
    const double fx = (double)net_w / w, fy = (double)net_h / h;
    double f = fx < fy ? fx : fy;
    if (f >= 1.0) { *dw = std::min(w, net_w); *dh = std::min(h, net_h); *scale = 1.0; return; }
    *dw = std::min((int)std::ceil(w * f - 1e-9), net_w);
    *dh = std::min((int)std::ceil(h * f - 1e-9), net_h);
    *scale = 1.0 / f;
}

float letterbox_fill(LbItem &it, const uint8_t *src, int w, int h, uint8_t *dst, int box_w, int box_h, int flip, int area) {
    it.src = src; it.dst = dst; it.sw = w; it.sh = h; it.flip = flip; it.area = area;
    if (area) letterbox_geometry_npp(w, h, box_w, box_h, &it.dw, &it.dh, &it.scale);
    else letterbox_geometry(w, h, box_w, box_h, &it.dw, &it.dh, &it.scale);
    it.identity = it.scale == 1.0 ? 1 : 0;
    if (it.identity) it.area = 0;
    // the reference's own float `scale` (RetinaFace.cpp:587-591), the factor its map-back multiplies by (:732-738)
    const float sw = (float)(1.0 * w / box_w), sh = (float)(1.0 * h / box_h);
    const float sc = sw > sh ? sw : sh;
    return sc > 1.0f ? sc : 1.0f;
}

cudaError_t launch_letterbox_batch(const LbItem *items, int n, int net_w, int net_h, cudaStream_t s) {
    for (int i0 = 0; i0 < n; i0 += LB_MAX_IMAGES) {
        const int m = std::min(LB_MAX_IMAGES, n - i0);
        LbBatch B{};
        for (int i = 0; i < m; i++) {
            const LbItem &it = items[i0 + i];
            B.img[i] = LbImage{it.src, it.dst, it.sw, it.sh, it.dw, it.dh, it.scale, it.identity, it.flip, it.area};
        }
        dim3 grid((net_w / 4 + 127) / 128, net_h, m);
        k_letterbox_batch<<<grid, 128, 0, s>>>(B, net_w, net_h);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

void launch_letterbox(const uint8_t *src, int w, int h, uint8_t *dst, int net_w, int net_h, cudaStream_t s) {
    launch_letterbox_view(src, w, h, dst, net_w, net_h, net_w, net_h, 0, s);
}

float launch_letterbox_view(const uint8_t *src, int w, int h, uint8_t *dst, int net_w, int net_h, int box_w, int box_h, int flip,
                            cudaStream_t s) {
    LbItem it;
    const float sc = letterbox_fill(it, src, w, h, dst, box_w, box_h, flip, 0);
    launch_letterbox_batch(&it, 1, net_w, net_h, s);
    return sc;
}

}  // namespace rf
