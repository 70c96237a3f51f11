// kernels_simt.cuh -- SIMT (CUDA-core) layer kernels of librf_b200: the FP32 tight-parity path
// and the FP16 fallback for layers that have no tcgen05 kernel.  NHWC activations, FP32
// accumulate, folded-BN bias + optional ReLU fused into every epilogue.
//
// What each kernel replaces in the reference: the corresponding layers that TensorRT executes
// inside context->enqueue (retinaface/tensorrt/trtretinafacenet.cpp:60) for the graph of
// model/mnet-deconv-0517.prototxt (layer line numbers given per kernel).
#pragma once
#include "common.cuh"

namespace rf {

// ------------------------------------------------------------------------------------------
// conv0: mobilenet0_conv0 (prototxt:11) 3x3 s2 p1, 3 -> 8, + BN + ReLU, consuming the u8 BGR
// HWC image directly: the BGR->RGB swap and u8->float of convertBGR2RGBfloat / imageSplit
// (resizeconvertion.cu:46-63,165-185) are folded into the weight order / the load.
//   wk: [27][8] floats, row = (ky*3+kx)*3 + c_bgr ; bias [8].
// One thread per output pixel (8 channels = one 16/32-byte store).
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_conv0(const PostParams *__restrict__ run, T *__restrict__ out,
                                               const float *__restrict__ wk, const float *__restrict__ bias,
                                               int n, int H, int W) {
    __shared__ float sw[27 * 8 + 8];
    pdl_trigger();
    for (int i = threadIdx.x; i < 27 * 8 + 8; i += blockDim.x) sw[i] = i < 216 ? wk[i] : bias[i - 216];
    __syncthreads();
    pdl_wait();
    const int OH = H >> 1, OW = W >> 1;
    const long total = (long)n * OH * OW;
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int ox = (int)(idx % OW);
    int oy = (int)((idx / OW) % OH);
    int b = (int)(idx / ((long)OW * OH));
    const uint8_t *__restrict__ base = run->input + (size_t)b * H * W * 3;
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; o++) acc[o] = sw[216 + o];
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
        int iy = oy * 2 + ky - 1;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
            int ix = ox * 2 + kx - 1;
            if (ix < 0 || ix >= W) continue;
            const uint8_t *p = base + ((size_t)iy * W + ix) * 3;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                float v = (float)p[c];
                const float *wr = &sw[((ky * 3 + kx) * 3 + c) * 8];
#pragma unroll
                for (int o = 0; o < 8; o++) acc[o] = fmaf(v, wr[o], acc[o]);
            }
        }
    }
#pragma unroll
    for (int o = 0; o < 8; o++) acc[o] = fmaxf(acc[o], 0.f);
    Vec8<T> v;
    v.from_float(acc);
    v.store(out + (size_t)idx * 8);
}

// ------------------------------------------------------------------------------------------
// Depthwise 3x3 (+BN+ReLU): mobilenet0_conv{1,3,...,25} (prototxt:55 ...).  One thread per
// (output pixel, 8-channel group); channels are the contiguous NHWC axis so a warp reads
// 32 x 16 B (FP16) of consecutive channels/pixels.   wd: [9][C] floats, bias [C].
// ------------------------------------------------------------------------------------------
template <typename T, int STRIDE>
__global__ void __launch_bounds__(256) k_dw3x3(const T *__restrict__ in, T *__restrict__ out,
                                               const float *__restrict__ wd, const float *__restrict__ bias,
                                               int n, int H, int W, int C) {
    pdl_trigger();
    pdl_wait();
    const int OH = H / STRIDE, OW = W / STRIDE;
    const int cg = C >> 3;
    const long total = (long)n * OH * OW * cg;
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int g = (int)(idx % cg);
    long pix = idx / cg;
    int ox = (int)(pix % OW);
    int oy = (int)((pix / OW) % OH);
    int b = (int)(pix / ((long)OW * OH));
    const int c0 = g * 8;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = __ldg(bias + c0 + i);
    const T *base = in + (size_t)b * H * W * C + c0;
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
        int iy = oy * STRIDE + ky - 1;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
            int ix = ox * STRIDE + kx - 1;
            if (ix < 0 || ix >= W) continue;
            Vec8<T> v;
            v.load(base + ((size_t)iy * W + ix) * C);
            float f[8];
            v.to_float(f);
            const float *wr = wd + (ky * 3 + kx) * C + c0;
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = fmaf(f[i], __ldg(wr + i), acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = fmaxf(acc[i], 0.f);
    Vec8<T> o;
    o.from_float(acc);
    o.store(out + (size_t)pix * C + c0);
}

// ------------------------------------------------------------------------------------------
// Implicit-GEMM convolution (1x1 or 3x3, stride 1, "same" padding) on CUDA cores:
//   D[m][n] = sum_k A[m][k] * Wk[k][n] + bias[n],  m = output pixel, k = (tap, cin), n = cout.
// Used for every pointwise conv (mobilenet0_conv{2,4,...,26}, rf_*_lateral / red_conv) and every
// full 3x3 conv (rf_c*_aggr, the SSH det/context convs) in FP32 mode.
// The epilogue can split the output channels over two destinations with their own pixel
// stride / ReLU flag -- that is how the SSH branches write straight into the concat buffer
// (prototxt:1418 Concat + :1427 ReLU) and how det_conv1 + context_conv1 share one launch.
// Tile: 64 pixels x BN channels per CTA (256 threads: 16 pixel-threads x 16 channel-threads,
// 4 x TN outputs each), K chunk 16.
// ------------------------------------------------------------------------------------------
template <typename T>
struct OutSplit {
    T *p0; int ld0; int n0; int relu0;     // channels [0, n0)   -> p0[m*ld0 + n]
    T *p1; int ld1; int relu1;             // channels [n0, N)   -> p1[m*ld1 + n - n0]
};

template <typename T, int BN, int KS>
__global__ void __launch_bounds__(256) k_conv_gemm(const T *__restrict__ in, int ldin, int Cin,
                                                   const float *__restrict__ wk, const float *__restrict__ bias,
                                                   int N, OutSplit<T> outs, int nimg, int H, int W) {
    constexpr int BM = 64, BK = 16, TN = BN / 16;
    pdl_trigger();
    pdl_wait();
    __shared__ float As[BK][BM + 2];
    __shared__ float Bs[BK][BN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long M = (long)nimg * H * W;
    const long m0 = (long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    // A-load role: 4 threads per pixel, 4 consecutive channels each
    const int lm = tid >> 2, lk = (tid & 3) * 4;
    const long gm = m0 + lm;
    int px = 0, py = 0, pb = 0;
    const bool mvalid = gm < M;
    if (mvalid) { px = (int)(gm % W); py = (int)((gm / W) % H); pb = (int)(gm / ((long)W * H)); }

    float acc[4][TN];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

    const int kc = Cin < BK ? Cin : BK;  // channels per chunk (8 for the Cin=8 layer, else 16)
    for (int tap = 0; tap < KS * KS; tap++) {
        const int dy = KS == 3 ? tap / 3 - 1 : 0, dx = KS == 3 ? tap % 3 - 1 : 0;
        const int iy = py + dy, ix = px + dx;
        const bool inb = mvalid && iy >= 0 && iy < H && ix >= 0 && ix < W;
        const T *src = in + ((size_t)((size_t)pb * H + iy) * W + ix) * ldin;
        for (int c0 = 0; c0 < Cin; c0 += kc) {
            // A tile: As[k][m]
            if (lk < kc) {
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (inb) {
#pragma unroll
                    for (int i = 0; i < 4; i++) v[i] = to_f(src[c0 + lk + i]);
                }
#pragma unroll
                for (int i = 0; i < 4; i++) As[lk + i][lm] = v[i];
            }
            // B tile: Bs[k][n]
            for (int i = tid; i < kc * BN; i += 256) {
                int k = i / BN, nn = i % BN;
                Bs[k][nn] = (n0 + nn < N) ? wk[(size_t)(tap * Cin + c0 + k) * N + n0 + nn] : 0.f;
            }
            __syncthreads();
#pragma unroll 4
            for (int k = 0; k < kc; k++) {
                float a[4], bv[TN];
#pragma unroll
                for (int i = 0; i < 4; i++) a[i] = As[k][ty + 16 * i];
#pragma unroll
                for (int j = 0; j < TN; j++) bv[j] = Bs[k][tx * TN + j];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
    // epilogue
#pragma unroll
    for (int i = 0; i < 4; i++) {
        long m = m0 + ty + 16 * i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; j++) {
            int nn = n0 + tx * TN + j;
            if (nn >= N) continue;
            float v = acc[i][j] + __ldg(bias + nn);
            if (nn < outs.n0) {
                if (outs.relu0) v = fmaxf(v, 0.f);
                outs.p0[(size_t)m * outs.ld0 + nn] = from_f<T>(v);
            } else {
                if (outs.relu1) v = fmaxf(v, 0.f);
                outs.p1[(size_t)m * outs.ld1 + (nn - outs.n0)] = from_f<T>(v);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// FPN merge: out = lateral + crop(deconv_k4s2p1_depthwise(up))   (prototxt:1553-1592, 1948-1987)
// Caffe Deconvolution: out[y] = sum_i in[i] * w[y - 2i + 1], out-of-range inputs contribute 0
// (NOT a plain bilinear resize -- borders are attenuated).  The crop to
// the lateral's H x W is the loop bound.  uw: [C][16] floats (the caffemodel's own kernel).
// One thread per (output pixel, 8-channel group).
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_upsample_add(const T *__restrict__ lateral, const T *__restrict__ up,
                                                      T *__restrict__ out, const float *__restrict__ uw,
                                                      int n, int H, int W, int C, int UH, int UW) {
    pdl_trigger();
    pdl_wait();
    const int cg = C >> 3;
    const long total = (long)n * H * W * cg;
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int g = (int)(idx % cg);
    long pix = idx / cg;
    int x = (int)(pix % W);
    int y = (int)((pix / W) % H);
    int b = (int)(pix / ((long)W * H));
    const int c0 = g * 8;
    Vec8<T> lv;
    lv.load(lateral + (size_t)pix * C + c0);
    float acc[8];
    lv.to_float(acc);
    // contributing input rows: ky = y - 2i + 1 in [0,3]
    const int i_hi = (y + 1) >> 1, j_hi = (x + 1) >> 1;
#pragma unroll
    for (int di = 0; di < 2; di++) {
        int i = i_hi - di;
        int ky = y - 2 * i + 1;
        if (i < 0 || i >= UH || ky < 0 || ky > 3) continue;
#pragma unroll
        for (int dj = 0; dj < 2; dj++) {
            int j = j_hi - dj;
            int kx = x - 2 * j + 1;
            if (j < 0 || j >= UW || kx < 0 || kx > 3) continue;
            Vec8<T> uv;
            uv.load(up + (((size_t)b * UH + i) * UW + j) * C + c0);
            float f[8];
            uv.to_float(f);
#pragma unroll
            for (int c = 0; c < 8; c++) acc[c] = fmaf(f[c], __ldg(uw + (c0 + c) * 16 + ky * 4 + kx), acc[c]);
        }
    }
    Vec8<T> o;
    o.from_float(acc);
    o.store(out + (size_t)pix * C + c0);
}

}  // namespace rf

namespace rf {

// ------------------------------------------------------------------------------------------
// Stem: mobilenet0_conv0 (3x3 s2, 3->8) + BN + ReLU -> conv1 (depthwise 3x3) + BN + ReLU -> conv2
// (pointwise 8->16) + BN + ReLU   (prototxt:11-141) in ONE kernel, u8 BGR image in, FP16 NHWC
// [n][H/2][W/2][16] out.  With 3/8/16 channels these layers are pure bandwidth/latency work (a 128x16x16
// tensor-core tile would be >90% padding), so they run on CUDA cores from shared memory:
//   per CTA: a 16x16 tile of the H/2 x W/2 map; the 37x37x3 u8 input patch is staged once, conv0 is
//   evaluated on the 18x18 halo ring into shared memory (zero outside the map = conv1's padding), then
//   each thread finishes one output pixel (depthwise + pointwise) in registers.
// Intermediate activations stay in FP32 (they never leave the SM), only the result is rounded to FP16.
//   w0: [27][8] (row = (ky*3+kx)*3 + c_bgr), b0[8]; wd: [9][8], bd[8]; wp: [8][16] (k-major), bp[16].
// ------------------------------------------------------------------------------------------
struct StemWeights { const float *w0, *b0, *wd, *bd, *wp, *bp; };

// OutT = __half (FP16 path) or int8_t (INT8 path: the result is quantised with out_inv_scale = 1/s(relu2)).
template <typename OutT>
__global__ void __launch_bounds__(256, 3) k_stem(const PostParams *__restrict__ run, OutT *__restrict__ out, StemWeights sw,
                                              int n, int H, int W, float out_inv_scale) {
    __shared__ __align__(4) uint8_t s_in[37][116];   // input patch rows: `mis` alignment bytes + 37 px * 3 B, as 29 words
    __shared__ __align__(16) float s_c0[2][18 * 18][4];   // [channels 0-3 | 4-7][ring position]: 16-byte stride, conflict-free LDS.128
    __shared__ __align__(16) float s_w0[27 * 8 + 8];
    __shared__ __align__(16) float s_wd[9 * 8 + 8];
    __shared__ __align__(16) float s_wp[8 * 16 + 16];
    const int tid = threadIdx.x;
    const int OH = H >> 1, OW = W >> 1;
    const int tiles_x = (OW + 15) >> 4, tiles_y = (OH + 15) >> 4;
    const int b = blockIdx.x / (tiles_x * tiles_y);
    const int trem = blockIdx.x - b * tiles_x * tiles_y;
    const int oy0 = (trem / tiles_x) << 4, ox0 = (trem % tiles_x) << 4;
    pdl_trigger();
    for (int i = tid; i < 27 * 8 + 8; i += 256) s_w0[i] = i < 216 ? sw.w0[i] : sw.b0[i - 216];
    for (int i = tid; i < 9 * 8 + 8; i += 256) s_wd[i] = i < 72 ? sw.wd[i] : sw.bd[i - 72];
    for (int i = tid; i < 8 * 16 + 16; i += 256) s_wp[i] = i < 128 ? sw.wp[i] : sw.bp[i - 128];
    pdl_wait();
    // ---- stage the u8 patch: input rows 2*oy0-3 .. 2*oy0+33, columns 2*ox0-3 .. 2*ox0+33 -------------------
    // 32-bit loads: the patch row starts at row byte cb0 = 3*ix0; words are taken from the 4-byte aligned
    // address below it (`mis` bytes of slack), so a row is 29 word loads instead of 111 byte loads.  W is a
    // multiple of 32, so every image row has the same alignment.  Words that straddle the image border are
    // assembled bytewise.
    const uint8_t *__restrict__ img = run->input + (size_t)b * H * W * 3;
    const int iy0 = 2 * oy0 - 3, cb0 = (2 * ox0 - 3) * 3;
    const int al0 = cb0 & ~3, mis = cb0 - al0;           // floor to a multiple of 4 (two's complement: also for cb0 < 0)
    const bool aligned = ((reinterpret_cast<uintptr_t>(img) & 3) == 0);
    const int rowbytes = W * 3;
    for (int i = tid; i < 37 * 29; i += 256) {
        const int r = i / 29, w = i - r * 29;
        const int iy = iy0 + r, gb = al0 + 4 * w;        // first row byte of this word
        uint32_t v = 0;
        if (iy >= 0 && iy < H) {
            const uint8_t *rowp = img + (size_t)iy * rowbytes;
            if (aligned && gb >= 0 && gb + 3 < rowbytes) {
                v = *reinterpret_cast<const uint32_t *>(rowp + gb);
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (gb + k >= 0 && gb + k < rowbytes) v |= (uint32_t)rowp[gb + k] << (8 * k);
            }
        }
        *reinterpret_cast<uint32_t *>(&s_in[r][4 * w]) = v;
    }
    __syncthreads();
    // ---- conv0 on the 18x18 ring (conv0 coordinates oy0-1 .. oy0+16) ---------------------------------------
    for (int p = tid; p < 18 * 18; p += 256) {
        const int py = p / 18, px = p - py * 18;
        const int cy = oy0 - 1 + py, cx = ox0 - 1 + px;
        float acc[8];
        if (cy < 0 || cy >= OH || cx < 0 || cx >= OW) {
#pragma unroll
            for (int o = 0; o < 8; o++) acc[o] = 0.f;            // zero padding seen by the depthwise conv
        } else {
#pragma unroll
            for (int o = 0; o < 8; o++) acc[o] = s_w0[216 + o];
            // input row of tap ky: 2*cy + ky - 1 -> patch row 2*py + ky ; column likewise
#pragma unroll 1
            for (int ky = 0; ky < 3; ky++) {
                const uint8_t *row = &s_in[2 * py + ky][mis + (2 * px) * 3];
#pragma unroll
                for (int j = 0; j < 9; j++) {                    // j = kx*3 + c_bgr: 9 consecutive bytes
                    const float v = (float)row[j];
                    const float4 wa = *reinterpret_cast<const float4 *>(&s_w0[(ky * 9 + j) * 8]);
                    const float4 wb = *reinterpret_cast<const float4 *>(&s_w0[(ky * 9 + j) * 8 + 4]);
                    acc[0] = fmaf(v, wa.x, acc[0]); acc[1] = fmaf(v, wa.y, acc[1]); acc[2] = fmaf(v, wa.z, acc[2]); acc[3] = fmaf(v, wa.w, acc[3]);
                    acc[4] = fmaf(v, wb.x, acc[4]); acc[5] = fmaf(v, wb.y, acc[5]); acc[6] = fmaf(v, wb.z, acc[6]); acc[7] = fmaf(v, wb.w, acc[7]);
                }
            }
#pragma unroll
            for (int o = 0; o < 8; o++) acc[o] = fmaxf(acc[o], 0.f);
        }
        *reinterpret_cast<float4 *>(&s_c0[0][p][0]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4 *>(&s_c0[1][p][0]) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    __syncthreads();
    // ---- depthwise 3x3 + pointwise 8->16 for this thread's pixel ------------------------------------------
    const int ty = tid >> 4, tx = tid & 15;
    const int oy = oy0 + ty, ox = ox0 + tx;
    if (oy >= OH || ox >= OW) return;
    float d[8];
#pragma unroll
    for (int c = 0; c < 8; c++) d[c] = s_wd[72 + c];
#pragma unroll 1
    for (int ky = 0; ky < 3; ky++)
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
            const int pos = (ty + ky) * 18 + tx + kx;
            const float4 a0 = *reinterpret_cast<const float4 *>(&s_c0[0][pos][0]), a1 = *reinterpret_cast<const float4 *>(&s_c0[1][pos][0]);
            const float4 w0 = *reinterpret_cast<const float4 *>(&s_wd[(ky * 3 + kx) * 8]);
            const float4 w1 = *reinterpret_cast<const float4 *>(&s_wd[(ky * 3 + kx) * 8 + 4]);
            d[0] = fmaf(a0.x, w0.x, d[0]); d[1] = fmaf(a0.y, w0.y, d[1]); d[2] = fmaf(a0.z, w0.z, d[2]); d[3] = fmaf(a0.w, w0.w, d[3]);
            d[4] = fmaf(a1.x, w1.x, d[4]); d[5] = fmaf(a1.y, w1.y, d[5]); d[6] = fmaf(a1.z, w1.z, d[6]); d[7] = fmaf(a1.w, w1.w, d[7]);
        }
    float o[16];
#pragma unroll
    for (int j = 0; j < 16; j++) o[j] = s_wp[128 + j];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const float v = fmaxf(d[c], 0.f);
#pragma unroll
        for (int j4 = 0; j4 < 4; j4++) {
            const float4 w = *reinterpret_cast<const float4 *>(&s_wp[c * 16 + j4 * 4]);
            o[j4 * 4 + 0] = fmaf(v, w.x, o[j4 * 4 + 0]); o[j4 * 4 + 1] = fmaf(v, w.y, o[j4 * 4 + 1]);
            o[j4 * 4 + 2] = fmaf(v, w.z, o[j4 * 4 + 2]); o[j4 * 4 + 3] = fmaf(v, w.w, o[j4 * 4 + 3]);
        }
    }
#pragma unroll
    for (int j = 0; j < 16; j++) o[j] = fmaxf(o[j], 0.f);
    OutT *dst = out + (((size_t)b * OH + oy) * OW + ox) * 16;
    if constexpr (sizeof(OutT) == 2) {
        Vec8<__half> v0, v1;
        v0.from_float(o);
        v1.from_float(o + 8);
        v0.store(reinterpret_cast<__half *>(dst));
        v1.store(reinterpret_cast<__half *>(dst) + 8);
    } else {
        uint32_t pk[4];
#pragma unroll
        for (int j4 = 0; j4 < 4; j4++) {
            uint32_t w = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                int q = __float2int_rn(__fmul_rn(o[j4 * 4 + k], out_inv_scale));
                q = max(-127, min(127, q));
                w |= (uint32_t)(q & 0xff) << (8 * k);
            }
            pk[j4] = w;
        }
        *reinterpret_cast<uint4 *>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

}  // namespace rf

namespace rf {

// ------------------------------------------------------------------------------------------
// FPN merge, FP16 fast path: out = lateral + crop(deconv_k4s2p1_depthwise(up)) (prototxt:1948-1987) with packed
// HFMA2 arithmetic (the deconvolution weights 1/16, 3/16, 9/16 are exact in FP16; the sum is stored as FP16
// anyway).  One thread per (pixel, 8 channels); the lateral vector and the (up to) four coarse vectors are
// loaded up front so all five global loads of a thread are in flight together.  Used where fusing the merge
// into the consumer's staging (tc_conv.cuh UPADD) would push that kernel beyond one wave of CTAs.
//   uwh: [16 taps][C] FP16 weights.
// ------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) k_fpn_merge_h2(const __half *__restrict__ lateral, const __half *__restrict__ up,
                                                      __half *__restrict__ out, const __half *__restrict__ uwh, int n, int H, int W, int C) {
    pdl_trigger();
    // grid = (row pieces, H, images): no division (the 64-bit ones of the first version were most of this kernel's instructions);
    // C / 8 is a power of two (host-checked)
    const int cg = C >> 3, lg = 31 - __clz(cg), UH = H >> 1, UW = W >> 1;
    const int t_row = blockIdx.x * blockDim.x + threadIdx.x;
    if (t_row >= W * cg) return;
    const int g = t_row & (cg - 1), x = t_row >> lg, y = blockIdx.y, b = blockIdx.z;
    const long pix = ((long)b * H + y) * W + x;
    (void)n;
    const int c0 = g * 8;
    const int i_hi = (y + 1) >> 1, j_hi = (x + 1) >> 1;
    int ci[2], cj[2], ky[2], kx[2];
    bool vi[2], vj[2];
#pragma unroll
    for (int d = 0; d < 2; d++) {
        ci[d] = i_hi - d; ky[d] = y - 2 * ci[d] + 1; vi[d] = ci[d] >= 0 && ci[d] < UH;
        cj[d] = j_hi - d; kx[d] = x - 2 * cj[d] + 1; vj[d] = cj[d] >= 0 && cj[d] < UW;
    }
    uint4 wv[4];
#pragma unroll
    for (int t = 0; t < 4; t++) wv[t] = __ldg(reinterpret_cast<const uint4 *>(uwh + (ky[t >> 1] * 4 + kx[t & 1]) * C + c0));   // weights: not produced by the previous kernel
    pdl_wait();
    uint4 accv = *reinterpret_cast<const uint4 *>(lateral + (size_t)pix * C + c0);
    uint4 uv[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const bool ok = vi[t >> 1] && vj[t & 1];
        uv[t] = ok ? *reinterpret_cast<const uint4 *>(up + (((size_t)b * UH + ci[t >> 1]) * UW + cj[t & 1]) * C + c0) : make_uint4(0, 0, 0, 0);
    }
    __half2 *acc = reinterpret_cast<__half2 *>(&accv);
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const __half2 *u2 = reinterpret_cast<const __half2 *>(&uv[t]), *w2 = reinterpret_cast<const __half2 *>(&wv[t]);
#pragma unroll
        for (int c = 0; c < 4; c++) acc[c] = __hfma2(u2[c], w2[c], acc[c]);
    }
    *reinterpret_cast<uint4 *>(out + (size_t)pix * C + c0) = accv;
}

}  // namespace rf

namespace rf {

// FPN merge, INT8 path (c1 level): out_q = rint(q_lat * (s_lat/s_out) + sum_taps q_up * (w * s_up/s_out)), one thread per
// (pixel, 16 channels).  wq: [16 taps][C] floats.
static __global__ void __launch_bounds__(256) k_fpn_merge_i8(const int8_t *__restrict__ lateral, const int8_t *__restrict__ up, int8_t *__restrict__ out,
                                                      const float *__restrict__ wq, float lat_mul, int n, int H, int W, int C) {
    pdl_trigger();
    const int cg = C >> 4, lg = 31 - __clz(cg), UH = H >> 1, UW = W >> 1;     // grid = (row pieces, H, images); C / 16 a power of two
    const int t_row = blockIdx.x * blockDim.x + threadIdx.x;
    if (t_row >= W * cg) return;
    const int g = t_row & (cg - 1), x = t_row >> lg, y = blockIdx.y, b = blockIdx.z;
    const long pix = ((long)b * H + y) * W + x;
    (void)n;
    const int c0 = g * 16;
    const int i_hi = (y + 1) >> 1, j_hi = (x + 1) >> 1;
    pdl_wait();
    auto unpack = [](const uint4 &v, float f[16]) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            f[4 * i + 0] = (float)(int8_t)(w[i] & 0xff); f[4 * i + 1] = (float)(int8_t)((w[i] >> 8) & 0xff);
            f[4 * i + 2] = (float)(int8_t)((w[i] >> 16) & 0xff); f[4 * i + 3] = (float)(int8_t)(w[i] >> 24);
        }
    };
    float acc[16];
    unpack(*reinterpret_cast<const uint4 *>(lateral + (size_t)pix * C + c0), acc);
#pragma unroll
    for (int c = 0; c < 16; c++) acc[c] = __fmul_rn(acc[c], lat_mul);
#pragma unroll
    for (int di = 0; di < 2; di++) {
        const int i = i_hi - di, ky = y - 2 * i + 1;
        if (i < 0 || i >= UH) continue;
#pragma unroll
        for (int dj = 0; dj < 2; dj++) {
            const int j = j_hi - dj, kx = x - 2 * j + 1;
            if (j < 0 || j >= UW) continue;
            float u[16];
            unpack(*reinterpret_cast<const uint4 *>(up + (((size_t)b * UH + i) * UW + j) * C + c0), u);
            const float *w = wq + (ky * 4 + kx) * C + c0;
#pragma unroll
            for (int c = 0; c < 16; c++) acc[c] = __fadd_rn(acc[c], __fmul_rn(u[c], __ldg(w + c)));   // no FMA: bit-identical to the integer oracle
        }
    }
    uint32_t pk[4];
#pragma unroll
    for (int j4 = 0; j4 < 4; j4++) {
        uint32_t w = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int q = max(-127, min(127, __float2int_rn(acc[j4 * 4 + k])));
            w |= (uint32_t)(q & 0xff) << (8 * k);
        }
        pk[j4] = w;
    }
    *reinterpret_cast<uint4 *>(out + (size_t)pix * C + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

}  // namespace rf
