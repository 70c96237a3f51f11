// tc_dwpw2d.cuh -- fused depthwise 3x3 (stride 1|2) + BN + ReLU -> pointwise 1x1 + BN + ReLU for the LARGE feature maps
// (112x112, 56x56 at 448x448 input: conv3..conv10, prototxt:143-488), 2-D tiles.
//
// Same arithmetic, rounding points and GEMM as k_tc_dwpw_staged (tc_conv.cuh): the depthwise stencil runs on CUDA cores
// from staged shared memory (FP32 accumulate, FP16 round) straight into the UMMA A operand, the pointwise GEMM on tcgen05
// with the accumulator in TMEM.  What differs is the tile: TH x TW output pixels of ONE image (TH*TW <= 128) instead of
// 128 consecutive pixels of the linearised map.  On a wide map the 1-D tile stages 128 + 2*(W+3) input positions for 128
// outputs (2.8x at W = 112); the 2-D tile stages ((TH-1)*S+3) x ((TW-1)*S+3) (1.4x), needs no position table (a staged
// row is a contiguous run of the NHWC input: the cp.async addresses are affine in the lane), and vertically adjacent
// outputs sit in one thread: at stride 1 a thread computes two output rows of one column from 4 input rows (12 loads +
// conversions instead of 18, one set of weight reads).
#pragma once
#include "tc_conv.cuh"

namespace rf {

struct TcDw2dArgs {
    const __half *in;       // NHWC dense [nimg][IH][IW][C], C in {16, 32, 64}
    int C, nimg, IH, IW, OH, OW, S;
    int N;                  // output channels (multiple of 16, <= 256)
    int TH, TW;             // output tile, TH * TW <= 128, TH even
    int tiles_x, tiles_y;
    int PH, PW;             // staged window: (TH-1)*S+3 x (TW-1)*S+3   (tc_dw2d_finish)
    uint32_t lbo_a;         // group stride of the A operand, bytes (tc_dw2d_finish)
    uint32_t mul_TW, mul_tiles_x, mul_tiles;   // fast_div multipliers (tc_dw2d_finish)
    const __half *wimg;     // [C/8][N][8]
    const float *bias;      // [N]
    const float *dw_w, *dw_b;   // [9][C], [C]
    __half *out;            // [nimg][OH][OW][N]
};

// Derived geometry, computed once on the host.  The staged window is PIXEL-major, [PH][PW][C] -- a byte-for-byte copy of the
// NHWC rows it comes from: consecutive cp.async lanes write consecutive shared addresses (one wavefront per 128 bytes; a
// channel-group-major layout scatters every 16-byte piece into its own wavefront), and the stencil's lanes -- (group, column)
// with the group fastest -- read consecutive 16-byte pieces.  The A operand's group stride in 16-byte units is 8/G mod 8, so
// that the 8 lanes of a quarter warp hit 8 different bank groups.
inline void tc_dw2d_finish(TcDw2dArgs &a) {
    const int G = a.C >> 3;
    a.PH = (a.TH - 1) * a.S + 3;
    a.PW = (a.TW - 1) * a.S + 3;
    a.tiles_x = (a.OW + a.TW - 1) / a.TW;
    a.tiles_y = (a.OH + a.TH - 1) / a.TH;
    a.lbo_a = (uint32_t)(128 + 8 / G) * 16;
    a.mul_TW = fast_div_mul((uint32_t)a.TW);
    a.mul_tiles_x = fast_div_mul((uint32_t)a.tiles_x);
    a.mul_tiles = fast_div_mul((uint32_t)(a.tiles_x * a.tiles_y));
}
inline size_t tc_dw2d_smem_bytes(const TcDw2dArgs &a) {
    return (size_t)a.PH * a.PW * a.C * 2 + (size_t)(a.C / 8) * a.lbo_a + (size_t)a.C * a.N * 2 + 128;
}

// resident CTAs per SM the register allocation aims at: 6 = 40 registers (24 bytes of spills), measured 81.6 -> 80.0 us per batch-8 step against 4 = 64 registers; the driver sizes the shared-memory carve-out to match
#ifndef RF_DW2D_OCC
#define RF_DW2D_OCC 6
#endif
template <int NT>
__global__ void __launch_bounds__(TC_THREADS, RF_DW2D_OCC) k_tc_dwpw_2d(const TcDw2dArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar_b, bar_done;
    __shared__ uint32_t s_tmem;
    __shared__ float s_bias[256];
    __shared__ __align__(16) __half s_dwh[9 * 64];    // [tap][C] folded depthwise weights, FP16 (fhfma8)
    __shared__ __align__(16) float s_dwb[64];         // [C] bias

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = a.C >> 3, lg = 31 - __clz(G);
    const int PH = a.PH, PW = a.PW;
    const uint32_t lbo_a = a.lbo_a;
    const int pix = a.C * 2;             // bytes per staged pixel
    unsigned char *sS = smem;
    unsigned char *sA = smem + (size_t)PH * PW * pix;
    unsigned char *sB = sA + (size_t)G * lbo_a;
    const int tiles = a.tiles_x * a.tiles_y;
    const int b = fast_div((int)blockIdx.x, a.mul_tiles), trem = blockIdx.x - b * tiles;
    const int ty0 = fast_div(trem, a.mul_tiles_x);
    const int oy0 = ty0 * a.TH, ox0 = (trem - ty0 * a.tiles_x) * a.TW;
    const int iy0 = oy0 * a.S - 1, ix0 = ox0 * a.S - 1;          // input coordinates of staged (0, 0)

    if (tid == 0) {
        tc::mbar_init(&bar_b, 1);
        tc::mbar_init(&bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned bytes = (unsigned)((size_t)a.C * a.N * 2);
        tc::mbar_expect_tx(&bar_b, bytes);
        tc::bulk_g2s(sB, a.wimg, bytes, &bar_b);
    }
    if (warp == 1) tc::tmem_alloc<NT>(&s_tmem);
    pdl_trigger();
    if (tid < a.N) s_bias[tid] = a.bias[tid];
    for (int i = tid; i < 9 * a.C; i += TC_THREADS) s_dwh[i] = __float2half_rn(a.dw_w[i]);
    if (tid < a.C) s_dwb[tid] = a.dw_b[tid];
    pdl_wait();
    // ---- stage the (PH x PW) input window: one warp per staged row, lanes over (column, channel group) -- a staged row
    // is PW*C contiguous halfs of the input (16 B per lane, fully coalesced); outside the map: zero fill ------------------
    {
        // item i of a row = (column px = i / G, group g = i % G): with C = 8 G its source is src_row + 8 i halfs -- affine in i
        const int per_row = PW << lg;
        const int px_lo = max(0, -ix0), px_hi = min(PW, a.IW - ix0);       // columns inside the map
        const unsigned px_n = (unsigned)max(px_hi - px_lo, 0);
        const uint32_t sS_s = tc::smem_u32(sS);
        for (int py = warp; py < PH; py += TC_THREADS / 32) {
            const int iy = iy0 + py;
            const bool rowok = iy >= 0 && iy < a.IH;
            const __half *src_row = a.in + (ptrdiff_t)(((b * a.IH + (rowok ? iy : 0)) * a.IW + ix0) * a.C);
            const uint32_t dst_row = sS_s + (uint32_t)(py * PW * pix);
            const __half *zsrc = a.in;      // any valid address: zero bytes are read from it
            for (int i = lane; i < per_row; i += 32) {
                const bool ok = rowok && (unsigned)((i >> lg) - px_lo) < px_n;
                cp_async16_zfill_s(dst_row + (uint32_t)i * 16u, ok ? src_row + i * 8 : zsrc, ok);
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();
    // ---- depthwise stencil -> A operand.  GEMM row r = ty * TW + tx ------------------------------------------------------
    const int rows = a.TH * a.TW;
    if (a.S == 1) {
        // item = (channel group, column, PAIR of output rows): 4 staged rows feed 2 outputs
        const int items = (a.TH >> 1) * a.TW << lg;
        for (int it = tid; it < items; it += TC_THREADS) {
            const int g = it & (G - 1), rest = it >> lg;
            const int typ = fast_div(rest, a.mul_TW), tx = rest - typ * a.TW;
            const int ty = typ * 2;
            float acc0[8], acc1[8];
            {
                const float4 b0 = *reinterpret_cast<const float4 *>(&s_dwb[g * 8]), b1 = *reinterpret_cast<const float4 *>(&s_dwb[g * 8 + 4]);
                acc0[0] = b0.x; acc0[1] = b0.y; acc0[2] = b0.z; acc0[3] = b0.w; acc0[4] = b1.x; acc0[5] = b1.y; acc0[6] = b1.z; acc0[7] = b1.w;
#pragma unroll
                for (int i = 0; i < 8; i++) acc1[i] = acc0[i];
            }
            const unsigned char *base = sS + (ty * PW + tx) * pix + g * 16;
#pragma unroll 1                     // (uniform branches on ry; keeps the 12 window loads from being hoisted into 48 registers)
            for (int ry = 0; ry < 4; ry++) {
#pragma unroll
                for (int kx = 0; kx < 3; kx++) {
                    const uint4 x = *reinterpret_cast<const uint4 *>(base + (ry * PW + kx) * pix);
                    if (ry < 3) fhfma8(acc0, x, *reinterpret_cast<const uint4 *>(&s_dwh[(ry * 3 + kx) * a.C + g * 8]));         // output row ty: kernel row ry
                    if (ry > 0) fhfma8(acc1, x, *reinterpret_cast<const uint4 *>(&s_dwh[((ry - 1) * 3 + kx) * a.C + g * 8]));   // output row ty + 1: kernel row ry - 1
                }
            }
#pragma unroll
            for (int i = 0; i < 8; i++) { acc0[i] = fmaxf(acc0[i], 0.f); acc1[i] = fmaxf(acc1[i], 0.f); }
            Vec8<__half> o0, o1;
            o0.from_float(acc0);
            o1.from_float(acc1);
            const int r = ty * a.TW + tx;
            *reinterpret_cast<uint4 *>(sA + (size_t)g * lbo_a + (size_t)r * 16) = o0.v;
            *reinterpret_cast<uint4 *>(sA + (size_t)g * lbo_a + (size_t)(r + a.TW) * 16) = o1.v;
        }
    } else {
        const int items = rows << lg;
        for (int it = tid; it < items; it += TC_THREADS) {
            const int g = it & (G - 1), r = it >> lg;
            const int ty = fast_div(r, a.mul_TW), tx = r - ty * a.TW;
            float acc[8];
            {
                const float4 b0 = *reinterpret_cast<const float4 *>(&s_dwb[g * 8]), b1 = *reinterpret_cast<const float4 *>(&s_dwb[g * 8 + 4]);
                acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
            }
            const unsigned char *base = sS + (ty * a.S * PW + tx * a.S) * pix + g * 16;
#pragma unroll
            for (int t = 0; t < 9; t++) {
                fhfma8(acc, *reinterpret_cast<const uint4 *>(base + ((t / 3) * PW + (t % 3)) * pix), *reinterpret_cast<const uint4 *>(&s_dwh[t * a.C + g * 8]));
            }
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = fmaxf(acc[i], 0.f);
            Vec8<__half> o;
            o.from_float(acc);
            *reinterpret_cast<uint4 *>(sA + (size_t)g * lbo_a + (size_t)r * 16) = o.v;
        }
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        tc::mbar_wait(&bar_b, 0);
        tc::tc_fence_after();
        const uint32_t idesc = (1u << 4) | ((uint32_t)(a.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_addr = tc::smem_u32(sA), b_addr = tc::smem_u32(sB);
        const uint32_t lbo_b = (uint32_t)a.N * 16;
        for (int ks = 0; ks < (a.C >> 4); ks++) {
            const uint64_t ad = tc::smem_desc(a_addr + (uint32_t)(2 * ks) * lbo_a, lbo_a, 128);
            const uint64_t bd = tc::smem_desc(b_addr + (uint32_t)(2 * ks) * lbo_b, lbo_b, 128);
            tc::mma_f16(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
        }
        tc::mma_commit(&bar_done);
    }
    if (warp == 0) tc::mbar_wait(&bar_done, 0);      // one warp polls; the block barrier (no issue slots) releases the rest
    __syncthreads();
    tc::tc_fence_after();
    {
        const int r = (warp & 3) * 32 + lane;
        const int ty = fast_div(r, a.mul_TW), tx = r - ty * a.TW;
        const int oy = oy0 + ty, ox = ox0 + tx;
        const bool ok = r < rows && oy < a.OH && ox < a.OW;
        TcOut o{a.out, a.N, a.N, 1, nullptr, 0, 0};
        tc_epilogue(tmem, a.N, s_bias, o, ok ? (long)((b * a.OH + oy) * a.OW + ox) : -1, 0);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<NT>(tmem);
}

}  // namespace rf
