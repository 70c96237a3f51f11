// tc_conv.cuh -- tcgen05 (5th-gen tensor core) convolution kernels of librf_b200, FP16 in / FP32
// accumulate in TMEM.  sm_100a only.
//
// One kernel template covers the three GEMM-shaped layer families of the network
// (model/mnet-deconv-0517.prototxt):
//   TC_PW   : pointwise 1x1 conv                       (rf_c*_lateral / rf_c1_red_conv)
//   TC_3X3  : full 3x3 conv, pad 1, stride 1           (rf_c*_aggr, SSH det/context convs)
//   TC_DWPW : depthwise 3x3 (stride 1|2) + BN + ReLU  fused with the following pointwise 1x1
//             conv + BN + ReLU (mobilenet0_conv{2k-1,2k}, k = 1..13): the depthwise stencil is
//             evaluated on CUDA cores straight into the shared-memory A operand of the tensor-core
//             GEMM, so the depthwise activation never touches global memory.
// GEMM view: D[128 pixels][N] += A[128][K] * W[N][K]^T, K = taps * Cin, one CTA per 128-pixel tile.
//   A operand: built in shared memory by all 128 threads ("im2col in registers": each thread gathers
//              16-byte channel groups of the shifted source pixels, or computes the depthwise result)
//              in the UMMA canonical K-major no-swizzle layout: 8x8 core matrices (8 rows x 16 B),
//              SBO = 128 B between 8-row groups, LBO = 128*16+16 B between 8-channel groups (the
//              16 B pad makes the 16-byte shared stores of a quarter-warp conflict free).
//   B operand: weights pre-packed on the host as the exact shared-memory image of each K chunk and
//              copied with a single cp.async.bulk (TMA bulk copy, mbarrier complete_tx) per chunk.
//   MMA      : tcgen05.mma.cta_group::1.kind::f16, M=128, N=Cout (16..256), K=16 per instruction,
//              issued by one thread; accumulator in TMEM (N fp32 columns x 128 lanes).
//   Pipeline : 2 shared-memory stages; tcgen05.commit -> mbarrier frees a stage / signals the epilogue.
//   Epilogue : tcgen05.ld 32x32b (thread = pixel row), + folded-BN bias, ReLU, FP16 pack, 16-byte
//              stores; output channels may be split over two destinations (SSH concat fusion).
#pragma once
#include "common.cuh"

namespace rf {

enum TcMode { TC_PW = 0, TC_3X3 = 1, TC_DWPW = 2 };

constexpr int TC_KC = 64;                      // K elements per pipeline chunk
constexpr int TC_LBO_A = 128 * 16 + 16;        // bytes between 8-channel groups of the A tile
constexpr int TC_A_STAGE = (TC_KC / 8) * TC_LBO_A;   // 16,512 B

struct TcOut {
    __half *p0; int ld0; int n0; int relu0;    // channels [0, n0)  -> p0[m*ld0 + n]
    __half *p1; int ld1; int relu1;            // channels [n0, N)  -> p1[m*ld1 + n - n0]
};

struct TcArgs {
    const __half *in;       // NHWC input, pixel stride = ldin
    int ldin, Cin;
    int nimg, IH, IW;       // input spatial size
    int OH, OW;             // output spatial size (== input except stride-2 depthwise)
    int N;                  // output channels
    int K;                  // taps * Cin (TC_DWPW: Cin), rounded up to a multiple of 16
    int Kreal;              // un-padded K: A groups at or beyond it are zero (Cin = 8 layer)
    const __half *wimg;     // packed B chunk images, chunk q at q * (N * TC_KC) halfs
    const float *bias;      // [N]
    const float *dw_w;      // TC_DWPW: [9][Cin] folded depthwise weights
    const float *dw_b;      // TC_DWPW: [Cin]
    int dw_stride;
    TcOut out;
};

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    unsigned done = 0;
    unsigned spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 24)) __trap();   // a lost arrive must fail loudly, never hang the GPU
    }
}
// TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1 = Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t r[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace tc

constexpr int tc_tmem_cols(int n) { return n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 128 ? 128 : 256)); }

// dynamic shared memory: 2 stages x (A tile + B chunk image)
inline size_t tc_smem_bytes(int N) { return 2 * ((size_t)TC_A_STAGE + (size_t)N * TC_KC * 2) + 128; }

template <int MODE, int NT /* TMEM columns: 32/64/128/256 */>
__global__ void __launch_bounds__(128) k_tc_conv(const TcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar_full[2];    // B chunk landed (complete_tx)
    __shared__ __align__(8) uint64_t bar_free[2];    // MMAs that read this stage have retired
    __shared__ __align__(8) uint64_t bar_done;       // all MMAs of the tile retired
    __shared__ uint32_t s_tmem;
    __shared__ int s_pb[128], s_py[128], s_px[128];  // output pixel -> (image, oy, ox); pb < 0: row beyond M
    __shared__ float s_dw[MODE == TC_DWPW ? 10 * 256 : 1];

    const int tid = threadIdx.x, warp = tid >> 5;
    const int N = a.N;
    const size_t b_stage_bytes = (size_t)N * TC_KC * 2;
    unsigned char *sA[2] = {smem, smem + TC_A_STAGE};
    unsigned char *sB[2] = {smem + 2 * TC_A_STAGE, smem + 2 * TC_A_STAGE + b_stage_bytes};

    const long M = (long)a.nimg * a.OH * a.OW;
    const long m0 = (long)blockIdx.x * 128;
    {
        long m = m0 + tid;
        if (m < M) {
            s_px[tid] = (int)(m % a.OW);
            s_py[tid] = (int)((m / a.OW) % a.OH);
            s_pb[tid] = (int)(m / ((long)a.OW * a.OH));
        } else {
            s_pb[tid] = -1; s_py[tid] = 0; s_px[tid] = 0;
        }
    }
    if (MODE == TC_DWPW) {
        for (int i = tid; i < 9 * a.Cin; i += 128) s_dw[i] = a.dw_w[i];
        for (int i = tid; i < a.Cin; i += 128) s_dw[9 * 256 + i] = a.dw_b[i];
    }
    if (tid == 0) {
        tc::mbar_init(&bar_full[0], 1); tc::mbar_init(&bar_full[1], 1);
        tc::mbar_init(&bar_free[0], 1); tc::mbar_init(&bar_free[1], 1);
        tc::mbar_init(&bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tc::tmem_alloc<NT>(&s_tmem);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = s_tmem;

    const int K = a.K;
    const int nchunks = (K + TC_KC - 1) / TC_KC;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    for (int q = 0; q < nchunks; q++) {
        const int st = q & 1;
        const int use = q >> 1;                          // how many times this stage was used before
        const int kc = min(TC_KC, K - q * TC_KC);        // K elements in this chunk (multiple of 16)
        if (use > 0) tc::mbar_wait(&bar_free[st], (use - 1) & 1);
        // ---- B chunk: one TMA bulk copy of the pre-packed image --------------------------------
        if (tid == 0) {
            const unsigned bytes = (unsigned)((size_t)N * kc * 2);
            tc::mbar_expect_tx(&bar_full[st], bytes);
            tc::bulk_g2s(sB[st], a.wimg + (size_t)q * N * TC_KC, bytes, &bar_full[st]);
        }
        // ---- A chunk: 128 rows x kc channels, 16-byte items (row, 8-channel group) ---------------
        const int G = kc >> 3;                           // groups in this chunk (2, 4, 8 ...)
        for (int it = tid; it < 128 * G; it += 128) {
            const int g = it % G, r = it / G;
            const int kidx = q * TC_KC + g * 8;          // global K index of this group
            uint4 v = make_uint4(0, 0, 0, 0);
            const int pb = s_pb[r];
            if (pb >= 0 && kidx < a.Kreal) {
                if (MODE == TC_DWPW) {
                    // depthwise 3x3 (+BN+ReLU) of channels [kidx, kidx+8) at output pixel r
                    const int c0 = kidx;
                    float acc[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) acc[i] = s_dw[9 * 256 + c0 + i];
                    const int S = a.dw_stride;
                    const __half *base = a.in + (size_t)pb * a.IH * a.IW * a.ldin + c0;
#pragma unroll
                    for (int ky = 0; ky < 3; ky++) {
                        const int iy = s_py[r] * S + ky - 1;
                        if (iy < 0 || iy >= a.IH) continue;
#pragma unroll
                        for (int kx = 0; kx < 3; kx++) {
                            const int ix = s_px[r] * S + kx - 1;
                            if (ix < 0 || ix >= a.IW) continue;
                            Vec8<__half> x;
                            x.load(base + ((size_t)iy * a.IW + ix) * a.ldin);
                            float f[8];
                            x.to_float(f);
                            const float *wr = &s_dw[(ky * 3 + kx) * a.Cin + c0];
#pragma unroll
                            for (int i = 0; i < 8; i++) acc[i] = fmaf(f[i], wr[i], acc[i]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 8; i++) acc[i] = fmaxf(acc[i], 0.f);
                    Vec8<__half> o;
                    o.from_float(acc);
                    v = o.v;
                } else {
                    const int tap = MODE == TC_3X3 ? kidx / a.Cin : 0;
                    const int c0 = MODE == TC_3X3 ? kidx - tap * a.Cin : kidx;
                    const int iy = s_py[r] + (MODE == TC_3X3 ? tap / 3 - 1 : 0);
                    const int ix = s_px[r] + (MODE == TC_3X3 ? tap % 3 - 1 : 0);
                    if (iy >= 0 && iy < a.IH && ix >= 0 && ix < a.IW)
                        v = *reinterpret_cast<const uint4 *>(a.in + (((size_t)pb * a.IH + iy) * a.IW + ix) * a.ldin + c0);
                }
            }
            *reinterpret_cast<uint4 *>(sA[st] + g * TC_LBO_A + r * 16) = v;
        }
        tc::fence_async_smem();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncthreads();
        // ---- MMA: one thread issues kc/16 instructions -------------------------------------------
        if (tid == 0) {
            tc::mbar_wait(&bar_full[st], use & 1);       // weights landed
            tc::tc_fence_after();
            const uint32_t a_addr = tc::smem_u32(sA[st]), b_addr = tc::smem_u32(sB[st]);
            const uint32_t lbo_b = (uint32_t)N * 16;
            for (int ks = 0; ks < kc / 16; ks++) {
                const uint64_t ad = tc::smem_desc(a_addr + ks * 2 * TC_LBO_A, TC_LBO_A, 128);
                const uint64_t bd = tc::smem_desc(b_addr + ks * 2 * lbo_b, lbo_b, 128);
                tc::mma_f16(tmem, ad, bd, idesc, (q > 0 || ks > 0) ? 1u : 0u);
            }
            tc::mma_commit(&bar_free[st]);               // frees this stage when the MMAs retire
            if (q == nchunks - 1) tc::mma_commit(&bar_done);
        }
    }
    // ---- epilogue: TMEM -> registers -> bias/ReLU -> FP16 -> global ---------------------------------
    tc::mbar_wait(&bar_done, 0);
    tc::tc_fence_after();
    const long m = m0 + tid;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int n0 = 0; n0 < N; n0 += 16) {
        uint32_t r[16];
        tc::tmem_ld16(lane_addr + n0, r);
        tc::tmem_ld_wait();
        if (m < M) {
            const bool first = n0 < a.out.n0;
            const int relu = first ? a.out.relu0 : a.out.relu1;
            __half *dst = first ? a.out.p0 + (size_t)m * a.out.ld0 + n0 : a.out.p1 + (size_t)m * a.out.ld1 + (n0 - a.out.n0);
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                f[i] = __uint_as_float(r[i]) + __ldg(a.bias + n0 + i);
                if (relu) f[i] = fmaxf(f[i], 0.f);
            }
            Vec8<__half> o0, o1;
            o0.from_float(f);
            o1.from_float(f + 8);
            o0.store(dst);
            o1.store(dst + 8);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<NT>(tmem);
}

}  // namespace rf
