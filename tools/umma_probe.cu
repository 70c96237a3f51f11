// umma_probe.cu -- hardware-semantics probe for the round-2 tile kernels (development tool, not product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o umma_probe umma_probe.cu ; ./umma_probe <test>
// Questions it answers on a B200 (each test is its own process: a trap in one cannot poison the next):
//   tma128 / tma64 / tma32 : where does a TMA tiled load (4-D NHWC map, OOB zero fill, SWIZZLE_128B/64B/32B) put the 16-byte
//                            chunk j of position p?   expected: p*ROW + ((j ^ f(p)) << 4)
//   conv128/64/32 v        : 3x3 conv as 9 row-shifted K-major swizzled descriptors over ONE staged tile
//                            (start address moved by shift*ROW bytes), base_offset variant v (0: always 0, 1: (addr>>7)&7)
//   dw                     : depthwise 3x3 as 9*C/16 diagonal MMAs (N=16) into TMEM column offsets, mid-epilogue
//                            TMEM -> bias+ReLU -> FP16 -> hand-swizzled SW128 A operand -> pointwise MMA
//   ts                     : same, but the pointwise A operand comes from TMEM (tcgen05.st FP16 pairs, .kind::f16 TS form)
//   s2                     : TMA with elementStrides 2 (parity planes for stride-2 depthwise)
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>

#include "../retinaface_b200/csrc/tc_conv.cuh"

using namespace rf;

#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled get_encode() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CKC(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(2); }
    return (EncodeTiled)fn;
}

// NHWC fp16 tensor [B][H][W][C] -> 4-D map {C, W, H, B}, box {bc, bw, bh, 1}
static CUtensorMap make_map(void *g, int C, int W, int H, int B, int bc, int bw, int bh, CUtensorMapSwizzle sw, int estride_w = 1, int estride_h = 1) {
    CUtensorMap m;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t es[4] = {1, (cuuint32_t)estride_w, (cuuint32_t)estride_h, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(2); }
    return m;
}

__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(tc::smem_u32(dst)), "l"(map), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// K-major swizzled descriptor: layout 2 = SW128, 4 = SW64, 6 = SW32; SBO = 8 rows; LBO unused (1)
__device__ __forceinline__ uint64_t sw_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t layout, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)layout << 61;
    return d;
}

struct ProbeArgs {
    int C, W, H, B;          // tensor
    int bw, bh;              // box (positions per row, rows)
    int x0, y0, b0;          // box origin
    int mode;                // 0: dump tile; 1: conv3x3; 2: dw+pw; 3: dw + pw(TS)
    int variant;             // base_offset variant
    int N;                   // conv / pw out channels
    int layout;              // 2/4/6
    const __half *wimg;      // B image(s)
    const float *bias;       // dw bias [C]
    unsigned char *dump;     // mode 0: raw tile bytes
    float *out;              // [128][N]
};

// One CTA of 128 threads.
__global__ void __launch_bounds__(128) k_probe(const __grid_constant__ CUtensorMap map, ProbeArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_t, bar_w, bar_m, bar_m2;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int ROW = a.C * 2 > 128 ? 128 : a.C * 2;         // bytes per position per slab
    const int slabs = (a.C * 2 + 127) / 128;
    const int P = a.bw * a.bh;
    const int tile_bytes = P * ROW;                        // per slab
    unsigned char *sT = smem;                              // [slab][P][ROW]
    unsigned char *sW = sT + ((slabs * tile_bytes + 1023) & ~1023) + 4096;   // weights (+ slack for row over-reads)
    unsigned char *sA2 = sW + 64 * 1024;                   // second A operand (dw -> pw), 128 rows x C*2 bytes, [slab][128][128]
    if (tid == 0) {
        tc::mbar_init(&bar_t, 1); tc::mbar_init(&bar_w, 1); tc::mbar_init(&bar_m, 1); tc::mbar_init(&bar_m2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tc::tmem_alloc<512>(&s_tmem);
    // zero the slack behind the tile (rows the shifted descriptors over-read)
    for (int i = tid; i < 4096 / 16; i += 128) reinterpret_cast<uint4 *>(sT + slabs * tile_bytes)[i] = make_uint4(0, 0, 0, 0);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = s_tmem;
    int wbytes = 0;
    if (a.mode == 1) wbytes = 9 * a.C * a.N * 2;
    if (a.mode >= 2) wbytes = 9 * (a.C / 16) * 512 + a.C * a.N * 2;
    if (tid == 0) {
        tc::mbar_expect_tx(&bar_t, (unsigned)(slabs * tile_bytes));
        for (int s = 0; s < slabs; s++) tma_load_4d(sT + s * tile_bytes, &map, &bar_t, s * 64, a.x0, a.y0, a.b0);
        if (wbytes) { tc::mbar_expect_tx(&bar_w, (unsigned)wbytes); tc::bulk_g2s(sW, a.wimg, (unsigned)wbytes, &bar_w); }
    }
    tc::mbar_wait(&bar_t, 0);
    if (a.mode == 0) {
        for (int i = tid; i < slabs * tile_bytes / 16; i += 128) reinterpret_cast<uint4 *>(a.dump)[i] = reinterpret_cast<const uint4 *>(sT)[i];
        __syncthreads();
        if (warp == 1) tc::tmem_dealloc<512>(tmem);
        return;
    }
    tc::mbar_wait(&bar_w, 0);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const int Wl = a.bw;
    const int row0 = Wl + 1;                               // first output position
    const uint32_t sbo = 8 * ROW;
    auto a_desc = [&](int slab, int pos, int kbyte) {
        const uint32_t addr = tc::smem_u32(sT) + slab * tile_bytes + pos * ROW + kbyte;
        return sw_desc(addr, sbo, a.layout, a.variant ? (addr >> 7) & 7 : 0);
    };
    if (a.mode == 1) {
        if (tid == 0) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(a.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t b_addr = tc::smem_u32(sW), lbo_b = (uint32_t)a.N * 16;
            uint32_t acc = 0;
            for (int t = 0; t < 9; t++) {
                const int shift = (t / 3 - 1) * Wl + (t % 3 - 1);
                for (int ks = 0; ks < a.C / 16; ks++) {
                    const uint64_t ad = a_desc(ks / 4, row0 + shift, (ks % 4) * 32);
                    const uint64_t bd = tc::smem_desc(b_addr + (uint32_t)(t * (a.C / 8) + 2 * ks) * lbo_b, lbo_b, 128);
                    tc::mma_f16(tmem, ad, bd, idesc, acc);
                    acc = 1;
                }
            }
            tc::mma_commit(&bar_m);
        }
        tc::mbar_wait(&bar_m, 0);
        tc::tc_fence_after();
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int n0 = 0; n0 < a.N; n0 += 16) {
            uint32_t r[16];
            tc::tmem_ld16(lane_addr + n0, r);
            tc::tmem_ld_wait();
            for (int i = 0; i < 16; i++) a.out[tid * a.N + n0 + i] = __uint_as_float(r[i]);
        }
    } else {
        // depthwise: 9 taps x C/16 slabs of diagonal 16x16 B tiles -> TMEM columns [0, C)
        const uint32_t idesc16 = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        if (tid == 0) {
            const uint32_t b_addr = tc::smem_u32(sW);
            for (int s = 0; s < a.C / 16; s++)
                for (int t = 0; t < 9; t++) {
                    const int shift = (t / 3 - 1) * Wl + (t % 3 - 1);
                    const uint64_t ad = a_desc(s / 4, row0 + shift, (s % 4) * 32);
                    const uint64_t bd = tc::smem_desc(b_addr + (uint32_t)(t * (a.C / 16) + s) * 512, 256, 128);
                    tc::mma_f16(tmem + s * 16, ad, bd, idesc16, t > 0);
                }
            tc::mma_commit(&bar_m);
        }
        tc::mbar_wait(&bar_m, 0);
        tc::tc_fence_after();
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        const uint32_t ACOL = 256;                          // TS: FP16 A operand at TMEM columns [256, 256 + C/2)
        for (int s = 0; s < a.C / 16; s++) {
            uint32_t r[16];
            tc::tmem_ld16(lane_addr + s * 16, r);
            tc::tmem_ld_wait();
            uint32_t pk[8];
            for (int i = 0; i < 8; i++) {
                float v0 = fmaxf(__uint_as_float(r[2 * i]) + a.bias[s * 16 + 2 * i], 0.f);
                float v1 = fmaxf(__uint_as_float(r[2 * i + 1]) + a.bias[s * 16 + 2 * i + 1], 0.f);
                __half2 h = __floats2half2_rn(v0, v1);
                pk[i] = *reinterpret_cast<uint32_t *>(&h);
            }
            if (a.mode == 2) {
                const int slab = s / 4, j0 = (s % 4) * 2;
                unsigned char *rowp = sA2 + slab * (128 * 128) + tid * 128;
                *reinterpret_cast<uint4 *>(rowp + (((j0) ^ (tid & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4 *>(rowp + (((j0 + 1) ^ (tid & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            } else {
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(lane_addr + ACOL + s * 8),
                             "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
            }
        }
        if (a.mode == 3) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc::fence_async_smem();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        const uint32_t DCOL = 128 + 0;                      // pointwise accumulator columns [DCOL, DCOL + N) -- C <= 128 here
        if (tid == 0) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(a.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t b_addr = tc::smem_u32(sW) + 9 * (a.C / 16) * 512, lbo_b = (uint32_t)a.N * 16;
            for (int ks = 0; ks < a.C / 16; ks++) {
                const uint64_t bd = tc::smem_desc(b_addr + (uint32_t)(2 * ks) * lbo_b, lbo_b, 128);
                if (a.mode == 2) {
                    const uint32_t addr = tc::smem_u32(sA2) + (ks / 4) * (128 * 128) + (ks % 4) * 32;
                    tc::mma_f16(tmem + DCOL, sw_desc(addr, 1024, 2, 0), bd, idesc, ks > 0);
                } else {
                    const uint32_t acc = ks > 0;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "setp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                        ::"r"(tmem + DCOL), "r"(tmem + ACOL + ks * 8), "l"(bd), "r"(idesc), "r"(acc)
                        : "memory");
                }
            }
            tc::mma_commit(&bar_m2);
        }
        tc::mbar_wait(&bar_m2, 0);
        tc::tc_fence_after();
        for (int n0 = 0; n0 < a.N; n0 += 16) {
            uint32_t r[16];
            tc::tmem_ld16(lane_addr + DCOL + n0, r);
            tc::tmem_ld_wait();
            for (int i = 0; i < 16; i++) a.out[tid * a.N + n0 + i] = __uint_as_float(r[i]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<512>(tmem);
}

// MMA latency / throughput: `count` MMAs (M=128, N=n, K=16, SS mode, SW128 A) issued by one thread round-robin over `nacc`
// independent accumulators; cycles from first issue to the commit's arrival
__global__ void __launch_bounds__(128) k_mma_rate(int n, int count, int nacc, int ts, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 64 * 1024 / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { tc::mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 1) tc::tmem_alloc<512>(&s_tmem);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_addr = tc::smem_u32(smem), b_addr = a_addr + 32 * 1024;
        const long long t0 = clock64();
        for (int i = 0; i < count; i++) {
            const uint64_t ad = sw_desc(a_addr + ((i % 9) * 7 + 3) * 128 + (i % 4) * 32, 1024, 2, 0);
            const uint64_t bd = tc::smem_desc(b_addr + (i % 8) * 1024, (uint32_t)n * 16, 128);
            const uint32_t d = tmem + (uint32_t)(i % nacc) * (uint32_t)n;
            if (ts) {
                const uint32_t acc = i >= nacc;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                             ::"r"(d), "r"(tmem + 448u), "l"(bd), "r"(idesc), "r"(acc) : "memory");
            } else {
                tc::mma_f16(d, ad, bd, idesc, i >= nacc);
            }
        }
        const long long t1 = clock64();
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<512>(tmem);
}

// MMA throughput with NOTHING but the instructions in the issuing thread: fully unrolled, every operand a compile-time offset
// from one base.  LAYOUT 0: SW128 A (64-channel rows, K-step = 32 B of each 128-byte row); 1: no-swizzle planes (16 B per
// row per K-chunk).  D rotates over NACC accumulators, A over 9 row shifts x 4 K-steps like a depthwise / 3x3 issue block.
template <int N, int NACC, int TS, int LAYOUT>
__global__ void __launch_bounds__(128) k_mma_rate2(long long *out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 96 * 1024 / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { tc::mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 1) tc::tmem_alloc<512>(&s_tmem);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = s_tmem;
    constexpr int COUNT = 144;
    if (warp == 0) {
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        if (leader) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t a_addr = tc::smem_u32(smem), b_addr = a_addr + 64 * 1024;
            const uint64_t a0 = LAYOUT == 0 ? sw_desc(a_addr, 1024, 2, 0) : tc::smem_desc(a_addr, 24 * 1024, 128);
            const uint64_t b0 = tc::smem_desc(b_addr, (uint32_t)N * 16, 128);
            const long long t0 = clock64();
#pragma unroll
            for (int i = 0; i < COUNT; i++) {
                const int t = (i / 4) % 9, k = i % 4;
                const uint64_t ad = a0 + (uint64_t)(LAYOUT == 0 ? ((t * 7 + 3) * 128 + k * 32) / 16 : ((t * 7 + 3) * 16 + (k & 1) * 8 * 1024) / 16);
                const uint64_t bd = b0 + (uint64_t)((i % 8) * 512 / 16);
                const uint32_t d = tmem + (uint32_t)((i % NACC) * N);
                if (TS) {
                    if (i >= NACC)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                     ::"r"(d), "r"(tmem + 448u), "l"(bd), "r"(idesc) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                     ::"r"(d), "r"(tmem + 448u), "l"(bd), "r"(idesc) : "memory");
                } else {
                    if (i >= NACC)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                     ::"r"(d), "l"(ad), "l"(bd), "r"(idesc) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                     ::"r"(d), "l"(ad), "l"(bd), "r"(idesc) : "memory");
                }
            }
            const long long t1 = clock64();
            tc::mma_commit(&bar);
            tc::mbar_wait(&bar, 0);
            const long long t2 = clock64();
            out[0] = t1 - t0; out[1] = t2 - t0;
        }
        __syncwarp();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<512>(tmem);
}
template <int N, int NACC, int TS, int LAYOUT>
static void run_rate2(long long *dout) {
    CKC(cudaFuncSetAttribute(k_mma_rate2<N, NACC, TS, LAYOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    long long best[2] = {1ll << 60, 1ll << 60};
    for (int rep = 0; rep < 3; rep++) {
        k_mma_rate2<N, NACC, TS, LAYOUT><<<1, 128, 100 * 1024>>>(dout);
        CKC(cudaDeviceSynchronize());
        long long h2[2];
        CKC(cudaMemcpy(h2, dout, 16, cudaMemcpyDeviceToHost));
        if (h2[1] < best[1]) { best[0] = h2[0]; best[1] = h2[1]; }
    }
    printf("rate2 %s %s N=%3d accumulators=%d: issue %5.1f cyc/MMA, complete %6.1f cyc/MMA (144 MMAs, M=128 K=16)\n", TS ? "TS" : "SS",
           LAYOUT ? "noswz" : "sw128", N, NACC, best[0] / 144.0, best[1] / 144.0);
}

struct StoreMaps { CUtensorMap in; CUtensorMap st[3]; };
// loads a box with maps.in and stores it back through maps.st[idx] at (0, sx, sy, b)
__global__ void __launch_bounds__(128) k_store_probe(const __grid_constant__ StoreMaps maps, int idx, int lx, int ly, int sx, int sy, int b, unsigned bytes) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        tc::mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tc::mbar_expect_tx(&bar, bytes);
        tma_load_4d(smem, &maps.in, &bar, 0, lx, ly, b);
    }
    __syncthreads();
    tc::mbar_wait(&bar, 0);
    tc::fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                     ::"l"(&maps.st[idx]), "r"(tc::smem_u32(smem)), "r"(0), "r"(sx), "r"(sy), "r"(b) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

static float h2f(__half h) { return __half2float(h); }

int main(int argc, char **argv) {
    std::string test = argc > 1 ? argv[1] : "tma128";
    const int variant = argc > 2 ? atoi(argv[2]) : 0;
    int C = 64, layout = 2;
    CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
    if (test.find("64") != std::string::npos && test != "tma128" && test != "conv128") { C = 32; layout = 4; sw = CU_TENSOR_MAP_SWIZZLE_64B; }
    if (test.find("32") != std::string::npos) { C = 16; layout = 6; sw = CU_TENSOR_MAP_SWIZZLE_32B; }
    if (test == "dw128" || test == "ts128") { C = 128; }
    const int W = 30, H = 28, B = 2, bw = 32, bh = 7, x0 = -1, y0 = 24, b0 = 1;     // box hangs over the right/bottom edge; x0 = -1: left pad
    const int ROW = C * 2 > 128 ? 128 : C * 2, slabs = (C * 2 + 127) / 128, P = bw * bh;
    std::vector<__half> hin((size_t)B * H * W * C);
    srand(1);
    for (auto &v : hin) v = __float2half((float)((rand() % 17) - 8) / 8.0f);
    __half *din;
    CKC(cudaMalloc(&din, hin.size() * 2));
    CKC(cudaMemcpy(din, hin.data(), hin.size() * 2, cudaMemcpyHostToDevice));
    auto in_at = [&](int p, int c) -> float {       // local position p of the box -> tensor value or 0 (OOB)
        const int lx = p % bw, ly = p / bw;
        const int x = x0 + lx, y = y0 + ly;
        if (ly >= bh || x < 0 || x >= W || y < 0 || y >= H) return 0.f;
        return h2f(hin[(((size_t)b0 * H + y) * W + x) * C + c]);
    };
    ProbeArgs a{};
    a.C = C; a.W = W; a.H = H; a.B = B; a.bw = bw; a.bh = bh; a.x0 = x0; a.y0 = y0; a.b0 = b0; a.variant = variant; a.layout = layout;
    CUtensorMap map = make_map(din, C, W, H, B, C > 64 ? 64 : C, bw, bh, sw);
    const size_t smem = 200 * 1024;
    CKC(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    if (test.rfind("tma", 0) == 0) {
        a.mode = 0;
        CKC(cudaMalloc(&a.dump, slabs * P * ROW));
        k_probe<<<1, 128, smem>>>(map, a);
        CKC(cudaDeviceSynchronize());
        std::vector<__half> d((size_t)slabs * P * ROW / 2);
        CKC(cudaMemcpy(d.data(), a.dump, d.size() * 2, cudaMemcpyDeviceToHost));
        // hypothesis: chunk j of position p at p*ROW + ((j ^ ((p >> sh) & m)) << 4), SW128: sh 0 m 7; SW64: sh 1 m 3; SW32: sh 2 m 1
        const int sh = layout == 2 ? 0 : (layout == 4 ? 1 : 2), m = layout == 2 ? 7 : (layout == 4 ? 3 : 1);
        long bad = 0, badplain = 0;
        for (int p = 0; p < P; p++)
            for (int c = 0; c < C; c++) {
                const int slab = c / 64, cc = c % 64, j = cc / 8;
                const float want = in_at(p, c);
                const float got = h2f(d[((size_t)slab * P * ROW + (size_t)p * ROW + (((j ^ ((p >> sh) & m)) << 4)) + (cc % 8) * 2) / 2]);
                const float plain = h2f(d[((size_t)slab * P * ROW + (size_t)p * ROW + (j << 4) + (cc % 8) * 2) / 2]);
                bad += got != want;
                badplain += plain != want;
            }
        printf("%s: C=%d ROW=%d: mismatches with swizzle hypothesis %ld, with plain layout %ld of %d\n", test.c_str(), C, ROW, bad, badplain, P * C);
        return bad ? 1 : 0;
    }
    if (test.rfind("conv", 0) == 0) {
        const int N = 32;
        a.mode = 1; a.N = N;
        std::vector<float> w((size_t)9 * C * N);
        for (auto &v : w) v = (float)((rand() % 9) - 4) / 4.0f;
        std::vector<__half> img((size_t)9 * C * N);
        for (int t = 0; t < 9; t++) for (int c = 0; c < C; c++) for (int n = 0; n < N; n++) {
            const int kk = t * C + c;
            img[((size_t)(kk / 8) * N + n) * 8 + kk % 8] = __float2half(w[((size_t)t * C + c) * N + n]);
        }
        __half *dw_; CKC(cudaMalloc(&dw_, img.size() * 2)); CKC(cudaMemcpy(dw_, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
        a.wimg = dw_;
        CKC(cudaMalloc(&a.out, 128 * N * 4));
        k_probe<<<1, 128, smem>>>(map, a);
        CKC(cudaDeviceSynchronize());
        std::vector<float> out(128 * N);
        CKC(cudaMemcpy(out.data(), a.out, out.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0; long bad = 0;
        for (int r = 0; r < 128; r++) for (int n = 0; n < N; n++) {
            const int p = bw + 1 + r;
            double ref = 0;
            for (int t = 0; t < 9; t++) { const int q = p + (t / 3 - 1) * bw + (t % 3 - 1); for (int c = 0; c < C; c++) ref += (double)in_at(q, c) * w[((size_t)t * C + c) * N + n]; }
            const double e = fabs(ref - out[r * N + n]);
            if (e > maxerr) maxerr = e;
            bad += e > 1e-2;
        }
        printf("%s variant %d: C=%d max err %.5f, bad %ld of %d\n", test.c_str(), variant, C, maxerr, bad, 128 * N);
        return bad ? 1 : 0;
    }
    if (test.rfind("dw", 0) == 0 || test.rfind("ts", 0) == 0) {
        const int N = 64;
        a.mode = test[0] == 'd' ? 2 : 3; a.N = N;
        std::vector<float> wd((size_t)9 * C), bd(C), wp((size_t)C * N);
        for (auto &v : wd) v = (float)((rand() % 9) - 4) / 8.0f;
        for (auto &v : bd) v = (float)((rand() % 9) - 4) / 4.0f;
        for (auto &v : wp) v = (float)((rand() % 9) - 4) / 8.0f;
        std::vector<__half> img((size_t)9 * (C / 16) * 256 + (size_t)C * N, __float2half(0.f));
        for (int t = 0; t < 9; t++) for (int s = 0; s < C / 16; s++) for (int i = 0; i < 16; i++)
            img[((size_t)(t * (C / 16) + s)) * 256 + ((size_t)(i / 8) * 16 + i) * 8 + i % 8] = __float2half(wd[t * C + s * 16 + i]);   // B[n=i][k=i]
        for (int c = 0; c < C; c++) for (int n = 0; n < N; n++) img[(size_t)9 * (C / 16) * 256 + ((size_t)(c / 8) * N + n) * 8 + c % 8] = __float2half(wp[(size_t)c * N + n]);
        __half *dw_; CKC(cudaMalloc(&dw_, img.size() * 2)); CKC(cudaMemcpy(dw_, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
        float *db; CKC(cudaMalloc(&db, C * 4)); CKC(cudaMemcpy(db, bd.data(), C * 4, cudaMemcpyHostToDevice));
        a.wimg = dw_; a.bias = db;
        CKC(cudaMalloc(&a.out, 128 * N * 4));
        k_probe<<<1, 128, smem>>>(map, a);
        CKC(cudaDeviceSynchronize());
        std::vector<float> out(128 * N);
        CKC(cudaMemcpy(out.data(), a.out, out.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0; long bad = 0;
        for (int r = 0; r < 128; r++) {
            const int p = bw + 1 + r;
            std::vector<float> mid(C);
            for (int c = 0; c < C; c++) {
                float acc = 0;
                for (int t = 0; t < 9; t++) acc += in_at(p + (t / 3 - 1) * bw + (t % 3 - 1), c) * wd[t * C + c];
                mid[c] = h2f(__float2half(fmaxf(acc + bd[c], 0.f)));
            }
            for (int n = 0; n < N; n++) {
                double ref = 0;
                for (int c = 0; c < C; c++) ref += (double)mid[c] * wp[(size_t)c * N + n];
                const double e = fabs(ref - out[r * N + n]);
                if (e > maxerr) maxerr = e;
                bad += e > 2e-2;
            }
        }
        printf("%s variant %d: C=%d max err %.5f, bad %ld of %d\n", test.c_str(), variant, C, maxerr, bad, 128 * N);
        return bad ? 1 : 0;
    }
    if (test == "s2") {
        // elementStrides 2 along W and H: box position (lx, ly) <- tensor (x0 + 2*lx, y0 + 2*ly)?  dump and report the mapping
        CUtensorMap m2 = make_map(din, C, W, H, B, C, 32, 8, CU_TENSOR_MAP_SWIZZLE_128B, 2, 2);   // boxDim counts UN-strided elements: 32x8 -> 16x4 loaded
        a.mode = 0; a.bw = 16; a.bh = 4; a.x0 = -1; a.y0 = 3; a.b0 = 0;
        CKC(cudaMalloc(&a.dump, 16 * 4 * ROW));
        k_probe<<<1, 128, smem>>>(m2, a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("s2: launch failed: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<__half> d((size_t)16 * 4 * ROW / 2);
        CKC(cudaMemcpy(d.data(), a.dump, d.size() * 2, cudaMemcpyDeviceToHost));
        long bad = 0;
        for (int p = 0; p < 64; p++) for (int c = 0; c < C; c++) {
            const int lx = p % 16, ly = p / 16, x = -1 + 2 * lx, y = 3 + 2 * ly;
            const float want = (x < 0 || x >= W || y < 0 || y >= H) ? 0.f : h2f(hin[(((size_t)0 * H + y) * W + x) * C + c]);
            const int j = c / 8;
            const float got = h2f(d[((size_t)p * ROW + ((j ^ (p & 7)) << 4) + (c % 8) * 2) / 2]);
            bad += got != want;
        }
        printf("s2: elementStrides {1,2,2,1}, hypothesis box(lx,ly) <- (x0+2lx, y0+2ly): mismatches %ld of %d\n", bad, 64 * C);
        return bad ? 1 : 0;
    }
    if (test == "rate2") {
        long long *dout;
        CKC(cudaMalloc(&dout, 16));
        run_rate2<16, 1, 0, 0>(dout); run_rate2<16, 4, 0, 0>(dout); run_rate2<16, 4, 0, 1>(dout); run_rate2<32, 4, 0, 0>(dout);
        run_rate2<48, 2, 0, 0>(dout); run_rate2<64, 1, 0, 0>(dout); run_rate2<64, 2, 0, 0>(dout); run_rate2<64, 2, 0, 1>(dout);
        run_rate2<128, 1, 0, 0>(dout); run_rate2<128, 2, 0, 0>(dout); run_rate2<256, 1, 0, 0>(dout);
        run_rate2<16, 4, 1, 0>(dout); run_rate2<32, 2, 1, 0>(dout); run_rate2<64, 1, 1, 0>(dout); run_rate2<64, 2, 1, 0>(dout);
        run_rate2<128, 2, 1, 0>(dout); run_rate2<256, 1, 1, 0>(dout);
        return 0;
    }
    if (test == "rate") {
        long long *dout;
        CKC(cudaMalloc(&dout, 16));
        CKC(cudaFuncSetAttribute(k_mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        for (int ts = 0; ts < 2; ts++)
            for (int n : {16, 32, 64, 128})
                for (int nacc : {1, 2, 3, 4, 8}) {
                    if (nacc * n > 448) continue;
                    const int count = 72;
                    k_mma_rate<<<1, 128, 80 * 1024>>>(n, count, nacc, ts, dout);
                    CKC(cudaDeviceSynchronize());
                    k_mma_rate<<<1, 128, 80 * 1024>>>(n, count, nacc, ts, dout);
                    CKC(cudaDeviceSynchronize());
                    long long h2[2];
                    CKC(cudaMemcpy(h2, dout, 16, cudaMemcpyDeviceToHost));
                    printf("rate %s N=%3d accumulators=%d: issue %5.1f cyc/MMA, complete %6.1f cyc/MMA (%d MMAs)\n", ts ? "TS" : "SS", n, nacc, (double)h2[0] / count, (double)h2[1] / count, count);
                }
        return 0;
    }
    if (test == "store") {
        // variant 0: in-bounds box; 1: box starting at x = -1 (pad column) and hanging over the bottom edge
        __half *dout;
        CKC(cudaMalloc(&dout, hin.size() * 2));
        CKC(cudaMemset(dout, 0, hin.size() * 2));
        StoreMaps sm;
        // variants: 0 in bounds; 1 in bounds, runtime map index 2; 2 box hangs over right + bottom edge; 3 box starts at x = -1
        const int sbw = 16, sbh = 4;
        sm.in = make_map(din, C, W, H, B, C, sbw, sbh, sw);
        for (int i = 0; i < 3; i++) sm.st[i] = make_map(dout, C, W, H, B, C, sbw, sbh, sw);
        const int lx = variant == 3 ? -1 : (variant == 2 ? 20 : 4), ly = variant == 2 ? 26 : 8;
        CKC(cudaFuncSetAttribute(k_store_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        k_store_probe<<<1, 128, 64 * 1024>>>(sm, variant == 1 ? 2 : 0, lx, ly, lx, ly, 1, (unsigned)(sbw * sbh * C * 2));
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("store variant %d: FAILED: %s\n", variant, cudaGetErrorString(e)); return 1; }
        std::vector<__half> ho(hin.size());
        CKC(cudaMemcpy(ho.data(), dout, ho.size() * 2, cudaMemcpyDeviceToHost));
        long bad = 0, wrote = 0;
        for (int b = 0; b < B; b++) for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) for (int c = 0; c < C; c++) {
            const size_t i = (((size_t)b * H + y) * W + x) * C + c;
            const bool inbox = b == 1 && x >= lx && x < lx + sbw && y >= ly && y < ly + sbh;   // (x, y range over the tensor: out-of-bounds box parts never count)
            const float want = inbox ? h2f(hin[i]) : 0.f;
            bad += h2f(ho[i]) != want;
            wrote += inbox;
        }
        printf("store variant %d: box at (%d,%d) %dx%d: mismatches %ld (box elements in bounds %ld)\n", variant, lx, ly, sbw, sbh, bad, wrote);
        return bad ? 1 : 0;
    }
    printf("unknown test %s\n", test.c_str());
    return 2;
}
