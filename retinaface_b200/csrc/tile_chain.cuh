// tile_chain.cuh -- persistent, warp-specialised tcgen05 "tile chain" kernel of librf_b200 (FP16 operands, FP32 accumulate).
// sm_100a only: cp.async.bulk.tensor (TMA tensor maps, SASS UTMALDG / UTMASTG), tcgen05.mma / ld / st (UTCHMMA, LDTM, STTM),
// mbarrier pipelines, griddepcontrol.
//
// One CTA owns a tile = TH full-width rows of one image and runs a CHAIN of layers on it without leaving the SM:
// every intermediate activation stays in shared memory (or TMEM), only the tensors other kernels need are written back,
// by TMA stores.  Layers the reference's graph (model/mnet-deconv-0517.prototxt) runs one by one inside TensorRT
// (retinaface/tensorrt/trtretinafacenet.cpp:60) become STAGES of one launch:
//   TCH_DWPW : depthwise 3x3 (stride 1 | 2) + BN + ReLU  ->  pointwise 1x1 + BN + ReLU   (mobilenet0_conv3 .. conv24)
//   TCH_CONV : 1x1 or 3x3 (pad 1) convolution + BN (+ ReLU)                                (laterals, rf_c*_aggr, SSH convs)
//   TCH_HEAD : the three predictor 1x1 convs of a level as ONE N = 32 GEMM (FP32-grade: hi + lo FP16 weight pieces) whose
//              epilogue is the post-process: 2-way softmax, threshold, anchor decode, clip, candidate append
//              (RetinaFace.cpp:666-723) -- and, in the last CTA that finishes an image, sort + greedy NMS (:434-492).
//   (pre-stage) FPN merge: lateral + crop(deconv_k4s2p1(coarser level)) (prototxt:1553-1592, :1948-1987) into the input tile.
//
// Geometry.  Local position space of a tile: width Wl = W + 2 (one zero column either side), rows = tile rows plus the
// halo the chain needs; position p = ly * Wl + lx.  In this space a 3x3 tap is the constant shift dy * Wl + dx, so the A
// operand of every tap is the SAME shared-memory buffer with the descriptor start address moved by shift * ROW bytes
// (K-major swizzled layouts swizzle on absolute address bits, so any row shift is legal: tools/umma_probe.cu).  A stage
// computes whole local rows, 128 consecutive positions per MMA tile; positions outside the image are written as zeros
// (they are the next stage's padding).  Halo rows are recomputed per tile.
//
// Data movement.  Activations enter by TMA tiled loads (4-D NHWC tensor map, box {<=64 channels, Wl, rows}, out-of-bounds
// fill = the zero padding) in SWIZZLE_128B / 64B / 32B mode (64 / 32 / 16 channels per row), which is exactly the UMMA
// K-major swizzled operand layout; stride-2 depthwise inputs arrive as four parity planes (tensor-map element strides 2).
// Epilogue threads write stage outputs into shared memory in the same swizzled layout (thread = position).
//
// Depthwise on tensor cores.  out[p][c] = sum_t in[p + shift_t][c] * w_t[c] is 9 * C/16 MMAs with M = 128, N = K = 16 and
// B = diag(w_t[16 channels]) accumulating into TMEM columns [16 s, 16 s + 16): no CUDA-core instruction per tap, but each
// such MMA occupies the tensor pipe for 37.7 cycles, not the 8 its math needs -- an SS-mode MMA re-reads its 4 KB A operand
// from shared memory whatever N is (tools/umma_probe.cu rate2: SS 37.7 / 46 / 61 cycles at N = 16 / 64 / 128, TS N/2), so a
// 64-channel depthwise stage is tensor-pipe bound at 36 x 37.7 = 1357 cycles per 128 positions (the chains' traces agree:
// DESIGN.md section 3).  The mid-epilogue (TMEM -> + bias, ReLU, FP16) writes the result back to TMEM
// in place as packed FP16 and the pointwise GEMM takes its A operand from TMEM (.kind::f16 with [a_tmem]).
//
// Roles (320 threads): warp 0 = TMA producer (tiles, per-stage weights, TMA stores), warp 1 = MMA issuer (one elected
// lane), warps 2-5 / 6-9 = two epilogue warpgroups, each bound to one of two TMEM accumulator sets so the MMAs of tile
// m+1 run while tile m is drained.  All hand-offs are mbarriers; the CTA loops over tiles (persistent).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "postproc_dev.cuh"
#include "tc_conv.cuh"

namespace rf {

constexpr int TCH_THREADS = 320;
constexpr int TCH_MAX_STAGES = 8;
constexpr int TCH_MAX_BUFS = 8;
constexpr int TCH_EPI_THREADS = 256;
enum { TCH_CONV = 0, TCH_DWPW = 1, TCH_HEAD = 2 };

struct TchBuf {
    int off;            // byte offset in dynamic shared memory (1024-aligned); position index 0 lives here
    int row;            // bytes per position per slab: 32 | 64 | 128 (16 | 32 | >= 64 channels)
    int slabs;          // 64-channel slabs (1 unless channels >= 128)
    int slab_stride;    // bytes between slabs
    int rows_lo;        // local row held at position index `slack`
    int nrows;          // rows held
    int slack;          // positions in front of row rows_lo (>= 1: the (-1, -1) tap of the first computed position reads one back)
};

struct TchStage {
    int type;
    int Cin, N;               // DWPW: depthwise channels / pointwise outputs; CONV, HEAD: input / output channels
    int taps, stride;         // CONV: 1 | 9;  DWPW: stride 1 | 2 (2: stage 0 only, input = parity planes)
    int in_buf;
    int rows_lo, nrows;       // local rows this stage computes
    int wd_off, wd_bytes;     // depthwise diagonal B tiles in the weight arena
    int wp_off, wp_bytes;     // B image [K/8][N][8] (HEAD: hi image then lo image)
    int wd_smem, wp_smem;     // resident weights: this stage's own regions in shared memory
    int bias_dw, bias_pw;     // float offsets into the bias arena
    int store_buf, store_map; // -1 | buffer whose owned rows [HT, HT + TH) are TMA-stored once the stage is complete
    unsigned char ob_buf[16], ob_c16[16], ob_relu[16];   // per 16-column output block: buffer, channel offset / 16, ReLU
};

struct TchHead {              // TCH_HEAD epilogue + last-block NMS
    LevelDesc lv;
    PostBuffers pb;
    const PostParams *params;
    int net_w, net_h;
    int *done;                // [max_batch] tiles finished per image (all levels); the last one runs the NMS
    int expected;             // tiles per image over all levels; 0: no fused NMS
    int nms_smem;             // byte offset of the NMS scratch in dynamic shared memory (never a TMA-stored buffer)
    float *blobs[3];          // rf_forward_heads: cls_prob / bbox_pred / landmark_pred of this level (NCHW f32) or NULL
};

struct TchArgs {
    int nstages, nbufs;
    TchStage st[TCH_MAX_STAGES];
    TchBuf buf[TCH_MAX_BUFS];
    int Wl, HT, TH;
    int W, H, nimg, tiles_per_img, ntiles;
    int nsets, set_cols;
    int wd_smem, wp_smem, bias_smem, smem_bytes;     // byte offsets in dynamic shared memory / total
    int resident;             // 1: the weights of ALL stages stay in shared memory for the CTA's lifetime (loaded once, before
                              // griddepcontrol.wait); 0: streamed per stage through the wd / wp buffers
    unsigned *dbg;            // host-mapped word: code of the hand-off a timed-out wait was stuck on (0: none)
    unsigned long long *trace;   // -DRF_TCH_TRACE builds: timeline of CTA 0 (count, then (code, ns) pairs)
    const unsigned char *warena;
    const float *bias;
    int bias_floats;
    int in_s2, plane_stride, in_C;
    unsigned in_bytes;
    // FPN merge pre-stage (merge_C > 0): buffer 0 += crop(deconv(coarse)); coarse tile in buffer `merge_buf`
    int merge_C, merge_buf, merge_rows, merge_w_bias;   // merge_w_bias: float offset of the [16 taps][C] FP16 weights in the bias arena
    unsigned merge_bytes;
    TchHead head;
};

struct TchMaps {
    CUtensorMap in;           // chain input (stride 2: element strides {1, 2, 2, 1})
    CUtensorMap aux;          // FPN merge: the coarser level
    CUtensorMap st[3];        // TMA stores
};

namespace tch {

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
// mbarrier wait that reports WHICH hand-off was lost before trapping (code -> host-mapped debug word): a lost arrive must
// fail loudly and say where, never hang the GPU
__device__ __forceinline__ void wait(uint64_t *bar, unsigned parity, unsigned *dbg, unsigned code) {
    unsigned done = 0, spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(tc::smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 22)) {
            if (dbg) { *reinterpret_cast<volatile unsigned *>(dbg) = code | (blockIdx.x << 20); __threadfence_system(); __nanosleep(2000000); }
            __trap();
        }
    }
}
// the same for a whole warp: lanes may leave the polling loop in different iterations; the warp-collective (.sync.aligned)
// tcgen05 instructions that follow need it converged again
__device__ __forceinline__ void wait_warp(uint64_t *bar, unsigned parity, unsigned *dbg, unsigned code) {
    wait(bar, parity, dbg, code);
    __syncwarp();
}
// Optional timeline of CTA 0 (build with -DRF_TCH_TRACE; tools/tile_bringup.py --trace): (event code, globaltimer ns) pairs
#ifdef RF_TCH_TRACE
__device__ __forceinline__ void trace(unsigned long long *t, unsigned code) {
    if (t && blockIdx.x == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        const unsigned i = atomicAdd(reinterpret_cast<unsigned *>(t), 1u);
        if (i < 500) { t[1 + 2 * i] = code; t[2 + 2 * i] = now; }
    }
}
#define TCH_TRACE(code) tch::trace(a.trace, (code))
#else
#define TCH_TRACE(code)
#endif
// one lane of a converged warp (elect.sync): ptxas then knows the guarded region runs with a single active thread and moves
// the MMA operands to uniform registers without a waterfall loop
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_map(const CUtensorMap *map) { asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_u32(uint32_t dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(tc::smem_u32(bar))
                 : "memory");
}
// K-major swizzled operand descriptor: rows of `row` bytes (128 / 64 / 32 -> layout 2 / 4 / 6), 8-row groups SBO = 8 * row
__device__ __forceinline__ uint64_t sw_desc(uint32_t addr, uint32_t row) {
    const uint64_t layout = row == 128 ? 2 : (row == 64 ? 4 : 6);
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(((8 * row) >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}
// Descriptor words.  A shared-memory matrix descriptor is {lo: (addr >> 4) [0,14) | LBO >> 4 [16,30)} {hi: SBO >> 4 [0,14) |
// version 1 [14] | layout [29,32)}: moving the operand by a multiple of 16 bytes is ONE 32-bit add on the low word (shared
// memory is < 256 KB, the address field cannot carry out), so the issue loops below precompute every tap / K-step offset.
__device__ __forceinline__ uint32_t a_desc_hi(uint32_t row) {
    const uint32_t layout = row == 128 ? 2u : (row == 64 ? 4u : 6u);
    return ((8u * row) >> 4) | (1u << 14) | (layout << 29);
}
__device__ __forceinline__ uint32_t a_desc_lo(uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint32_t b_desc_hi() { return (128u >> 4) | (1u << 14); }                       // K-major, no swizzle, SBO 128
__device__ __forceinline__ uint32_t b_desc_lo(uint32_t addr, uint32_t lbo) { return ((addr >> 4) & 0x3FFFu) | ((lbo >> 4) << 16); }
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// The issuing thread runs DEPENDENT scalar instructions at ~5 cycles each (tools/umma_probe.cu "rate": a loop that computes its
// descriptors per MMA issues one MMA per ~200 cycles, whatever N): the issue blocks below are fully unrolled over taps and K
// steps with every operand offset precomputed, so that each MMA costs two independent adds and the instruction itself.
template <bool ACC>
__device__ __forceinline__ void mma_ss_c(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc) {
    if (ACC)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.eq.u32 p, 0, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.u32 p, 0, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
template <bool ACC>
__device__ __forceinline__ void mma_ts_c(uint32_t d_tmem, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc) {
    if (ACC)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\tsetp.eq.u32 p, 0, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.u32 p, 0, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
// operand geometry of one stage, in 16-byte units (see the MMA issuer)
struct IssueGeo {
    uint32_t ahi, bhi, slabq, idesc;
    int lkpr, kpr;
};
// 3x3 convolution: 9 taps x NK K-steps into one accumulator; B image tap-major: tap t at t * tapb, K-step k at k * kb
template <int NK>
__device__ __forceinline__ void issue_conv9(uint32_t d0, uint32_t am, const uint32_t (&tapq)[9], uint32_t wplo0, uint32_t tapb, uint32_t kb, const IssueGeo &g) {
    uint32_t ka[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) ka[k] = (uint32_t)(k >> g.lkpr) * g.slabq + (uint32_t)(k & (g.kpr - 1)) * 2u;
#pragma unroll
    for (int t = 0; t < 9; t++) {
        const uint32_t at = am + tapq[t], bt = wplo0 + (uint32_t)t * tapb;
#pragma unroll
        for (int k = 0; k < NK; k++) {
            if (t == 0 && k == 0) mma_ss_c<false>(d0, at + ka[k], g.ahi, bt + (uint32_t)k * kb, g.bhi, g.idesc);
            else mma_ss_c<true>(d0, at + ka[k], g.ahi, bt + (uint32_t)k * kb, g.bhi, g.idesc);
        }
    }
}
// 1x1 convolution / predictors: NK K-steps x PIECES weight pieces (hi, lo) into one accumulator
template <int NK, int PIECES>
__device__ __forceinline__ void issue_conv1(uint32_t d0, uint32_t am, uint32_t wplo0, uint32_t pieceq, uint32_t kb, const IssueGeo &g) {
#pragma unroll
    for (int k = 0; k < NK; k++) {
        const uint32_t ak = am + (uint32_t)(k >> g.lkpr) * g.slabq + (uint32_t)(k & (g.kpr - 1)) * 2u;
#pragma unroll
        for (int pc = 0; pc < PIECES; pc++) {
            if (k == 0 && pc == 0) mma_ss_c<false>(d0, ak, g.ahi, wplo0 + (uint32_t)pc * pieceq + (uint32_t)k * kb, g.bhi, g.idesc);
            else mma_ss_c<true>(d0, ak, g.ahi, wplo0 + (uint32_t)pc * pieceq + (uint32_t)k * kb, g.bhi, g.idesc);
        }
    }
}
// depthwise: tap-major so that consecutive MMAs accumulate into DIFFERENT 16-column slabs; diagonal tile of (t, k) at (t * nk + k) * 32
template <int NK>
__device__ __forceinline__ void issue_dw9(uint32_t d0, uint32_t am, const uint32_t (&tapq)[9], uint32_t wdlo0, int nk, int k0, const IssueGeo &g) {
    uint32_t ka[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) ka[k] = (uint32_t)((k0 + k) >> g.lkpr) * g.slabq + (uint32_t)((k0 + k) & (g.kpr - 1)) * 2u;
#pragma unroll
    for (int t = 0; t < 9; t++) {
        const uint32_t at = am + tapq[t], bt = wdlo0 + (uint32_t)(t * nk + k0) * 32u;
#pragma unroll
        for (int k = 0; k < NK; k++) {
            if (t == 0) mma_ss_c<false>(d0 + (uint32_t)(k0 + k) * 16u, at + ka[k], g.ahi, bt + (uint32_t)k * 32u, g.bhi, g.idesc);
            else mma_ss_c<true>(d0 + (uint32_t)(k0 + k) * 16u, at + ka[k], g.ahi, bt + (uint32_t)k * 32u, g.bhi, g.idesc);
        }
    }
}
// pointwise with A from TMEM: NK K-steps into one accumulator
template <int NK>
__device__ __forceinline__ void issue_pw_ts(uint32_t dacc, uint32_t a_tmem, uint32_t wplo0, uint32_t kb, uint32_t bhi, uint32_t idesc) {
#pragma unroll
    for (int k = 0; k < NK; k++) {
        if (k == 0) mma_ts_c<false>(dacc, a_tmem, wplo0, bhi, idesc);
        else mma_ts_c<true>(dacc, a_tmem + (uint32_t)k * 8u, wplo0 + (uint32_t)k * kb, bhi, idesc);
    }
}
// D[tmem] (+)= A[tmem, packed FP16] * B[smem desc]
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t r[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(TCH_EPI_THREADS) : "memory"); }
__device__ __forceinline__ uint32_t idesc_f16(int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
// byte offset of 16-byte chunk j of position p inside a buffer (swizzle on address bits [4,7) ^ [7,10), buffers are 1024-aligned)
__device__ __forceinline__ uint32_t chunk_off(int row, int p, int j) {
    const int sh = row == 128 ? 0 : (row == 64 ? 1 : 2), m = (row >> 4) - 1;
    return (uint32_t)p * row + (uint32_t)((j ^ ((p >> sh) & m)) << 4);
}
__device__ __forceinline__ void tmem_alloc_n(uint32_t *slot, int cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_n(uint32_t addr, int cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

}  // namespace tch

__host__ __device__ inline int tch_tmem_cols(int n) { return n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 128 ? 128 : (n <= 256 ? 256 : 512))); }

template <int UNUSED>
__global__ void __launch_bounds__(TCH_THREADS, 2) k_tile_chain(const __grid_constant__ TchMaps maps, const __grid_constant__ TchArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar_in, bar_bias, bar_merge, bar_stage, bar_tile;
    __shared__ __align__(8) uint64_t bar_wd_full, bar_wd_empty, bar_wp_full, bar_wp_empty;
    __shared__ __align__(8) uint64_t bar_dw_full[2], bar_a16_full[2], bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t s_tmem;
    __shared__ int s_last;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *const smem = smem_raw + (sbase - tc::smem_u32(smem_raw));
    const int Wl = a.Wl;
    const int tmem_cols = tch_tmem_cols(a.nsets * a.set_cols);

    if (tid == 0) {
        tc::mbar_init(&bar_in, 1); tc::mbar_init(&bar_bias, 1); tc::mbar_init(&bar_merge, 8); tc::mbar_init(&bar_stage, 8); tc::mbar_init(&bar_tile, 8);
        tc::mbar_init(&bar_wd_full, 1); tc::mbar_init(&bar_wd_empty, 1); tc::mbar_init(&bar_wp_full, 1); tc::mbar_init(&bar_wp_empty, 1);
        for (int i = 0; i < 2; i++) {
            tc::mbar_init(&bar_dw_full[i], 1); tc::mbar_init(&bar_a16_full[i], 4);
            tc::mbar_init(&bar_acc_full[i], 1); tc::mbar_init(&bar_acc_empty[i], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tch::tmem_alloc_n(&s_tmem, tmem_cols);
    pdl_trigger();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (tid == 0) TCH_TRACE(1);

    if (warp == 0) {
        // =========================================== TMA producer ===========================================
        if (lane == 0) {
            const unsigned bias_bytes = (unsigned)a.bias_floats * 4u;
            unsigned const_bytes = bias_bytes;
            if (a.resident) for (int s = 0; s < a.nstages; s++) const_bytes += (unsigned)(a.st[s].wd_bytes + a.st[s].wp_bytes);
            tc::mbar_expect_tx(&bar_bias, const_bytes);
            tch::bulk_g2s_u32(sbase + a.bias_smem, a.bias, bias_bytes, &bar_bias);       // constants: independent of earlier kernels
            if (a.resident)
                for (int s = 0; s < a.nstages; s++) {
                    if (a.st[s].wd_bytes) tch::bulk_g2s_u32(sbase + a.st[s].wd_smem, a.warena + a.st[s].wd_off, (unsigned)a.st[s].wd_bytes, &bar_bias);
                    tch::bulk_g2s_u32(sbase + a.st[s].wp_smem, a.warena + a.st[s].wp_off, (unsigned)a.st[s].wp_bytes, &bar_bias);
                }
            tch::prefetch_map(&maps.in);
            pdl_wait();                                                                   // activations of earlier kernels from here on
            unsigned wdc = 0, wpc = 0, sc = 0;
            const TchBuf &B0 = a.buf[0];
            // owned rows of a finished buffer -> global, one box {<= 64 channels, W, 1} per row and slab starting at image column 0
            // (TMA stores reject negative coordinates -- tools/umma_probe.cu "store" -- so the zero column at lx = 0 stays behind)
            auto store_rows = [&](const TchBuf &BS, int map, int ty, int b) {
                for (int rr = 0; rr < a.TH; rr++) {
                    const int y = ty * a.TH + rr;
                    if (y >= a.H) break;
                    for (int k = 0; k < BS.slabs; k++)
                        tch::tma_store_4d(&maps.st[map], sbase + BS.off + k * BS.slab_stride + (BS.slack + (a.HT - BS.rows_lo + rr) * Wl + 1) * BS.row, k * 64, 0, y, b);
                }
                tch::bulk_commit();
            };
            for (int tile = blockIdx.x, it = 0; tile < a.ntiles; tile += gridDim.x, it++) {
                const int b = tile / a.tiles_per_img, ty = tile - b * a.tiles_per_img;
                const int Y0 = ty * a.TH - a.HT;                 // image row of local row 0; image column of local column 0 is -1
                if (it > 0) { tch::wait(&bar_tile, (it - 1) & 1, a.dbg, __LINE__); tch::bulk_wait_read0(); }   // buffers free, stores have read them
                tc::mbar_expect_tx(&bar_in, a.in_bytes + a.merge_bytes);
                const int in_slabs = (a.in_C + 63) >> 6;
                if (!a.in_s2) {
                    for (int s = 0; s < in_slabs; s++)
                        tch::tma_load_4d(sbase + B0.off + s * B0.slab_stride + B0.slack * B0.row, &maps.in, &bar_in, s * 64, -1, Y0 + B0.rows_lo, b);
                } else {
                    // plane (py, px) element (pr, lx) = input(2 * (Y0 + rows_lo0 - 1 + pr) + py, 2 * (lx - 1) + px)
                    for (int k = 0; k < 4; k++)
                        for (int s = 0; s < in_slabs; s++)
                            tch::tma_load_4d(sbase + B0.off + k * a.plane_stride + s * B0.slab_stride + B0.slack * B0.row, &maps.in, &bar_in, s * 64,
                                             -2 + (k & 1), 2 * (Y0 + a.st[0].rows_lo - 1) + (k >> 1), b);
                }
                if (a.merge_C) {
                    const TchBuf &BM = a.buf[a.merge_buf];
                    // coarse rows ((Y0 + rows_lo + 1) >> 1) - 1 ..., coarse columns -1 .. W/2
                    const int cy0 = ((Y0 + B0.rows_lo + 1) >> 1) - 1;
                    tch::tma_load_4d(sbase + BM.off, &maps.aux, &bar_in, 0, -1, cy0, b);
                }
                for (int s = 0; s < a.nstages; s++, sc++) {
                    const TchStage &st = a.st[s];
                    if (!a.resident) {
                        if (st.wd_bytes) {
                            tch::wait(&bar_wd_empty, (wdc & 1) ^ 1, a.dbg, __LINE__);
                            tc::mbar_expect_tx(&bar_wd_full, (unsigned)st.wd_bytes);
                            tch::bulk_g2s_u32(sbase + a.wd_smem, a.warena + st.wd_off, (unsigned)st.wd_bytes, &bar_wd_full);
                            wdc++;
                        }
                        tch::wait(&bar_wp_empty, (wpc & 1) ^ 1, a.dbg, __LINE__);
                        tc::mbar_expect_tx(&bar_wp_full, (unsigned)st.wp_bytes);
                        tch::bulk_g2s_u32(sbase + a.wp_smem, a.warena + st.wp_off, (unsigned)st.wp_bytes, &bar_wp_full);
                        wpc++;
                    }
                    // every phase of bar_stage is observed in order; stage s-1 is complete -> its TMA store
                    if (s > 0) {
                        tch::wait(&bar_stage, (sc - 1) & 1, a.dbg, __LINE__);
                        const TchStage &sp = a.st[s - 1];
                        if (sp.store_buf >= 0) store_rows(a.buf[sp.store_buf], sp.store_map, ty, b);
                    }
                }
                tch::wait(&bar_stage, (sc - 1) & 1, a.dbg, __LINE__);
                {
                    const TchStage &sp = a.st[a.nstages - 1];
                    if (sp.store_buf >= 0) store_rows(a.buf[sp.store_buf], sp.store_map, ty, b);
                }
            }
            tch::bulk_wait0();       // global writes of the last stores are complete before the CTA exits
        }
    } else if (warp == 1) {
        // =========================================== MMA issuer =============================================
        unsigned wdc = 0, wpc = 0, sc = 0, g = 0;
        unsigned use[2] = {0, 0}, dwuse[2] = {0, 0};
        if (a.resident) tch::wait_warp(&bar_bias, 0, a.dbg, __LINE__);       // every stage's weights are in shared memory
        for (int tile = blockIdx.x, it = 0; tile < a.ntiles; tile += gridDim.x, it++) {
            tch::wait_warp(&bar_in, it & 1, a.dbg, __LINE__);
            if (lane == 0) TCH_TRACE(3);
            if (a.merge_C) tch::wait_warp(&bar_merge, it & 1, a.dbg, __LINE__);
            for (int s = 0; s < a.nstages; s++, sc++) {
                const TchStage &st = a.st[s];
                const TchBuf &BI = a.buf[st.in_buf];
                if (sc > 0) tch::wait_warp(&bar_stage, (sc - 1) & 1, a.dbg, __LINE__);      // inputs of this stage are in shared memory
                if (lane == 0) TCH_TRACE(100 + s);
                if (!a.resident) {
                    if (st.wd_bytes) { tch::wait_warp(&bar_wd_full, wdc & 1, a.dbg, __LINE__); wdc++; }
                    tch::wait_warp(&bar_wp_full, wpc & 1, a.dbg, __LINE__); wpc++;
                }
                tc::tc_fence_after();
                const int npos = st.nrows * Wl, ntile = (npos + 127) >> 7;
                const int kpr = BI.row >> 5;                             // 16-channel K steps per row: 1 | 2 | 4
                const int lkpr = kpr == 4 ? 2 : (kpr == 2 ? 1 : 0);
                // position index (in the input buffer) of this stage's position 0
                const int pos0 = st.type == TCH_DWPW && st.stride == 2 ? BI.slack : BI.slack + (st.rows_lo - BI.rows_lo) * Wl;
                const uint32_t rowq = (uint32_t)BI.row >> 4;              // 16-byte units per position
                const uint32_t ahi = tch::a_desc_hi(BI.row), bhi = tch::b_desc_hi();
                const uint32_t alo0 = tch::a_desc_lo(sbase + BI.off) + (uint32_t)pos0 * rowq;
                const uint32_t tile_step = 128u * rowq;                   // one MMA tile further
                const uint32_t slabq = (uint32_t)BI.slab_stride >> 4;
                const uint32_t wp_addr = sbase + (a.resident ? st.wp_smem : a.wp_smem), wd_addr = sbase + (a.resident ? st.wd_smem : a.wd_smem);
                if (st.type == TCH_DWPW) {
                    const int nk = st.Cin >> 4;
                    const uint32_t idesc16 = tch::idesc_f16(16), idescN = tch::idesc_f16(st.N);
                    // per tap: operand shift (+ parity plane for stride 2), in 16-byte units
                    uint32_t tapq[9];
#pragma unroll
                    for (int t = 0; t < 9; t++) {
                        const int dy = t / 3 - 1, dx = t % 3 - 1;
                        int shift, plane = 0;
                        if (st.stride == 2) { plane = ((dy & 1) << 1) | (dx & 1); shift = (dy >= 0 ? Wl : 0) + (dx < 0 ? -1 : 0); }
                        else shift = dy * Wl + dx;
                        tapq[t] = (uint32_t)(shift * (int)rowq + plane * (a.plane_stride >> 4));
                    }
                    const uint32_t wdlo0 = tch::b_desc_lo(wd_addr, 256), wplo0 = tch::b_desc_lo(wp_addr, (uint32_t)st.N * 16);
                    auto issue_dw = [&](int m) {
                        const int set = a.nsets == 2 ? ((g + m) & 1) : 0;
                        tch::wait_warp(&bar_acc_empty[set], (use[set] & 1) ^ 1, a.dbg, __LINE__);
                        use[set]++;
                        tc::tc_fence_after();
                        if (tch::elect_one()) {
                            const uint32_t d0 = tmem + set * a.set_cols;
                            const uint32_t am = alo0 + (uint32_t)m * tile_step;
                            const tch::IssueGeo geo{ahi, bhi, slabq, idesc16, lkpr, kpr};
                            switch (nk) {
                                case 1: tch::issue_dw9<1>(d0, am, tapq, wdlo0, nk, 0, geo); break;
                                case 2: tch::issue_dw9<2>(d0, am, tapq, wdlo0, nk, 0, geo); break;
                                case 4: tch::issue_dw9<4>(d0, am, tapq, wdlo0, nk, 0, geo); break;
                                default: for (int k0 = 0; k0 < nk; k0 += 8) tch::issue_dw9<8>(d0, am, tapq, wdlo0, nk, k0, geo); break;     // nk = 8, 16
                            }
                            tc::mma_commit(&bar_dw_full[set]);
                            if (m == ntile - 1 && !a.resident) tc::mma_commit(&bar_wd_empty);
                        }
                        __syncwarp();
                    };
                    auto issue_pw = [&](int m) {
                        const int set = a.nsets == 2 ? ((g + m) & 1) : 0;
                        tch::wait_warp(&bar_a16_full[set], dwuse[set] & 1, a.dbg, __LINE__);
                        dwuse[set]++;
                        tc::tc_fence_after();
                        if (tch::elect_one()) {
                            const uint32_t d0 = tmem + set * a.set_cols;
                            const uint32_t kb = 2u * (uint32_t)st.N;
                            switch (nk) {
                                case 1: tch::issue_pw_ts<1>(d0 + st.Cin, d0, wplo0, kb, bhi, idescN); break;
                                case 2: tch::issue_pw_ts<2>(d0 + st.Cin, d0, wplo0, kb, bhi, idescN); break;
                                case 4: tch::issue_pw_ts<4>(d0 + st.Cin, d0, wplo0, kb, bhi, idescN); break;
                                case 8: tch::issue_pw_ts<8>(d0 + st.Cin, d0, wplo0, kb, bhi, idescN); break;
                                default: for (int k = 0; k < nk; k++) tch::mma_ts(d0 + st.Cin, d0 + k * 8, wplo0 + (uint32_t)k * kb, bhi, idescN, k > 0); break;
                            }
                            tc::mma_commit(&bar_acc_full[set]);
                            if (m == ntile - 1 && !a.resident) tc::mma_commit(&bar_wp_empty);
                        }
                        __syncwarp();
                    };
                    if (a.nsets == 2) {
                        for (int m = 0; m <= ntile; m++) {
                            if (m < ntile) issue_dw(m);
                            if (m > 0) issue_pw(m - 1);
                        }
                    } else {
                        for (int m = 0; m < ntile; m++) { issue_dw(m); issue_pw(m); }
                    }
                } else {
                    const int nk = st.Cin >> 4;
                    const uint32_t idescN = tch::idesc_f16(st.N);
                    const int pieces = st.type == TCH_HEAD ? 2 : 1;
                    const uint32_t pieceq = ((uint32_t)st.taps * st.Cin * st.N * 2) >> 4;
                    const uint32_t wplo0 = tch::b_desc_lo(wp_addr, (uint32_t)st.N * 16);
                    const uint32_t tapb = (uint32_t)(st.Cin >> 3) * st.N;      // B image: 16-byte units per tap
                    uint32_t tapc[9];
#pragma unroll
                    for (int t = 0; t < 9; t++) tapc[t] = (uint32_t)(((t / 3 - 1) * Wl + (t % 3 - 1)) * (int)rowq);
                    for (int m = 0; m < ntile; m++) {
                        const int set = a.nsets == 2 ? ((g + m) & 1) : 0;
                        tch::wait_warp(&bar_acc_empty[set], (use[set] & 1) ^ 1, a.dbg, __LINE__);
                        use[set]++;
                        tc::tc_fence_after();
                        if (tch::elect_one()) {
                            const uint32_t d0 = tmem + set * a.set_cols;
                            const uint32_t am = alo0 + (uint32_t)m * tile_step;
                            const tch::IssueGeo geo{ahi, bhi, slabq, idescN, lkpr, kpr};
                            const uint32_t kb = 2u * (uint32_t)st.N;
                            if (st.taps == 9) {
                                switch (nk) {
                                    case 1: tch::issue_conv9<1>(d0, am, tapc, wplo0, tapb, kb, geo); break;
                                    case 2: tch::issue_conv9<2>(d0, am, tapc, wplo0, tapb, kb, geo); break;
                                    default: tch::issue_conv9<4>(d0, am, tapc, wplo0, tapb, kb, geo); break;       // 64 input channels
                                }
                            } else if (pieces == 2) {
                                tch::issue_conv1<4, 2>(d0, am, wplo0, pieceq, kb, geo);                               // predictors: 64 channels, hi + lo
                            } else {
                                switch (nk) {
                                    case 4: tch::issue_conv1<4, 1>(d0, am, wplo0, pieceq, kb, geo); break;
                                    case 8: tch::issue_conv1<8, 1>(d0, am, wplo0, pieceq, kb, geo); break;
                                    default: {
                                        uint32_t acc = 0;
                                        for (int k = 0; k < nk; k++) {
                                            tch::mma_ss(d0, am + (uint32_t)(k >> lkpr) * slabq + (uint32_t)(k & (kpr - 1)) * 2u, ahi, wplo0 + (uint32_t)k * kb, bhi, idescN, acc);
                                            acc = 1;
                                        }
                                    }
                                }
                            }
                            tc::mma_commit(&bar_acc_full[set]);
                            if (m == ntile - 1 && !a.resident) tc::mma_commit(&bar_wp_empty);
                        }
                        __syncwarp();
                    }
                }
                g += ntile;
                if (lane == 0) TCH_TRACE(200 + s);
            }
        }
    } else {
        // =========================================== epilogue warpgroups ====================================
        const int wg = (warp - 2) >> 2, quad = warp & 3;
        const int r = quad * 32 + lane;                         // GEMM row = TMEM lane
        const int etid = tid - 64;                              // 0..255 over both warpgroups
        const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
        const float *s_bias = reinterpret_cast<const float *>(smem + a.bias_smem);
        tch::wait_warp(&bar_bias, 0, a.dbg, __LINE__);
        pdl_wait();
        unsigned g = 0, sc = 0, ca[2] = {0, 0}, cd[2] = {0, 0};
        for (int tile = blockIdx.x, it = 0; tile < a.ntiles; tile += gridDim.x, it++) {
            const int b = tile / a.tiles_per_img, ty = tile - b * a.tiles_per_img;
            const int Y0 = ty * a.TH - a.HT;
            if (a.merge_C) {
                // ---- FPN merge: buffer 0 (lateral) += crop(deconv_k4s2p1(coarse)) at in-image positions; packed HFMA2, the
                //      sum of <= 5 terms is stored as FP16 anyway (same operation order as k_fpn_merge_h2)
                tch::wait_warp(&bar_in, it & 1, a.dbg, __LINE__);
                const TchBuf &B0 = a.buf[0], &BM = a.buf[a.merge_buf];
                const int UH = a.H >> 1, UW = a.W >> 1, CW = UW + 2;
                const int cy0 = ((Y0 + B0.rows_lo + 1) >> 1) - 1;
                const __half *uw = reinterpret_cast<const __half *>(s_bias + a.merge_w_bias);     // [16 taps][64]
                const int items = B0.nrows * Wl * 8;
                for (int i = etid; i < items; i += TCH_EPI_THREADS) {
                    const int j = i & 7, p = i >> 3;
                    const int ly = p / Wl, lx = p - ly * Wl;
                    const int y = Y0 + B0.rows_lo + ly, x = lx - 1;
                    if (x < 0 || x >= a.W || y < 0 || y >= a.H) continue;
                    unsigned char *slot = smem + B0.off + tch::chunk_off(128, B0.slack + p, j);
                    uint4 accv = *reinterpret_cast<const uint4 *>(slot);
                    __half2 *acc = reinterpret_cast<__half2 *>(&accv);
                    const int i_hi = (y + 1) >> 1, j_hi = (x + 1) >> 1;
#pragma unroll
                    for (int di = 0; di < 2; di++) {
                        const int ci = i_hi - di, ky = y - 2 * ci + 1;
                        if (ci < 0 || ci >= UH) continue;
#pragma unroll
                        for (int dj = 0; dj < 2; dj++) {
                            const int cj = j_hi - dj, kx = x - 2 * cj + 1;
                            if (cj < 0 || cj >= UW) continue;
                            const int cp = (ci - cy0) * CW + (cj + 1);
                            const uint4 uv = *reinterpret_cast<const uint4 *>(smem + BM.off + tch::chunk_off(128, cp, j));
                            const uint4 wv = *reinterpret_cast<const uint4 *>(uw + (ky * 4 + kx) * 64 + j * 8);
                            const __half2 *u2 = reinterpret_cast<const __half2 *>(&uv), *w2 = reinterpret_cast<const __half2 *>(&wv);
#pragma unroll
                            for (int c = 0; c < 4; c++) acc[c] = __hfma2(u2[c], w2[c], acc[c]);
                        }
                    }
                    *reinterpret_cast<uint4 *>(slot) = accv;
                }
                tc::fence_async_smem();
                __syncwarp();
                if (lane == 0) tch::mbar_arrive(&bar_merge);
            }
            for (int s = 0; s < a.nstages; s++, sc++) {
                const TchStage &st = a.st[s];
                const int npos = st.nrows * Wl, ntile = (npos + 127) >> 7;
                // a warp with no tile in this stage must not arrive for stage sc before phase sc-1 of bar_stage has completed
                if (sc > 0) tch::wait_warp(&bar_stage, (sc - 1) & 1, a.dbg, __LINE__);
                for (int m = 0; m < ntile; m++) {
                    const int set = a.nsets == 2 ? ((g + m) & 1) : 0;
                    const unsigned pa = ca[set] & 1, pd = cd[set] & 1;
                    ca[set]++;
                    if (st.type == TCH_DWPW) cd[set]++;
                    if (wg != set) continue;                   // one accumulator set per warpgroup (a single set: warpgroup 0 only)
                    const int q = m * 128 + r;
                    const bool valid = q < npos;
                    const int qy = q / Wl;
                    const int ly = st.rows_lo + qy, lx = q - qy * Wl;
                    const int gy = Y0 + ly, gx = lx - 1;
                    const bool inimg = valid && gx >= 0 && gx < a.W && gy >= 0 && gy < a.H;
                    const uint32_t t0 = lane_base + set * a.set_cols;
                    uint32_t acc_col = 0;
                    if (st.type == TCH_DWPW) {
                        // ---- mid-epilogue: depthwise accumulators -> + bias, ReLU -> packed FP16, back into TMEM in place
                        tch::wait_warp(&bar_dw_full[set], pd, a.dbg, __LINE__);
                        tc::tc_fence_after();
                        const float *bd = s_bias + st.bias_dw;
                        for (int k = 0; k < (st.Cin >> 4); k++) {
                            uint32_t v[16];
                            tc::tmem_ld16(t0 + k * 16, v);
                            tc::tmem_ld_wait();
                            uint32_t pk[8];
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                const float2 bb = *reinterpret_cast<const float2 *>(bd + k * 16 + 2 * i);
                                pk[i] = tch::pack_h2(fmaxf(__uint_as_float(v[2 * i]) + bb.x, 0.f), fmaxf(__uint_as_float(v[2 * i + 1]) + bb.y, 0.f));
                            }
                            tch::tmem_st8(t0 + k * 8, pk);
                        }
                        tch::tmem_st_wait();
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tch::mbar_arrive(&bar_a16_full[set]);
                        acc_col = st.Cin;
                    }
                    tch::wait_warp(&bar_acc_full[set], pa, a.dbg, __LINE__);
                    tc::tc_fence_after();
                    const float *bp = s_bias + st.bias_pw;
                    if (st.type != TCH_HEAD) {
                        for (int jb = 0; jb < (st.N >> 4); jb++) {
                            uint32_t v[16];
                            tc::tmem_ld16(t0 + acc_col + jb * 16, v);
                            tc::tmem_ld_wait();
                            const TchBuf &BO = a.buf[st.ob_buf[jb]];
                            const int prow = ly - BO.rows_lo;
                            if (valid && prow >= 0 && prow < BO.nrows) {
                                uint32_t pk[8];
                                const bool relu = st.ob_relu[jb] != 0;
#pragma unroll
                                for (int i = 0; i < 8; i++) {
                                    const float2 bb = *reinterpret_cast<const float2 *>(bp + jb * 16 + 2 * i);
                                    float f0 = __uint_as_float(v[2 * i]) + bb.x, f1 = __uint_as_float(v[2 * i + 1]) + bb.y;
                                    if (relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
                                    pk[i] = inimg ? tch::pack_h2(f0, f1) : 0u;
                                }
                                const int p = BO.slack + prow * Wl + lx;
                                const int c16 = st.ob_c16[jb];
                                unsigned char *base = smem + BO.off + (c16 >> 2) * BO.slab_stride;
                                const int j0 = (c16 & 3) * 2;
                                *reinterpret_cast<uint4 *>(base + tch::chunk_off(BO.row, p, j0)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                                *reinterpret_cast<uint4 *>(base + tch::chunk_off(BO.row, p, j0 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                            }
                        }
                    } else {
                        // ---- predictor GEMM -> softmax, threshold, decode, candidate append (postproc_dev.cuh)
                        const TchHead &hd = a.head;
                        uint32_t v[16], v2[16];
                        tc::tmem_ld16(t0, v);
                        tc::tmem_ld_wait();
                        const bool owned = inimg && ly >= a.HT && ly < a.HT + a.TH;
                        float sc4[4], pf[2], pbg[2];
#pragma unroll
                        for (int i = 0; i < 4; i++) sc4[i] = __fadd_rn(__uint_as_float(v[i]), bp[i]);
#pragma unroll
                        for (int an = 0; an < 2; an++) softmax_pair(sc4[an], sc4[an + 2], pbg[an], pf[an]);
                        const float thr = hd.params->score_thr;
                        const bool wb = hd.blobs[0] != nullptr;
                        const bool pass0 = owned && !(pf[0] <= thr), pass1 = owned && !(pf[1] <= thr);
                        const bool need = __any_sync(0xffffffffu, pass0 || pass1 || (wb && owned));
                        if (need) {
                            tc::tmem_ld16(t0 + 16, v2);
                            tc::tmem_ld_wait();
                            const int hw = hd.lv.h * hd.lv.w, jpix = gy * hd.lv.w + gx;
                            float reg[8], lm[20];
#pragma unroll
                            for (int i = 0; i < 8; i++) reg[i] = __fadd_rn(__uint_as_float(v[4 + i]), bp[4 + i]);
#pragma unroll
                            for (int i = 0; i < 4; i++) lm[i] = __fadd_rn(__uint_as_float(v[12 + i]), bp[12 + i]);
#pragma unroll
                            for (int i = 0; i < 16; i++) lm[4 + i] = __fadd_rn(__uint_as_float(v2[i]), bp[16 + i]);
                            if (wb && owned) {
                                float *cls = hd.blobs[0] + (size_t)b * 4 * hw, *bb = hd.blobs[1] + (size_t)b * 8 * hw, *lb = hd.blobs[2] + (size_t)b * 20 * hw;
                                cls[0 * hw + jpix] = pbg[0]; cls[1 * hw + jpix] = pbg[1]; cls[2 * hw + jpix] = pf[0]; cls[3 * hw + jpix] = pf[1];
#pragma unroll
                                for (int i = 0; i < 8; i++) bb[i * hw + jpix] = reg[i];
#pragma unroll
                                for (int i = 0; i < 20; i++) lb[i * hw + jpix] = lm[i];
                            }
#pragma unroll
                            for (int an = 0; an < 2; an++) {
                                if (!(an ? pass1 : pass0)) continue;
                                rf_det d;
                                decode_one(pf[an], reg + 4 * an, lm + 10 * an, hd.lv, an, gy, gx, hd.net_w, hd.net_h, hd.lv.anchor_base + an * hw + jpix, d);
                                append_candidate(hd.pb, b, d);
                            }
                        }
                    }
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tch::mbar_arrive(&bar_acc_empty[set]);
                }
                g += ntile;
                tc::fence_async_smem();          // this stage's shared-memory writes -> visible to UMMA / TMA (async proxy)
                __syncwarp();
                if (lane == 0) tch::mbar_arrive(&bar_stage);
            }
            if (a.head.expected > 0) {
                // ---- last-block NMS: the CTA that completes an image's last tile (over all three levels) sorts and suppresses it
                __threadfence();
                tch::epi_bar();
                if (etid == 0) s_last = atomicAdd(&a.head.done[b], 1) == a.head.expected - 1;
                tch::epi_bar();
                if (s_last) {
                    __threadfence();
                    NmsSmem &S = *reinterpret_cast<NmsSmem *>(smem + a.head.nms_smem);
                    int *s_kept = reinterpret_cast<int *>(smem + a.head.nms_smem + ((sizeof(NmsSmem) + 15) & ~15));
                    nms_image<TCH_EPI_THREADS, true>(b, etid, a.head.params->nms_thr, a.head.params, a.head.pb, S, s_kept, [] { tch::epi_bar(); });
                    if (etid == 0) a.head.done[b] = 0;           // self-cleaning for the next forward
                }
            }
            __syncwarp();
            if (lane == 0) tch::mbar_arrive(&bar_tile);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) TCH_TRACE(9);
    if (warp == 1) tch::tmem_dealloc_n(tmem, tmem_cols);
}

}  // namespace rf
