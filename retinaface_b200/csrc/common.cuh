// common.cuh -- shared device/host helpers of librf_b200 (sm_100a only).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rf_b200.h"

namespace rf {

// ---- element type helpers (activations are NHWC in T = float or __half) --------------------
template <typename T> struct Vec8;  // 8 consecutive channels
template <> struct Vec8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float *p) { a = *reinterpret_cast<const float4 *>(p); b = *reinterpret_cast<const float4 *>(p + 4); }
    __device__ __forceinline__ void store(float *p) const { *reinterpret_cast<float4 *>(p) = a; *reinterpret_cast<float4 *>(p + 4) = b; }
    __device__ __forceinline__ void to_float(float f[8]) const { f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w; }
    __device__ __forceinline__ void from_float(const float f[8]) { a = make_float4(f[0], f[1], f[2], f[3]); b = make_float4(f[4], f[5], f[6], f[7]); }
};
template <> struct Vec8<__half> {
    uint4 v;
    __device__ __forceinline__ void load(const __half *p) { v = *reinterpret_cast<const uint4 *>(p); }
    __device__ __forceinline__ void store(__half *p) const { *reinterpret_cast<uint4 *>(p) = v; }
    __device__ __forceinline__ void to_float(float f[8]) const {
        const __half2 *h = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
        for (int i = 0; i < 4; i++) { float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
    __device__ __forceinline__ void from_float(const float f[8]) {
        __half2 *h = reinterpret_cast<__half2 *>(&v);
#pragma unroll
        for (int i = 0; i < 4; i++) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    }
};

// int8 activations (INT8 path): 8 consecutive channels = 8 bytes; to_float yields the raw integer values, the
// caller applies the tensor's scale.
template <> struct Vec8<int8_t> {
    uint2 v;
    __device__ __forceinline__ void load(const int8_t *p) { v = *reinterpret_cast<const uint2 *>(p); }
    __device__ __forceinline__ void to_float(float f[8]) const {
        f[0] = (float)(int8_t)(v.x & 0xff); f[1] = (float)(int8_t)((v.x >> 8) & 0xff); f[2] = (float)(int8_t)((v.x >> 16) & 0xff); f[3] = (float)(int8_t)(v.x >> 24);
        f[4] = (float)(int8_t)(v.y & 0xff); f[5] = (float)(int8_t)((v.y >> 8) & 0xff); f[6] = (float)(int8_t)((v.y >> 16) & 0xff); f[7] = (float)(int8_t)(v.y >> 24);
    }
};

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__half x) { return __half2float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __half from_f<__half>(float x) { return __float2half_rn(x); }

// ---- division by a launch constant: q = umulhi(n, ceil(2^32 / d)), exact while n * d < 2^32 (n, d >= 1; d == 1: mul = 0 marks
// the identity).  An integer division by a run-time value costs ~25 instructions; the 2-D depthwise kernels did five per thread.
inline uint32_t fast_div_mul(uint32_t d) { return d <= 1 ? 0u : (uint32_t)((((uint64_t)1 << 32) + d - 1) / d); }
__device__ __forceinline__ int fast_div(int n, uint32_t mul) { return mul ? (int)__umulhi((uint32_t)n, mul) : n; }
__device__ __forceinline__ int fast_floor_div(int n, int d, uint32_t mul) { return n >= 0 ? fast_div(n, mul) : -fast_div(-n + d - 1, mul); }

// ---- depthwise inner product without conversions: acc[i] += x[i] * w[i] for 8 packed FP16 pairs with FHFMA (fma.rn.f32.f16:
// FP16 x FP16 multiplicands, exact product, FP32 addend and result -- one instruction per MAC where cvt + FFMA took two).
// The depthwise weights are therefore FP16-rounded, like every other weight of the FP16 engine; the bias stays FP32.
__device__ __forceinline__ void fhfma2(float &a0, float &a1, uint32_t x2, uint32_t w2) {
    asm("{\n\t.reg .f16 xl, xh, wl, wh;\n\tmov.b32 {xl, xh}, %2;\n\tmov.b32 {wl, wh}, %3;\n\tfma.rn.f32.f16 %0, xl, wl, %0;\n\tfma.rn.f32.f16 %1, xh, wh, %1;\n\t}"
        : "+f"(a0), "+f"(a1) : "r"(x2), "r"(w2));
}
__device__ __forceinline__ void fhfma8(float (&acc)[8], const uint4 &x, const uint4 &w) {
    fhfma2(acc[0], acc[1], x.x, w.x); fhfma2(acc[2], acc[3], x.y, w.y); fhfma2(acc[4], acc[5], x.z, w.z); fhfma2(acc[6], acc[7], x.w, w.w);
}
__device__ __forceinline__ uint4 pack_half8(const float4 &a, const float4 &b) {
    const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w), h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
    return make_uint4(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1), *reinterpret_cast<const uint32_t *>(&h2),
                      *reinterpret_cast<const uint32_t *>(&h3));
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// Every kernel of the forward pass is launched with cudaLaunchAttributeProgrammaticStreamSerialization:
// its CTAs may start while the previous kernel drains.  pdl_trigger() lets the NEXT kernel start its
// own prologue (barrier init, TMEM allocation, weight copies -- nothing that depends on activations);
// pdl_wait() blocks until the PREVIOUS kernel has completed and its global writes are visible, and
// must precede the first read of an activation and the first global write.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// RF_NO_PDL=1 (read once): launch without the attribute -- griddepcontrol.* are then no-ops (A/B measurements)
inline bool pdl_allowed() {
    static const bool on = [] { const char *e = getenv("RF_NO_PDL"); return !(e && e[0] == '1'); }();
    return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_allowed() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---- post-process shared structures -------------------------------------------------------
struct LevelDesc {           // one FPN level of one launch
    int stride, h, w;        // feature map size
    int anchor_base;         // emission index of (num 0, j 0) of this level
    int pix_base;            // first pixel id of this level in the fused per-image pixel range
    float base[8];           // 2 base anchors x (x1,y1,x2,y2)
};
struct PostParams {          // device-resident: a replayed CUDA graph picks up new values without re-capture
    float score_thr;
    float nms_thr;
    const uint8_t *input;    // [n][net_h][net_w][3] u8 BGR images of this run (library buffer or caller's)
    unsigned comm_seq;       // multi-GPU exchange (comm.cu): sequence number of this step (0: no exchange) ...
    unsigned comm_slot;      // ... and its slot in every rank's gather window
};

}  // namespace rf
