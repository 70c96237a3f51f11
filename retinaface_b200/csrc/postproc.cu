// postproc.cu -- GPU post-process of librf_b200: predictor 1x1 convs + softmax + threshold +
// anchor decode + clip (fused, all FPN levels in one launch) and sort + greedy NMS.
//
// Replaces the reference's HOST post-process, retinaface/RetinaFace.cpp:
//   :666-687  blob gather (second half of cls_prob = P(face))         -> k_head_decode / k_blob_decode
//   :689-723  threshold-first decode loop                              -> decode_one()
//   :378-398  bbox_pred, :179-199 clip_boxes, :418-432 landmark_pred   -> decode_one()
//   :127-154  anchors_plane (computed on the fly from (k, ih, iw))     -> decode_one()
//   :434-492  CompareBBox sort + greedy nms                            -> k_nms
// so that no head tensor (527 KB / image at 448x448) ever leaves the GPU.
//
// Bit-exactness contract: given identical head values, candidate selection, kept-face
// selection and order are identical to the reference's; scores and landmarks are bit-identical;
// box corners may differ by <= 1 ulp where the reference's expf (libm) is not correctly rounded
// (we round exp() computed in double).  Every float operation below therefore spells its
// rounding (__fmul_rn/__fadd_rn: no FMA contraction; the file is also built with -fmad=false),
// and the `0.5 * (x - 1.0)` sub-expressions run in double like the reference's C++ does.
// Ties in score (std::sort leaves them unspecified) are broken by emission order.
#include "postproc_dev.cuh"

namespace rf {

namespace {

constexpr int NMS_THREADS = 512;

struct HeadLaunch {
    const void *feat[3];
    HeadWeights hw[3];
    LevelDesc lv[3];
    int blk_base[4];   // first block of each level (level-aligned blocks)
    float *blobs[9];
};

// One thread per feature-map pixel; a block never straddles levels so its shared memory holds
// exactly one level's 32x64 predictor weights.
template <typename T, bool WRITE_BLOBS>
__global__ void __launch_bounds__(128) k_head_decode(HeadLaunch L, int net_w, int net_h,
                                                     const PostParams *__restrict__ params, PostBuffers pb, int fuse_nms) {
    __shared__ __align__(16) float sw[32 * 64];
    __shared__ float sb[32];
    __shared__ int s_last;
    const int blk = blockIdx.x;
    const int l = blk >= L.blk_base[2] ? 2 : (blk >= L.blk_base[1] ? 1 : 0);
    const LevelDesc lv = L.lv[l];
    pdl_trigger();
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) sw[i] = L.hw[l].w[i];   // weights: independent of the previous kernel
    if (threadIdx.x < 32) sb[threadIdx.x] = L.hw[l].b[threadIdx.x];
    __syncthreads();
    pdl_wait();
    const int hw = lv.h * lv.w;
    const int j = (blk - L.blk_base[l]) * blockDim.x + threadIdx.x;
    const int img = blockIdx.y;
    if (j < hw) {
    const float thr = params->score_thr;
    const T *f = reinterpret_cast<const T *>(L.feat[l]) + ((size_t)img * hw + j) * 64;
    float x[64];
#pragma unroll
    for (int g = 0; g < 8; g++) {
        Vec8<T> v;
        v.load(f + g * 8);
        v.to_float(&x[g * 8]);
    }
    if (sizeof(T) == 1) {                 // int8 features: dequantise with the concat tensor's scale
        const float sc = L.hw[l].in_scale;
#pragma unroll
        for (int c = 0; c < 64; c++) x[c] = __fmul_rn(x[c], sc);
    }
    auto dot = [&](int o) -> float {
        float acc = sb[o];
        const float4 *w4 = reinterpret_cast<const float4 *>(&sw[o * 64]);
#pragma unroll
        for (int c = 0; c < 16; c++) {
            float4 w = w4[c];
            acc = __fmaf_rn(x[4 * c + 0], w.x, acc);
            acc = __fmaf_rn(x[4 * c + 1], w.y, acc);
            acc = __fmaf_rn(x[4 * c + 2], w.z, acc);
            acc = __fmaf_rn(x[4 * c + 3], w.w, acc);
        }
        return acc;
    };
    float s[4];
#pragma unroll
    for (int o = 0; o < 4; o++) s[o] = dot(o);
    // Softmax over the (N,2,2h,w) view (prototxt:1448-1483): anchor a pairs channel a (bg) with a+2 (face).
    float pf[2], pbg[2];
#pragma unroll
    for (int a = 0; a < 2; a++) softmax_pair(s[a], s[a + 2], pbg[a], pf[a]);
    const int ih = j / lv.w, iw = j % lv.w;
    if (WRITE_BLOBS) {
        float *cls = L.blobs[3 * l] + (size_t)img * 4 * hw;
        cls[0 * hw + j] = pbg[0]; cls[1 * hw + j] = pbg[1]; cls[2 * hw + j] = pf[0]; cls[3 * hw + j] = pf[1];
    }
#pragma unroll
    for (int a = 0; a < 2; a++) {
        const bool pass = !(pf[a] <= thr);   // reference: `if (conf <= threshold) continue;`
        if (!pass && !WRITE_BLOBS) continue;
        float reg[4], lm[10];
#pragma unroll
        for (int c = 0; c < 4; c++) reg[c] = dot(4 + a * 4 + c);
#pragma unroll
        for (int c = 0; c < 10; c++) lm[c] = dot(12 + a * 10 + c);
        if (WRITE_BLOBS) {
            float *bb = L.blobs[3 * l + 1] + (size_t)img * 8 * hw;
            float *lb = L.blobs[3 * l + 2] + (size_t)img * 20 * hw;
#pragma unroll
            for (int c = 0; c < 4; c++) bb[(a * 4 + c) * hw + j] = reg[c];
#pragma unroll
            for (int c = 0; c < 10; c++) lb[(a * 10 + c) * hw + j] = lm[c];
        }
        if (pass) {
            rf_det d;
            decode_one(pf[a], reg, lm, lv, a, ih, iw, net_w, net_h, lv.anchor_base + a * hw + j, d);
            append_candidate(pb, img, d);
        }
    }
    }
    if (fuse_nms) {
        // decode -> NMS in ONE launch: the block that finishes an image's last pixels sorts and suppresses it (last-block pattern;
        // the counter cleans itself for the next forward)
        __shared__ NmsSmem S;
        extern __shared__ int s_kept[];
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(&pb.tile_done[img], 1) == (int)gridDim.x - 1;
        __syncthreads();
        if (s_last) {
            __threadfence();
            nms_image<128, true>(img, threadIdx.x, params->nms_thr, params, pb, S, s_kept, [] { __syncthreads(); });
            if (threadIdx.x == 0) pb.tile_done[img] = 0;
        }
    }
}

struct BlobLaunch {
    const float *blobs[9];
    LevelDesc lv[3];
};

// rf_postprocess: one thread per anchor, reading caller-supplied NCHW head blobs.
__global__ void __launch_bounds__(256) k_blob_decode(BlobLaunch L, int net_w, int net_h,
                                                     const PostParams *__restrict__ params, PostBuffers pb) {
    pdl_trigger();
    pdl_wait();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;  // emission index within the image
    if (e >= pb.anchors_per_image) return;
    const int img = blockIdx.y;
    const int l = e >= L.lv[2].anchor_base ? 2 : (e >= L.lv[1].anchor_base ? 1 : 0);
    const LevelDesc lv = L.lv[l];
    const int hw = lv.h * lv.w;
    const int r = e - lv.anchor_base;
    const int num = r / hw, j = r % hw;
    const float *cls = L.blobs[3 * l] + (size_t)img * 4 * hw;
    const float conf = cls[(2 + num) * hw + j];
    if (conf <= params->score_thr) return;
    const float *bb = L.blobs[3 * l + 1] + (size_t)img * 8 * hw;
    const float *lb = L.blobs[3 * l + 2] + (size_t)img * 20 * hw;
    float reg[4], lm[10];
#pragma unroll
    for (int c = 0; c < 4; c++) reg[c] = bb[(num * 4 + c) * hw + j];
#pragma unroll
    for (int c = 0; c < 10; c++) lm[c] = lb[(num * 10 + c) * hw + j];
    rf_det d;
    decode_one(conf, reg, lm, lv, num, j / lv.w, j % lv.w, net_w, net_h, e, d);
    append_candidate(pb, img, d);
}

// One CTA per image: sort + greedy NMS (postproc_dev.cuh nms_image).
__global__ void __launch_bounds__(NMS_THREADS) k_nms(const PostParams *__restrict__ params, PostBuffers pb) {
    extern __shared__ int s_kept[];                         // [max_faces]
    __shared__ NmsSmem S;
    pdl_trigger();
    pdl_wait();
    nms_image<NMS_THREADS, false>(blockIdx.x, threadIdx.x, params->nms_thr, params, pb, S, s_kept, [] { __syncthreads(); });
}


// ---- test-time augmentation: views -> one candidate list in image coordinates (postproc.cuh) -------------------------
__global__ void __launch_bounds__(256) k_merge_views(PostBuffers src, ViewSet vs, PostBuffers dst) {
    const int v = blockIdx.x;
    const int n = min(src.out_counts[v], src.max_faces);
    const float sc = vs.scale[v], wm1 = vs.img_w_minus1;
    const bool flip = vs.flip[v] != 0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const rf_face f = src.out_dets[(size_t)v * src.max_faces + j].face;
        rf_det d;
        d.face.score = f.score;
        const float x1 = __fmul_rn(f.x1, sc), x2 = __fmul_rn(f.x2, sc);          // RetinaFace.cpp:733-734: rect.x1 * scale ...
        d.face.y1 = __fmul_rn(f.y1, sc);
        d.face.y2 = __fmul_rn(f.y2, sc);
        d.face.x1 = flip ? __fsub_rn(wm1, x2) : x1;
        d.face.x2 = flip ? __fsub_rn(wm1, x1) : x2;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            // mirrored view: the detector's "left eye" is the subject's right one -- swap 0<->1 and 3<->4 (2 = nose)
            const int ks = flip ? (k == 0 ? 1 : k == 1 ? 0 : k == 3 ? 4 : k == 4 ? 3 : 2) : k;
            const float x = __fmul_rn(f.lx[ks], sc);                                // :739
            d.face.lx[k] = flip ? __fsub_rn(wm1, x) : x;
            d.face.ly[k] = __fmul_rn(f.ly[ks], sc);
        }
        d.anchor_index = v * src.max_faces + j;
        append_candidate(dst, 0, d);
    }
}

size_t nms_smem_bytes(int max_faces) { return sizeof(int) * (size_t)max_faces; }

}  // namespace

template <typename T>
void launch_head_decode(const T *const feat[3], const HeadWeights hw[3], const LevelDesc lv[3], int n,
                        int net_w, int net_h, const PostParams *params, const PostBuffers &pb,
                        float *const blobs[9], cudaStream_t s, bool fuse_nms) {
    HeadLaunch L;
    int blk = 0;
    bool write = blobs && blobs[0];
    for (int l = 0; l < 3; l++) {
        L.feat[l] = feat[l];
        L.hw[l] = hw[l];
        L.lv[l] = lv[l];
        L.blk_base[l] = blk;
        blk += (lv[l].h * lv[l].w + 127) / 128;
    }
    L.blk_base[3] = blk;
    for (int i = 0; i < 9; i++) L.blobs[i] = write ? blobs[i] : nullptr;
    dim3 grid(blk, n);
    const size_t dyn = fuse_nms ? nms_smem_bytes(pb.max_faces) : 0;
    if (write) launch_k(k_head_decode<T, true>, grid, dim3(128), dyn, s, L, net_w, net_h, params, pb, fuse_nms ? 1 : 0);
    else launch_k(k_head_decode<T, false>, grid, dim3(128), dyn, s, L, net_w, net_h, params, pb, fuse_nms ? 1 : 0);
}
template void launch_head_decode<float>(const float *const[3], const HeadWeights[3], const LevelDesc[3], int, int, int,
                                        const PostParams *, const PostBuffers &, float *const[9], cudaStream_t, bool);
template void launch_head_decode<__half>(const __half *const[3], const HeadWeights[3], const LevelDesc[3], int, int, int,
                                         const PostParams *, const PostBuffers &, float *const[9], cudaStream_t, bool);
template void launch_head_decode<int8_t>(const int8_t *const[3], const HeadWeights[3], const LevelDesc[3], int, int, int,
                                         const PostParams *, const PostBuffers &, float *const[9], cudaStream_t, bool);

void launch_blob_decode(const float *const blobs[9], const LevelDesc lv[3], int n, int net_w, int net_h,
                        const PostParams *params, const PostBuffers &pb, cudaStream_t s) {
    BlobLaunch L;
    for (int i = 0; i < 9; i++) L.blobs[i] = blobs[i];
    for (int l = 0; l < 3; l++) L.lv[l] = lv[l];
    dim3 grid((pb.anchors_per_image + 255) / 256, n);
    launch_k(k_blob_decode, grid, dim3(256), 0, s, L, net_w, net_h, params, pb);
}

void launch_nms(int n, const PostParams *params, const PostBuffers &pb, cudaStream_t s) {
    launch_k(k_nms, dim3(n), dim3(NMS_THREADS), nms_smem_bytes(pb.max_faces), s, params, pb);
}

void launch_merge_views(const PostBuffers &src, const ViewSet &vs, const PostBuffers &dst, cudaStream_t s) {
    // dst.cand_count[0] is zero here: cleared at allocation and by every k_nms on dst (self-cleaning)
    k_merge_views<<<vs.nviews, 256, 0, s>>>(src, vs, dst);
}

cudaError_t postproc_init() {
    // static 27 KB + up to 32 KB dynamic (max_faces <= 8192) exceeds the 48 KB default: opt in once
    cudaError_t e;
    const int dyn = (int)nms_smem_bytes(8192);
#define RF_HD_ATTR(T_, W_) if ((e = cudaFuncSetAttribute(k_head_decode<T_, W_>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn))) return e
    RF_HD_ATTR(float, true); RF_HD_ATTR(float, false); RF_HD_ATTR(__half, true); RF_HD_ATTR(__half, false); RF_HD_ATTR(int8_t, true); RF_HD_ATTR(int8_t, false);
#undef RF_HD_ATTR
    return cudaFuncSetAttribute(k_nms, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
}

}  // namespace rf
