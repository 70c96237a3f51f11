// postproc.cuh -- launch wrappers of the GPU post-process (implemented in postproc.cu).
#pragma once
#include "common.cuh"

namespace rf {

// Multi-GPU exchange of the final detections, fused into the NMS: the CTA that finishes an image stores its
// kept records straight into the gather window of EVERY rank (peer device memory over NVLink, mapped through CUDA IPC),
// then raises that image's flag there.  Window of one rank: [ring][world][max_batch] x {count, flag, max_faces records}.
constexpr int RF_COMM_MAX_WORLD = RF_COMM_MAX_WORLD_SIZE;
struct CommView {
    int world, rank, ring;
    rf_det *dets[RF_COMM_MAX_WORLD];       // rank p's window: [ring][world][max_batch][max_faces]
    int *counts[RF_COMM_MAX_WORLD];        //                  [ring][world][max_batch]
    unsigned *flags[RF_COMM_MAX_WORLD];    //                  [ring][world][max_batch]  == seq once the image's records have landed
};

struct PostBuffers {
    // per image i (capacity = anchors_per_image A):
    unsigned long long *cand_keys;  // [B][A]   sort keys of candidates in append order
    rf_det *cand_recs;              // [B][A]   decoded record, indexed by anchor emission index
    int *cand_count;                // [B]      number appended (reset by the head kernel's launch wrapper)
    unsigned long long *sort_scratch;  // [B][A_pow2] global scratch for sorts that do not fit in smem
    unsigned char *flag_scratch;    // [B][A_pow2]
    rf_det *out_dets;               // [B][max_faces]
    int *out_counts;                // [B]   kept (clamped to max_faces)
    int *out_total_kept;            // [B]   kept before clamping
    int *tile_done;                 // [B]   tiles of the image finished (tile_chain.cuh last-block NMS; self-cleaning)
    int anchors_per_image;
    int anchors_pow2;
    int max_faces;
    int max_batch;
    CommView comm;                  // world <= 1: single GPU, no exchange
};

struct HeadWeights {
    const float *w;     // [32][64]: rows 0-3 cls_score, 4-11 bbox_pred, 12-31 landmark_pred
    const float *b;     // [32]
    float in_scale;     // 1 for float/half features; the concat tensor's quantisation scale for int8 features
};

// fuse_nms: the block that completes an image also sorts + suppresses it (decode -> NMS in one launch; pb.tile_done counts).
// Fused per-level predictor + decode: 1x1 convs (cls 4, bbox 8, landmark 20), the 2-way softmax,
// threshold, anchor decode, clip -> candidate append.  One launch covers all three levels.
// feat[l]: NHWC [n][h][w][64] SSH output (post concat+ReLU) in T.  blobs (optional, may be all
// NULL): the 9 NCHW float32 head blobs in engine order, for rf_forward_heads.
template <typename T>
void launch_head_decode(const T *const feat[3], const HeadWeights hw[3], const LevelDesc lv[3], int n,
                        int net_w, int net_h, const PostParams *params, const PostBuffers &pb,
                        float *const blobs[9], cudaStream_t s, bool fuse_nms = false);

// Decode from caller-provided head blobs (device, NCHW f32, engine order): rf_postprocess.
void launch_blob_decode(const float *const blobs[9], const LevelDesc lv[3], int n, int net_w, int net_h,
                        const PostParams *params, const PostBuffers &pb, cudaStream_t s);

// Sort candidates by (score desc, emission index asc) and run greedy NMS; one CTA per image.
void launch_nms(int n, const PostParams *params, const PostBuffers &pb, cudaStream_t s);

// Test-time augmentation: gather the kept detections of `nviews` views (src.out_dets / out_counts, network
// coordinates of each view) into ONE candidate list in original-image coordinates: x *= scale[v] (the reference's map-back,
// RetinaFace.cpp:732-738), mirrored views are un-mirrored (x -> img_w-1 - x, box corners and left/right landmarks swapped).
// Candidate id (rf_det::anchor_index of the merged records) = view * max_faces + rank in the view.  launch_nms(1, ..., dst)
// then selects across views.
constexpr int RF_MAX_VIEWS_DEV = 16;
struct ViewSet {
    int nviews;
    float img_w_minus1;
    float scale[RF_MAX_VIEWS_DEV];
    int flip[RF_MAX_VIEWS_DEV];
};
void launch_merge_views(const PostBuffers &src, const ViewSet &vs, const PostBuffers &dst, cudaStream_t s);

// dynamic shared memory the NMS kernel wants (set once at init)
cudaError_t postproc_init();

}  // namespace rf
