// engine.cu -- the handle behind include/rf_b200.h: model upload, layer plan, activation arena,
// CUDA-graph executor and the C-ABI entry points.
//
// Replaces the reference's engine slot: TrtNetBase / TrtRetinaFaceNet
// (retinaface/tensorrt/trtnetbase.cpp:199-330, trtretinafacenet.cpp:48-210) and the detect
// orchestration of RetinaFace::detect / detectBatchImages (retinaface/RetinaFace.cpp:576-940).
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#include "calibrate.cuh"
#include "common.cuh"
#include "kernels_simt.cuh"
#include "model.h"
#include "postproc.cuh"
#include "preprocess.cuh"
#include "tc_conv.cuh"
#include "tc_conv_i8.cuh"
#include "stem_tc.cuh"
#include "tc_dwpw2d.cuh"
#include "tc_dwpw2d_i8.cuh"

using namespace rf;

#define RF_STR2(x) #x
#define RF_STR(x) RF_STR2(x)

namespace {

thread_local std::string g_create_error;

struct CudaFail { cudaError_t e; const char *what; const char *file; int line; };
#define CK(call)                                                        \
    do {                                                                \
        cudaError_t _e = (call);                                        \
        if (_e != cudaSuccess) throw CudaFail{_e, #call, __FILE__, __LINE__}; \
    } while (0)

std::string fmt(const char *f, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof buf, f, ap);
    va_end(ap);
    return buf;
}

struct TensorInfo {
    std::string name;
    int h = 0, w = 0, c = 0;
    size_t bytes_per_img = 0;
    int first = -1, last = -1;
    size_t offset = 0;  // bytes into the arena (already scaled by max_batch)
};

struct Step {
    std::string name;
    std::vector<int> in, out;
    std::function<void(int /*n*/, cudaStream_t)> launch;
    double flops_per_img = 0, bytes_per_img = 0;  // algorithmic
    int lane = 0;                 // 0 = main stream; 1, 2 = side branches of the forward graph
    std::vector<int> deps;        // producer steps in OTHER lanes this step must wait for (filled by link_steps)
    bool signals = false;         // some step in another lane waits for this one
};

}  // namespace

struct rf_handle_s {
    rf_config cfg{};
    std::string caffemodel, table;
    std::string err;
    Model model;
    std::map<std::string, float> int8_scales;
    int device = 0;
    int elem = 4;  // bytes per activation element
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    std::vector<TensorInfo> tensors;
    std::map<std::string, int> tensor_by_name;
    std::vector<Step> steps;
    int head_step = -1;
    unsigned char *arena = nullptr;
    size_t arena_bytes = 0;

    // weights
    std::vector<float> wstage;  // host staging of all fp32 weights
    float *d_weights = nullptr;
    std::vector<__half> wstage_h;  // FP16 tensor-core weight images (tc_conv.cuh B chunks)
    __half *d_weights_h = nullptr;
    std::vector<int8_t> wstage_q;  // INT8 tensor-core weight images (tc_conv_i8.cuh)
    int8_t *d_weights_q = nullptr;
    bool use_tc = false;

    // io
    uint8_t *d_input = nullptr;       // [max_batch][H][W][3] u8 BGR
    uint8_t *h_input = nullptr;       // pinned mirror
    PostBuffers pb_merge{};           // rf_detect_views: candidates of all views of one image (lazily allocated)
    uint8_t *d_raw = nullptr;         // one raw caller image (max_image) for the letterbox kernel
    uint8_t *h_raw = nullptr;         // pinned
    size_t raw_bytes = 0;
    PostParams *d_params = nullptr, *h_params = nullptr;
    PostBuffers pb{};
    LevelDesc lv[3];
    HeadWeights hw[3];
    int feat_tensor[3] = {-1, -1, -1};
    float *d_blobs[9] = {nullptr};    // rf_forward_heads / rf_postprocess staging (device)
    size_t blob_elems[9] = {0};       // per image
    rf_det *h_dets = nullptr;         // pinned [max_batch][max_faces]
    int *h_counts = nullptr;          // pinned [2*max_batch]: kept, candidates
    std::map<int, cudaGraphExec_t> graphs;
    // pipelined end-to-end path (rf_submit_batch / rf_collect_batch)
    struct Slot {
        uint8_t *d_in = nullptr, *h_in = nullptr;     // device input, pinned staging for pageable sources
        rf_det *h_dets = nullptr;                      // pinned results
        int *h_counts = nullptr;
        cudaEvent_t ev_h2d = nullptr, ev_done = nullptr;
        int n = 0;
        bool busy = false;
    } slots[RF_PIPELINE_DEPTH];
    cudaStream_t copy_stream = nullptr;
    unsigned submit_seq = 0, collect_seq = 0;
    cudaStream_t lane_stream[3] = {nullptr, nullptr, nullptr};   // [0] unused (the caller's stream is lane 0)
    std::vector<cudaEvent_t> step_event;
    bool blobs_in_plan = false;       // head step writes blobs (forward_heads path)
    static constexpr int kParamSlots = 1024;
    unsigned param_seq = 0;
    float cur_thr = 0.5f, cur_nms = 0.4f;

    void *tptr(int id) const { return arena + tensors[id].offset; }

    // Execution contexts.  Everything a forward pass writes (activation arena, candidate / output buffers,
    // run parameters) and everything it is issued on (stream, lane streams, events, captured graphs) exists
    // once per context; the asynchronous entry points rotate through the contexts so that consecutive batches
    // overlap on the GPU (most kernels of one batch-8 step fill well under one wave of the 148 SMs).  The
    // members above always hold the ACTIVE context; switch_ctx() swaps them with a saved one.
    struct Ctx {
        cudaStream_t stream = nullptr, lane_stream[3] = {nullptr, nullptr, nullptr};
        std::vector<cudaEvent_t> step_event;
        unsigned char *arena = nullptr;
        PostBuffers pb{};
        PostParams *d_params = nullptr, *h_params = nullptr;
        unsigned param_seq = 0;
        float cur_thr = 0.5f, cur_nms = 0.4f;
        std::map<int, cudaGraphExec_t> graphs;
        cudaEvent_t fence = nullptr;
    };
    std::vector<Ctx> saved;
    int active = 0, nctx = 1;
    unsigned next_dev_ctx = 0;
    cudaStream_t last_stream = nullptr;
    cudaEvent_t fence = nullptr;
};

namespace {
void switch_ctx(rf_handle h, int i) {
    if (i == h->active) return;
    auto xchg = [&](rf_handle_s::Ctx &c) {
        std::swap(c.stream, h->stream);
        for (int l = 0; l < 3; l++) std::swap(c.lane_stream[l], h->lane_stream[l]);
        std::swap(c.step_event, h->step_event);
        std::swap(c.arena, h->arena);
        std::swap(c.pb, h->pb);
        std::swap(c.d_params, h->d_params);
        std::swap(c.h_params, h->h_params);
        std::swap(c.param_seq, h->param_seq);
        std::swap(c.cur_thr, h->cur_thr);
        std::swap(c.cur_nms, h->cur_nms);
        std::swap(c.graphs, h->graphs);
        std::swap(c.fence, h->fence);
    };
    xchg(h->saved[h->active]);   // park the active state in its slot
    xchg(h->saved[i]);           // and bring context i in
    h->active = i;
}
}  // namespace

namespace {

int fail(rf_handle h, int code, const std::string &msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}
int fail_cuda(rf_handle h, const CudaFail &f) {
    return fail(h, RF_ERR_CUDA, fmt("%s failed: %s (%s:%d)", f.what, cudaGetErrorString(f.e), f.file, f.line));
}

// ---------------------------------------------------------------------------------------------
// Plan builder
// ---------------------------------------------------------------------------------------------
struct Builder {
    rf_handle h;
    int H, W;
    size_t add_weights(const std::vector<float> &v) {
        size_t off = h->wstage.size();
        h->wstage.insert(h->wstage.end(), v.begin(), v.end());
        while (h->wstage.size() % 4) h->wstage.push_back(0.f);  // keep float4 alignment
        return off;
    }
    size_t add_weights_h(const std::vector<__half> &v) {
        size_t off = h->wstage_h.size();
        h->wstage_h.insert(h->wstage_h.end(), v.begin(), v.end());
        while (h->wstage_h.size() % 64) h->wstage_h.push_back(__float2half(0.f));  // 128-byte alignment for bulk copies
        return off;
    }
    size_t add_weights_q(const std::vector<int8_t> &v) {
        size_t off = h->wstage_q.size();
        h->wstage_q.insert(h->wstage_q.end(), v.begin(), v.end());
        while (h->wstage_q.size() % 128) h->wstage_q.push_back(0);
        return off;
    }
    int tensor(const std::string &name, int hh, int ww, int c) {
        TensorInfo t;
        t.name = name; t.h = hh; t.w = ww; t.c = c;
        t.bytes_per_img = (size_t)hh * ww * c * h->elem;
        h->tensors.push_back(t);
        h->tensor_by_name[name] = (int)h->tensors.size() - 1;
        return (int)h->tensors.size() - 1;
    }
    void step(Step s) { h->steps.push_back(std::move(s)); }
};

// GEMM weight matrix [K = (tap, cin)][N] from conv weights [cout][cin][k][k]; several convs that
// share an input are concatenated along N (det_conv1 + context_conv1, context_conv2 + conv3_1).
std::vector<float> pack_gemm(const std::vector<const FoldedConv *> &cs, std::vector<float> &bias) {
    const int cin = cs[0]->cin, k = cs[0]->k;
    int N = 0;
    for (auto c : cs) N += c->cout;
    std::vector<float> w((size_t)k * k * cin * N);
    bias.assign(N, 0.f);
    int n0 = 0;
    for (auto c : cs) {
        for (int o = 0; o < c->cout; o++) {
            bias[n0 + o] = c->b[o];
            for (int ci = 0; ci < cin; ci++)
                for (int t = 0; t < k * k; t++)
                    w[((size_t)t * cin + ci) * N + n0 + o] = c->w[((size_t)o * cin + ci) * k * k + t];
        }
        n0 += c->cout;
    }
    return w;
}

template <typename T>
void launch_gemm(const T *in, int ldin, int cin, const float *wk, const float *bias, int N, int ks, OutSplit<T> outs,
                 int n, int H, int W, cudaStream_t s) {
    long M = (long)n * H * W;
    int bn = (N % 64 == 0) ? 64 : (N % 32 == 0 ? 32 : 16);
    dim3 grid((unsigned)((M + 63) / 64), (N + bn - 1) / bn);
#define RF_GEMM(BN_, KS_) launch_k(k_conv_gemm<T, BN_, KS_>, grid, dim3(256), 0, s, in, ldin, cin, wk, bias, N, outs, n, H, W)
    if (ks == 1) { if (bn == 64) RF_GEMM(64, 1); else if (bn == 32) RF_GEMM(32, 1); else RF_GEMM(16, 1); }
    else { if (bn == 64) RF_GEMM(64, 3); else if (bn == 32) RF_GEMM(32, 3); else RF_GEMM(16, 3); }
#undef RF_GEMM
}

// ---- tcgen05 path helpers --------------------------------------------------------------------
// B operand image [K/8][n][8] halfs (UMMA K-major no-swizzle, LBO = n*16 B), K ordered (tap, cin) and
// zero-padded to a multiple of 16; convs sharing an input are concatenated along N; `nsplit` slices
// of N each get their own image (slice s at s * Kpad * (N/nsplit)).
std::vector<__half> pack_tc_weights(const std::vector<const FoldedConv *> &cs, std::vector<float> &bias, int &Kpad, int nsplit = 1) {
    const int cin = cs[0]->cin, k = cs[0]->k;
    int N = 0;
    for (auto c : cs) N += c->cout;
    const int K = k * k * cin;
    Kpad = (K + 15) / 16 * 16;
    const int Ns = N / nsplit;
    std::vector<__half> img((size_t)Kpad * N, __float2half(0.f));
    bias.assign(N, 0.f);
    int n0 = 0;
    for (auto c : cs) {
        for (int o = 0; o < c->cout; o++) {
            const int n = n0 + o, sl = n / Ns, nl = n % Ns;
            bias[n] = c->b[o];
            for (int ci = 0; ci < cin; ci++)
                for (int t = 0; t < k * k; t++) {
                    const int kk = t * cin + ci;
                    img[(size_t)sl * Kpad * Ns + ((size_t)(kk / 8) * Ns + nl) * 8 + (kk % 8)] =
                        __float2half(c->w[((size_t)o * cin + ci) * k * k + t]);
                }
        }
        n0 += c->cout;
    }
    return img;
}

void launch_tc_conv(const TcConvArgs &a, cudaStream_t s) {
    const long P = (long)a.nimg * a.Hp * a.Wp;
    const unsigned grid = (unsigned)((P + 127) / 128);
    const size_t smem = tc_conv_smem_bytes(a);
    switch (tc_tmem_cols(a.N)) {
        case 32: if (a.up) launch_k(k_tc_conv_staged<32, true>, dim3(grid), dim3(TC_THREADS), smem, s, a); else launch_k(k_tc_conv_staged<32, false>, dim3(grid), dim3(TC_THREADS), smem, s, a); break;
        case 64: if (a.up) launch_k(k_tc_conv_staged<64, true>, dim3(grid), dim3(TC_THREADS), smem, s, a); else launch_k(k_tc_conv_staged<64, false>, dim3(grid), dim3(TC_THREADS), smem, s, a); break;
        case 128: if (a.up) launch_k(k_tc_conv_staged<128, true>, dim3(grid), dim3(TC_THREADS), smem, s, a); else launch_k(k_tc_conv_staged<128, false>, dim3(grid), dim3(TC_THREADS), smem, s, a); break;
        default: if (a.up) launch_k(k_tc_conv_staged<256, true>, dim3(grid), dim3(TC_THREADS), smem, s, a); else launch_k(k_tc_conv_staged<256, false>, dim3(grid), dim3(TC_THREADS), smem, s, a); break;
    }
}
void launch_tc_dwpw(const TcDwArgs &a, int nsplit, cudaStream_t s) {
    const long M = (long)a.nimg * a.OH * a.OW;
    dim3 grid((unsigned)((M + a.rows - 1) / a.rows), nsplit);
    const size_t smem = tc_dw_smem_bytes(a);
    switch (tc_tmem_cols(a.N)) {
        case 32: if (a.C >= 64) launch_k(k_tc_dwpw_staged<32, true>, grid, dim3(TC_THREADS), smem, s, a); else launch_k(k_tc_dwpw_staged<32, false>, grid, dim3(TC_THREADS), smem, s, a); break;
        case 64: if (a.C >= 64) launch_k(k_tc_dwpw_staged<64, true>, grid, dim3(TC_THREADS), smem, s, a); else launch_k(k_tc_dwpw_staged<64, false>, grid, dim3(TC_THREADS), smem, s, a); break;
        case 128: if (a.C >= 64) launch_k(k_tc_dwpw_staged<128, true>, grid, dim3(TC_THREADS), smem, s, a); else launch_k(k_tc_dwpw_staged<128, false>, grid, dim3(TC_THREADS), smem, s, a); break;
        default: if (a.C >= 64) launch_k(k_tc_dwpw_staged<256, true>, grid, dim3(TC_THREADS), smem, s, a); else launch_k(k_tc_dwpw_staged<256, false>, grid, dim3(TC_THREADS), smem, s, a); break;
    }
}
void launch_tc_dwpw_2d(const TcDw2dArgs &a, cudaStream_t s) {
    const dim3 grid((unsigned)(a.tiles_x * a.tiles_y * a.nimg));
    const size_t smem = tc_dw2d_smem_bytes(a);
    switch (tc_tmem_cols(a.N)) {
        case 32: launch_k(k_tc_dwpw_2d<32>, grid, dim3(TC_THREADS), smem, s, a); break;
        case 64: launch_k(k_tc_dwpw_2d<64>, grid, dim3(TC_THREADS), smem, s, a); break;
        case 128: launch_k(k_tc_dwpw_2d<128>, grid, dim3(TC_THREADS), smem, s, a); break;
        default: launch_k(k_tc_dwpw_2d<256>, grid, dim3(TC_THREADS), smem, s, a); break;
    }
}

constexpr int TC_SMEM_LIMIT = 200 * 1024;   // dynamic; the kernels also hold ~20 KB static
cudaError_t tc_init() {
    cudaError_t e;
#define RF_TC_ATTR(K_) if ((e = cudaFuncSetAttribute(K_, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT))) return e
    RF_TC_ATTR((k_tc_conv_staged<32, false>)); RF_TC_ATTR((k_tc_conv_staged<64, false>)); RF_TC_ATTR((k_tc_conv_staged<128, false>)); RF_TC_ATTR((k_tc_conv_staged<256, false>));
    RF_TC_ATTR((k_tc_conv_staged<32, true>)); RF_TC_ATTR((k_tc_conv_staged<64, true>)); RF_TC_ATTR((k_tc_conv_staged<128, true>)); RF_TC_ATTR((k_tc_conv_staged<256, true>));
    RF_TC_ATTR((k_tc_dwpw_staged<32, true>)); RF_TC_ATTR((k_tc_dwpw_staged<64, true>)); RF_TC_ATTR((k_tc_dwpw_staged<128, true>)); RF_TC_ATTR((k_tc_dwpw_staged<256, true>));
    RF_TC_ATTR((k_tc_dwpw_staged<32, false>)); RF_TC_ATTR((k_tc_dwpw_staged<64, false>)); RF_TC_ATTR((k_tc_dwpw_staged<128, false>)); RF_TC_ATTR((k_tc_dwpw_staged<256, false>));
    RF_TC_ATTR(k_tc_dwpw_2d<32>); RF_TC_ATTR(k_tc_dwpw_2d<64>); RF_TC_ATTR(k_tc_dwpw_2d<128>); RF_TC_ATTR(k_tc_dwpw_2d<256>);
#undef RF_TC_ATTR
    return cudaSuccess;
}

// Tile geometry of one fused depthwise+pointwise layer: rows per CTA, N slices and the exact upper
// bound of the staged range, so that everything fits in shared memory.
struct DwGeom { int rows, nsplit, Rmax; };
DwGeom dw_geometry(int C, int N, int IH, int IW, int S) {
    const int OH = IH / S, OW = IW / S, Wp = IW + 2, Hp = IH + 1, Kpad = (C + 15) / 16 * 16;
    auto centre = [&](long m) { long ox = m % OW, oy = (m / OW) % OH, b = m / ((long)OW * OH); return (b * Hp + oy * S) * Wp + ox * S + 1; };
    for (int rows : {128, 64}) {
        if (rows == 128 && OH * OW <= 28 * 28) continue;   // small maps: more, smaller CTAs (latency bound)
        for (int nsplit : {1, 2, 4}) {
            if ((N / nsplit) % 16) continue;
            // tile starts shift against image boundaries with period lcm(rows, OH*OW): scan one full period
            // (+1 image) so that every alignment, including tiles straddling two images, is covered
            long g = rows, t = (long)OH * OW;
            while (t) { long u = g % t; g = t; t = u; }
            const long M = ((long)rows / g + 1) * OH * OW;
            int R = 0;
            for (long m0 = 0; m0 < M; m0 += rows) {
                long ml = std::min(m0 + rows, M) - 1;
                R = std::max(R, (int)(centre(ml) - centre(m0) + 2 * (Wp + 1) + 1));
            }
            R |= 1;
            TcDwArgs a{};
            a.C = C; a.Rmax = R; a.Kpad = Kpad; a.N = N / nsplit; a.rows = rows;
            if (R <= TC_MAX_R && tc_dw_smem_bytes(a) <= (size_t)TC_SMEM_LIMIT) return {rows, nsplit, R};
        }
    }
    return {0, 0, 0};
}

// Constants of the tensor-core stem (stem_tc.cuh) as one blob: conv0's folded FP32 weights as two FP16 pieces (hi + lo), the
// pointwise B image, then the FP32 constants.  w0: [27][8] (k = (tap*3 + c_bgr), out channel), wd: [9][8], wp: [8][16].
static std::vector<__half> make_stem_blob(const std::vector<float> &w0, const std::vector<float> &b0, const std::vector<float> &wd,
                                          const std::vector<float> &bd, const std::vector<float> &wp, const std::vector<float> &bp) {
    std::vector<__half> b0img(2 * 4 * 16 * 8, __float2half(0.f)), b1img(2 * 16 * 8, __float2half(0.f));
    for (int k = 0; k < 27; k++)
        for (int o = 0; o < 8; o++) {
            const float wv = w0[k * 8 + o];
            const __half hi = __float2half(wv);
            b0img[((k / 8) * 16 + o) * 8 + (k % 8)] = hi;                                            // w = hi + lo
            b0img[((4 + k / 8) * 16 + o) * 8 + (k % 8)] = __float2half(wv - __half2float(hi));
        }
    for (int c = 0; c < 8; c++)
        for (int o = 0; o < 16; o++) b1img[(0 * 16 + o) * 8 + c] = __float2half(wp[c * 16 + o]);
    std::vector<__half> blob(STEM_CONST_BYTES / 2, __float2half(0.f));
    memcpy(blob.data(), b0img.data(), STEM_B0_BYTES);
    memcpy(reinterpret_cast<unsigned char *>(blob.data()) + STEM_B0_BYTES, b1img.data(), STEM_B1_BYTES);
    std::vector<float> fl;
    fl.insert(fl.end(), b0.begin(), b0.begin() + 8);
    fl.insert(fl.end(), wd.begin(), wd.begin() + 72);
    fl.insert(fl.end(), bd.begin(), bd.begin() + 8);
    fl.insert(fl.end(), bp.begin(), bp.begin() + 16);
    fl.insert(fl.end(), wp.begin(), wp.begin() + 128);
    memcpy(reinterpret_cast<unsigned char *>(blob.data()) + STEM_B0_BYTES + STEM_B1_BYTES, fl.data(), STEM_F_FLOATS * 4);
    return blob;
}

template <typename T>
void build_plan(rf_handle h) {
    Builder B{h, h->cfg.net_h, h->cfg.net_w};
    const Model &m = h->model;
    const int H = h->cfg.net_h, W = h->cfg.net_w;
    auto T_ = [h](int id) { return reinterpret_cast<T *>(h->tptr(id)); };
    auto Wd = [h](size_t off) { return h->d_weights + off; };
    const double es = h->elem;

    // ---- stem ------------------------------------------------------------------------------------
    int first_pair = 1;
    int cur_h = H / 2, cur_w = W / 2, cur_c = 8;
    int cur = -1;
    bool stem_done = false;
    if constexpr (std::is_same<T, __half>::value) {
        if (h->use_tc) {
            // conv0 + dw1 + pw2 fused: the two dense layers on tensor cores (stem_tc.cuh), or all on CUDA cores
            // (kernels_simt.cuh k_stem) with RF_FLAG_SIMT_STEM
            const FoldedConv &c0 = m.conv("mobilenet0_conv0_fwd"), &dw = m.conv("mobilenet0_conv1_fwd"), &pw = m.conv("mobilenet0_conv2_fwd");
            std::vector<float> w0(27 * 8), wd(72), wp(128);
            for (int o = 0; o < 8; o++)
                for (int cb = 0; cb < 3; cb++)
                    for (int t = 0; t < 9; t++) w0[(t * 3 + cb) * 8 + o] = c0.w[((size_t)o * 3 + (2 - cb)) * 9 + t];
            for (int c = 0; c < 8; c++)
                for (int t = 0; t < 9; t++) wd[t * 8 + c] = dw.w[(size_t)c * 9 + t];
            for (int o = 0; o < 16; o++)
                for (int c = 0; c < 8; c++) wp[c * 16 + o] = pw.w[(size_t)o * 8 + c];
            size_t ow0 = B.add_weights(w0), ob0 = B.add_weights(c0.b), owd = B.add_weights(wd), obd = B.add_weights(dw.b),
                   owp = B.add_weights(wp), obp = B.add_weights(pw.b);
            std::vector<__half> blob = make_stem_blob(w0, c0.b, wd, dw.b, wp, pw.b);
            size_t oblob = B.add_weights_h(blob);
            const bool simt_stem = (h->cfg.flags & (RF_FLAG_SIMT_STEM | RF_FLAG_NO_TENSORCORE)) != 0;
            cur = B.tensor("mobilenet0_relu2_fwd", cur_h, cur_w, 16);
            int out = cur;
            Step s;
            s.name = simt_stem ? "stem_conv0+dw1+pw2_u8_to_16ch" : "tc_stem_conv0+dw1+pw2_u8_to_16ch";
            s.out = {out};
            s.flops_per_img = 2.0 * cur_h * cur_w * (8 * 27 + 8 * 9 + 8 * 16);
            s.bytes_per_img = (double)H * W * 3 + (double)cur_h * cur_w * 16 * es;
            s.launch = [=](int n, cudaStream_t st) {
                const int tiles = ((H / 2 + 15) / 16) * ((W / 2 + 15) / 16);
                if (simt_stem) {
                    StemWeights sw{Wd(ow0), Wd(ob0), Wd(owd), Wd(obd), Wd(owp), Wd(obp)};
                    launch_k(k_stem<__half>, dim3((unsigned)(tiles * n)), dim3(256), 0, st, (const PostParams *)h->d_params, (__half *)T_(out), sw, n, H, W, 1.0f);
                } else {
                    StemTcArgs a{reinterpret_cast<const unsigned char *>(h->d_weights_h + oblob)};
                    launch_k(k_stem_tc<__half>, dim3((unsigned)((W / 2 + 15) / 16), (unsigned)((H / 2 + 15) / 16), (unsigned)n), dim3(256), 0, st, (const PostParams *)h->d_params, (__half *)T_(out), a, n, H, W, 1.0f);
                }
            };
            B.step(std::move(s));
            cur_c = 16;
            first_pair = 3;
            stem_done = true;
        }
    }
    if (!stem_done) {
    cur = B.tensor("mobilenet0_relu0_fwd", cur_h, cur_w, 8);
    {
        const FoldedConv &c = m.conv("mobilenet0_conv0_fwd");
        std::vector<float> wk(27 * 8);
        for (int o = 0; o < 8; o++)
            for (int cb = 0; cb < 3; cb++)       // cb: BGR channel of the u8 image; network channel = 2 - cb (RGB)
                for (int t = 0; t < 9; t++) wk[(t * 3 + cb) * 8 + o] = c.w[((size_t)o * 3 + (2 - cb)) * 9 + t];
        size_t ow = B.add_weights(wk), ob = B.add_weights(c.b);
        int out = cur;
        Step s;
        s.name = "conv0_u8_3x3s2_bn_relu";
        s.out = {out};
        s.flops_per_img = 2.0 * cur_h * cur_w * 8 * 27;
        s.bytes_per_img = (double)H * W * 3 + (double)cur_h * cur_w * 8 * es;
        s.launch = [=](int n, cudaStream_t st) {
            long total = (long)n * (H / 2) * (W / 2);
            launch_k(k_conv0<T>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, (const PostParams *)h->d_params, T_(out), Wd(ow), Wd(ob), n, H, W);
        };
        B.step(std::move(s));
    }
    }
    // ---- 13 x (depthwise 3x3, pointwise 1x1) (prototxt:55-1192) -----------------------------
    int c1 = -1, c2 = -1, c3 = -1;
    for (int i = first_pair; i <= 26; i += 2) {
        const FoldedConv &dw = m.conv("mobilenet0_conv" + std::to_string(i) + "_fwd");
        const FoldedConv &pw = m.conv("mobilenet0_conv" + std::to_string(i + 1) + "_fwd");
        const int C = dw.cout, S = dw.stride;
        std::vector<float> wd(9 * C);
        for (int c = 0; c < C; c++)
            for (int t = 0; t < 9; t++) wd[t * C + c] = dw.w[(size_t)c * 9 + t];
        size_t owd = B.add_weights(wd), obd = B.add_weights(dw.b);
        const int ih = cur_h, iw = cur_w, oh = cur_h / S, ow_ = cur_w / S;
        int tin = cur;
        if constexpr (std::is_same<T, __half>::value) {
            if (h->use_tc) {
                // depthwise + pointwise fused: stencil from staged shared memory -> tcgen05 GEMM (tc_conv.cuh)
                const int N = pw.cout;
                const DwGeom geo = dw_geometry(C, N, ih, iw, S);
                if (geo.rows == 0) throw CudaFail{cudaErrorInvalidConfiguration, "dw_geometry: layer does not fit shared memory", __FILE__, __LINE__};
                std::vector<float> bias;
                int Kpad = 0;
                std::vector<__half> img = pack_tc_weights({&pw}, bias, Kpad, geo.nsplit);
                size_t oimg = B.add_weights_h(img), obp = B.add_weights(bias);
                int tpw = B.tensor("mobilenet0_relu" + std::to_string(i + 1) + "_fwd", oh, ow_, N);
                Step s;
                s.name = fmt("tc_dw%d+pw%d_s%d_%dto%d", i, i + 1, S, C, N);
                s.in = {tin}; s.out = {tpw};
                s.flops_per_img = 2.0 * oh * ow_ * C * 9 + 2.0 * oh * ow_ * C * N;
                s.bytes_per_img = ((double)ih * iw * C + (double)oh * ow_ * N) * es;
                // large maps (> 56x56 outputs; measured: no gain below): 2-D tiles (tc_dwpw2d.cuh) -- half the staged halo, no position
                // table, vertical reuse
                const bool tiles2d = oh * ow_ > 56 * 56 && C >= 16 && C <= 64 && geo.nsplit == 1 && !(h->cfg.flags & RF_FLAG_DW_1D);
                if (tiles2d) s.name = fmt("tc2d_dw%d+pw%d_s%d_%dto%d", i, i + 1, S, C, N);
                s.launch = [=](int n, cudaStream_t st) {
                    if (tiles2d) {
                        TcDw2dArgs a{};
                        a.in = T_(tin); a.C = C; a.nimg = n; a.IH = ih; a.IW = iw; a.OH = oh; a.OW = ow_; a.S = S; a.N = N;
                        a.TH = 8;
                        const int t16 = (ow_ + 15) / 16, t14 = (ow_ + 13) / 14;
                        a.TW = t14 < t16 ? 14 : 16;
                        tc_dw2d_finish(a);
                        a.wimg = h->d_weights_h + oimg; a.bias = Wd(obp); a.dw_w = Wd(owd); a.dw_b = Wd(obd); a.out = T_(tpw);
                        launch_tc_dwpw_2d(a, st);
                        return;
                    }
                    TcDwArgs a{};
                    a.in = T_(tin); a.C = C; a.nimg = n; a.IH = ih; a.IW = iw; a.OH = oh; a.OW = ow_; a.S = S;
                    a.N = N / geo.nsplit; a.Ntotal = N; a.Kpad = Kpad; a.rows = geo.rows; a.Wp = iw + 2; a.Hp = ih + 1; a.Rmax = geo.Rmax;
                    a.wimg = h->d_weights_h + oimg; a.bias = Wd(obp); a.dw_w = Wd(owd); a.dw_b = Wd(obd); a.out = T_(tpw);
                    launch_tc_dwpw(a, geo.nsplit, st);
                };
                B.step(std::move(s));
                cur = tpw; cur_h = oh; cur_w = ow_; cur_c = N;
                if (i + 1 == 10) c1 = cur;
                if (i + 1 == 22) c2 = cur;
                if (i + 1 == 26) c3 = cur;
                continue;
            }
        }
        int tdw = B.tensor("mobilenet0_relu" + std::to_string(i) + "_fwd", oh, ow_, C);
        {
            Step s;
            s.name = fmt("dw%d_3x3s%d_c%d", i, S, C);
            s.in = {tin}; s.out = {tdw};
            s.flops_per_img = 2.0 * oh * ow_ * C * 9;
            s.bytes_per_img = ((double)ih * iw * C + (double)oh * ow_ * C) * es;
            s.launch = [=](int n, cudaStream_t st) {
                long total = (long)n * oh * ow_ * (C / 8);
                unsigned g = (unsigned)((total + 255) / 256);
                if (S == 1) launch_k(k_dw3x3<T, 1>, dim3(g), dim3(256), 0, st, (const T *)T_(tin), T_(tdw), Wd(owd), Wd(obd), n, ih, iw, C);
                else launch_k(k_dw3x3<T, 2>, dim3(g), dim3(256), 0, st, (const T *)T_(tin), T_(tdw), Wd(owd), Wd(obd), n, ih, iw, C);
            };
            B.step(std::move(s));
        }
        std::vector<float> bias;
        std::vector<float> wk = pack_gemm({&pw}, bias);
        size_t owp = B.add_weights(wk), obp = B.add_weights(bias);
        const int N = pw.cout;
        int tpw = B.tensor("mobilenet0_relu" + std::to_string(i + 1) + "_fwd", oh, ow_, N);
        {
            Step s;
            s.name = fmt("pw%d_1x1_%dto%d", i + 1, C, N);
            s.in = {tdw}; s.out = {tpw};
            s.flops_per_img = 2.0 * oh * ow_ * C * N;
            s.bytes_per_img = ((double)oh * ow_ * C + (double)oh * ow_ * N) * es;
            s.launch = [=](int n, cudaStream_t st) {
                OutSplit<T> o{T_(tpw), N, N, 1, nullptr, 0, 0};
                launch_gemm<T>(T_(tdw), C, C, Wd(owp), Wd(obp), N, 1, o, n, oh, ow_, st);
            };
            B.step(std::move(s));
        }
        cur = tpw; cur_h = oh; cur_w = ow_; cur_c = N;
        if (i + 1 == 10) c1 = cur;
        if (i + 1 == 22) c2 = cur;
        if (i + 1 == 26) c3 = cur;
    }
    (void)cur_c;

    // ---- FPN + SSH (prototxt:1199-2302) -----------------------------------------------------
    auto conv_step = [&](const std::string &sname, std::vector<const FoldedConv *> cs, int tin, int ih, int iw,
                         int t0, int ld0, int off0, int n0, int relu0, int t1, int ld1, int off1, int relu1, int lane = 0,
                         int tup = -1, int up_which = 0) {
        if constexpr (std::is_same<T, __half>::value) {
            if (h->use_tc) {
                std::vector<float> bias;
                int Kpad = 0;
                std::vector<__half> img = pack_tc_weights(cs, bias, Kpad);
                size_t oimg = B.add_weights_h(img), ob = B.add_weights(bias);
                const int N = (int)bias.size(), cin = cs[0]->cin, ks = cs[0]->k;
                size_t oup = tup >= 0 ? B.add_weights(m.up_w[up_which]) : 0;
                Step s;
                s.name = "tc_" + sname;
                s.lane = lane;
                s.in = {tin};
                if (tup >= 0) s.in.push_back(tup);
                s.out = {t0};
                if (t1 >= 0) s.out.push_back(t1);
                s.flops_per_img = 2.0 * ih * iw * cin * ks * ks * N + (tup >= 0 ? 2.0 * ih * iw * cin * 4 : 0.0);
                s.bytes_per_img = ((double)ih * iw * cin + (double)ih * iw * N + (tup >= 0 ? (double)(ih / 2) * (iw / 2) * cin : 0.0)) * es;
                s.launch = [=](int n, cudaStream_t st) {
                    TcConvArgs a{};
                    a.in = T_(tin); a.Cin = cin; a.nimg = n; a.H = ih; a.W = iw; a.taps = ks * ks; a.N = N;
                    a.Wp = ks == 3 ? iw + 2 : iw; a.Hp = ks == 3 ? ih + 1 : ih;
                    a.R = (ks == 3 ? 128 + 2 * (iw + 3) : 128) | 1;
                    a.wimg = h->d_weights_h + oimg; a.bias = Wd(ob);
                    a.out = TcOut{T_(t0) + off0, ld0, n0, relu0, t1 >= 0 ? T_(t1) + off1 : nullptr, ld1, relu1};
                    if (tup >= 0) { a.up = T_(tup); a.up_w = Wd(oup); a.Cmax = (((a.R / a.Wp + 2) / 2 + 3) * (iw / 2)) | 1; }
                    launch_tc_conv(a, st);
                };
                B.step(std::move(s));
                return;
            }
        }
        std::vector<float> bias;
        std::vector<float> wk = pack_gemm(cs, bias);
        size_t ow = B.add_weights(wk), ob = B.add_weights(bias);
        const int N = (int)bias.size(), cin = cs[0]->cin, ks = cs[0]->k;
        const int ldin = h->tensors[tin].c;
        Step s;
        s.name = sname;
        s.lane = lane;
        s.in = {tin};
        s.out = {t0};
        if (t1 >= 0) s.out.push_back(t1);
        s.flops_per_img = 2.0 * ih * iw * cin * ks * ks * N;
        s.bytes_per_img = ((double)ih * iw * cin + (double)ih * iw * N) * es;
        s.launch = [=](int n, cudaStream_t st) {
            OutSplit<T> o{T_(t0) + off0, ld0, n0, relu0, t1 >= 0 ? T_(t1) + off1 : nullptr, ld1, relu1};
            launch_gemm<T>(T_(tin), ldin, cin, Wd(ow), Wd(ob), N, ks, o, n, ih, iw, st);
        };
        B.step(std::move(s));
    };
    auto ssh = [&](const std::string &lvname, int tin, int fh, int fw, int level, int lane) {
        const std::string p = "rf_" + lvname + "_det";
        int cat = B.tensor(p + "_concat_relu", fh, fw, 64);
        int ctx1 = B.tensor(p + "_context_conv1_relu", fh, fw, 16);
        int ctx31 = B.tensor(p + "_context_conv3_1_relu", fh, fw, 16);
        // det_conv1 (64->32, BN, ReLU after concat) + context_conv1 (64->16, BN, ReLU): one launch
        conv_step("ssh_" + lvname + "_conv1+ctx1_3x3_64to48", {&m.conv(p + "_conv1"), &m.conv(p + "_context_conv1")}, tin, fh,
                  fw, cat, 64, 0, 32, 1, ctx1, 16, 0, 1, lane);
        // context_conv2 (16->16 -> concat[32:48]) + context_conv3_1 (16->16, ReLU): one launch
        conv_step("ssh_" + lvname + "_ctx2+ctx3_1_3x3_16to32", {&m.conv(p + "_context_conv2"), &m.conv(p + "_context_conv3_1")},
                  ctx1, fh, fw, cat, 64, 32, 16, 1, ctx31, 16, 0, 1, lane);
        // context_conv3_2 (16->16 -> concat[48:64])
        conv_step("ssh_" + lvname + "_ctx3_2_3x3_16to16", {&m.conv(p + "_context_conv3_2")}, ctx31, fh, fw, cat, 64, 48, 16, 1,
                  -1, 0, 0, 0, lane);
        h->feat_tensor[level] = cat;
        // the concat tensor is written by three steps: make it live from the first of them
    };
    auto upadd = [&](const std::string &name, int tlat, int tup, int fh, int fw, int which) {
        size_t ow = B.add_weights(m.up_w[which]);
        int out = B.tensor(name, fh, fw, 64);
        Step s;
        s.name = "upsample_add" + name;
        s.in = {tlat, tup}; s.out = {out};
        s.flops_per_img = 2.0 * fh * fw * 64 * 4;
        s.bytes_per_img = ((double)fh * fw * 64 * 2 + (double)(fh / 2) * (fw / 2) * 64) * es;
        s.launch = [=](int n, cudaStream_t st) {
            long total = (long)n * fh * fw * 8;
            launch_k(k_upsample_add<T>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, (const T *)T_(tlat), (const T *)T_(tup), T_(out), Wd(ow), n,
                     fh, fw, 64, fh / 2, fw / 2);
        };
        B.step(std::move(s));
        return out;
    };
    const int h32 = H / 32, w32 = W / 32, h16 = H / 16, w16 = W / 16, h8 = H / 8, w8 = W / 8;
    // Lanes: the forward graph is not a chain.  rf_c1_red_conv only needs C1 and rf_c2_lateral only C2,
    // so they run on side lanes while the backbone continues; each level's SSH head runs on a side lane
    // while the main lane walks the top-down path lat3 -> aggr2 -> aggr1 -> ssh_c1 (the critical path).
    // In TC mode the FPN merge (deconv-upsample + add) is fused into the aggr conv's staging.
    const bool lanes = h->use_tc;
    // A side-lane step may start as soon as its producer finishes, i.e. EARLIER than later main-lane steps:
    // the step list (which the arena's liveness analysis walks in order) must show it right after that
    // producer, otherwise its output could be placed on memory a concurrently running main step still uses.
    auto move_last_step_after_producer = [&](int tensor_id) {
        int pos = 0;
        for (int i = (int)h->steps.size() - 2; i >= 0 && !pos; i--)
            for (int t : h->steps[i].out) if (t == tensor_id) { pos = i + 1; break; }
        Step st = std::move(h->steps.back());
        h->steps.pop_back();
        h->steps.insert(h->steps.begin() + pos, std::move(st));
    };
    int lat3 = B.tensor("rf_c3_lateral_relu", h32, w32, 64);
    int lat2 = B.tensor("rf_c2_lateral_relu", h16, w16, 64);
    int lat1 = B.tensor("rf_c1_red_conv_relu", h8, w8, 64);
    conv_step("c1_red_1x1_64to64", {&m.conv("rf_c1_red_conv")}, c1, h8, w8, lat1, 64, 0, 64, 1, -1, 0, 0, 0, lanes ? 1 : 0);
    if (lanes) move_last_step_after_producer(c1);
    conv_step("c2_lateral_1x1_128to64", {&m.conv("rf_c2_lateral")}, c2, h16, w16, lat2, 64, 0, 64, 1, -1, 0, 0, 0, lanes ? 2 : 0);
    if (lanes) move_last_step_after_producer(c2);
    conv_step("c3_lateral_1x1_256to64", {&m.conv("rf_c3_lateral")}, c3, h32, w32, lat3, 64, 0, 64, 1, -1, 0, 0, 0);
    ssh("c3", lat3, h32, w32, 0, lanes ? 1 : 0);
    int aggr2 = B.tensor("rf_c2_aggr_relu", h16, w16, 64);
    if (h->use_tc) {
        conv_step("c2_upsample+add+aggr_3x3_64to64", {&m.conv("rf_c2_aggr")}, lat2, h16, w16, aggr2, 64, 0, 64, 1, -1, 0, 0, 0, 0, lat3, 0);
    } else {
        int plus0 = upadd("_plus0", lat2, lat3, h16, w16, 0);
        conv_step("c2_aggr_3x3_64to64", {&m.conv("rf_c2_aggr")}, plus0, h16, w16, aggr2, 64, 0, 64, 1, -1, 0, 0, 0);
    }
    ssh("c2", aggr2, h16, w16, 1, lanes ? 2 : 0);
    int aggr1 = B.tensor("rf_c1_aggr_relu", h8, w8, 64);
    // Fusing the merge into the aggr conv costs ~50 KB of shared memory: fine while the conv's tiles fit one
    // wave (c2 level), a loss once it forces a second wave (c1 level at batch 8: 207 tiles, 1 CTA/SM).
    const long c1_tiles = ((long)h->cfg.max_batch * (h8 + 1) * (w8 + 2) + 127) / 128;
    if (h->use_tc && c1_tiles <= 148) {
        conv_step("c1_upsample+add+aggr_3x3_64to64", {&m.conv("rf_c1_aggr")}, lat1, h8, w8, aggr1, 64, 0, 64, 1, -1, 0, 0, 0, 0, aggr2, 1);
    } else if (h->use_tc) {
        if constexpr (std::is_same<T, __half>::value) {
            std::vector<__half> uwh(16 * 64);
            for (int c = 0; c < 64; c++)
                for (int t = 0; t < 16; t++) uwh[t * 64 + c] = __float2half(m.up_w[1][c * 16 + t]);
            size_t ouw = B.add_weights_h(uwh);
            int plus1 = B.tensor("_plus1", h8, w8, 64);
            Step s;
            s.name = "fpn_merge_c1_upsample+add_h2";
            s.in = {lat1, aggr2}; s.out = {plus1};
            s.flops_per_img = 2.0 * h8 * w8 * 64 * 4;
            s.bytes_per_img = ((double)h8 * w8 * 64 * 2 + (double)(h8 / 2) * (w8 / 2) * 64) * es;
            s.launch = [=](int n, cudaStream_t st) {
                long total = (long)n * h8 * w8 * 8;
                launch_k(k_fpn_merge_h2, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, (const __half *)T_(lat1), (const __half *)T_(aggr2),
                         (__half *)T_(plus1), (const __half *)(h->d_weights_h + ouw), n, h8, w8, 64);
            };
            B.step(std::move(s));
            conv_step("c1_aggr_3x3_64to64", {&m.conv("rf_c1_aggr")}, plus1, h8, w8, aggr1, 64, 0, 64, 1, -1, 0, 0, 0);
        }
    } else {
        int plus1 = upadd("_plus1", lat1, aggr2, h8, w8, 1);
        conv_step("c1_aggr_3x3_64to64", {&m.conv("rf_c1_aggr")}, plus1, h8, w8, aggr1, 64, 0, 64, 1, -1, 0, 0, 0);
    }
    ssh("c1", aggr1, h8, w8, 2, 0);

    // ---- predictors + decode (fused) and NMS -------------------------------------------------
    size_t hw_off[3], hb_off[3];
    const int strides[3] = {32, 16, 8};
    for (int l = 0; l < 3; l++) {
        std::string st = "_stride" + std::to_string(strides[l]);
        const FoldedConv *cs[3] = {&m.conv("face_rpn_cls_score" + st), &m.conv("face_rpn_bbox_pred" + st),
                                   &m.conv("face_rpn_landmark_pred" + st)};
        std::vector<float> w(32 * 64), b(32);
        int r = 0;
        for (auto c : cs)
            for (int o = 0; o < c->cout; o++, r++) {
                b[r] = c->b[o];
                for (int ci = 0; ci < 64; ci++) w[r * 64 + ci] = c->w[(size_t)o * 64 + ci];
            }
        hw_off[l] = B.add_weights(w);
        hb_off[l] = B.add_weights(b);
    }
    {
        Step s;
        s.name = "heads_1x1+softmax+decode_all_levels";
        s.in = {h->feat_tensor[0], h->feat_tensor[1], h->feat_tensor[2]};
        double px = (double)h32 * w32 + (double)h16 * w16 + (double)h8 * w8;
        s.flops_per_img = 2.0 * px * 64 * 4;   // threshold-first: only cls logits are computed for every pixel
        s.bytes_per_img = px * 64 * es;
        int f0 = h->feat_tensor[0], f1 = h->feat_tensor[1], f2 = h->feat_tensor[2];
        size_t w0 = hw_off[0], w1 = hw_off[1], w2 = hw_off[2], b0 = hb_off[0], b1 = hb_off[1], b2 = hb_off[2];
        s.launch = [=](int n, cudaStream_t st) {
            const T *feat[3] = {T_(f0), T_(f1), T_(f2)};
            HeadWeights hws[3] = {{Wd(w0), Wd(b0), 1.f}, {Wd(w1), Wd(b1), 1.f}, {Wd(w2), Wd(b2), 1.f}};
            launch_head_decode<T>(feat, hws, h->lv, n, W, H, h->d_params, h->pb, h->blobs_in_plan ? h->d_blobs : nullptr, st);
        };
        h->head_step = (int)h->steps.size();
        B.step(std::move(s));
    }
    {
        Step s;
        s.name = "sort+nms";
        s.flops_per_img = 0;
        s.bytes_per_img = 0;
        s.launch = [=](int n, cudaStream_t st) { launch_nms(n, h->d_params, h->pb, st); };
        B.step(std::move(s));
    }
}

// =============================================================================================
// INT8 plan (RF_PREC_INT8): same graph, int8 activations with the calibration table's scales.
// The integer scheme is restated in oracle/mnet_int8.py (the checker); tensor scales are looked up by
// the Caffe top name each tensor carries.
// =============================================================================================
struct QWeights { std::vector<int8_t> img; std::vector<float> mult, bq; };

// cs: convs sharing an input, concatenated along N.  s_out[n]: quantisation scale of output channel n.
// Image: nsplit slices of N/nsplit channels, each [taps * groups][Ns][16] with `groups` 16-channel groups per tap
// (zero padded beyond cin).
QWeights pack_tc_weights_i8(const std::vector<const FoldedConv *> &cs, float s_in, const std::vector<float> &s_out, int groups, int nsplit) {
    const int cin = cs[0]->cin, k = cs[0]->k, taps = k * k;
    int N = 0;
    for (auto c : cs) N += c->cout;
    const int Ns = N / nsplit;
    QWeights q;
    q.img.assign((size_t)taps * groups * 16 * N, 0);
    q.mult.resize(N); q.bq.resize(N);
    int n0 = 0;
    for (auto c : cs) {
        const size_t per = (size_t)cin * taps;
        for (int o = 0; o < c->cout; o++) {
            const int n = n0 + o, sl = n / Ns, nl = n % Ns;
            float mx = 0.f;
            for (size_t i = 0; i < per; i++) mx = std::max(mx, std::fabs(c->w[o * per + i]));
            const float sw = mx > 0.f ? mx / 127.0f : 1.0f;
            q.mult[n] = (float)((double)s_in * (double)sw / (double)s_out[n]);
            q.bq[n] = (float)((double)c->b[o] / (double)s_out[n]);
            for (int ci = 0; ci < cin; ci++)
                for (int t = 0; t < taps; t++) {
                    double v = std::nearbyint((double)c->w[((size_t)o * cin + ci) * taps + t] / (double)sw);
                    v = std::max(-127.0, std::min(127.0, v));
                    const int g = t * groups + ci / 16;
                    q.img[(size_t)sl * taps * groups * 16 * Ns + ((size_t)g * Ns + nl) * 16 + (ci % 16)] = (int8_t)v;
                }
        }
        n0 += c->cout;
    }
    return q;
}

void launch_tc_conv_i8(const TcConvArgsI8 &a, cudaStream_t s) {
    const long P = (long)a.nimg * a.Hp * a.Wp;
    const dim3 grid((unsigned)((P + 127) / 128));
    const size_t smem = tc_conv_i8_smem_bytes(a);
#define RF_I8C(NT_) if (a.up) launch_k(k_tc_conv_staged_i8<NT_, true>, grid, dim3(TC_THREADS), smem, s, a); else launch_k(k_tc_conv_staged_i8<NT_, false>, grid, dim3(TC_THREADS), smem, s, a)
    switch (tc_tmem_cols(a.N)) {
        case 32: RF_I8C(32); break;
        case 64: RF_I8C(64); break;
        case 128: RF_I8C(128); break;
        default: RF_I8C(256); break;
    }
#undef RF_I8C
}
void launch_tc_dwpw_2d_i8(const TcDw2dArgsI8 &a, cudaStream_t s) {
    const dim3 grid((unsigned)a.tiles_x, (unsigned)a.tiles_y, (unsigned)a.nimg);
    const size_t smem = tc_dw2d_i8_smem_bytes(a);
    switch (tc_tmem_cols(a.N)) {
        case 32: launch_k(k_tc_dwpw_2d_i8<32>, grid, dim3(TC_THREADS), smem, s, a); break;
        case 64: launch_k(k_tc_dwpw_2d_i8<64>, grid, dim3(TC_THREADS), smem, s, a); break;
        case 128: launch_k(k_tc_dwpw_2d_i8<128>, grid, dim3(TC_THREADS), smem, s, a); break;
        default: launch_k(k_tc_dwpw_2d_i8<256>, grid, dim3(TC_THREADS), smem, s, a); break;
    }
}

void launch_tc_dwpw_i8(const TcDwArgsI8 &a, int nsplit, cudaStream_t s) {
    const long M = (long)a.nimg * a.OH * a.OW;
    const dim3 grid((unsigned)((M + a.rows - 1) / a.rows), nsplit);
    const size_t smem = tc_dw_i8_smem_bytes(a);
    switch (tc_tmem_cols(a.N)) {
        case 32: if (a.C >= 64) launch_k(k_tc_dwpw_staged_i8<32, true>, grid, dim3(TC_THREADS), smem, s, a); else launch_k(k_tc_dwpw_staged_i8<32, false>, grid, dim3(TC_THREADS), smem, s, a); break;
        case 64: if (a.C >= 64) launch_k(k_tc_dwpw_staged_i8<64, true>, grid, dim3(TC_THREADS), smem, s, a); else launch_k(k_tc_dwpw_staged_i8<64, false>, grid, dim3(TC_THREADS), smem, s, a); break;
        case 128: if (a.C >= 64) launch_k(k_tc_dwpw_staged_i8<128, true>, grid, dim3(TC_THREADS), smem, s, a); else launch_k(k_tc_dwpw_staged_i8<128, false>, grid, dim3(TC_THREADS), smem, s, a); break;
        default: if (a.C >= 64) launch_k(k_tc_dwpw_staged_i8<256, true>, grid, dim3(TC_THREADS), smem, s, a); else launch_k(k_tc_dwpw_staged_i8<256, false>, grid, dim3(TC_THREADS), smem, s, a); break;
    }
}
cudaError_t tc_init_i8() {
    cudaError_t e;
#define RF_TC_ATTR(K_) if ((e = cudaFuncSetAttribute(K_, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT))) return e
    RF_TC_ATTR((k_tc_conv_staged_i8<32, false>)); RF_TC_ATTR((k_tc_conv_staged_i8<64, false>)); RF_TC_ATTR((k_tc_conv_staged_i8<128, false>)); RF_TC_ATTR((k_tc_conv_staged_i8<256, false>));
    RF_TC_ATTR((k_tc_conv_staged_i8<32, true>)); RF_TC_ATTR((k_tc_conv_staged_i8<64, true>)); RF_TC_ATTR((k_tc_conv_staged_i8<128, true>)); RF_TC_ATTR((k_tc_conv_staged_i8<256, true>));
    RF_TC_ATTR((k_tc_dwpw_staged_i8<32, true>)); RF_TC_ATTR((k_tc_dwpw_staged_i8<64, true>)); RF_TC_ATTR((k_tc_dwpw_staged_i8<128, true>)); RF_TC_ATTR((k_tc_dwpw_staged_i8<256, true>));
    RF_TC_ATTR((k_tc_dwpw_staged_i8<32, false>)); RF_TC_ATTR((k_tc_dwpw_staged_i8<64, false>)); RF_TC_ATTR((k_tc_dwpw_staged_i8<128, false>)); RF_TC_ATTR((k_tc_dwpw_staged_i8<256, false>));
    RF_TC_ATTR(k_tc_dwpw_2d_i8<32>); RF_TC_ATTR(k_tc_dwpw_2d_i8<64>); RF_TC_ATTR(k_tc_dwpw_2d_i8<128>); RF_TC_ATTR(k_tc_dwpw_2d_i8<256>);
#undef RF_TC_ATTR
    return cudaSuccess;
}

DwGeom dw_geometry_i8(int C, int N, int IH, int IW, int S) {
    const int OH = IH / S, OW = IW / S, Wp = IW + 2, Hp = IH + 1, Kpad = (C + 31) / 32 * 32;
    auto centre = [&](long m) { long ox = m % OW, oy = (m / OW) % OH, b = m / ((long)OW * OH); return (b * Hp + oy * S) * Wp + ox * S + 1; };
    for (int rows : {128, 64}) {
        if (rows == 128 && OH * OW <= 28 * 28) continue;
        for (int nsplit : {1, 2, 4}) {
            if ((N / nsplit) % 16) continue;
            long g = rows, t = (long)OH * OW;
            while (t) { long u = g % t; g = t; t = u; }
            const long M = ((long)rows / g + 1) * OH * OW;
            int R = 0;
            for (long m0 = 0; m0 < M; m0 += rows) {
                long ml = std::min(m0 + rows, M) - 1;
                R = std::max(R, (int)(centre(ml) - centre(m0) + 2 * (Wp + 1) + 1));
            }
            R |= 1;
            TcDwArgsI8 a{};
            a.C = C; a.Rmax = R; a.Kpad = Kpad; a.N = N / nsplit; a.rows = rows;
            if (R <= TC_MAX_R && tc_dw_i8_smem_bytes(a) <= (size_t)TC_SMEM_LIMIT) return {rows, nsplit, R};
        }
    }
    return {0, 0, 0};
}

void build_plan_i8(rf_handle h) {
    Builder B{h, h->cfg.net_h, h->cfg.net_w};
    const Model &m = h->model;
    const int H = h->cfg.net_h, W = h->cfg.net_w;
    auto Q_ = [h](int id) { return reinterpret_cast<int8_t *>(h->tptr(id)); };
    auto Wd = [h](size_t off) { return h->d_weights + off; };
    auto scale_of = [h](const std::string &name) -> float {
        auto it = h->int8_scales.find(name);
        if (it == h->int8_scales.end()) { h->err = "INT8 calibration table lacks the scale of tensor '" + name + "'"; throw CudaFail{cudaErrorInvalidValue, "INT8 calibration table lookup", __FILE__, __LINE__}; }
        return it->second;
    };
    auto tscale = [&](int id) { return scale_of(h->tensors[id].name); };

    // ---- stem: FP32 inside, output quantised with s(relu2) ---------------------------------------------------
    int cur_h = H / 2, cur_w = W / 2;
    int cur = B.tensor("mobilenet0_relu2_fwd", cur_h, cur_w, 16);
    {
        const FoldedConv &c0 = m.conv("mobilenet0_conv0_fwd"), &dw = m.conv("mobilenet0_conv1_fwd"), &pw = m.conv("mobilenet0_conv2_fwd");
        std::vector<float> w0(27 * 8), wd(72), wp(128);
        for (int o = 0; o < 8; o++)
            for (int cb = 0; cb < 3; cb++)
                for (int t = 0; t < 9; t++) w0[(t * 3 + cb) * 8 + o] = c0.w[((size_t)o * 3 + (2 - cb)) * 9 + t];
        for (int c = 0; c < 8; c++)
            for (int t = 0; t < 9; t++) wd[t * 8 + c] = dw.w[(size_t)c * 9 + t];
        for (int o = 0; o < 16; o++)
            for (int c = 0; c < 8; c++) wp[c * 16 + o] = pw.w[(size_t)o * 8 + c];
        size_t ow0 = B.add_weights(w0), ob0 = B.add_weights(c0.b), owd = B.add_weights(wd), obd = B.add_weights(dw.b),
               owp = B.add_weights(wp), obp = B.add_weights(pw.b);
        const float inv = 1.0f / tscale(cur);
        int out = cur;
        Step s;
        s.name = "stem_conv0+dw1+pw2_u8_to_16ch_i8";
        s.out = {out};
        s.flops_per_img = 2.0 * cur_h * cur_w * (8 * 27 + 8 * 9 + 8 * 16);
        s.bytes_per_img = (double)H * W * 3 + (double)cur_h * cur_w * 16;
        // conv0 on tensor cores, depthwise + pointwise in FP32 on CUDA cores (stem_tc.cuh, OutT = int8_t); RF_FLAG_SIMT_STEM:
        // all three layers on CUDA cores (k_stem)
        const bool simt_stem = (h->cfg.flags & (RF_FLAG_SIMT_STEM | RF_FLAG_NO_TENSORCORE)) != 0;
        size_t oblob = B.add_weights_h(make_stem_blob(w0, c0.b, wd, dw.b, wp, pw.b));
        if (!simt_stem) s.name = "tc_stem_conv0+dw1+pw2_u8_to_16ch_i8";
        s.launch = [=](int n, cudaStream_t st) {
            if (simt_stem) {
                StemWeights sw{Wd(ow0), Wd(ob0), Wd(owd), Wd(obd), Wd(owp), Wd(obp)};
                const int tiles = ((H / 2 + 15) / 16) * ((W / 2 + 15) / 16);
                launch_k(k_stem<int8_t>, dim3((unsigned)(tiles * n)), dim3(256), 0, st, (const PostParams *)h->d_params, Q_(out), sw, n, H, W, inv);
            } else {
                StemTcArgs a{reinterpret_cast<const unsigned char *>(h->d_weights_h + oblob)};
                launch_k(k_stem_tc<int8_t>, dim3((unsigned)((W / 2 + 15) / 16), (unsigned)((H / 2 + 15) / 16), (unsigned)n), dim3(256), 0, st,
                         (const PostParams *)h->d_params, Q_(out), a, n, H, W, inv);
            }
        };
        B.step(std::move(s));
    }
    // ---- 12 x (depthwise + pointwise) --------------------------------------------------------------------------
    int c1 = -1, c2 = -1, c3 = -1;
    for (int i = 3; i <= 26; i += 2) {
        const FoldedConv &dw = m.conv("mobilenet0_conv" + std::to_string(i) + "_fwd");
        const FoldedConv &pw = m.conv("mobilenet0_conv" + std::to_string(i + 1) + "_fwd");
        const int C = dw.cout, S = dw.stride, N = pw.cout;
        const int ih = cur_h, iw = cur_w, oh = cur_h / S, ow_ = cur_w / S;
        const float s_in = tscale(cur), s_mid = scale_of("mobilenet0_relu" + std::to_string(i) + "_fwd");
        std::vector<float> wd(9 * C);
        for (int c = 0; c < C; c++)
            for (int t = 0; t < 9; t++) wd[t * C + c] = dw.w[(size_t)c * 9 + t] * s_in;     // float32 product, as the oracle
        size_t owd = B.add_weights(wd), obd = B.add_weights(dw.b);
        const DwGeom geo = dw_geometry_i8(C, N, ih, iw, S);
        if (geo.rows == 0) throw CudaFail{cudaErrorInvalidConfiguration, "dw_geometry_i8: layer does not fit shared memory", __FILE__, __LINE__};
        int tin = cur;
        int tpw = B.tensor("mobilenet0_relu" + std::to_string(i + 1) + "_fwd", oh, ow_, N);
        const int Kpad = (C + 31) / 32 * 32;
        std::vector<float> s_out(N, tscale(tpw));
        QWeights q = pack_tc_weights_i8({&pw}, s_mid, s_out, Kpad / 16, geo.nsplit);
        size_t oimg = B.add_weights_q(q.img), omul = B.add_weights(q.mult), obq = B.add_weights(q.bq);
        const float inv_mid = 1.0f / s_mid;
        Step s;
        s.name = fmt("i8_dw%d+pw%d_s%d_%dto%d", i, i + 1, S, C, N);
        s.in = {tin}; s.out = {tpw};
        s.flops_per_img = 2.0 * oh * ow_ * C * 9 + 2.0 * oh * ow_ * C * N;
        s.bytes_per_img = (double)ih * iw * C + (double)oh * ow_ * N;
        const bool tiles2d = oh * ow_ > 56 * 56 && C >= 16 && C <= 64 && geo.nsplit == 1 && !(h->cfg.flags & RF_FLAG_DW_1D);   // as the FP16 plan
        if (tiles2d) s.name = fmt("i8_2d_dw%d+pw%d_s%d_%dto%d", i, i + 1, S, C, N);
        s.launch = [=](int n, cudaStream_t st) {
            if (tiles2d) {
                TcDw2dArgsI8 a{};
                a.in = Q_(tin); a.C = C; a.nimg = n; a.IH = ih; a.IW = iw; a.OH = oh; a.OW = ow_; a.S = S; a.N = N; a.Kpad = Kpad;
                a.TH = 8;
                a.TW = (ow_ + 13) / 14 < (ow_ + 15) / 16 ? 14 : 16;
                tc_dw2d_i8_finish(a);
                a.wimg = h->d_weights_q + oimg; a.mult = Wd(omul); a.bq = Wd(obq); a.dw_w = Wd(owd); a.dw_b = Wd(obd); a.inv_mid = inv_mid;
                a.out = Q_(tpw);
                launch_tc_dwpw_2d_i8(a, st);
                return;
            }
            TcDwArgsI8 a{};
            a.in = Q_(tin); a.C = C; a.nimg = n; a.IH = ih; a.IW = iw; a.OH = oh; a.OW = ow_; a.S = S;
            a.N = N / geo.nsplit; a.Ntotal = N; a.Kpad = Kpad; a.rows = geo.rows; a.Wp = iw + 2; a.Hp = ih + 1; a.Rmax = geo.Rmax;
            a.wimg = h->d_weights_q + oimg; a.mult = Wd(omul); a.bq = Wd(obq); a.dw_w = Wd(owd); a.dw_b = Wd(obd); a.inv_mid = inv_mid;
            a.out = Q_(tpw);
            launch_tc_dwpw_i8(a, geo.nsplit, st);
        };
        B.step(std::move(s));
        cur = tpw; cur_h = oh; cur_w = ow_;
        if (i + 1 == 10) c1 = cur;
        if (i + 1 == 22) c2 = cur;
        if (i + 1 == 26) c3 = cur;
    }
    // ---- FPN + SSH ------------------------------------------------------------------------------------------------
    auto conv_step = [&](const std::string &sname, std::vector<const FoldedConv *> cs, int tin, int ih, int iw, int t0, int ld0, int off0,
                         int n0, int relu0, int t1, int ld1, int off1, int relu1, int lane, int tup, int up_which, int tlat_for_up) {
        (void)tlat_for_up;
        const int cin = cs[0]->cin, ks = cs[0]->k;
        int N = 0;
        for (auto c : cs) N += c->cout;
        std::vector<float> s_out(N);
        for (int n = 0; n < N; n++) s_out[n] = n < n0 ? tscale(t0) : tscale(t1);
        // with the FPN merge fused in, the conv's input tensor is the (never materialised) sum: its scale is the table's
        const float s_in = tup >= 0 ? scale_of(up_which == 0 ? "_plus0" : "_plus1") : tscale(tin);
        QWeights q = pack_tc_weights_i8(cs, s_in, s_out, tc_i8_gs(cin), 1);
        size_t oimg = B.add_weights_q(q.img), omul = B.add_weights(q.mult), obq = B.add_weights(q.bq);
        size_t oup = 0;
        float lat_mul = 0.f;
        if (tup >= 0) {
            std::vector<float> wq(16 * cin);
            const float s_up = tscale(tup);
            for (int c = 0; c < cin; c++)
                for (int t = 0; t < 16; t++) wq[t * cin + c] = (float)((double)m.up_w[up_which][c * 16 + t] * (double)s_up / (double)s_in);
            oup = B.add_weights(wq);
            lat_mul = (float)((double)tscale(tin) / (double)s_in);
        }
        Step s;
        s.name = "i8_" + sname;
        s.lane = lane;
        s.in = {tin};
        if (tup >= 0) s.in.push_back(tup);
        s.out = {t0};
        if (t1 >= 0) s.out.push_back(t1);
        s.flops_per_img = 2.0 * ih * iw * cin * ks * ks * N;
        s.bytes_per_img = (double)ih * iw * cin + (double)ih * iw * N + (tup >= 0 ? (double)(ih / 2) * (iw / 2) * cin : 0.0);
        s.launch = [=](int n, cudaStream_t st) {
            TcConvArgsI8 a{};
            a.in = Q_(tin); a.Cin = cin; a.nimg = n; a.H = ih; a.W = iw; a.taps = ks * ks; a.N = N;
            a.Wp = ks == 3 ? iw + 2 : iw; a.Hp = ks == 3 ? ih + 1 : ih;
            a.R = (ks == 3 ? 128 + 2 * (iw + 3) : 128) | 1;
            a.wimg = h->d_weights_q + oimg; a.mult = Wd(omul); a.bq = Wd(obq);
            a.out = TcOutI8{Q_(t0) + off0, ld0, n0, relu0, t1 >= 0 ? Q_(t1) + off1 : nullptr, ld1, relu1};
            if (tup >= 0) { a.up = Q_(tup); a.up_wq = Wd(oup); a.lat_mul = lat_mul; a.Cmax = (((a.R / a.Wp + 2) / 2 + 3) * (iw / 2)) | 1; }
            launch_tc_conv_i8(a, st);
        };
        B.step(std::move(s));
    };
    auto move_last_step_after_producer = [&](int tensor_id) {
        int pos = 0;
        for (int i = (int)h->steps.size() - 2; i >= 0 && !pos; i--)
            for (int t : h->steps[i].out) if (t == tensor_id) { pos = i + 1; break; }
        Step st = std::move(h->steps.back());
        h->steps.pop_back();
        h->steps.insert(h->steps.begin() + pos, std::move(st));
    };
    auto ssh = [&](const std::string &lvname, int tin, int fh, int fw, int level, int lane) {
        const std::string p = "rf_" + lvname + "_det";
        int cat = B.tensor(p + "_concat_relu", fh, fw, 64);
        int ctx1 = B.tensor(p + "_context_conv1_relu", fh, fw, 16);
        int ctx31 = B.tensor(p + "_context_conv3_1_relu", fh, fw, 16);
        conv_step("ssh_" + lvname + "_conv1+ctx1_3x3_64to48", {&m.conv(p + "_conv1"), &m.conv(p + "_context_conv1")}, tin, fh, fw, cat, 64, 0, 32, 1,
                  ctx1, 16, 0, 1, lane, -1, 0, -1);
        conv_step("ssh_" + lvname + "_ctx2+ctx3_1_3x3_16to32", {&m.conv(p + "_context_conv2"), &m.conv(p + "_context_conv3_1")}, ctx1, fh, fw, cat,
                  64, 32, 16, 1, ctx31, 16, 0, 1, lane, -1, 0, -1);
        conv_step("ssh_" + lvname + "_ctx3_2_3x3_16to16", {&m.conv(p + "_context_conv3_2")}, ctx31, fh, fw, cat, 64, 48, 16, 1, -1, 0, 0, 0, lane, -1,
                  0, -1);
        h->feat_tensor[level] = cat;
    };
    const int h32 = H / 32, w32 = W / 32, h16 = H / 16, w16 = W / 16, h8 = H / 8, w8 = W / 8;
    int lat3 = B.tensor("rf_c3_lateral_relu", h32, w32, 64);
    int lat2 = B.tensor("rf_c2_lateral_relu", h16, w16, 64);
    int lat1 = B.tensor("rf_c1_red_conv_relu", h8, w8, 64);
    conv_step("c1_red_1x1_64to64", {&m.conv("rf_c1_red_conv")}, c1, h8, w8, lat1, 64, 0, 64, 1, -1, 0, 0, 0, 1, -1, 0, -1);
    move_last_step_after_producer(c1);
    conv_step("c2_lateral_1x1_128to64", {&m.conv("rf_c2_lateral")}, c2, h16, w16, lat2, 64, 0, 64, 1, -1, 0, 0, 0, 2, -1, 0, -1);
    move_last_step_after_producer(c2);
    conv_step("c3_lateral_1x1_256to64", {&m.conv("rf_c3_lateral")}, c3, h32, w32, lat3, 64, 0, 64, 1, -1, 0, 0, 0, 0, -1, 0, -1);
    ssh("c3", lat3, h32, w32, 0, 1);
    int aggr2 = B.tensor("rf_c2_aggr_relu", h16, w16, 64);
    conv_step("c2_upsample+add+aggr_3x3_64to64", {&m.conv("rf_c2_aggr")}, lat2, h16, w16, aggr2, 64, 0, 64, 1, -1, 0, 0, 0, 0, lat3, 0, lat2);
    ssh("c2", aggr2, h16, w16, 1, 2);
    int aggr1 = B.tensor("rf_c1_aggr_relu", h8, w8, 64);
    const long c1_tiles = ((long)h->cfg.max_batch * (h8 + 1) * (w8 + 2) + 127) / 128;
    if (c1_tiles <= 148) {
        conv_step("c1_upsample+add+aggr_3x3_64to64", {&m.conv("rf_c1_aggr")}, lat1, h8, w8, aggr1, 64, 0, 64, 1, -1, 0, 0, 0, 0, aggr2, 1, lat1);
    } else {
        int plus1 = B.tensor("_plus1", h8, w8, 64);
        const float s_out = tscale(plus1), s_up = tscale(aggr2), s_lat = tscale(lat1);
        std::vector<float> wq(16 * 64);
        for (int c = 0; c < 64; c++)
            for (int t = 0; t < 16; t++) wq[t * 64 + c] = (float)((double)m.up_w[1][c * 16 + t] * (double)s_up / (double)s_out);
        size_t owq = B.add_weights(wq);
        const float lat_mul = (float)((double)s_lat / (double)s_out);
        Step s;
        s.name = "i8_fpn_merge_c1_upsample+add";
        s.in = {lat1, aggr2}; s.out = {plus1};
        s.flops_per_img = 2.0 * h8 * w8 * 64 * 4;
        s.bytes_per_img = (double)h8 * w8 * 64 * 2 + (double)(h8 / 2) * (w8 / 2) * 64;
        s.launch = [=](int n, cudaStream_t st) {
            long total = (long)n * h8 * w8 * 4;
            launch_k(k_fpn_merge_i8, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, (const int8_t *)Q_(lat1), (const int8_t *)Q_(aggr2), Q_(plus1),
                     Wd(owq), lat_mul, n, h8, w8, 64);
        };
        B.step(std::move(s));
        conv_step("c1_aggr_3x3_64to64", {&m.conv("rf_c1_aggr")}, plus1, h8, w8, aggr1, 64, 0, 64, 1, -1, 0, 0, 0, 0, -1, 0, -1);
    }
    ssh("c1", aggr1, h8, w8, 2, 0);
    // ---- predictors + decode (FP32 on the dequantised concat tensors) and NMS -------------------------------------
    size_t hw_off[3], hb_off[3];
    float hs[3];
    const int strides[3] = {32, 16, 8};
    for (int l = 0; l < 3; l++) {
        std::string st = "_stride" + std::to_string(strides[l]);
        const FoldedConv *cs[3] = {&m.conv("face_rpn_cls_score" + st), &m.conv("face_rpn_bbox_pred" + st), &m.conv("face_rpn_landmark_pred" + st)};
        std::vector<float> w(32 * 64), b(32);
        int r = 0;
        for (auto c : cs)
            for (int o = 0; o < c->cout; o++, r++) {
                b[r] = c->b[o];
                for (int ci = 0; ci < 64; ci++) w[r * 64 + ci] = c->w[(size_t)o * 64 + ci];
            }
        hw_off[l] = B.add_weights(w);
        hb_off[l] = B.add_weights(b);
        hs[l] = tscale(h->feat_tensor[l]);
    }
    {
        Step s;
        s.name = "i8_heads_1x1+softmax+decode_all_levels";
        s.in = {h->feat_tensor[0], h->feat_tensor[1], h->feat_tensor[2]};
        double px = (double)h32 * w32 + (double)h16 * w16 + (double)h8 * w8;
        s.flops_per_img = 2.0 * px * 64 * 4;
        s.bytes_per_img = px * 64;
        int f0 = h->feat_tensor[0], f1 = h->feat_tensor[1], f2 = h->feat_tensor[2];
        size_t w0 = hw_off[0], w1 = hw_off[1], w2 = hw_off[2], b0 = hb_off[0], b1 = hb_off[1], b2 = hb_off[2];
        float s0 = hs[0], s1 = hs[1], s2 = hs[2];
        s.launch = [=](int n, cudaStream_t st) {
            const int8_t *feat[3] = {Q_(f0), Q_(f1), Q_(f2)};
            HeadWeights hws[3] = {{Wd(w0), Wd(b0), s0}, {Wd(w1), Wd(b1), s1}, {Wd(w2), Wd(b2), s2}};
            launch_head_decode<int8_t>(feat, hws, h->lv, n, W, H, h->d_params, h->pb, h->blobs_in_plan ? h->d_blobs : nullptr, st);
        };
        h->head_step = (int)h->steps.size();
        B.step(std::move(s));
    }
    {
        Step s;
        s.name = "sort+nms";
        s.launch = [=](int n, cudaStream_t st) { launch_nms(n, h->d_params, h->pb, st); };
        B.step(std::move(s));
    }
}

// Cross-lane dependencies: a step waits (event) for the producers of its inputs that live in another lane.
void link_steps(rf_handle h) {
    auto &st = h->steps;
    for (int i = 0; i < (int)st.size(); i++) {
        st[i].deps.clear();
        for (int t : st[i].in) {
            for (int j = i - 1; j >= 0; j--) {
                bool writes = false;
                for (int o : st[j].out) writes |= (o == t);
                if (!writes) continue;
                if (st[j].lane != st[i].lane) {
                    if (std::find(st[i].deps.begin(), st[i].deps.end(), j) == st[i].deps.end()) st[i].deps.push_back(j);
                    st[j].signals = true;
                }
                break;   // the last writer before i (its lane orders earlier writers of the same tensor)
            }
        }
    }
}

// Liveness-based first-fit placement of activation tensors in one arena.  Steps on side lanes run
// concurrently with later main-lane steps: every tensor such a step touches stays live until the
// first step of another lane that waits for its lane (the join), so no concurrent writer can land on it.
void place_tensors(rf_handle h, bool keep_all) {
    auto &ts = h->tensors;
    auto &st = h->steps;
    for (int si = 0; si < (int)st.size(); si++) {
        for (int t : st[si].out) { if (ts[t].first < 0) ts[t].first = si; ts[t].last = std::max(ts[t].last, si); }
        for (int t : st[si].in) ts[t].last = std::max(ts[t].last, si);
    }
    for (int k = 0; k < (int)st.size(); k++) {
        if (st[k].lane == 0) continue;
        int join = (int)st.size() - 1;
        for (int j = k + 1; j < (int)st.size() && join == (int)st.size() - 1; j++)
            for (int d : st[j].deps)
                if (st[j].lane != st[k].lane && st[d].lane == st[k].lane && d >= k) { join = j; break; }
        for (int t : st[k].in) ts[t].last = std::max(ts[t].last, join);
        for (int t : st[k].out) ts[t].last = std::max(ts[t].last, join);
    }
    // a main-lane step that runs while a side lane is still reading must not overwrite those inputs either:
    // covered above because the side step's inputs stay live until the join.
    const size_t B = (size_t)h->cfg.max_batch;
    size_t top = 0;
    std::vector<int> order(ts.size());
    for (size_t i = 0; i < ts.size(); i++) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ts[a].first < ts[b].first; });
    std::vector<int> placed;
    for (int id : order) {
        size_t sz = (ts[id].bytes_per_img * B + 255) / 256 * 256;
        size_t off = 0;
        if (keep_all) {
            off = top;
        } else {
            bool moved = true;
            while (moved) {
                moved = false;
                for (int p : placed) {
                    bool live_overlap = !(ts[p].last < ts[id].first || ts[id].last < ts[p].first);
                    size_t psz = (ts[p].bytes_per_img * B + 255) / 256 * 256;
                    bool mem_overlap = off < ts[p].offset + psz && ts[p].offset < off + sz;
                    if (live_overlap && mem_overlap) { off = ts[p].offset + psz; moved = true; }
                }
            }
        }
        ts[id].offset = off;
        top = std::max(top, off + sz);
        placed.push_back(id);
    }
    h->arena_bytes = top;
}

// Issue the forward pass: lane 0 on `s`, side lanes on their own streams, joined by events.  Works both
// under stream capture (the side streams fork from / join into the capturing stream) and eagerly.
void run_steps(rf_handle h, int n, cudaStream_t s, bool use_lanes = true) {
    for (size_t i = 0; i < h->steps.size(); i++) {
        Step &st = h->steps[i];
        cudaStream_t cs = (use_lanes && st.lane) ? h->lane_stream[st.lane] : s;
        if (use_lanes)
            for (int d : st.deps) CK(cudaStreamWaitEvent(cs, h->step_event[d], 0));
        st.launch(n, cs);
        if (use_lanes && st.signals) CK(cudaEventRecord(h->step_event[i], cs));
    }
}

void forward_graph(rf_handle h, int n) {
    if (h->cfg.flags & RF_FLAG_NO_GRAPH) {
        run_steps(h, n, h->stream);
        CK(cudaGetLastError());
        return;
    }
    auto it = h->graphs.find(n);
    if (it == h->graphs.end()) {
        cudaGraph_t g = nullptr;
        CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        run_steps(h, n, h->stream);
        cudaError_t e = cudaStreamEndCapture(h->stream, &g);
        if (e != cudaSuccess) throw CudaFail{e, "cudaStreamEndCapture", __FILE__, __LINE__};
        cudaGraphExec_t ge = nullptr;
        e = cudaGraphInstantiate(&ge, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) throw CudaFail{e, "cudaGraphInstantiate", __FILE__, __LINE__};
        it = h->graphs.emplace(n, ge).first;
    }
    CK(cudaGraphLaunch(it->second, h->stream));
}

// Run parameters travel through a small ring of pinned slots so that an asynchronous caller
// (rf_detect_batch_device) can queue several runs without overwriting a copy still in flight.
void set_params(rf_handle h, float thr, float nms, const uint8_t *input = nullptr) {
    PostParams *slot = h->h_params + (h->param_seq++ % rf_handle_s::kParamSlots);
    slot->score_thr = thr;
    slot->nms_thr = nms;
    slot->input = input ? input : h->d_input;
    h->cur_thr = thr;
    h->cur_nms = nms;
    CK(cudaMemcpyAsync(h->d_params, slot, sizeof(PostParams), cudaMemcpyHostToDevice, h->stream));
}

void destroy(rf_handle h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->saved.empty()) h->saved.resize(1);
    for (int c = 0; c < (int)h->saved.size(); c++) {
        switch_ctx(h, c);
        if (h->stream) cudaStreamSynchronize(h->stream);
        for (auto &g : h->graphs) cudaGraphExecDestroy(g.second);
        h->graphs.clear();
        cudaFree(h->arena); cudaFree(h->d_params); cudaFreeHost(h->h_params);
        cudaFree(h->pb.cand_keys); cudaFree(h->pb.cand_recs); cudaFree(h->pb.cand_count); cudaFree(h->pb.sort_scratch);
        cudaFree(h->pb.flag_scratch); cudaFree(h->pb.out_dets); cudaFree(h->pb.out_counts); cudaFree(h->pb.out_total_kept);
        for (auto e : h->step_event) if (e) cudaEventDestroy(e);
        for (int l = 1; l < 3; l++) if (h->lane_stream[l]) cudaStreamDestroy(h->lane_stream[l]);
        if (h->fence) cudaEventDestroy(h->fence);
        if (h->stream) cudaStreamDestroy(h->stream);
    }
    cudaFree(h->pb_merge.cand_keys); cudaFree(h->pb_merge.cand_recs); cudaFree(h->pb_merge.cand_count); cudaFree(h->pb_merge.sort_scratch);
    cudaFree(h->pb_merge.flag_scratch); cudaFree(h->pb_merge.out_dets); cudaFree(h->pb_merge.out_counts); cudaFree(h->pb_merge.out_total_kept);
    cudaFree(h->d_weights); cudaFree(h->d_weights_h); cudaFree(h->d_weights_q); cudaFree(h->d_input); cudaFree(h->d_raw);
    for (auto p : h->d_blobs) cudaFree(p);
    cudaFreeHost(h->h_input); cudaFreeHost(h->h_raw); cudaFreeHost(h->h_dets); cudaFreeHost(h->h_counts);
    for (auto &sl : h->slots) {
        cudaFree(sl.d_in); cudaFreeHost(sl.h_in); cudaFreeHost(sl.h_dets); cudaFreeHost(sl.h_counts);
        if (sl.ev_h2d) cudaEventDestroy(sl.ev_h2d);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    delete h;
}

int check_n(rf_handle h, int n) {
    if (!h) return RF_ERR_INVALID_ARG;
    if (n < 0) return fail(h, RF_ERR_INVALID_ARG, "negative batch size");
    if (n > h->cfg.max_batch) return fail(h, RF_ERR_CAPACITY, fmt("batch %d exceeds max_batch %d", n, h->cfg.max_batch));
    return RF_OK;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int rf_abi_version(void) { return RF_B200_ABI_VERSION; }

const char *rf_build_info(void) {
    return "librf_b200 (RetinaFace mnet25 detect path) built for sm_100a, CUDA " RF_STR(CUDART_VERSION);
}

const char *rf_status_string(int s) {
    switch (s) {
        case RF_OK: return "ok";
        case RF_ERR_INVALID_ARG: return "invalid argument";
        case RF_ERR_IO: return "i/o error";
        case RF_ERR_MODEL: return "model error";
        case RF_ERR_CUDA: return "CUDA error";
        case RF_ERR_NO_DEVICE: return "no usable CUDA device (the library has no CPU path)";
        case RF_ERR_CAPACITY: return "capacity exceeded";
        case RF_ERR_UNSUPPORTED: return "unsupported";
    }
    return "unknown status";
}

const char *rf_last_error(rf_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int rf_create(const rf_config *cfg, rf_handle *out) {
    if (out) *out = nullptr;
    if (!cfg || !out) return fail(nullptr, RF_ERR_INVALID_ARG, "rf_create: cfg and out must be non-NULL");
    if (!cfg->caffemodel_path) return fail(nullptr, RF_ERR_INVALID_ARG, "rf_create: caffemodel_path is NULL");
    if (cfg->net_w <= 0 || cfg->net_h <= 0 || cfg->net_w % 32 || cfg->net_h % 32)
        return fail(nullptr, RF_ERR_INVALID_ARG, fmt("rf_create: net size %dx%d must be positive multiples of 32", cfg->net_w, cfg->net_h));
    if (cfg->max_batch <= 0 || cfg->max_batch > 4096) return fail(nullptr, RF_ERR_INVALID_ARG, "rf_create: max_batch must be in [1, 4096]");
    if (cfg->precision != RF_PREC_FP32 && cfg->precision != RF_PREC_FP16 && cfg->precision != RF_PREC_INT8)
        return fail(nullptr, RF_ERR_INVALID_ARG, "rf_create: unknown precision");
    if (cfg->precision == RF_PREC_INT8 && !cfg->int8_table_path)
        return fail(nullptr, RF_ERR_INVALID_ARG, "rf_create: RF_PREC_INT8 needs int8_table_path (the TensorRT calibration cache of this caffemodel)");

    std::unique_ptr<rf_handle_s, void (*)(rf_handle)> H(new rf_handle_s, destroy);
    rf_handle h = H.get();
    h->cfg = *cfg;
    h->caffemodel = cfg->caffemodel_path;
    h->cfg.caffemodel_path = h->caffemodel.c_str();
    if (cfg->int8_table_path) { h->table = cfg->int8_table_path; h->cfg.int8_table_path = h->table.c_str(); }
    if (h->cfg.max_faces <= 0) h->cfg.max_faces = 256;
    if (h->cfg.max_faces > 8192) return fail(nullptr, RF_ERR_INVALID_ARG, "rf_create: max_faces must be <= 8192");
    if (h->cfg.max_image_w <= 0) h->cfg.max_image_w = h->cfg.net_w;
    if (h->cfg.max_image_h <= 0) h->cfg.max_image_h = h->cfg.net_h;
    h->cfg.max_image_w = std::max(h->cfg.max_image_w, h->cfg.net_w);
    h->cfg.max_image_h = std::max(h->cfg.max_image_h, h->cfg.net_h);
    h->device = cfg->device;
    h->elem = cfg->precision == RF_PREC_FP32 ? 4 : (cfg->precision == RF_PREC_FP16 ? 2 : 1);

    // ---- model (host) --------------------------------------------------------------------
    {
        std::vector<RawLayer> layers;
        std::string err;
        bool io = false;
        if (!read_caffemodel(h->caffemodel, layers, err, io)) return fail(nullptr, io ? RF_ERR_IO : RF_ERR_MODEL, err);
        if (!build_mnet_model(layers, h->model, err)) return fail(nullptr, RF_ERR_MODEL, err);
        if (!h->table.empty() && !read_int8_table(h->table, h->int8_scales, err)) return fail(nullptr, RF_ERR_IO, err);
    }
    // ---- device ----------------------------------------------------------------------------
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, RF_ERR_NO_DEVICE, fmt("no CUDA device (%s); librf_b200 has no CPU path", e == cudaSuccess ? "count 0" : cudaGetErrorString(e)));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, RF_ERR_INVALID_ARG, fmt("device %d out of range (%d devices)", cfg->device, ndev));
    try {
        CK(cudaSetDevice(h->device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, h->device));
        if (prop.major != 10)
            return fail(nullptr, RF_ERR_NO_DEVICE, fmt("device %d is sm_%d%d; librf_b200 is built for sm_100a only", h->device, prop.major, prop.minor));
        CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&h->ev0));
        CK(cudaEventCreate(&h->ev1));
        CK(postproc_init());

        const int Hn = h->cfg.net_h, Wn = h->cfg.net_w, Bm = h->cfg.max_batch;
        // levels / anchors
        const int strides[3] = {32, 16, 8};
        int abase = 0, pbase = 0;
        for (int l = 0; l < 3; l++) {
            LevelDesc &lv = h->lv[l];
            lv.stride = strides[l]; lv.h = Hn / strides[l]; lv.w = Wn / strides[l];
            lv.anchor_base = abase; lv.pix_base = pbase;
            base_anchors_net3(strides[l], lv.base);
            abase += 2 * lv.h * lv.w; pbase += lv.h * lv.w;
        }
        const int A = abase;
        int ap2 = 1;
        while (ap2 < A) ap2 <<= 1;

        h->use_tc = h->cfg.precision == RF_PREC_INT8 || (h->cfg.precision == RF_PREC_FP16 && !(h->cfg.flags & RF_FLAG_NO_TENSORCORE));
        if (h->use_tc) CK(tc_init());
        if (h->cfg.precision == RF_PREC_INT8) { CK(tc_init_i8()); build_plan_i8(h); }
        else if (h->cfg.precision == RF_PREC_FP32) build_plan<float>(h);
        else build_plan<__half>(h);
        link_steps(h);
        place_tensors(h, false);
        CK(cudaMalloc(&h->d_weights, h->wstage.size() * sizeof(float)));
        CK(cudaMemcpy(h->d_weights, h->wstage.data(), h->wstage.size() * sizeof(float), cudaMemcpyHostToDevice));
        if (!h->wstage_h.empty()) {
            CK(cudaMalloc(&h->d_weights_h, h->wstage_h.size() * sizeof(__half)));
            CK(cudaMemcpy(h->d_weights_h, h->wstage_h.data(), h->wstage_h.size() * sizeof(__half), cudaMemcpyHostToDevice));
        }
        if (!h->wstage_q.empty()) {
            CK(cudaMalloc(&h->d_weights_q, h->wstage_q.size()));
            CK(cudaMemcpy(h->d_weights_q, h->wstage_q.data(), h->wstage_q.size(), cudaMemcpyHostToDevice));
        }
        const size_t in_bytes = (size_t)Bm * Hn * Wn * 3;
        CK(cudaMalloc(&h->d_input, in_bytes));
        CK(cudaHostAlloc(&h->h_input, in_bytes, cudaHostAllocDefault));
        h->raw_bytes = (size_t)h->cfg.max_image_w * h->cfg.max_image_h * 3;
        CK(cudaMalloc(&h->d_raw, h->raw_bytes));
        CK(cudaHostAlloc(&h->h_raw, h->raw_bytes, cudaHostAllocDefault));
        CK(cudaHostAlloc(&h->h_dets, sizeof(rf_det) * (size_t)Bm * h->cfg.max_faces, cudaHostAllocDefault));
        CK(cudaHostAlloc(&h->h_counts, sizeof(int) * 2 * Bm, cudaHostAllocDefault));
        // ---- per-context resources ----
        h->nctx = h->cfg.streams <= 0 ? 4 : std::min(h->cfg.streams, 4);
        h->saved.resize(h->nctx);
        for (int c = 0; c < h->nctx; c++) {
            switch_ctx(h, c);
            if (c > 0) CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));   // context 0 keeps the stream created above
            for (int l = 1; l < 3; l++) CK(cudaStreamCreateWithFlags(&h->lane_stream[l], cudaStreamNonBlocking));
            h->step_event.assign(h->steps.size(), nullptr);
            for (size_t i = 0; i < h->steps.size(); i++)
                if (h->steps[i].signals) CK(cudaEventCreateWithFlags(&h->step_event[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&h->fence, cudaEventDisableTiming));
            CK(cudaMalloc(&h->arena, h->arena_bytes));
            CK(cudaMalloc(&h->d_params, sizeof(PostParams)));
            CK(cudaHostAlloc(&h->h_params, sizeof(PostParams) * rf_handle_s::kParamSlots, cudaHostAllocDefault));
            PostBuffers &pb = h->pb;
            pb.anchors_per_image = A; pb.anchors_pow2 = ap2; pb.max_faces = h->cfg.max_faces;
            CK(cudaMalloc(&pb.cand_keys, sizeof(unsigned long long) * (size_t)Bm * A));
            CK(cudaMalloc(&pb.cand_recs, sizeof(rf_det) * (size_t)Bm * A));
            CK(cudaMalloc(&pb.cand_count, sizeof(int) * Bm));
            CK(cudaMemset(pb.cand_count, 0, sizeof(int) * Bm));
            CK(cudaMalloc(&pb.sort_scratch, sizeof(unsigned long long) * (size_t)Bm * ap2));
            CK(cudaMalloc(&pb.flag_scratch, (size_t)Bm * ap2));
            CK(cudaMalloc(&pb.out_dets, sizeof(rf_det) * (size_t)Bm * pb.max_faces));
            CK(cudaMalloc(&pb.out_counts, sizeof(int) * Bm));
            CK(cudaMalloc(&pb.out_total_kept, sizeof(int) * Bm));
            CK(cudaMemset(pb.out_counts, 0, sizeof(int) * Bm));
        }
        switch_ctx(h, 0);
        for (int l = 0; l < 3; l++) {
            const int ch[3] = {4, 8, 20};
            for (int k = 0; k < 3; k++) h->blob_elems[3 * l + k] = (size_t)ch[k] * h->lv[l].h * h->lv[l].w;
        }
        CK(cudaDeviceSynchronize());
    } catch (const CudaFail &f) {
        return fail_cuda(nullptr, f);
    }
    *out = H.release();
    return RF_OK;
}

void rf_destroy(rf_handle h) { destroy(h); }

uint8_t *rf_pinned_input(rf_handle h) { return h ? h->h_input : nullptr; }
uint8_t *rf_device_input(rf_handle h) { return h ? h->d_input : nullptr; }

int rf_get_net_size(rf_handle h, int *net_w, int *net_h, int *max_batch, int *max_faces) {
    if (!h) return RF_ERR_INVALID_ARG;
    if (net_w) *net_w = h->cfg.net_w;
    if (net_h) *net_h = h->cfg.net_h;
    if (max_batch) *max_batch = h->cfg.max_batch;
    if (max_faces) *max_faces = h->cfg.max_faces;
    return RF_OK;
}
int rf_num_anchors(rf_handle h) { return h ? h->pb.anchors_per_image : RF_ERR_INVALID_ARG; }
void *rf_stream(rf_handle h) { if (!h) return nullptr; switch_ctx(h, 0); return (void *)h->stream; }

int rf_synchronize(rf_handle h) {
    if (!h) return RF_ERR_INVALID_ARG;
    try {
        CK(cudaSetDevice(h->device));
        const int keep = h->active;
        for (int c = 0; c < h->nctx; c++) { switch_ctx(h, c); CK(cudaStreamSynchronize(h->stream)); }
        switch_ctx(h, keep);
    } catch (const CudaFail &f) { return fail_cuda(h, f); }
    return RF_OK;
}

// Orders context 0's stream (the one rf_stream returns) after everything queued so far on every context.
int rf_fence(rf_handle h) {
    if (!h) return RF_ERR_INVALID_ARG;
    try {
        CK(cudaSetDevice(h->device));
        switch_ctx(h, 0);
        cudaStream_t s0 = h->stream;
        for (int c = 1; c < h->nctx; c++) {
            switch_ctx(h, c);
            CK(cudaEventRecord(h->fence, h->stream));
            CK(cudaStreamWaitEvent(s0, h->fence, 0));
        }
        switch_ctx(h, 0);
    } catch (const CudaFail &f) { return fail_cuda(h, f); }
    return RF_OK;
}

void *rf_last_stream(rf_handle h) { return h ? (void *)(h->last_stream ? h->last_stream : h->stream) : nullptr; }

int rf_launches_per_batch(rf_handle h, int n) {
    (void)n;
    return h ? (int)h->steps.size() : RF_ERR_INVALID_ARG;
}

int rf_detect_batch_device(rf_handle h, const uint8_t *dev_bgr, int n, float thr, float nms, const rf_det **dev_dets,
                           const int32_t **dev_counts) {
    int rc = check_n(h, n);
    if (rc) return rc;
    try {
        CK(cudaSetDevice(h->device));
        switch_ctx(h, (int)(h->next_dev_ctx++ % (unsigned)h->nctx));   // consecutive batches overlap on different contexts
        h->last_stream = h->stream;
        // the caller's device images are read in place (conv0 takes the pointer from the run parameters)
        if (h->param_seq && h->param_seq % rf_handle_s::kParamSlots == 0) CK(cudaStreamSynchronize(h->stream));
        set_params(h, thr, nms, dev_bgr);
        if (n > 0) forward_graph(h, n);
    } catch (const CudaFail &f) { return fail_cuda(h, f); }
    if (dev_dets) *dev_dets = h->pb.out_dets;
    if (dev_counts) *dev_counts = h->pb.out_counts;
    return RF_OK;
}

static int fetch_results(rf_handle h, int n, rf_face *out_faces, int *out_counts, int32_t *out_idx, int *out_ncand) {
    const int mf = h->cfg.max_faces;
    CK(cudaMemcpyAsync(h->h_counts, h->pb.out_counts, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(h->h_dets, h->pb.out_dets, sizeof(rf_det) * (size_t)n * mf, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n; i++) {
        int k = h->h_counts[i];
        if (out_counts) out_counts[i] = k;
        for (int j = 0; j < k; j++) {
            const rf_det &d = h->h_dets[(size_t)i * mf + j];
            if (out_faces) out_faces[(size_t)i * mf + j] = d.face;
            if (out_idx) out_idx[(size_t)i * mf + j] = d.anchor_index;
        }
    }
    (void)out_ncand;
    return RF_OK;
}

// One caller image of arbitrary size -> d_raw (packed rows) on the handle's stream.  Pinned sources (cudaHostAlloc /
// cudaHostRegister) are copied straight from the caller's memory, row stride and all, with no host synchronisation: the
// stream orders the copy behind the letter-box kernel that still reads the previous image.  Pageable sources go through
// the library's single pinned buffer (one host memcpy per image: ~10x the cost of the DMA itself).
static void upload_raw(rf_handle h, const uint8_t *src, int width, int height, int row_stride) {
    cudaPointerAttributes at{};
    const bool pinned = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost;
    if (pinned) {
        CK(cudaMemcpy2DAsync(h->d_raw, (size_t)width * 3, src, (size_t)row_stride, (size_t)width * 3, (size_t)height, cudaMemcpyHostToDevice, h->stream));
        return;
    }
    cudaGetLastError();
    CK(cudaStreamSynchronize(h->stream));  // h_raw is single-buffered
    for (int y = 0; y < height; y++) memcpy(h->h_raw + (size_t)y * width * 3, src + (size_t)y * row_stride, (size_t)width * 3);
    CK(cudaMemcpyAsync(h->d_raw, h->h_raw, (size_t)width * height * 3, cudaMemcpyHostToDevice, h->stream));
}

int rf_detect_batch(rf_handle h, const uint8_t *const *imgs, const int *widths, const int *heights, const int *row_strides,
                    int n, float thr, float nms, rf_face *out_faces, int *out_counts, int32_t *out_idx) {
    int rc = check_n(h, n);
    if (rc) return rc;
    if (n == 0) return RF_OK;
    if (!imgs || !widths || !heights) return fail(h, RF_ERR_INVALID_ARG, "rf_detect_batch: NULL image arrays");
    const int Hn = h->cfg.net_h, Wn = h->cfg.net_w;
    const size_t img_bytes = (size_t)Hn * Wn * 3;
    try {
        CK(cudaSetDevice(h->device));
        switch_ctx(h, 0);
        // Network-sized packed images are copied H2D straight from the caller's memory when it is
        // pinned (cudaHostAlloc / cudaHostRegister / the library's own rf_pinned_input), otherwise via the
        // library's pinned mirror; runs of adjacent sources collapse into one copy.  Other sizes are
        // letter-boxed on the GPU one by one (preprocess.cuh).
        const uint8_t *run_src = nullptr;
        int run_start = -1, run_len = 0;
        auto flush = [&]() {
            if (run_start < 0) return;
            CK(cudaMemcpyAsync(h->d_input + (size_t)run_start * img_bytes, run_src, (size_t)run_len * img_bytes,
                               cudaMemcpyHostToDevice, h->stream));
            run_start = -1;
        };
        bool staging_dirty = false;
        for (int i = 0; i < n; i++) {
            if (!imgs[i] || widths[i] <= 0 || heights[i] <= 0) { return fail(h, RF_ERR_INVALID_ARG, fmt("rf_detect_batch: image %d is empty", i)); }
            const int rs = row_strides && row_strides[i] ? row_strides[i] : widths[i] * 3;
            if (widths[i] == Wn && heights[i] == Hn && rs == Wn * 3) {
                const uint8_t *src = imgs[i];
                const bool in_mirror = src >= h->h_input && src < h->h_input + (size_t)h->cfg.max_batch * img_bytes;
                if (!in_mirror) {
                    cudaPointerAttributes at{};
                    bool pinned = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost;
                    if (!pinned) {
                        cudaGetLastError();
                        if (!staging_dirty) { CK(cudaStreamSynchronize(h->stream)); staging_dirty = true; }
                        uint8_t *slot = h->h_input + (size_t)i * img_bytes;
                        memcpy(slot, src, img_bytes);
                        src = slot;
                    }
                }
                if (run_start >= 0 && src == run_src + (size_t)run_len * img_bytes) { run_len++; }
                else { flush(); run_start = i; run_src = src; run_len = 1; }
            } else {
                flush();
                if (widths[i] > h->cfg.max_image_w || heights[i] > h->cfg.max_image_h)
                    return fail(h, RF_ERR_CAPACITY, fmt("image %d is %dx%d, larger than max_image %dx%d", i, widths[i], heights[i],
                                                        h->cfg.max_image_w, h->cfg.max_image_h));
                upload_raw(h, imgs[i], widths[i], heights[i], rs);
                launch_letterbox(h->d_raw, widths[i], heights[i], h->d_input + (size_t)i * img_bytes, Wn, Hn, h->stream);
            }
        }
        flush();
        set_params(h, thr, nms);
        forward_graph(h, n);
        fetch_results(h, n, out_faces, out_counts, out_idx, nullptr);
    } catch (const CudaFail &f) { return fail_cuda(h, f); }
    return RF_OK;
}

static void ensure_slots(rf_handle h) {
    if (h->copy_stream) return;
    const size_t in_bytes = (size_t)h->cfg.max_batch * h->cfg.net_h * h->cfg.net_w * 3;
    CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (auto &sl : h->slots) {
        CK(cudaMalloc(&sl.d_in, in_bytes));
        CK(cudaHostAlloc(&sl.h_in, in_bytes, cudaHostAllocDefault));
        CK(cudaHostAlloc(&sl.h_dets, sizeof(rf_det) * (size_t)h->cfg.max_batch * h->cfg.max_faces, cudaHostAllocDefault));
        CK(cudaHostAlloc(&sl.h_counts, sizeof(int) * h->cfg.max_batch, cudaHostAllocDefault));
        CK(cudaEventCreateWithFlags(&sl.ev_h2d, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
    }
}

int rf_submit_batch(rf_handle h, const uint8_t *const *imgs, int n, float thr, float nms, int *ticket) {
    int rc = check_n(h, n);
    if (rc) return rc;
    if (!imgs || !ticket || n == 0) return fail(h, RF_ERR_INVALID_ARG, "rf_submit_batch: NULL argument or empty batch");
    const size_t img_bytes = (size_t)h->cfg.net_h * h->cfg.net_w * 3;
    try {
        CK(cudaSetDevice(h->device));
        ensure_slots(h);
        switch_ctx(h, (int)(h->submit_seq % (unsigned)h->nctx));
        rf_handle_s::Slot &sl = h->slots[h->submit_seq % RF_PIPELINE_DEPTH];
        if (sl.busy) return fail(h, RF_ERR_CAPACITY, "rf_submit_batch: RF_PIPELINE_DEPTH batches already in flight; collect one first");
        // H2D on the copy stream: adjacent sources collapse into one copy
        const uint8_t *run_src = nullptr;
        int run_start = -1, run_len = 0;
        auto flush = [&]() {
            if (run_start < 0) return;
            CK(cudaMemcpyAsync(sl.d_in + (size_t)run_start * img_bytes, run_src, (size_t)run_len * img_bytes, cudaMemcpyHostToDevice,
                               h->copy_stream));
            run_start = -1;
        };
        for (int i = 0; i < n; i++) {
            if (!imgs[i]) return fail(h, RF_ERR_INVALID_ARG, fmt("rf_submit_batch: image %d is NULL", i));
            const uint8_t *src = imgs[i];
            cudaPointerAttributes at{};
            bool pinned = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost;
            if (!pinned) {
                cudaGetLastError();
                memcpy(sl.h_in + (size_t)i * img_bytes, src, img_bytes);   // slot is free: its previous H2D completed before collect
                src = sl.h_in + (size_t)i * img_bytes;
            }
            if (run_start >= 0 && src == run_src + (size_t)run_len * img_bytes) run_len++;
            else { flush(); run_start = i; run_src = src; run_len = 1; }
        }
        flush();
        CK(cudaEventRecord(sl.ev_h2d, h->copy_stream));
        CK(cudaStreamWaitEvent(h->stream, sl.ev_h2d, 0));
        if (h->param_seq && h->param_seq % rf_handle_s::kParamSlots == 0) CK(cudaStreamSynchronize(h->stream));
        set_params(h, thr, nms, sl.d_in);
        forward_graph(h, n);
        CK(cudaMemcpyAsync(sl.h_counts, h->pb.out_counts, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(sl.h_dets, h->pb.out_dets, sizeof(rf_det) * (size_t)n * h->cfg.max_faces, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaEventRecord(sl.ev_done, h->stream));
        sl.n = n;
        sl.busy = true;
        *ticket = (int)h->submit_seq++;
    } catch (const CudaFail &f) { return fail_cuda(h, f); }
    return RF_OK;
}

int rf_collect_batch(rf_handle h, int ticket, rf_face *out_faces, int *out_counts, int32_t *out_idx) {
    if (!h) return RF_ERR_INVALID_ARG;
    if ((unsigned)ticket != h->collect_seq) return fail(h, RF_ERR_INVALID_ARG, fmt("rf_collect_batch: ticket %d out of order (next is %u)", ticket, h->collect_seq));
    rf_handle_s::Slot &sl = h->slots[h->collect_seq % RF_PIPELINE_DEPTH];
    if (!sl.busy) return fail(h, RF_ERR_INVALID_ARG, "rf_collect_batch: nothing submitted under this ticket");
    try {
        CK(cudaSetDevice(h->device));
        CK(cudaEventSynchronize(sl.ev_done));
    } catch (const CudaFail &f) { return fail_cuda(h, f); }
    const int mf = h->cfg.max_faces;
    for (int i = 0; i < sl.n; i++) {
        const int k = sl.h_counts[i];
        if (out_counts) out_counts[i] = k;
        for (int j = 0; j < k; j++) {
            const rf_det &d = sl.h_dets[(size_t)i * mf + j];
            if (out_faces) out_faces[(size_t)i * mf + j] = d.face;
            if (out_idx) out_idx[(size_t)i * mf + j] = d.anchor_index;
        }
    }
    sl.busy = false;
    h->collect_seq++;
    return RF_OK;
}

static void ensure_merge_buffers(rf_handle h) {
    PostBuffers &pb = h->pb_merge;
    if (pb.cand_keys) return;
    const int A = h->cfg.max_batch * h->cfg.max_faces;       // every view may contribute max_faces candidates
    int ap2 = 1;
    while (ap2 < A) ap2 <<= 1;
    pb.anchors_per_image = A; pb.anchors_pow2 = ap2; pb.max_faces = h->cfg.max_faces;
    CK(cudaMalloc(&pb.cand_keys, sizeof(unsigned long long) * (size_t)A));
    CK(cudaMalloc(&pb.cand_recs, sizeof(rf_det) * (size_t)A));
    CK(cudaMalloc(&pb.cand_count, sizeof(int)));
    CK(cudaMemset(pb.cand_count, 0, sizeof(int)));
    CK(cudaMalloc(&pb.sort_scratch, sizeof(unsigned long long) * (size_t)ap2));
    CK(cudaMalloc(&pb.flag_scratch, (size_t)ap2));
    CK(cudaMalloc(&pb.out_dets, sizeof(rf_det) * (size_t)pb.max_faces));
    CK(cudaMalloc(&pb.out_counts, sizeof(int)));
    CK(cudaMalloc(&pb.out_total_kept, sizeof(int)));
    CK(cudaMemset(pb.out_counts, 0, sizeof(int)));
}

int rf_detect_views(rf_handle h, const uint8_t *bgr, int width, int height, int row_stride, const rf_view *views, int nviews, float thr,
                    float nms, rf_face *out_faces, int *out_count, int32_t *out_view_of, float *out_view_scales) {
    if (!h || !bgr || !views || !out_count || width <= 0 || height <= 0) return fail(h, RF_ERR_INVALID_ARG, "rf_detect_views: bad arguments");
    if (nviews < 1 || nviews > RF_MAX_VIEWS || nviews > h->cfg.max_batch)
        return fail(h, RF_ERR_CAPACITY, fmt("rf_detect_views: %d views, limit min(RF_MAX_VIEWS = %d, max_batch = %d)", nviews, RF_MAX_VIEWS, h->cfg.max_batch));
    if (width > h->cfg.max_image_w || height > h->cfg.max_image_h) return fail(h, RF_ERR_CAPACITY, "rf_detect_views: image larger than max_image");
    for (int v = 0; v < nviews; v++)
        if (!(views[v].shrink > 0.f && views[v].shrink <= 1.f)) return fail(h, RF_ERR_INVALID_ARG, fmt("rf_detect_views: view %d: shrink must be in (0, 1]", v));
    static_assert(RF_MAX_VIEWS == RF_MAX_VIEWS_DEV, "view capacity of the merge kernel");
    const int Hn = h->cfg.net_h, Wn = h->cfg.net_w, mf = h->cfg.max_faces;
    const size_t img_bytes = (size_t)Hn * Wn * 3;
    const int rs = row_stride ? row_stride : width * 3;
    try {
        CK(cudaSetDevice(h->device));
        switch_ctx(h, 0);
        ensure_merge_buffers(h);
        upload_raw(h, bgr, width, height, rs);
        ViewSet vs{};
        vs.nviews = nviews;
        vs.img_w_minus1 = (float)(width - 1);
        for (int v = 0; v < nviews; v++) {
            const int bw = std::max(1, (int)(Wn * views[v].shrink)), bh = std::max(1, (int)(Hn * views[v].shrink));
            vs.flip[v] = views[v].flip ? 1 : 0;
            vs.scale[v] = launch_letterbox_view(h->d_raw, width, height, h->d_input + (size_t)v * img_bytes, Wn, Hn, bw, bh, vs.flip[v], h->stream);
            if (out_view_scales) out_view_scales[v] = vs.scale[v];
        }
        set_params(h, thr, nms);
        forward_graph(h, nviews);
        launch_merge_views(h->pb, vs, h->pb_merge, h->stream);
        launch_nms(1, h->d_params, h->pb_merge, h->stream);
        CK(cudaMemcpyAsync(h->h_counts, h->pb_merge.out_counts, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(h->h_dets, h->pb_merge.out_dets, sizeof(rf_det) * (size_t)mf, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        const int k = h->h_counts[0];
        *out_count = k;
        for (int j = 0; j < k; j++) {
            if (out_faces) out_faces[j] = h->h_dets[j].face;
            if (out_view_of) out_view_of[j] = h->h_dets[j].anchor_index / mf;
        }
    } catch (const CudaFail &f) { return fail_cuda(h, f); }
    return RF_OK;
}

int rf_preprocess(rf_handle h, const uint8_t *bgr, int width, int height, int row_stride, uint8_t *out) {
    if (!h || !bgr || !out || width <= 0 || height <= 0) return fail(h, RF_ERR_INVALID_ARG, "rf_preprocess: bad arguments");
    if (width > h->cfg.max_image_w || height > h->cfg.max_image_h) return fail(h, RF_ERR_CAPACITY, "rf_preprocess: image larger than max_image");
    const int Hn = h->cfg.net_h, Wn = h->cfg.net_w;
    const int rs = row_stride ? row_stride : width * 3;
    try {
        CK(cudaSetDevice(h->device));
        switch_ctx(h, 0);
        upload_raw(h, bgr, width, height, rs);
        launch_letterbox(h->d_raw, width, height, h->d_input, Wn, Hn, h->stream);
        CK(cudaMemcpyAsync(h->h_input, h->d_input, (size_t)Hn * Wn * 3, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        memcpy(out, h->h_input, (size_t)Hn * Wn * 3);
    } catch (const CudaFail &f) { return fail_cuda(h, f); }
    return RF_OK;
}

static void ensure_blobs(rf_handle h) {
    if (h->d_blobs[0]) return;
    for (int i = 0; i < 9; i++) CK(cudaMalloc(&h->d_blobs[i], sizeof(float) * h->blob_elems[i] * h->cfg.max_batch));
}

int rf_forward_heads(rf_handle h, const uint8_t *bgr, int n, float *const heads_out[9]) {
    int rc = check_n(h, n);
    if (rc) return rc;
    if (n == 0) return RF_OK;
    if (!bgr || !heads_out) return fail(h, RF_ERR_INVALID_ARG, "rf_forward_heads: NULL argument");
    try {
        CK(cudaSetDevice(h->device));
        switch_ctx(h, 0);
        ensure_blobs(h);
        const size_t bytes = (size_t)n * h->cfg.net_h * h->cfg.net_w * 3;
        CK(cudaStreamSynchronize(h->stream));
        memcpy(h->h_input, bgr, bytes);
        CK(cudaMemcpyAsync(h->d_input, h->h_input, bytes, cudaMemcpyHostToDevice, h->stream));
        set_params(h, h->cur_thr, h->cur_nms);
        h->blobs_in_plan = true;
        try { run_steps(h, n, h->stream); } catch (...) { h->blobs_in_plan = false; throw; }
        h->blobs_in_plan = false;
        CK(cudaGetLastError());
        for (int i = 0; i < 9; i++)
            CK(cudaMemcpyAsync(heads_out[i], h->d_blobs[i], sizeof(float) * h->blob_elems[i] * n, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    } catch (const CudaFail &f) { return fail_cuda(h, f); }
    return RF_OK;
}

int rf_postprocess(rf_handle h, const float *const heads[9], int n, float thr, float nms, rf_face *out_faces, int *out_counts,
                   int32_t *out_idx, int *out_ncand) {
    int rc = check_n(h, n);
    if (rc) return rc;
    if (n == 0) return RF_OK;
    if (!heads) return fail(h, RF_ERR_INVALID_ARG, "rf_postprocess: NULL heads");
    try {
        CK(cudaSetDevice(h->device));
        switch_ctx(h, 0);
        ensure_blobs(h);
        for (int i = 0; i < 9; i++)
            CK(cudaMemcpyAsync(h->d_blobs[i], heads[i], sizeof(float) * h->blob_elems[i] * n, cudaMemcpyHostToDevice, h->stream));
        set_params(h, thr, nms);
        launch_blob_decode(h->d_blobs, h->lv, n, h->cfg.net_w, h->cfg.net_h, h->d_params, h->pb, h->stream);
        if (out_ncand) CK(cudaMemcpyAsync(h->h_counts + h->cfg.max_batch, h->pb.cand_count, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
        launch_nms(n, h->d_params, h->pb, h->stream);
        CK(cudaGetLastError());
        fetch_results(h, n, out_faces, out_counts, out_idx, nullptr);
        if (out_ncand) for (int i = 0; i < n; i++) out_ncand[i] = h->h_counts[h->cfg.max_batch + i];
    } catch (const CudaFail &f) { return fail_cuda(h, f); }
    return RF_OK;
}

// Debug / parity: any materialised activation by its Caffe top name, as NCHW float32.
int rf_debug_get_tensor(rf_handle h, const char *name, int n, float *out_nchw, int *c, int *hh, int *ww) {
    int rc = check_n(h, n);
    if (rc) return rc;
    auto it = h->tensor_by_name.find(name ? name : "");
    if (it == h->tensor_by_name.end()) return fail(h, RF_ERR_INVALID_ARG, fmt("unknown tensor '%s'", name ? name : "(null)"));
    const TensorInfo &t = h->tensors[it->second];
    if (c) *c = t.c;
    if (hh) *hh = t.h;
    if (ww) *ww = t.w;
    if (!out_nchw) return RF_OK;
    try {
        CK(cudaSetDevice(h->device));
        switch_ctx(h, 0);
        CK(cudaStreamSynchronize(h->stream));
        size_t elems = (size_t)n * t.h * t.w * t.c;
        std::vector<unsigned char> host(elems * h->elem);
        CK(cudaMemcpy(host.data(), h->tptr(it->second), host.size(), cudaMemcpyDeviceToHost));
        for (int b = 0; b < n; b++)
            for (int y = 0; y < t.h; y++)
                for (int x = 0; x < t.w; x++)
                    for (int ch = 0; ch < t.c; ch++) {
                        size_t src = (((size_t)b * t.h + y) * t.w + x) * t.c + ch;
                        float v = h->elem == 4 ? reinterpret_cast<float *>(host.data())[src]
                                : h->elem == 2 ? __half2float(reinterpret_cast<__half *>(host.data())[src])
                                               : (float)reinterpret_cast<int8_t *>(host.data())[src];   // INT8: raw quantised values
                        out_nchw[(((size_t)b * t.c + ch) * t.h + y) * t.w + x] = v;
                    }
    } catch (const CudaFail &f) { return fail_cuda(h, f); }
    return RF_OK;
}

// Re-places activations without buffer reuse so that rf_debug_get_tensor sees every tensor of
// the last forward (debug only; call before the first forward).
int rf_debug_keep_all(rf_handle h) {
    if (!h) return RF_ERR_INVALID_ARG;
    try {
        CK(cudaSetDevice(h->device));
        for (auto &t : h->tensors) { t.first = -1; t.last = -1; }
        place_tensors(h, true);
        for (int c = 0; c < h->nctx; c++) {
            switch_ctx(h, c);
            CK(cudaStreamSynchronize(h->stream));
            for (auto &g : h->graphs) cudaGraphExecDestroy(g.second);
            h->graphs.clear();
            CK(cudaFree(h->arena));
            h->arena = nullptr;
            CK(cudaMalloc(&h->arena, h->arena_bytes));
        }
        switch_ctx(h, 0);
    } catch (const CudaFail &f) { return fail_cuda(h, f); }
    return RF_OK;
}

// INT8 entropy calibration (replaces INT8-Calibration-Tool/calibrationtable.cpp:399-583).  `h` must be an
// RF_PREC_FP32 handle (its SIMT plan materialises every tensor the INT8 plan quantises, including the depthwise outputs
// and the FPN sums).  Two passes over the n network-sized images: absmax, then 2048-bin histograms; then the KL threshold
// search per tensor on the host; the table is written in the reference's TensorRT cache format.
int rf_calibrate_int8(rf_handle h, const uint8_t *bgr_net_sized, int n_images, const char *out_table_path) {
    if (!h || !bgr_net_sized || n_images <= 0 || !out_table_path) return fail(h, RF_ERR_INVALID_ARG, "rf_calibrate_int8: bad arguments");
    if (h->cfg.precision != RF_PREC_FP32) return fail(h, RF_ERR_UNSUPPORTED, "rf_calibrate_int8: create the handle with RF_PREC_FP32 (every tensor must be materialised)");
    const size_t img_bytes = (size_t)h->cfg.net_h * h->cfg.net_w * 3;
    const int T = (int)h->tensors.size();
    float *d_max = nullptr;
    unsigned *d_hist = nullptr;
    try {
        CK(cudaSetDevice(h->device));
        int rc = rf_debug_keep_all(h);
        if (rc) return rc;
        switch_ctx(h, 0);
        CK(cudaMalloc(&d_max, sizeof(float) * T));
        CK(cudaMalloc(&d_hist, sizeof(unsigned) * (size_t)T * CALIB_BINS));
        CK(cudaMemsetAsync(d_max, 0, sizeof(float) * T, h->stream));
        CK(cudaMemsetAsync(d_hist, 0, sizeof(unsigned) * (size_t)T * CALIB_BINS, h->stream));
        std::vector<float> hmax(T, 0.f);
        for (int pass = 0; pass < 2; pass++) {
            for (int i0 = 0; i0 < n_images; i0 += h->cfg.max_batch) {
                const int n = std::min(h->cfg.max_batch, n_images - i0);
                CK(cudaStreamSynchronize(h->stream));
                memcpy(h->h_input, bgr_net_sized + (size_t)i0 * img_bytes, (size_t)n * img_bytes);
                CK(cudaMemcpyAsync(h->d_input, h->h_input, (size_t)n * img_bytes, cudaMemcpyHostToDevice, h->stream));
                set_params(h, h->cur_thr, h->cur_nms);
                run_steps(h, n, h->stream, false);
                for (int t = 0; t < T; t++) {
                    const TensorInfo &ti = h->tensors[t];
                    const size_t elems = (size_t)n * ti.h * ti.w * ti.c;
                    const float *x = reinterpret_cast<const float *>(h->tptr(t));
                    if (pass == 0) launch_absmax<float>(x, elems, d_max + t, h->stream);
                    else if (hmax[t] > 0.f) launch_hist<float>(x, elems, (float)CALIB_BINS / hmax[t], d_hist + (size_t)t * CALIB_BINS, h->stream);
                }
                CK(cudaGetLastError());
            }
            if (pass == 0) {
                CK(cudaMemcpyAsync(hmax.data(), d_max, sizeof(float) * T, cudaMemcpyDeviceToHost, h->stream));
                CK(cudaStreamSynchronize(h->stream));
            }
        }
        std::vector<unsigned> hist((size_t)T * CALIB_BINS);
        CK(cudaMemcpyAsync(hist.data(), d_hist, sizeof(unsigned) * hist.size(), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        cudaFree(d_max); cudaFree(d_hist);
        d_max = nullptr; d_hist = nullptr;
        std::vector<std::pair<std::string, float>> scales;
        scales.emplace_back("data", 255.0f / 127.0f);          // u8 input range; the engine consumes the u8 image directly
        for (int t = 0; t < T; t++) {
            if (hmax[t] <= 0.f) { scales.emplace_back(h->tensors[t].name, 1.0f / 127.0f); continue; }
            const double bins = kl_threshold_bins(hist.data() + (size_t)t * CALIB_BINS);
            const double thr = bins * (double)hmax[t] / CALIB_BINS;
            scales.emplace_back(h->tensors[t].name, (float)(thr / 127.0));
        }
        std::string err;
        if (!write_int8_table(out_table_path, scales, err)) return fail(h, RF_ERR_IO, err);
    } catch (const CudaFail &f) {
        cudaFree(d_max); cudaFree(d_hist);
        return fail_cuda(h, f);
    }
    return RF_OK;
}

// Host-only: the KL threshold search on a caller-supplied histogram (for CPU-side tests of the calibrator).
double rf_kl_threshold_bins(const unsigned *hist, int bins, int levels) { return kl_threshold_bins(hist, bins, levels); }

// Host-only (no GPU needed): folded FP32 weights/bias of one convolution as the engine will hold
// them (BatchNorm + Scale + bias folded).  dims = {cout, cin/groups, k, k}.  Lets CPU-only tests
// check the model front end against the oracle's fold.
int rf_model_inspect(const char *caffemodel_path, const char *layer, float *w, int wcap, float *b, int bcap, int dims[4]) {
    if (!caffemodel_path || !layer) return fail(nullptr, RF_ERR_INVALID_ARG, "rf_model_inspect: NULL argument");
    std::vector<RawLayer> layers;
    std::string err;
    bool io = false;
    Model m;
    if (!read_caffemodel(caffemodel_path, layers, err, io)) return fail(nullptr, io ? RF_ERR_IO : RF_ERR_MODEL, err);
    if (!build_mnet_model(layers, m, err)) return fail(nullptr, RF_ERR_MODEL, err);
    auto it = m.convs.find(layer);
    if (it == m.convs.end()) return fail(nullptr, RF_ERR_INVALID_ARG, std::string("no convolution '") + layer + "'");
    const FoldedConv &c = it->second;
    if (dims) { dims[0] = c.cout; dims[1] = c.cin / c.groups; dims[2] = c.k; dims[3] = c.k; }
    if (w) { if ((size_t)wcap < c.w.size()) return fail(nullptr, RF_ERR_CAPACITY, "w buffer too small"); memcpy(w, c.w.data(), c.w.size() * 4); }
    if (b) { if ((size_t)bcap < c.b.size()) return fail(nullptr, RF_ERR_CAPACITY, "b buffer too small"); memcpy(b, c.b.data(), c.b.size() * 4); }
    return RF_OK;
}

int rf_profile_layers(rf_handle h, int n, int iters, char (*names)[64], float *ms, double *bytes, double *flops, int cap) {
    int rc = check_n(h, n);
    if (rc) return rc;
    if (n == 0 || iters <= 0) return fail(h, RF_ERR_INVALID_ARG, "rf_profile_layers: n and iters must be positive");
    int cnt = 0;
    try {
        CK(cudaSetDevice(h->device));
        switch_ctx(h, 0);
        set_params(h, h->cur_thr, h->cur_nms);
        run_steps(h, n, h->stream, false);  // warm everything once (also leaves consistent inputs for every step)
        CK(cudaStreamSynchronize(h->stream));
        for (size_t si = 0; si < h->steps.size(); si++) {
            auto &st = h->steps[si];
            if (cnt >= cap) break;
            if ((int)si == h->head_step + 1) {
                // sort+nms consumes the candidate list: give every timed launch a fresh one
                float acc = 0;
                for (int i = 0; i < iters; i++) {
                    h->steps[h->head_step].launch(n, h->stream);
                    CK(cudaEventRecord(h->ev0, h->stream));
                    st.launch(n, h->stream);
                    CK(cudaEventRecord(h->ev1, h->stream));
                    CK(cudaEventSynchronize(h->ev1));
                    float t = 0;
                    CK(cudaEventElapsedTime(&t, h->ev0, h->ev1));
                    acc += t;
                }
                snprintf(names[cnt], 64, "%s", st.name.c_str());
                ms[cnt] = acc / iters;
                if (bytes) bytes[cnt] = st.bytes_per_img * n;
                if (flops) flops[cnt] = st.flops_per_img * n;
                cnt++;
                continue;
            }
            st.launch(n, h->stream);
            CK(cudaEventRecord(h->ev0, h->stream));
            for (int i = 0; i < iters; i++) st.launch(n, h->stream);
            CK(cudaEventRecord(h->ev1, h->stream));
            CK(cudaEventSynchronize(h->ev1));
            float t = 0;
            CK(cudaEventElapsedTime(&t, h->ev0, h->ev1));
            snprintf(names[cnt], 64, "%s", st.name.c_str());
            ms[cnt] = t / iters;
            if (bytes) bytes[cnt] = st.bytes_per_img * n;
            if (flops) flops[cnt] = st.flops_per_img * n;
            cnt++;
        }
        CK(cudaGetLastError());
    } catch (const CudaFail &f) { return fail_cuda(h, f); }
    return cnt;
}

}  // extern "C"
