// oracle/ref_driver.cpp -- harness that links against the reference's own, UNMODIFIED
// retinaface/RetinaFace.cpp (compiled from where it lies in /root/reference by
// oracle/build_ref.sh) so its post-process can be run on injected head tensors.
//
// TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Output goes to oracle/_ref/ (git-ignored).
//
// What is the reference's and what is ours:
//   * reference (executed verbatim): generate_anchors_fpn / anchors_plane / clip_boxes
//     (RetinaFace.cpp:9-199), the RetinaFace constructor's anchor setup (:205-302),
//     RetinaFace::postProcess (:494-574: gather -> threshold -> bbox_pred -> clip ->
//     landmark_pred -> nms(0.4)), RetinaFace::nms (:439-492).
//   * ours (this file): a fake `TrtRetinaFaceNet` engine.  The real one wraps TensorRT 5
//     (absent here); the fake implements the same member functions the detector calls
//     (tensorrt/trtretinafacenet.h:24-46, tensorrt/trtnetbase.h:57-164) and simply serves
//     head blobs that the test injected, in the engine's documented blob order
//     (tensorrt/trtretinafacenet.cpp:23-31).
#define private public   // reach RetinaFace::postProcess / nms / _anchors (test harness only)
#include "RetinaFace.h"
#undef private

#include <cstring>

namespace {
int g_net_w = 448, g_net_h = 448;
const char *kBlobNames[9] = {
    "face_rpn_cls_prob_reshape_stride32", "face_rpn_bbox_pred_stride32", "face_rpn_landmark_pred_stride32",
    "face_rpn_cls_prob_reshape_stride16", "face_rpn_bbox_pred_stride16", "face_rpn_landmark_pred_stride16",
    "face_rpn_cls_prob_reshape_stride8", "face_rpn_bbox_pred_stride8", "face_rpn_landmark_pred_stride8"};
const int kBlobCh[9] = {4, 8, 20, 4, 8, 20, 4, 8, 20};
const int kBlobStride[9] = {32, 32, 32, 16, 16, 16, 8, 8, 8};
}  // namespace

// ---- fake engine: TrtNetBase -------------------------------------------------------------
TrtNetBase::TrtNetBase(std::string name)
    : pLogger(nullptr), profiler(nullptr), runtime(nullptr), engine(nullptr), context(nullptr),
      useFp32(true), workSpaceSize(0), maxBatchSize(1), batchSize(1), channel(3),
      netWidth(g_net_w), netHeight(g_net_h), numBinding(10), inputBuffer(nullptr), buffers(nullptr),
      dumpResult(false), enableTrtProfiler(false), netWorkName(name) {}
TrtNetBase::~TrtNetBase() {}
uint32_t TrtNetBase::getMaxBatchSize() const { return maxBatchSize; }
int TrtNetBase::getNetWidth() const { return netWidth; }
int TrtNetBase::getNetHeight() const { return netHeight; }
int TrtNetBase::getChannel() const { return channel; }
void *&TrtNetBase::getBuffer(const int &index) { static void *none = nullptr; (void)index; return none; }
void TrtNetBase::buildTrtContext(const std::string &, const std::string &, bool) {
    netWidth = g_net_w;
    netHeight = g_net_h;
    allocateMemory(false);
}

// ---- fake engine: TrtRetinaFaceNet -------------------------------------------------------
TrtRetinaFaceNet::TrtRetinaFaceNet(std::string name) : TrtNetBase(name) {
    results.resize(9);
    for (int i = 0; i < 9; i++) results[i].layer_name = kBlobNames[i];
}
TrtRetinaFaceNet::~TrtRetinaFaceNet() {}
void TrtRetinaFaceNet::allocateMemory(bool) {
    outputDims.clear();
    for (int i = 0; i < 9; i++) {
        DimsCHW d(kBlobCh[i], netHeight / kBlobStride[i], netWidth / kBlobStride[i]);
        outputDims.push_back(d);
        results[i].outputDims = d;
        results[i].outputSize = d.c() * d.h() * d.w();
        results[i].batchsize = 1;
        results[i].result.assign(1, std::vector<float>(results[i].outputSize, 0.f));
    }
}
void TrtRetinaFaceNet::releaseMemory(bool) {}
void TrtRetinaFaceNet::doInference(int, float *) {}
TrtBlob *TrtRetinaFaceNet::blob_by_name(string layer_name) {
    for (auto &b : results)
        if (b.layer_name == layer_name) return &b;
    return nullptr;
}
vector<int> TrtRetinaFaceNet::getOutputWidth() {
    return {outputDims[0].w(), outputDims[3].w(), outputDims[6].w()};
}
vector<int> TrtRetinaFaceNet::getOutputHeight() {
    return {outputDims[0].h(), outputDims[3].h(), outputDims[6].h()};
}

// ---- C entry points used by oracle/postproc.py ------------------------------------------
extern "C" {

// Construct the reference detector ("net3", nms 0.4 as in main.cpp:15 / RetinaFace.h:66) for
// a network input of net_w x net_h.  The constructor prints nothing and builds its anchors
// with the reference's own generate_anchors_fpn / anchors_plane.
void *ref_create(int net_w, int net_h) {
    g_net_w = net_w;
    g_net_h = net_h;
    std::string model = "unused";
    return new RetinaFace(model, "net3");
}

void ref_destroy(void *h) { delete static_cast<RetinaFace *>(h); }

// Base anchors of one level as the reference's constructor computed them (x1,y1,x2,y2 per anchor).
int ref_base_anchors(void *h, int stride, float *out8) {
    RetinaFace *rf = static_cast<RetinaFace *>(h);
    auto &v = rf->_anchors_fpn["stride" + std::to_string(stride)];
    for (size_t i = 0; i < v.size() && i < 2; i++) {
        out8[4 * i + 0] = v[i].x1; out8[4 * i + 1] = v[i].y1; out8[4 * i + 2] = v[i].x2; out8[4 * i + 3] = v[i].y2;
    }
    return (int)v.size();
}

// Full anchor plane of one level (anchors_plane output, 4 floats per anchor).
int ref_anchor_plane(void *h, int stride, float *out, int cap) {
    RetinaFace *rf = static_cast<RetinaFace *>(h);
    auto &v = rf->_anchors["stride" + std::to_string(stride)];
    int n = (int)v.size() < cap ? (int)v.size() : cap;
    for (int i = 0; i < n; i++) { out[4 * i] = v[i].x1; out[4 * i + 1] = v[i].y1; out[4 * i + 2] = v[i].x2; out[4 * i + 3] = v[i].y2; }
    return (int)v.size();
}

// Inject the 9 head blobs (engine blob order, NCHW, one image) and run the reference's own
// RetinaFace::postProcess (NMS threshold hard-coded 0.4 there).  out: 15 floats per face.
int ref_postprocess(void *h, const float *const heads[9], float threshold, float *out, int cap) {
    RetinaFace *rf = static_cast<RetinaFace *>(h);
    for (int i = 0; i < 9; i++) {
        TrtBlob *b = rf->trtNet->blob_by_name(kBlobNames[i]);
        std::memcpy(b->result[0].data(), heads[i], sizeof(float) * b->outputSize);
    }
    // postProcess prints timing lines to std::cout; silence them.
    std::streambuf *old = std::cout.rdbuf(nullptr);
    std::vector<FaceDetectInfo> faces =
        rf->postProcess(rf->trtNet->getNetWidth(), rf->trtNet->getNetHeight(), threshold);
    std::cout.rdbuf(old);
    int n = (int)faces.size() < cap ? (int)faces.size() : cap;
    static_assert(sizeof(FaceDetectInfo) == 15 * sizeof(float), "FaceDetectInfo is 15 floats");
    std::memcpy(out, faces.data(), sizeof(FaceDetectInfo) * n);
    return (int)faces.size();
}

// The reference's RetinaFace::nms on caller-provided candidates (15 floats each), any threshold.
int ref_nms(void *h, const float *cands, int n, float threshold, float *out) {
    RetinaFace *rf = static_cast<RetinaFace *>(h);
    std::vector<FaceDetectInfo> v(n);
    std::memcpy(v.data(), cands, sizeof(FaceDetectInfo) * n);
    std::vector<FaceDetectInfo> kept = rf->nms(v, threshold);
    std::memcpy(out, kept.data(), sizeof(FaceDetectInfo) * kept.size());
    return (int)kept.size();
}

}  // extern "C"
