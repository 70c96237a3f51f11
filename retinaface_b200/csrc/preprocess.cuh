// preprocess.cuh -- the single preprocess kernel of librf_b200.
//
// Replaces, in one coalesced pass over the network-sized output, the reference's preprocess
// chain (retinaface/RetinaFace.cpp:593-647): cudaMemset of the resize buffer (:598), the
// resize (NPP imageROIResize8U3C, resizeconvertion.cu:279-316, or cv::resize :613-617) and
// copyMakeBorder (:614-623).  The remaining stages of that chain -- u8->f32, BGR->RGB
// (convertBGR2RGBfloat, resizeconvertion.cu:46-63) and HWC->CHW (imageSplit, :165-185) -- do
// not exist here at all: conv0 consumes the letter-boxed u8 BGR HWC image directly
// (kernels_simt.cuh k_conv0), so no float image is ever materialised.
//
// Resize definition: the reference's OpenCV branch, cv::resize(img, Size(), 1/scale, 1/scale)
// with INTER_LINEAR on 8UC3 -- OpenCV's fixed-point bilinear (coefficients rounded to 11 bits,
// horizontal pass in int, vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2 >> 2).
// Restated so results are BIT-IDENTICAL to cv2.resize (tests/test_preprocess.py).
// RF_FLAG_NPP_RESIZE selects the reference's OTHER branch instead (USE_NPP: imageROIResize8U3C -> nppiResizeSqrPixel_8u_C3R with
// NPPI_INTER_SUPER): coverage-weighted super-sampling with NPP's measured edge rule (preprocess.cu area_pixel; checked against
// NPP itself on a GPU box, oracle/npp_oracle.cu).
#pragma once
#include "common.cuh"

namespace rf {

// Letter-box one device u8 BGR HWC image (packed rows) into dst[net_h][net_w][3]:
// shrink by max(w/net_w, h/net_h, 1) (never up-scales), anchor top-left, zero the rest.
void launch_letterbox(const uint8_t *src, int w, int h, uint8_t *dst, int net_w, int net_h, cudaStream_t s);

// One view of a test-time-augmentation set: the image, optionally mirrored horizontally, fitted into the
// top-left box_w x box_h corner of the network input (box <= net: a smaller box is a smaller test scale), zero elsewhere.
// Returns the reference's map-back factor for this view (RetinaFace.cpp:587-591,732-738).
float launch_letterbox_view(const uint8_t *src, int w, int h, uint8_t *dst, int net_w, int net_h, int box_w, int box_h, int flip,
                            cudaStream_t s);

// Host helper: output size + scale the way RetinaFace::detect + cv::resize compute them.
void letterbox_geometry(int w, int h, int net_w, int net_h, int *dw, int *dh, double *inv_scale);
// ... and the way imageROIResize8U3C + NPP do (resizeconvertion.cu:296-310): extent ceil(w f) x ceil(h f)
void letterbox_geometry_npp(int w, int h, int net_w, int net_h, int *dw, int *dh, double *inv_scale);

// Batched letter-box: ONE launch for up to LB_MAX_IMAGES images per call chunk, each with its own source / destination buffer;
// 4 pixels (three 32-bit stores) per thread.
constexpr int LB_MAX_IMAGES = 64;
struct LbItem {
    const uint8_t *src; uint8_t *dst;
    int sw, sh, dw, dh;
    double scale;
    int identity, flip, area;
};
// fills one item (geometry of either resize definition); returns the reference's map-back factor
float letterbox_fill(LbItem &it, const uint8_t *src, int w, int h, uint8_t *dst, int box_w, int box_h, int flip, int area);
cudaError_t launch_letterbox_batch(const LbItem *items, int n, int net_w, int net_h, cudaStream_t s);

}  // namespace rf
