// model.h -- host-side model front end of librf_b200: caffemodel reader, mnet25 topology,
// BatchNorm/Scale/bias folding.  Replaces what the reference delegates to TensorRT's Caffe
// parser (retinaface/tensorrt/trtnetbase.cpp:257-330) for this one network family.
#pragma once
#include <map>
#include <string>
#include <vector>

namespace rf {

struct RawBlob {
    std::vector<long long> dims;
    std::vector<float> data;
};
struct RawLayer {
    std::string name, type;
    std::vector<RawBlob> blobs;
};

// Convolution with BatchNorm + Scale (+ conv bias) folded in FP32/FP64 at load time:
//   w'[o] = w[o] * gamma[o] / sqrt(var[o]/sf + eps),  b'[o] = (bias[o] - mean[o]/sf) * that + beta[o]
struct FoldedConv {
    std::string name;
    int cin = 0, cout = 0, k = 1, stride = 1, groups = 1;
    bool relu = false;
    std::vector<float> w;  // [cout][cin/groups][k][k]
    std::vector<float> b;  // [cout]
};

struct Model {
    std::map<std::string, FoldedConv> convs;
    std::vector<float> up_w[2];  // rf_c3_upsampling, rf_c2_upsampling: [64][4][4] depthwise deconv kernels
    const FoldedConv &conv(const std::string &n) const { return convs.at(n); }
};

// Parses a BVLC .caffemodel (protobuf wire format).  Returns false + err.
bool read_caffemodel(const std::string &path, std::vector<RawLayer> &layers, std::string &err, bool &io_error);

// Builds the folded mnet25 / mnet-deconv-0517 model (both share one topology,
// model/mnet-deconv-0517.prototxt) and validates every expected layer + blob shape.
bool build_mnet_model(const std::vector<RawLayer> &layers, Model &m, std::string &err);

// TensorRT EntropyCalibration2 cache: "<tensor>: <8 hex digits>" big-endian float32 scale
// (retinaface/tensorrt/trtnetbase.cpp:31-44 hands these bytes to TensorRT).
bool read_int8_table(const std::string &path, std::map<std::string, float> &scales, std::string &err);

// Base anchors of one FPN level for the "net3" configuration, computed the way
// generate_anchors does (retinaface/RetinaFace.cpp:35-104, config :245-268).
void base_anchors_net3(int stride, float out[8]);

}  // namespace rf
