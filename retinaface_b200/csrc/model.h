// model.h -- host-side model front end of librf_b200: caffemodel reader, mnet25 topology,
// BatchNorm/Scale/bias folding.  Replaces what the reference delegates to TensorRT's Caffe
// parser (retinaface/tensorrt/trtnetbase.cpp:257-330) for this one network family.
#pragma once
#include <cstdint>
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

// ---- prototxt front end (frontend.cpp) -------------------------------------------------------------------------------------
struct ProtoNode {               // one message of a protobuf text file
    std::vector<std::pair<std::string, std::string>> scalars;
    std::vector<std::pair<std::string, ProtoNode>> children;
    const ProtoNode *child(const std::string &name) const;
    std::vector<std::string> all(const std::string &name) const;
    std::string get(const std::string &name, const std::string &dflt) const;
};
struct ProtoLayer {
    std::string name, type;
    std::vector<std::string> bottom, top;
    int num_output = 0, kernel = 1, stride = 1, pad = 0, group = 1, axis = 1;
    bool bias_term = true, use_global_stats = true;
    double eps = 1e-5;
    std::string op = "SUM";
    std::vector<int> dims;       // Reshape
};
struct NetGraph {
    std::string name;
    int input_dims[4] = {0, 0, 0, 0};   // N, C, H, W
    std::vector<ProtoLayer> layers;
};
// what the prototxt says about one convolution and the BatchNorm / Scale / ReLU chain behind it: drives the folding
struct ConvSpec {
    std::string name, bn, scale;
    int cout = 0, k = 1, stride = 1, groups = 1, pad = 0;
    bool bias = true, relu = false;
    double eps = 1e-5;
};
bool read_prototxt(const std::string &path, NetGraph &g, std::string &err, bool &io_error);
bool graph_conv_specs(const NetGraph &g, std::vector<ConvSpec> &specs, std::string &err);
// wiring of the RetinaFace mnet25 family (backbone chain, FPN merge, SSH branches and concat order, softmax view)
bool check_mnet_topology(const NetGraph &g, std::string &err);

// network name -> FPN levels / anchors, the switch of RetinaFace::RetinaFace (RetinaFace.cpp:211-268)
struct NetworkConfig {
    int fmc = 3, base_size = 16, allowed_border = 9999;
    float pixel_means[3] = {0.f, 0.f, 0.f};
    std::vector<float> ratios;
    std::vector<int> strides;
    std::vector<std::vector<int>> scales;
};
bool network_config(const std::string &network, NetworkConfig &c, std::string &err);

// folded-model cache: valid only for the exact bytes of the caffemodel (+ prototxt) it was made from
struct ModelCacheKey { uint32_t version = 0; uint64_t size = 0, hash = 0; };
enum { CACHE_NONE = 0, CACHE_MISS = 1, CACHE_HIT = 2, CACHE_STALE = 3 };   // MISS / STALE: (re)written
struct Model;
bool model_cache_key(const std::string &caffemodel, const std::string &prototxt, ModelCacheKey &k);
bool save_model_cache(const std::string &path, const ModelCacheKey &k, const Model &m);
int load_model_cache(const std::string &path, const ModelCacheKey &k, Model &m);

// Parses a BVLC .caffemodel (protobuf wire format).  Returns false + err.
bool read_caffemodel(const std::string &path, std::vector<RawLayer> &layers, std::string &err, bool &io_error);

// Builds the folded mnet25 / mnet-deconv-0517 model (both share one topology,
// model/mnet-deconv-0517.prototxt) and validates every expected layer + blob shape.
// `graph_specs` (from a prototxt): every convolution's kernel / stride / group / bias_term / BatchNorm eps / ReLU come from the FILE
// and must describe the same layers the engine's plan expects; NULL: the built-in description of mnet25.
bool build_mnet_model(const std::vector<RawLayer> &layers, Model &m, std::string &err, const std::vector<ConvSpec> *graph_specs = nullptr);

// The whole front end: (cache ->) caffemodel [+ prototxt: parse, topology check, file-driven folding] (-> cache).
// status: an rf_status value on failure.  graph (optional) receives the parsed prototxt.
bool load_model(const std::string &caffemodel, const std::string &prototxt, const std::string &cache_path, Model &m, NetGraph *graph,
                int *cache_status, std::string &err, int &status);

// TensorRT EntropyCalibration2 cache: "<tensor>: <8 hex digits>" big-endian float32 scale
// (retinaface/tensorrt/trtnetbase.cpp:31-44 hands these bytes to TensorRT).
bool read_int8_table(const std::string &path, std::map<std::string, float> &scales, std::string &err);

// Base anchors of one FPN level for the "net3" configuration, computed the way
// generate_anchors does (retinaface/RetinaFace.cpp:35-104, config :245-268).
void base_anchors_net3(int stride, float out[8]);

}  // namespace rf
