// frontend.cpp -- model-format front end of librf_b200: a Caffe prototxt (protobuf text format) reader, the
// graph checks that tie a prototxt to the RetinaFace mnet25 family this engine executes, the network-name / anchor
// configuration switch of the reference's constructor, and a cache of the folded model with a staleness check.
//
// What it replaces in the reference:
//   TrtNetBase::parseNet (tensorrt/trtnetbase.cpp:149-197)      reads N,C,H,W from the text line after `input_param` by fixed
//                                                               character offsets (1-digit N/C, 3-digit H only; loops forever
//                                                               when `input_param` is missing) -> a real text-format parser
//   nvcaffeparser (trtnetbase.cpp:261-285)                      prototxt + caffemodel -> layers: here prototxt -> graph ->
//                                                               BatchNorm/Scale/bias folding driven by the FILE's parameters
//   RetinaFace::RetinaFace network switch (RetinaFace.cpp:211-268)   network name -> FPN levels, anchor scales / ratios
//   engine cache (trtnetbase.cpp:205-243: `<name>.cache`, reused with NO staleness check)   -> folded-model cache keyed by
//                                                               the caffemodel's size + FNV-1a hash + format version
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <set>
#include <sstream>

#include "model.h"

namespace rf {

// ---- protobuf text format -----------------------------------------------------------------------------------------------
namespace {

struct Tok { enum Kind { IDENT, STRING, NUMBER, LBRACE, RBRACE, COLON, END } kind; std::string text; int line; };

bool tokenize(const std::string &src, std::vector<Tok> &out, std::string &err) {
    size_t i = 0;
    int line = 1;
    while (i < src.size()) {
        const char c = src[i];
        if (c == '\n') { line++; i++; continue; }
        if (c == ' ' || c == '\t' || c == '\r' || c == ',' || c == ';') { i++; continue; }
        if (c == '#') { while (i < src.size() && src[i] != '\n') i++; continue; }
        if (c == '{' || c == '<') { out.push_back({Tok::LBRACE, "{", line}); i++; continue; }
        if (c == '}' || c == '>') { out.push_back({Tok::RBRACE, "}", line}); i++; continue; }
        if (c == ':') { out.push_back({Tok::COLON, ":", line}); i++; continue; }
        if (c == '"' || c == '\'') {
            std::string s;
            size_t j = i + 1;
            while (j < src.size() && src[j] != c) {
                if (src[j] == '\\' && j + 1 < src.size()) j++;
                if (src[j] == '\n') line++;
                s.push_back(src[j++]);
            }
            if (j >= src.size()) { err = "unterminated string at line " + std::to_string(line); return false; }
            out.push_back({Tok::STRING, s, line});
            i = j + 1;
            continue;
        }
        if (isalpha((unsigned char)c) || c == '_') {
            size_t j = i;
            while (j < src.size() && (isalnum((unsigned char)src[j]) || src[j] == '_' || src[j] == '.')) j++;
            out.push_back({Tok::IDENT, src.substr(i, j - i), line});
            i = j;
            continue;
        }
        if (isdigit((unsigned char)c) || c == '-' || c == '+' || c == '.') {
            size_t j = i + 1;
            while (j < src.size() && (isalnum((unsigned char)src[j]) || src[j] == '.' || src[j] == '-' || src[j] == '+')) j++;
            out.push_back({Tok::NUMBER, src.substr(i, j - i), line});
            i = j;
            continue;
        }
        err = std::string("unexpected character '") + c + "' at line " + std::to_string(line);
        return false;
    }
    out.push_back({Tok::END, "", line});
    return true;
}

bool parse_message(const std::vector<Tok> &t, size_t &i, ProtoNode &node, bool top, std::string &err) {
    while (true) {
        const Tok &k = t[i];
        if (k.kind == Tok::END) { if (!top) { err = "missing '}' at end of file"; return false; } return true; }
        if (k.kind == Tok::RBRACE) { if (top) { err = "unmatched '}' at line " + std::to_string(k.line); return false; } i++; return true; }
        if (k.kind != Tok::IDENT) { err = "field name expected at line " + std::to_string(k.line) + ", got '" + k.text + "'"; return false; }
        i++;
        bool colon = false;
        if (t[i].kind == Tok::COLON) { colon = true; i++; }
        if (t[i].kind == Tok::LBRACE) {
            i++;
            ProtoNode child;
            if (!parse_message(t, i, child, false, err)) return false;
            node.children.emplace_back(k.text, std::move(child));
        } else if (colon && (t[i].kind == Tok::STRING || t[i].kind == Tok::NUMBER || t[i].kind == Tok::IDENT)) {
            node.scalars.emplace_back(k.text, t[i].text);
            i++;
        } else {
            err = "value expected for field '" + k.text + "' at line " + std::to_string(k.line);
            return false;
        }
    }
}

}  // namespace

const ProtoNode *ProtoNode::child(const std::string &name) const {
    for (auto &c : children) if (c.first == name) return &c.second;
    return nullptr;
}
std::vector<std::string> ProtoNode::all(const std::string &name) const {
    std::vector<std::string> v;
    for (auto &s : scalars) if (s.first == name) v.push_back(s.second);
    return v;
}
std::string ProtoNode::get(const std::string &name, const std::string &dflt) const {
    for (auto &s : scalars) if (s.first == name) return s.second;
    return dflt;
}

bool read_prototxt(const std::string &path, NetGraph &g, std::string &err, bool &io_error) {
    io_error = false;
    std::ifstream f(path);
    if (!f) { err = "cannot open prototxt '" + path + "'"; io_error = true; return false; }
    std::stringstream ss;
    ss << f.rdbuf();
    std::vector<Tok> toks;
    ProtoNode root;
    size_t i = 0;
    if (!tokenize(ss.str(), toks, err) || !parse_message(toks, i, root, true, err)) { err = "prototxt '" + path + "': " + err; return false; }
    g = NetGraph{};
    g.name = root.get("name", "");
    // legacy input declaration: input: "data" input_dim: 1 ...   |   input_shape { dim: ... }
    std::vector<std::string> idims = root.all("input_dim");
    if (const ProtoNode *sh = root.child("input_shape")) idims = sh->all("dim");
    for (auto &c : root.children) {
        if (c.first != "layer" && c.first != "layers") continue;
        const ProtoNode &n = c.second;
        ProtoLayer L;
        L.name = n.get("name", "");
        L.type = n.get("type", "");
        L.bottom = n.all("bottom");
        L.top = n.all("top");
        if (L.name.empty() || L.type.empty()) { err = "prototxt '" + path + "': a layer without name or type"; return false; }
        if (const ProtoNode *p = n.child("convolution_param")) {
            L.num_output = atoi(p->get("num_output", "0").c_str());
            L.kernel = atoi(p->get("kernel_size", p->get("kernel_h", "1")).c_str());
            L.stride = atoi(p->get("stride", "1").c_str());
            L.pad = atoi(p->get("pad", "0").c_str());
            L.group = atoi(p->get("group", "1").c_str());
            L.bias_term = p->get("bias_term", "true") != "false";
        }
        if (const ProtoNode *p = n.child("batch_norm_param")) {
            L.eps = atof(p->get("eps", "1e-5").c_str());
            L.use_global_stats = p->get("use_global_stats", "true") != "false";
        }
        if (const ProtoNode *p = n.child("scale_param")) L.bias_term = p->get("bias_term", "false") == "true";
        if (const ProtoNode *p = n.child("eltwise_param")) L.op = p->get("operation", "SUM");
        if (const ProtoNode *p = n.child("concat_param")) L.axis = atoi(p->get("axis", "1").c_str());
        if (const ProtoNode *p = n.child("softmax_param")) L.axis = atoi(p->get("axis", "1").c_str());
        if (const ProtoNode *p = n.child("crop_param")) L.axis = atoi(p->get("axis", "2").c_str());
        if (const ProtoNode *p = n.child("reshape_param")) {
            L.axis = atoi(p->get("axis", "0").c_str());
            if (const ProtoNode *sh = p->child("shape")) for (auto &d : sh->all("dim")) L.dims.push_back(atoi(d.c_str()));
        }
        if (const ProtoNode *p = n.child("input_param"))
            if (const ProtoNode *sh = p->child("shape")) idims = sh->all("dim");
        g.layers.push_back(std::move(L));
    }
    if (g.layers.empty()) { err = "prototxt '" + path + "' holds no layers"; return false; }
    if (idims.size() != 4) { err = "prototxt '" + path + "': no 4-D input shape (input_param / input_shape / input_dim)"; return false; }
    for (int k = 0; k < 4; k++) g.input_dims[k] = atoi(idims[k].c_str());
    if (g.input_dims[1] != 3 || g.input_dims[2] <= 0 || g.input_dims[3] <= 0) { err = "prototxt '" + path + "': input must be N x 3 x H x W"; return false; }
    return true;
}

// ---- graph -> folding specification + topology check ------------------------------------------------------------------------
namespace {

// the layer that produced the DATA of a blob as a consumer sees it: walks back through in-place / pass-through layers
// (BatchNorm, Scale, ReLU, Crop, Reshape, Softmax, Split) to the Convolution / Deconvolution / Eltwise / Concat / Input
struct Wiring {
    std::vector<std::vector<int>> inputs;      // per layer: the layers that last wrote each of its bottoms
    const NetGraph &g;
    explicit Wiring(const NetGraph &gr) : g(gr) {
        std::map<std::string, int> writer;
        inputs.resize(g.layers.size());
        for (size_t i = 0; i < g.layers.size(); i++) {
            for (auto &b : g.layers[i].bottom) {
                auto it = writer.find(b);
                inputs[i].push_back(it == writer.end() ? -1 : it->second);
            }
            for (auto &t : g.layers[i].top) writer[t] = (int)i;
        }
    }
    static bool passthrough(const std::string &t) { return t == "BatchNorm" || t == "Scale" || t == "ReLU" || t == "Crop" || t == "Reshape" || t == "Softmax" || t == "Split"; }
    int root(int layer) const {
        while (layer >= 0 && passthrough(g.layers[layer].type)) layer = inputs[layer].empty() ? -1 : inputs[layer][0];
        return layer;
    }
    // root producers of layer i's bottoms
    std::vector<std::string> roots_of(int i) const {
        std::vector<std::string> v;
        for (int in : inputs[i]) { int r = root(in); v.push_back(r < 0 ? "?" : g.layers[r].name); }
        return v;
    }
};

}  // namespace

bool graph_conv_specs(const NetGraph &g, std::vector<ConvSpec> &specs, std::string &err) {
    Wiring w(g);
    std::map<std::string, int> by_name;
    for (size_t i = 0; i < g.layers.size(); i++) by_name[g.layers[i].name] = (int)i;
    // consumers of each layer's output (following in-place chains by writer order)
    for (size_t i = 0; i < g.layers.size(); i++) {
        const ProtoLayer &L = g.layers[i];
        if (L.type != "Convolution") continue;
        ConvSpec s;
        s.name = L.name; s.cout = L.num_output; s.k = L.kernel; s.stride = L.stride; s.groups = L.group; s.bias = L.bias_term; s.pad = L.pad;
        if (s.cout <= 0 || s.k <= 0) { err = "prototxt: convolution '" + L.name + "' lacks num_output / kernel_size"; return false; }
        // the chain behind it: BatchNorm -> Scale -> ReLU (each optional), found as the layers whose first input is the previous link
        int cur = (int)i;
        for (size_t j = i + 1; j < g.layers.size(); j++) {
            const ProtoLayer &N = g.layers[j];
            if (w.inputs[j].empty() || w.inputs[j][0] != cur) continue;
            if (N.type == "BatchNorm" && s.bn.empty() && !s.relu) { s.bn = N.name; s.eps = N.eps; if (!N.use_global_stats) { err = "prototxt: BatchNorm '" + N.name + "' without use_global_stats"; return false; } cur = (int)j; }
            else if (N.type == "Scale" && !s.bn.empty() && s.scale.empty()) { s.scale = N.name; cur = (int)j; }
            else if (N.type == "ReLU" && !s.relu) { s.relu = true; cur = (int)j; break; }
            else if (N.type != "BatchNorm" && N.type != "Scale" && N.type != "ReLU") break;
        }
        if (!s.bn.empty() && s.scale.empty()) { err = "prototxt: BatchNorm '" + s.bn + "' is not followed by a Scale layer"; return false; }
        specs.push_back(s);
    }
    return true;
}

bool check_mnet_topology(const NetGraph &g, std::string &err) {
    Wiring w(g);
    std::map<std::string, int> by_name;
    for (size_t i = 0; i < g.layers.size(); i++) by_name[g.layers[i].name] = (int)i;
    auto fail = [&](const std::string &m) { err = "prototxt is not the RetinaFace mnet25 graph this engine runs: " + m; return false; };
    auto fed_by = [&](const std::string &layer, std::vector<std::string> want, const char *type) -> bool {
        auto it = by_name.find(layer);
        if (it == by_name.end()) return fail("layer '" + layer + "' is missing");
        if (g.layers[it->second].type != type) return fail("layer '" + layer + "' is a " + g.layers[it->second].type + ", expected " + type);
        std::vector<std::string> got = w.roots_of(it->second);
        if (got != want) {
            std::string a, b;
            for (auto &x : got) a += x + " ";
            for (auto &x : want) b += x + " ";
            return fail("layer '" + layer + "' reads from [ " + a + "], expected [ " + b + "]");
        }
        return true;
    };
    // backbone chain (prototxt:11-1192)
    std::string prev = "data";
    for (int i = 0; i <= 26; i++) {
        const std::string n = "mobilenet0_conv" + std::to_string(i) + "_fwd";
        if (!fed_by(n, {prev}, "Convolution")) return false;
        prev = n;
    }
    // FPN (prototxt:1199-1237, 1513-1632, 1908-2027): laterals, bilinear deconvolution, crop, sum, aggregation
    if (!fed_by("rf_c3_lateral", {"mobilenet0_conv26_fwd"}, "Convolution") || !fed_by("rf_c2_lateral", {"mobilenet0_conv22_fwd"}, "Convolution") ||
        !fed_by("rf_c1_red_conv", {"mobilenet0_conv10_fwd"}, "Convolution"))
        return false;
    if (!fed_by("rf_c3_upsampling", {"rf_c3_lateral"}, "Deconvolution") || !fed_by("rf_c2_upsampling", {"rf_c2_aggr"}, "Deconvolution")) return false;
    auto sum_of = [&](const std::string &consumer, const std::string &a, const std::string &b) -> bool {
        auto it = by_name.find(consumer);
        if (it == by_name.end()) return fail("layer '" + consumer + "' is missing");
        const int e = w.root(w.inputs[it->second].empty() ? -1 : w.inputs[it->second][0]);
        if (e < 0 || g.layers[e].type != "Eltwise" || (g.layers[e].op != "SUM" && g.layers[e].op != "1")) return fail("'" + consumer + "' does not read an Eltwise SUM");
        std::vector<std::string> got = w.roots_of(e);
        std::multiset<std::string> gs(got.begin(), got.end()), ws{a, b};
        if (gs != ws) return fail("the Eltwise in front of '" + consumer + "' does not add " + a + " and " + b);
        return true;
    };
    if (!sum_of("rf_c2_aggr", "rf_c2_lateral", "rf_c3_upsampling") || !sum_of("rf_c1_aggr", "rf_c1_red_conv", "rf_c2_upsampling")) return false;
    // SSH heads + predictors (prototxt:1239-1511, 1634-1906, 2029-2302)
    const char *lvl[3] = {"c3", "c2", "c1"};
    const char *src[3] = {"rf_c3_lateral", "rf_c2_aggr", "rf_c1_aggr"};
    const int strides[3] = {32, 16, 8};
    for (int l = 0; l < 3; l++) {
        const std::string p = std::string("rf_") + lvl[l] + "_det";
        if (!fed_by(p + "_conv1", {src[l]}, "Convolution") || !fed_by(p + "_context_conv1", {src[l]}, "Convolution") ||
            !fed_by(p + "_context_conv2", {p + "_context_conv1"}, "Convolution") || !fed_by(p + "_context_conv3_1", {p + "_context_conv1"}, "Convolution") ||
            !fed_by(p + "_context_conv3_2", {p + "_context_conv3_1"}, "Convolution"))
            return false;
        const std::string st = "_stride" + std::to_string(strides[l]);
        for (const char *hd : {"face_rpn_cls_score", "face_rpn_bbox_pred", "face_rpn_landmark_pred"}) {
            auto it = by_name.find(std::string(hd) + st);
            if (it == by_name.end() || g.layers[it->second].type != "Convolution") return fail(std::string("predictor '") + hd + st + "' is missing");
            const int c = w.root(w.inputs[it->second].empty() ? -1 : w.inputs[it->second][0]);
            if (c < 0 || g.layers[c].type != "Concat") return fail(std::string("predictor '") + hd + st + "' does not read the SSH concat");
            if (w.roots_of(c) != std::vector<std::string>{p + "_conv1", p + "_context_conv2", p + "_context_conv3_2"})
                return fail("the SSH concat of level " + std::string(lvl[l]) + " is not [conv1, context_conv2, context_conv3_2]");
        }
        // softmax over the (N, 2, 2h, w) view of the class scores (prototxt:1448-1483)
        bool softmax = false;
        for (size_t i = 0; i < g.layers.size(); i++)
            if (g.layers[i].type == "Softmax" && w.root((int)i) >= 0 && g.layers[w.root((int)i)].name == "face_rpn_cls_score" + st) {
                softmax = true;
                const int r0 = w.inputs[i].empty() ? -1 : w.inputs[i][0];
                // either spelling: shape {0, 2, -1, 0}  (mnet-deconv-0517.prototxt)  or  shape {2, -1, 0} from axis 1  (mnet25.prototxt)
                bool view_ok = false;
                if (r0 >= 0 && g.layers[r0].type == "Reshape") {
                    const ProtoLayer &R = g.layers[r0];
                    const int at1 = 1 - R.axis;            // index of the dimension that lands on axis 1
                    view_ok = at1 >= 0 && at1 < (int)R.dims.size() && R.dims[at1] == 2 && (int)R.dims.size() == 4 - R.axis;
                }
                if (!view_ok) return fail("the class-score Softmax of stride " + std::to_string(strides[l]) + " is not taken over a (N, 2, -1, w) reshape");
            }
        if (!softmax) return fail("no Softmax on face_rpn_cls_score" + st);
    }
    return true;
}

// ---- network name -> FPN / anchor configuration (RetinaFace.cpp:211-268) ----------------------------------------------------
bool network_config(const std::string &network, NetworkConfig &c, std::string &err) {
    c = NetworkConfig{};
    c.ratios = {1.0f};
    c.fmc = 3;
    if (network == "ssh" || network == "vgg") { c.pixel_means[0] = 103.939f; c.pixel_means[1] = 116.779f; c.pixel_means[2] = 123.68f; }   // :211-215
    else if (network == "net3") c.ratios = {1.0f};                                   // :216-218
    else if (network == "net3a") c.ratios = {1.0f, 1.5f};                            // :219-221
    else if (network == "net6") c.fmc = 6;                                           // :222-224
    else if (network == "net5") c.fmc = 5;                                           // :225-227
    else if (network == "net5a") { c.fmc = 5; c.ratios = {1.0f, 1.5f}; }             // :228-231 (the second net5a branch, :236-239, is unreachable)
    else if (network == "net4") c.fmc = 4;                                           // :233-235
    else { err = "network setting error: '" + network + "'"; return false; }         // :240-242
    if (c.fmc != 3) { err = "network '" + network + "' wants " + std::to_string(c.fmc) + " FPN levels: the reference has no anchor configuration for it either (RetinaFace.cpp:266-268)"; return false; }
    // :245-264
    c.strides = {32, 16, 8};
    c.scales = {{32, 16}, {8, 4}, {2, 1}};
    c.base_size = 16;
    c.allowed_border = 9999;
    return true;
}

// ---- folded-model cache ------------------------------------------------------------------------------------------------------
namespace {
constexpr uint32_t CACHE_MAGIC = 0x4d434652;      // "RFCM"
constexpr uint32_t CACHE_VERSION = 2;

uint64_t fnv1a(const std::vector<uint8_t> &b, uint64_t h = 1469598103934665603ull) {
    for (uint8_t c : b) { h ^= c; h *= 1099511628211ull; }
    return h;
}
bool slurp(const std::string &path, std::vector<uint8_t> &out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    out.assign((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return true;
}
void put(std::vector<uint8_t> &o, const void *p, size_t n) { o.insert(o.end(), (const uint8_t *)p, (const uint8_t *)p + n); }
template <typename T> void put(std::vector<uint8_t> &o, T v) { put(o, &v, sizeof v); }
struct In {
    const uint8_t *p, *e;
    bool ok = true;
    template <typename T> T get() { T v{}; if ((size_t)(e - p) < sizeof v) { ok = false; return v; } memcpy(&v, p, sizeof v); p += sizeof v; return v; }
    bool bytes(void *dst, size_t n) { if ((size_t)(e - p) < n) { ok = false; return false; } memcpy(dst, p, n); p += n; return true; }
};
}  // namespace

bool model_cache_key(const std::string &caffemodel, const std::string &prototxt, ModelCacheKey &k) {
    std::vector<uint8_t> a, b;
    if (!slurp(caffemodel, a)) return false;
    k.size = a.size();
    k.hash = fnv1a(a);
    if (!prototxt.empty()) { if (!slurp(prototxt, b)) return false; k.hash = fnv1a(b, k.hash); k.size += b.size(); }
    k.version = CACHE_VERSION;
    return true;
}

bool save_model_cache(const std::string &path, const ModelCacheKey &k, const Model &m) {
    std::vector<uint8_t> o;
    put(o, CACHE_MAGIC); put(o, k.version); put(o, k.size); put(o, k.hash);
    put<uint32_t>(o, (uint32_t)m.convs.size());
    for (auto &kv : m.convs) {
        const FoldedConv &c = kv.second;
        put<uint32_t>(o, (uint32_t)c.name.size()); put(o, c.name.data(), c.name.size());
        int32_t hdr[6] = {c.cin, c.cout, c.k, c.stride, c.groups, c.relu ? 1 : 0};
        put(o, hdr, sizeof hdr);
        put<uint64_t>(o, c.w.size()); put(o, c.w.data(), c.w.size() * 4);
        put<uint64_t>(o, c.b.size()); put(o, c.b.data(), c.b.size() * 4);
    }
    for (int i = 0; i < 2; i++) { put<uint64_t>(o, m.up_w[i].size()); put(o, m.up_w[i].data(), m.up_w[i].size() * 4); }
    put<uint64_t>(o, fnv1a(o));                      // trailer: the cache file's own checksum
    const std::string tmp = path + ".tmp";
    {
        std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
        if (!f) return false;
        f.write((const char *)o.data(), (std::streamsize)o.size());
        if (!f) return false;
    }
    return rename(tmp.c_str(), path.c_str()) == 0;   // readers never see a half-written cache
}

int load_model_cache(const std::string &path, const ModelCacheKey &k, Model &m) {
    std::vector<uint8_t> buf;
    if (!slurp(path, buf) || buf.size() < 40) return CACHE_MISS;
    std::vector<uint8_t> body(buf.begin(), buf.end() - 8);
    uint64_t sum;
    memcpy(&sum, buf.data() + buf.size() - 8, 8);
    if (fnv1a(body) != sum) return CACHE_STALE;       // truncated / corrupted
    In in{body.data(), body.data() + body.size()};
    if (in.get<uint32_t>() != CACHE_MAGIC) return CACHE_STALE;
    if (in.get<uint32_t>() != k.version || in.get<uint64_t>() != k.size || in.get<uint64_t>() != k.hash) return CACHE_STALE;   // another model / format
    Model out;
    const uint32_t n = in.get<uint32_t>();
    for (uint32_t i = 0; i < n && in.ok; i++) {
        FoldedConv c;
        const uint32_t ln = in.get<uint32_t>();
        if (ln > 256) return CACHE_STALE;
        c.name.resize(ln);
        in.bytes(&c.name[0], ln);
        int32_t hdr[6];
        in.bytes(hdr, sizeof hdr);
        c.cin = hdr[0]; c.cout = hdr[1]; c.k = hdr[2]; c.stride = hdr[3]; c.groups = hdr[4]; c.relu = hdr[5] != 0;
        uint64_t wn = in.get<uint64_t>();
        if (!in.ok || wn > (1u << 26)) return CACHE_STALE;
        c.w.resize(wn); in.bytes(c.w.data(), wn * 4);
        uint64_t bn = in.get<uint64_t>();
        if (!in.ok || bn > (1u << 20)) return CACHE_STALE;
        c.b.resize(bn); in.bytes(c.b.data(), bn * 4);
        out.convs[c.name] = std::move(c);
    }
    for (int i = 0; i < 2 && in.ok; i++) {
        uint64_t un = in.get<uint64_t>();
        if (!in.ok || un > (1u << 20)) return CACHE_STALE;
        out.up_w[i].resize(un); in.bytes(out.up_w[i].data(), un * 4);
    }
    if (!in.ok) return CACHE_STALE;
    m = std::move(out);
    return CACHE_HIT;
}

}  // namespace rf
