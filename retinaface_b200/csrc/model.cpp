// model.cpp -- see model.h.
#include "model.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace rf {
namespace {

struct Reader {
    const uint8_t *p, *end;
    bool ok = true;
    uint64_t varint() {
        uint64_t v = 0;
        int shift = 0;
        while (p < end) {
            uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7F) << shift;
            if (!(b & 0x80)) return v;
            shift += 7;
            if (shift > 63) break;
        }
        ok = false;
        return 0;
    }
    // next field: returns false at end.  For length-delimited fields sub = [data, data+len).
    bool next(int &field, int &wt, uint64_t &val, Reader &sub) {
        if (p >= end || !ok) return false;
        uint64_t key = varint();
        if (!ok) return false;
        field = (int)(key >> 3);
        wt = (int)(key & 7);
        switch (wt) {
            case 0: val = varint(); break;
            case 1: if (end - p < 8) { ok = false; return false; } memcpy(&val, p, 8); p += 8; break;
            case 5: { if (end - p < 4) { ok = false; return false; } uint32_t t; memcpy(&t, p, 4); val = t; p += 4; break; }
            case 2: {
                uint64_t len = varint();
                if (!ok || (uint64_t)(end - p) < len) { ok = false; return false; }
                sub.p = p; sub.end = p + len; sub.ok = true;
                p += len;
                break;
            }
            default: ok = false; return false;
        }
        return ok;
    }
};

void parse_blob(Reader r, RawBlob &b) {
    int f, wt; uint64_t v; Reader s{nullptr, nullptr};
    long long legacy[4] = {0, 0, 0, 0};
    bool has_legacy = false;
    while (r.next(f, wt, v, s)) {
        if (f == 7 && wt == 2) {  // BlobShape
            int f2, w2; uint64_t v2; Reader s2{nullptr, nullptr};
            while (s.next(f2, w2, v2, s2)) {
                if (f2 == 1 && w2 == 2) { while (s2.p < s2.end && s2.ok) b.dims.push_back((long long)s2.varint()); }
                else if (f2 == 1 && w2 == 0) b.dims.push_back((long long)v2);
            }
        } else if (f == 5 && wt == 2) {  // packed float32, little endian
            size_t n = (size_t)(s.end - s.p) / 4;
            b.data.resize(n);
            memcpy(b.data.data(), s.p, n * 4);
        } else if (f == 5 && wt == 5) {
            uint32_t t = (uint32_t)v; float x; memcpy(&x, &t, 4); b.data.push_back(x);
        } else if (f >= 1 && f <= 4 && wt == 0) {
            legacy[f - 1] = (long long)v; has_legacy = true;
        }
    }
    if (b.dims.empty() && has_legacy) b.dims.assign(legacy, legacy + 4);
}

}  // namespace

bool read_caffemodel(const std::string &path, std::vector<RawLayer> &layers, std::string &err, bool &io_error) {
    io_error = false;
    std::ifstream f(path, std::ios::binary);
    if (!f) { err = "cannot open caffemodel '" + path + "'"; io_error = true; return false; }
    std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (buf.empty()) { err = "caffemodel '" + path + "' is empty"; io_error = true; return false; }
    Reader r{buf.data(), buf.data() + buf.size()};
    int fno, wt; uint64_t v; Reader s{nullptr, nullptr};
    while (r.next(fno, wt, v, s)) {
        if (fno != 100 || wt != 2) continue;  // NetParameter.layer
        RawLayer L;
        int f2, w2; uint64_t v2; Reader s2{nullptr, nullptr};
        while (s.next(f2, w2, v2, s2)) {
            if (f2 == 1 && w2 == 2) L.name.assign((const char *)s2.p, s2.end - s2.p);
            else if (f2 == 2 && w2 == 2) L.type.assign((const char *)s2.p, s2.end - s2.p);
            else if (f2 == 7 && w2 == 2) { L.blobs.emplace_back(); parse_blob(s2, L.blobs.back()); }
        }
        if (!s.ok) { err = "malformed LayerParameter in '" + path + "'"; return false; }
        layers.push_back(std::move(L));
    }
    if (!r.ok) { err = "malformed protobuf in '" + path + "'"; return false; }
    if (layers.empty()) { err = "'" + path + "' holds no NetParameter.layer entries (V1 caffemodels are not supported)"; return false; }
    return true;
}

namespace {

struct Spec { std::string name, bn; int cin, cout, k, stride, groups; bool bias, relu; double eps; std::string scale; };

void head_conv(std::vector<Spec> &v, const std::string &name, int cin, int cout, int k, bool relu) {
    v.push_back({name, name + "_bn", cin, cout, k, 1, 1, true, relu, 2e-5});
}

std::vector<Spec> mnet_specs() {
    // (cout, kind 0=full 1=dw 2=pw, stride) of mobilenet0_conv{i}_fwd  -- prototxt:11-1192
    static const int bb[27][3] = {
        {8, 0, 2}, {8, 1, 1}, {16, 2, 1}, {16, 1, 2}, {32, 2, 1}, {32, 1, 1}, {32, 2, 1}, {32, 1, 2}, {64, 2, 1},
        {64, 1, 1}, {64, 2, 1}, {64, 1, 2}, {128, 2, 1}, {128, 1, 1}, {128, 2, 1}, {128, 1, 1}, {128, 2, 1},
        {128, 1, 1}, {128, 2, 1}, {128, 1, 1}, {128, 2, 1}, {128, 1, 1}, {128, 2, 1}, {128, 1, 2}, {256, 2, 1},
        {256, 1, 1}, {256, 2, 1}};
    std::vector<Spec> v;
    int cin = 3;
    for (int i = 0; i < 27; i++) {
        int cout = bb[i][0], kind = bb[i][1];
        Spec s;
        s.name = "mobilenet0_conv" + std::to_string(i) + "_fwd";
        s.bn = "mobilenet0_batchnorm" + std::to_string(i) + "_fwd";
        s.cin = cin; s.cout = cout; s.k = kind == 2 ? 1 : 3; s.stride = bb[i][2];
        s.groups = kind == 1 ? cin : 1; s.bias = false; s.relu = true; s.eps = 1e-5;
        v.push_back(s);
        cin = cout;
    }
    head_conv(v, "rf_c3_lateral", 256, 64, 1, true);
    head_conv(v, "rf_c2_lateral", 128, 64, 1, true);
    head_conv(v, "rf_c1_red_conv", 64, 64, 1, true);
    head_conv(v, "rf_c2_aggr", 64, 64, 3, true);
    head_conv(v, "rf_c1_aggr", 64, 64, 3, true);
    for (const char *lv : {"c3", "c2", "c1"}) {
        std::string p = std::string("rf_") + lv + "_det";
        head_conv(v, p + "_conv1", 64, 32, 3, false);
        head_conv(v, p + "_context_conv1", 64, 16, 3, true);
        head_conv(v, p + "_context_conv2", 16, 16, 3, false);
        head_conv(v, p + "_context_conv3_1", 16, 16, 3, true);
        head_conv(v, p + "_context_conv3_2", 16, 16, 3, false);
    }
    for (int s : {32, 16, 8}) {
        std::string st = "_stride" + std::to_string(s);
        v.push_back({"face_rpn_cls_score" + st, "", 64, 4, 1, 1, 1, true, false, 0});
        v.push_back({"face_rpn_bbox_pred" + st, "", 64, 8, 1, 1, 1, true, false, 0});
        v.push_back({"face_rpn_landmark_pred" + st, "", 64, 20, 1, 1, 1, true, false, 0});
    }
    return v;
}

}  // namespace

bool build_mnet_model(const std::vector<RawLayer> &layers, Model &m, std::string &err, const std::vector<ConvSpec> *graph_specs) {
    std::map<std::string, const RawLayer *> by_name;
    for (auto &L : layers) by_name[L.name] = &L;
    std::map<std::string, const ConvSpec *> from_file;
    if (graph_specs) for (auto &g : *graph_specs) from_file[g.name] = &g;
    auto need = [&](const std::string &n, const char *type, size_t nblobs) -> const RawLayer * {
        auto it = by_name.find(n);
        if (it == by_name.end()) { err = "caffemodel lacks layer '" + n + "' (not an mnet25 RetinaFace model)"; return nullptr; }
        if (it->second->type != type) { err = "layer '" + n + "' has type " + it->second->type + ", expected " + type; return nullptr; }
        if (it->second->blobs.size() < nblobs) { err = "layer '" + n + "' holds " + std::to_string(it->second->blobs.size()) + " blobs, expected " + std::to_string(nblobs); return nullptr; }
        return it->second;
    };
    for (Spec s : mnet_specs()) {
        if (graph_specs) {
            // the prototxt's own description of this layer drives the folding; it must be the layer the plan expects
            auto it = from_file.find(s.name);
            if (it == from_file.end()) { err = "prototxt lacks convolution '" + s.name + "'"; return false; }
            const ConvSpec &g = *it->second;
            if (g.cout != s.cout || g.k != s.k || g.stride != s.stride || g.groups != s.groups || g.relu != s.relu || g.bn.empty() != s.bn.empty() || g.pad != (s.k == 3 ? 1 : 0)) {
                err = "prototxt: convolution '" + s.name + "' is " + std::to_string(g.cout) + " outputs, k" + std::to_string(g.k) + " s" + std::to_string(g.stride) + " g" +
                      std::to_string(g.groups) + " pad" + std::to_string(g.pad) + (g.relu ? " +ReLU" : "") + (g.bn.empty() ? "" : " +BN") + "; the engine's plan expects " + std::to_string(s.cout) +
                      " outputs, k" + std::to_string(s.k) + " s" + std::to_string(s.stride) + " g" + std::to_string(s.groups) + (s.relu ? " +ReLU" : "") + (s.bn.empty() ? "" : " +BN");
                return false;
            }
            s.bias = g.bias;
            s.eps = g.eps;
            if (!g.bn.empty()) { s.bn = g.bn; s.scale = g.scale; }
        }
        const RawLayer *L = need(s.name, "Convolution", s.bias ? 2 : 1);
        if (!L) return false;
        size_t wn = (size_t)s.cout * (s.cin / s.groups) * s.k * s.k;
        if (L->blobs[0].data.size() != wn) {
            err = "layer '" + s.name + "': weight blob has " + std::to_string(L->blobs[0].data.size()) + " values, expected " + std::to_string(wn);
            return false;
        }
        if (s.bias && L->blobs[1].data.size() != (size_t)s.cout) { err = "layer '" + s.name + "': bad bias size"; return false; }
        FoldedConv c;
        c.name = s.name; c.cin = s.cin; c.cout = s.cout; c.k = s.k; c.stride = s.stride; c.groups = s.groups; c.relu = s.relu;
        c.w.resize(wn); c.b.resize(s.cout);
        std::vector<double> scale(s.cout, 1.0), shift(s.cout, 0.0);
        if (!s.bn.empty()) {
            const RawLayer *B = need(s.bn, "BatchNorm", 3);
            if (!B) return false;
            const RawLayer *S = need(s.scale.empty() ? s.bn + "_scale" : s.scale, "Scale", 2);
            if (!S) return false;
            for (int i = 0; i < 2; i++)
                if (B->blobs[i].data.size() != (size_t)s.cout || S->blobs[i].data.size() != (size_t)s.cout) {
                    err = "layer '" + s.bn + "': BatchNorm/Scale blob size mismatch"; return false;
                }
            double sf = B->blobs[2].data.empty() ? 1.0 : (double)B->blobs[2].data[0];
            double inv = sf == 0.0 ? 0.0 : 1.0 / sf;  // batch_norm_layer.cpp: scale_factor == 0 ? 0 : 1/scale_factor
            for (int o = 0; o < s.cout; o++) {
                double mean = (double)B->blobs[0].data[o] * inv, var = (double)B->blobs[1].data[o] * inv;
                double k = (double)S->blobs[0].data[o] / std::sqrt(var + s.eps);
                scale[o] = k;
                shift[o] = (double)S->blobs[1].data[o] - mean * k;
            }
        }
        size_t per = wn / s.cout;
        for (int o = 0; o < s.cout; o++) {
            for (size_t i = 0; i < per; i++) c.w[o * per + i] = (float)((double)L->blobs[0].data[o * per + i] * scale[o]);
            double bias = s.bias ? (double)L->blobs[1].data[o] : 0.0;
            c.b[o] = (float)(bias * scale[o] + shift[o]);
        }
        m.convs[s.name] = std::move(c);
    }
    const char *ups[2] = {"rf_c3_upsampling", "rf_c2_upsampling"};
    for (int i = 0; i < 2; i++) {
        const RawLayer *L = need(ups[i], "Deconvolution", 1);
        if (!L) return false;
        if (L->blobs[0].data.size() != 64 * 16) { err = std::string("layer '") + ups[i] + "': expected 64x1x4x4 weights"; return false; }
        m.up_w[i] = L->blobs[0].data;
    }
    return true;
}

bool load_model(const std::string &caffemodel, const std::string &prototxt, const std::string &cache_path, Model &m, NetGraph *graph,
                int *cache_status, std::string &err, int &status) {
    // rf_status values (include/rf_b200.h): IO -2, MODEL -3
    if (cache_status) *cache_status = CACHE_NONE;
    NetGraph g;
    std::vector<ConvSpec> specs;
    bool io = false;
    if (!prototxt.empty()) {
        if (!read_prototxt(prototxt, g, err, io)) { status = io ? -2 : -3; return false; }
        if (!check_mnet_topology(g, err) || !graph_conv_specs(g, specs, err)) { status = -3; return false; }
        if (graph) *graph = g;
    }
    ModelCacheKey key;
    if (!cache_path.empty()) {
        if (!model_cache_key(caffemodel, prototxt, key)) { err = "cannot open caffemodel '" + caffemodel + "'"; status = -2; return false; }
        const int st = load_model_cache(cache_path, key, m);
        if (cache_status) *cache_status = st;
        if (st == CACHE_HIT) return true;
    }
    std::vector<RawLayer> layers;
    if (!read_caffemodel(caffemodel, layers, err, io)) { status = io ? -2 : -3; return false; }
    if (!build_mnet_model(layers, m, err, prototxt.empty() ? nullptr : &specs)) { status = -3; return false; }
    if (!cache_path.empty() && !save_model_cache(cache_path, key, m)) { err = "cannot write the model cache '" + cache_path + "'"; status = -2; return false; }
    return true;
}

bool read_int8_table(const std::string &path, std::map<std::string, float> &scales, std::string &err) {
    std::ifstream f(path);
    if (!f) { err = "cannot open INT8 calibration table '" + path + "'"; return false; }
    std::string line;
    if (!std::getline(f, line) || line.compare(0, 4, "TRT-") != 0) { err = "'" + path + "' is not a TensorRT calibration cache"; return false; }
    while (std::getline(f, line)) {
        size_t c = line.rfind(": ");
        if (c == std::string::npos) continue;
        std::string hex = line.substr(c + 2);
        while (!hex.empty() && (hex.back() == '\r' || hex.back() == ' ')) hex.pop_back();
        if (hex.size() != 8) continue;
        uint32_t bits = (uint32_t)std::strtoul(hex.c_str(), nullptr, 16);
        float v; memcpy(&v, &bits, 4);
        scales[line.substr(0, c)] = v;
    }
    if (scales.empty()) { err = "'" + path + "' holds no tensor scales"; return false; }
    return true;
}

void base_anchors_net3(int stride, float out[8]) {
    // generate_anchors(base_size 16, ratios {1}, scales) -- RetinaFace.cpp:35-104.  float storage,
    // double for the 0.5*(w-1) terms, like the reference expressions.
    int scales[2] = {0, 0};
    if (stride == 32) { scales[0] = 32; scales[1] = 16; }
    else if (stride == 16) { scales[0] = 8; scales[1] = 4; }
    else { scales[0] = 2; scales[1] = 1; }
    float bx1 = 0.f, by1 = 0.f, bx2 = 15.f, by2 = 15.f;
    float w = bx2 - bx1 + 1, h = by2 - by1 + 1;
    float xc = (float)(bx1 + 0.5 * (w - 1)), yc = (float)(by1 + 0.5 * (h - 1));
    float size = w * h, sc = size / 1.0f;
    float rw = std::round((float)std::sqrt((double)sc)), rh = std::round(rw * 1.0f);
    float rx1 = (float)(xc - 0.5 * (rw - 1)), ry1 = (float)(yc - 0.5 * (rh - 1));
    float rx2 = (float)(xc + 0.5 * (rw - 1)), ry2 = (float)(yc + 0.5 * (rh - 1));
    for (int i = 0; i < 2; i++) {
        float ww = rx2 - rx1 + 1, hh = ry2 - ry1 + 1;
        float cx = (float)(rx1 + 0.5 * (ww - 1)), cy = (float)(ry1 + 0.5 * (hh - 1));
        ww = ww * scales[i]; hh = hh * scales[i];
        out[4 * i + 0] = (float)(cx - 0.5 * (ww - 1));
        out[4 * i + 1] = (float)(cy - 0.5 * (hh - 1));
        out[4 * i + 2] = (float)(cx + 0.5 * (ww - 1));
        out[4 * i + 3] = (float)(cy + 0.5 * (hh - 1));
    }
}

}  // namespace rf
