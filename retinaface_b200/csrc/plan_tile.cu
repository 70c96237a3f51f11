// plan_tile.cu -- the FP16 tensor-core layer plan built from tile chains (tile_chain.cuh): which layers of the reference's graph
// (model/mnet-deconv-0517.prototxt) run as stages of which persistent kernel, the shared-memory / TMEM budget of every
// chain, the packed weights, the TMA tensor maps.  Where a chain does not fit (wide maps, 256-channel layers) the round-1
// per-layer kernels of plan_fp.cu take over, layer by layer.
//
//   stem (round-1 k_stem_tc)                          conv0 + dw1 + pw2
//   chain A  @ /4    dw3+pw4 (s2) -> dw5+pw6                                        -> relu6
//   chain B  @ /8    dw7+pw8 (s2) -> dw9+pw10 -> rf_c1_red_conv                      -> relu10 (C1), rf_c1_red_conv_relu
//   chain C  @ /16   dw11+pw12 (s2) -> dw13+pw14 -> dw15+pw16                        -> relu16
//   chain D  @ /16   dw17+pw18 -> dw19+pw20 -> dw21+pw22 -> rf_c2_lateral            -> relu22 (C2), rf_c2_lateral_relu
//   /32              dw23+pw24, dw25+pw26, rf_c3_lateral: tile chain when it fits, else round-1 kernels
//   level kernels    [FPN merge + rf_c*_aggr]  and  [SSH det/context convs + predictors + decode (+ last-block NMS)]
#include <cstdlib>

#include "engine_internal.cuh"
#include "tile_chain.cuh"

namespace rf_eng {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

constexpr int TCH_SMEM_LIMIT = 226 * 1024;    // dynamic shared memory a chain may use (227 KB per CTA minus the static part)

// NHWC FP16 tensor [n][H][W][C] as a 4-D map {C, W, H, n}; box {bc, bw, bh, 1}; element strides {1, es, es, 1};
// swizzle mode by the bytes of one box row (bc * 2: 32 / 64 / 128)
CUtensorMap make_map(const void *base, int C, int W, int H, int n, int bc, int bw, int bh, int es) {
    CUtensorMap m;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t est[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
    const CUtensorMapSwizzle sw = bc * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (bc * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(base), dims, strides, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw PlanFail{RF_ERR_CUDA, fmt("cuTensorMapEncodeTiled failed (%d) for a %dx%dx%d tensor, box %dx%dx%d", (int)r, C, W, H, bc, bw, bh)};
    return m;
}

int round_up(int v, int a) { return (v + a - 1) / a * a; }

}  // namespace

// Host description of one chain: logical buffers + stages (ChainSpec), then the finalised kernel arguments.
struct TileChain {
    struct LBuf {
        int C = 0;
        int halo = 0;                 // rows beyond the tile the consumers need, each side
        bool stored = false;          // the owned rows are TMA-stored into an arena tensor ...
        int store_tensor = -1;        // ... this one (assigned once the chain is known to fit)
        int first = 1 << 30, last = -1;   // stage indices (input: first = -1)
        bool is_input = false, is_merge = false;
    };
    struct LStage {
        int type = TCH_CONV, Cin = 0, N = 0, taps = 1, stride = 1, in_buf = 0;
        std::vector<int> ob_buf, ob_c16, ob_relu;
        std::vector<__half> wd_img, wp_img;
        std::vector<float> bd, bp;
        int hs = 0;
        int store_buf = -1;
        double flops = 0;             // per output position
    };
    std::string name;
    std::vector<LBuf> bufs;
    std::vector<LStage> stages;
    int in_tensor = -1, in_C = 0, in_W = 0, in_H = 0, in_s2 = 0;   // chain input (arena tensor) and its map size
    int W = 0, H = 0;                 // resolution the chain works at
    int merge_tensor = -1;            // FPN merge: coarser level (64 channels, W/2 x H/2)
    std::vector<__half> merge_w;      // [16 taps][64]
    int level = -1;                   // TCH_HEAD: FPN level (0: stride 32)
    bool fused_nms = false;
    // finalised
    TchArgs args{};
    size_t bias_off = 0;              // float offset into d_weights
    int store_buf_of[3] = {-1, -1, -1}, store_C[3] = {0, 0, 0};     // store map i <- logical buffer
    int nstores = 0;
    int TH = 0, mtiles = 0;           // rows per tile, MMA tiles per CTA tile (cost proxy)
};

namespace {

// ---- chain construction helpers ----------------------------------------------------------------------------------------
int add_buf(TileChain &c, int C, bool stored = false, int store_tensor = -1) {
    TileChain::LBuf b;
    b.C = C;
    b.stored = stored;
    b.store_tensor = store_tensor;
    c.bufs.push_back(b);
    return (int)c.bufs.size() - 1;
}

// depthwise diagonal B tiles: tap t, 16-channel slab k -> 16x16 K-major no-swizzle tile [k/8][n][8] with w on the diagonal
std::vector<__half> pack_dw_tiles(const FoldedConv &dw) {
    const int C = dw.cout, nk = C / 16;
    std::vector<__half> img((size_t)9 * nk * 256, __float2half(0.f));
    for (int t = 0; t < 9; t++)
        for (int k = 0; k < nk; k++)
            for (int i = 0; i < 16; i++)
                img[((size_t)(t * nk + k)) * 256 + ((size_t)(i / 8) * 16 + i) * 8 + i % 8] = __float2half(dw.w[(size_t)(k * 16 + i) * 9 + t]);
    return img;
}

int add_dwpw(TileChain &c, int in_buf, const FoldedConv &dw, const FoldedConv &pw, int out_buf) {
    TileChain::LStage s;
    s.type = TCH_DWPW; s.Cin = dw.cout; s.N = pw.cout; s.taps = 9; s.stride = dw.stride; s.in_buf = in_buf;
    s.wd_img = pack_dw_tiles(dw);
    s.bd = dw.b;
    int Kpad = 0;
    s.wp_img = pack_tc_weights({&pw}, s.bp, Kpad);
    for (int j = 0; j < s.N / 16; j++) { s.ob_buf.push_back(out_buf); s.ob_c16.push_back(j); s.ob_relu.push_back(1); }
    s.flops = 2.0 * s.Cin * 9 + 2.0 * s.Cin * s.N;
    c.stages.push_back(std::move(s));
    return (int)c.stages.size() - 1;
}

// convolution (all `cs` share the input and are concatenated along N); per conv: destination buffer, channel offset, ReLU
struct ConvDst { int buf, coff, relu; };
int add_conv(TileChain &c, int in_buf, const std::vector<const FoldedConv *> &cs, const std::vector<ConvDst> &dst) {
    TileChain::LStage s;
    s.type = TCH_CONV; s.Cin = cs[0]->cin; s.taps = cs[0]->k * cs[0]->k; s.in_buf = in_buf;
    int Kpad = 0;
    s.wp_img = pack_tc_weights(cs, s.bp, Kpad);
    s.N = (int)s.bp.size();
    for (size_t i = 0; i < cs.size(); i++)
        for (int j = 0; j < cs[i]->cout / 16; j++) { s.ob_buf.push_back(dst[i].buf); s.ob_c16.push_back(dst[i].coff / 16 + j); s.ob_relu.push_back(dst[i].relu); }
    s.flops = 2.0 * s.Cin * s.taps * s.N;
    c.stages.push_back(std::move(s));
    return (int)c.stages.size() - 1;
}

// the three predictor convs of one level as one N = 32 GEMM with hi + lo FP16 weight pieces (FP32-grade products)
int add_head(TileChain &c, int in_buf, const FoldedConv *cs[3]) {
    TileChain::LStage s;
    s.type = TCH_HEAD; s.Cin = 64; s.N = 32; s.taps = 1; s.in_buf = in_buf;
    std::vector<__half> hi((size_t)64 * 32), lo((size_t)64 * 32);
    int r = 0;
    for (int q = 0; q < 3; q++)
        for (int o = 0; o < cs[q]->cout; o++, r++) {
            s.bp.push_back(cs[q]->b[o]);
            for (int ci = 0; ci < 64; ci++) {
                const float w = cs[q]->w[(size_t)o * 64 + ci];
                const __half wh = __float2half(w);
                hi[((size_t)(ci / 8) * 32 + r) * 8 + ci % 8] = wh;
                lo[((size_t)(ci / 8) * 32 + r) * 8 + ci % 8] = __float2half(w - __half2float(wh));
            }
        }
    s.wp_img = hi;
    s.wp_img.insert(s.wp_img.end(), lo.begin(), lo.end());
    s.flops = 2.0 * 64 * 32 * 2;
    c.stages.push_back(std::move(s));
    return (int)c.stages.size() - 1;
}

// ---- finalisation: halos, rows, shared-memory placement, TMEM sets, kernel arguments --------------------------------------
// returns false when the chain does not fit with TH rows per tile
bool finalize_chain(rf_handle h, TileChain &c, int TH, int max_faces, bool resident) {
    const int ns = (int)c.stages.size(), nb = (int)c.bufs.size();
    if (ns > TCH_MAX_STAGES || nb > TCH_MAX_BUFS) return false;
    const int Wl = c.W + 2;
    if (Wl > 256 || (c.in_s2 && 2 * Wl > 256)) return false;
    // halos (reverse stage order: all consumers of a buffer come after its producers)
    for (auto &b : c.bufs) { b.halo = 0; b.first = 1 << 30; b.last = -1; }
    for (int s = ns - 1; s >= 0; s--) {
        auto &st = c.stages[s];
        int hs = 0;
        for (int b : st.ob_buf) hs = std::max(hs, c.bufs[b].halo);
        st.hs = hs;
        const bool k3 = st.type == TCH_DWPW || st.taps == 9;
        c.bufs[st.in_buf].halo = std::max(c.bufs[st.in_buf].halo, hs + (k3 ? 1 : 0));
    }
    const int HT = c.bufs[0].halo;       // buffer 0 is the chain input
    // lifetimes
    c.bufs[0].first = -1;
    for (int b = 0; b < nb; b++) if (c.bufs[b].is_merge) { c.bufs[b].first = -1; c.bufs[b].last = -1; }
    for (int s = 0; s < ns; s++) {
        auto &st = c.stages[s];
        c.bufs[st.in_buf].last = std::max(c.bufs[st.in_buf].last, s);
        for (int b : st.ob_buf) { c.bufs[b].first = std::min(c.bufs[b].first, s); c.bufs[b].last = std::max(c.bufs[b].last, s); }
    }
    for (auto &b : c.bufs) if (b.stored) b.last = ns;     // TMA stores read the buffer until the next tile starts
    if (c.merge_tensor >= 0) c.bufs[0].last = std::max(c.bufs[0].last, 0);

    TchArgs &a = c.args;
    a = TchArgs{};
    a.nstages = ns; a.nbufs = nb;
    a.Wl = Wl; a.HT = HT; a.TH = TH;
    a.W = c.W; a.H = c.H;
    a.tiles_per_img = (c.H + TH - 1) / TH;
    a.in_s2 = c.in_s2; a.in_C = c.in_C;
    // buffers
    std::vector<int> bytes(nb);
    for (int b = 0; b < nb; b++) {
        auto &lb = c.bufs[b];
        TchBuf &tb = a.buf[b];
        tb.row = lb.C >= 64 ? 128 : lb.C * 2;
        tb.slabs = std::max(1, lb.C / 64);
        tb.rows_lo = HT - lb.halo;
        tb.nrows = TH + 2 * lb.halo;
        // positions in front of the first row: the (-1, -1) tap of the first computed position reads one back; TMA-loaded
        // buffers need their row 0 128-byte aligned (8), TMA-stored ones their position (row, lx = 1) (7 for rows < 128 bytes)
        tb.slack = (lb.stored && tb.row < 128) ? 7 : 8;
        if (lb.is_merge) {
            a.merge_rows = (TH + 2 * HT) / 2 + 3;
            tb.rows_lo = 0; tb.nrows = a.merge_rows;
            tb.slab_stride = round_up(a.merge_rows * (c.W / 2 + 2) * 128, 1024);
            bytes[b] = tb.slab_stride;
            continue;
        }
        if (b == 0 && c.in_s2) {
            const int Hp = TH + 2 * c.stages[0].hs + 1;           // rows of each parity plane
            if (2 * Hp > 256) return false;
            tb.rows_lo = HT - c.stages[0].hs - 1; tb.nrows = Hp;
            tb.slab_stride = round_up((tb.slack + Hp * Wl + 8) * tb.row, 1024);
            a.plane_stride = tb.slabs * tb.slab_stride;
            bytes[b] = 4 * a.plane_stride;
            a.in_bytes = 4u * (unsigned)tb.slabs * (unsigned)(std::min(lb.C, 64) * 2 * Wl * Hp);
            continue;
        }
        tb.slab_stride = round_up((tb.slack + tb.nrows * Wl + 8) * tb.row, 1024);
        bytes[b] = tb.slabs * tb.slab_stride;
        if (b == 0) a.in_bytes = (unsigned)tb.slabs * (unsigned)(std::min(lb.C, 64) * 2 * Wl * tb.nrows);
    }
    if (c.merge_tensor >= 0) a.merge_bytes = (unsigned)(64 * 2 * (c.W / 2 + 2) * a.merge_rows);
    // first-fit placement by first use
    std::vector<int> order(nb);
    for (int i = 0; i < nb; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return c.bufs[x].first < c.bufs[y].first; });
    std::vector<int> placed;
    int top = 0;
    for (int id : order) {
        int off = 0;
        bool moved = true;
        while (moved) {
            moved = false;
            for (int p : placed) {
                const bool live = !(c.bufs[p].last < c.bufs[id].first || c.bufs[id].last < c.bufs[p].first);
                const bool mem = off < a.buf[p].off + bytes[p] && a.buf[p].off < off + bytes[id];
                if (live && mem) { off = a.buf[p].off + bytes[p]; moved = true; }
            }
        }
        a.buf[id].off = off;
        top = std::max(top, off + bytes[id]);
        placed.push_back(id);
    }
    // stages
    int set_cols = 0, wd_max = 0, wp_max = 0, read_end = 0, mt = 0;
    std::vector<float> bias;
    for (int s = 0; s < ns; s++) {
        auto &ls = c.stages[s];
        TchStage &st = a.st[s];
        st.type = ls.type; st.Cin = ls.Cin; st.N = ls.N; st.taps = ls.taps; st.stride = ls.stride; st.in_buf = ls.in_buf;
        if (ls.stride == 2 && (s != 0 || !c.in_s2)) return false;
        st.rows_lo = HT - ls.hs; st.nrows = TH + 2 * ls.hs;
        if ((int)ls.ob_buf.size() > 16 || ls.N % 16 || ls.Cin % 16 || ls.N > 256) return false;
        for (size_t j = 0; j < ls.ob_buf.size(); j++) { st.ob_buf[j] = (unsigned char)ls.ob_buf[j]; st.ob_c16[j] = (unsigned char)ls.ob_c16[j]; st.ob_relu[j] = (unsigned char)ls.ob_relu[j]; }
        st.wd_bytes = (int)ls.wd_img.size() * 2; st.wp_bytes = (int)ls.wp_img.size() * 2;
        wd_max = std::max(wd_max, st.wd_bytes); wp_max = std::max(wp_max, st.wp_bytes);
        st.bias_dw = (int)bias.size(); bias.insert(bias.end(), ls.bd.begin(), ls.bd.end());
        while (bias.size() % 4) bias.push_back(0.f);
        st.bias_pw = (int)bias.size(); bias.insert(bias.end(), ls.bp.begin(), ls.bp.end());
        while (bias.size() % 4) bias.push_back(0.f);
        set_cols = std::max(set_cols, ls.type == TCH_DWPW ? ls.Cin + ls.N : ls.N);
        st.store_buf = -1; st.store_map = -1;
        // furthest byte a (partial) MMA tile of this stage may read: rows past the range + one tap
        const TchBuf &bi = a.buf[ls.in_buf];
        const int ntile = (st.nrows * Wl + 127) / 128;
        mt += ntile * (ls.type == TCH_DWPW ? 2 : 1);
        const int pos0 = (ls.type == TCH_DWPW && ls.stride == 2) ? bi.slack : bi.slack + (st.rows_lo - bi.rows_lo) * Wl;
        const int last_plane = (ls.type == TCH_DWPW && ls.stride == 2) ? 3 * a.plane_stride : 0;
        read_end = std::max(read_end, bi.off + last_plane + (bi.slabs - 1) * bi.slab_stride + (pos0 + ntile * 128 + Wl + 2) * bi.row);
    }
    c.mtiles = mt;
    // a stage's outputs are stored once the LAST stage writing the buffer is complete
    c.nstores = 0;
    for (int b = 0; b < nb; b++) {
        if (!c.bufs[b].stored) continue;
        if (c.nstores == 3) return false;
        int last_writer = -1;
        for (int s = 0; s < ns; s++) for (int ob : c.stages[s].ob_buf) if (ob == b) last_writer = s;
        if (last_writer < 0 || a.st[last_writer].store_buf >= 0) return false;
        a.st[last_writer].store_buf = b; a.st[last_writer].store_map = c.nstores;
        c.store_buf_of[c.nstores] = b; c.store_C[c.nstores] = c.bufs[b].C;
        // TMA store sources (position (row, lx = 1) of every owned row) must be 128-byte aligned
        if (((a.buf[b].slack + 1) * a.buf[b].row) % 128 || (Wl * a.buf[b].row) % 128) return false;
        c.nstores++;
    }
    if (c.merge_tensor >= 0) {
        a.merge_C = 64;
        for (int b = 0; b < nb; b++) if (c.bufs[b].is_merge) a.merge_buf = b;
        a.merge_w_bias = (int)bias.size();
        const float *mw = reinterpret_cast<const float *>(c.merge_w.data());
        bias.insert(bias.end(), mw, mw + c.merge_w.size() / 2);
    }
    a.nsets = 2 * set_cols <= 512 ? 2 : 1;
    a.set_cols = set_cols;
    if (set_cols > 512) return false;
    // weights, bias arena, NMS scratch behind the buffers
    const int nms_need = c.fused_nms ? (int)((sizeof(NmsSmem) + 15) / 16 * 16 + sizeof(int) * (size_t)max_faces) : 0;
    a.resident = resident ? 1 : 0;
    int cursor = round_up(top, 1024);
    if (resident) {
        // every stage keeps its own weight regions for the CTA's lifetime
        for (int s = 0; s < ns; s++) {
            a.st[s].wd_smem = cursor; cursor += round_up(a.st[s].wd_bytes, 128);
            a.st[s].wp_smem = cursor; cursor += round_up(a.st[s].wp_bytes, 128);
        }
        a.wd_smem = a.wp_smem = cursor;
        // NMS scratch: the input tile's region (dead by then, never TMA-stored) when large enough
        if (nms_need && bytes[0] >= nms_need) a.head.nms_smem = a.buf[0].off;
        else { a.head.nms_smem = cursor; cursor += round_up(nms_need, 128); }
    } else {
        a.wd_smem = cursor; cursor += round_up(wd_max, 128);
        a.wp_smem = cursor; cursor += std::max(round_up(wp_max, 128), round_up(nms_need, 128));
        a.head.nms_smem = a.wp_smem;
    }
    a.bias_smem = cursor;
    a.bias_floats = (int)bias.size();
    a.smem_bytes = std::max(a.bias_smem + a.bias_floats * 4, read_end) + 1024;      // + alignment slack of the dynamic base
    if (a.smem_bytes > TCH_SMEM_LIMIT) return false;
    c.TH = TH;
    c.bias_off = (size_t)-1;
    // (weights go to the handle's arenas once the geometry is final: commit_chain)
    c.args.bias_floats = (int)bias.size();
    // stash the bias vector in the chain until commit
    h->tile_bias_tmp = bias;
    return true;
}

// copies the packed weights + bias arena of a finalised chain into the handle's upload staging
void commit_chain(Builder &B, TileChain &c) {
    rf_handle h = B.h;
    for (size_t s = 0; s < c.stages.size(); s++) {
        auto &ls = c.stages[s];
        if (!ls.wd_img.empty()) c.args.st[s].wd_off = (int)(B.add_weights_h(ls.wd_img) * 2);
        c.args.st[s].wp_off = (int)(B.add_weights_h(ls.wp_img) * 2);
    }
    c.bias_off = B.add_weights(h->tile_bias_tmp);
}

int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// picks the tile height: among the heights that fit, the one with the least (waves x MMA tiles per CTA), ties to the taller
bool choose_tile(rf_handle h, TileChain &c, int max_batch, int max_faces, int force_th) {
    double best = 1e30;
    int best_th = 0;
    bool best_res = false;
    const bool allow_res = env_int("RF_TILE_RESIDENT", 1) != 0;
    for (int th = 1; th <= std::min(c.H, 32); th++) {
        if (force_th > 0 && th != force_th) continue;
        for (int res = allow_res ? 1 : 0; res >= 0; res--) {
            if (!finalize_chain(h, c, th, max_faces, res != 0)) continue;
            const long tiles = (long)max_batch * c.args.tiles_per_img;
            // two CTAs share an SM when shared memory (and TMEM: 512 columns) allows: they fill each other's hand-off gaps
            const int occ = (2 * (c.args.smem_bytes + 1024) <= 232448 && 2 * tch_tmem_cols(c.args.nsets * c.args.set_cols) <= 512) ? 2 : 1;
            const double waves = std::ceil((double)tiles / (148.0 * (occ == 2 ? 1.6 : 1.0)));
            // MMA tiles + per-stage hand-offs (streamed weights: one exposed load per stage) + per-tile set-up
            const double cost = waves * (c.mtiles + (res ? 1.0 : 4.0) * c.stages.size() + 4.0);
            if (cost < best || (cost == best && th > best_th)) { best = cost; best_th = th; best_res = res != 0; }
            break;      // resident fits: no need to look at streaming for this height
        }
    }
    if (!best_th) return false;
    return finalize_chain(h, c, best_th, max_faces, best_res);
}

int forced_th(const std::string &chain) {
    // RF_TILE_TH="A=4,B=2,ssh2=7": per-chain tile heights for experiments
    const char *v = getenv("RF_TILE_TH");
    if (!v) return 0;
    std::string s(v);
    size_t p = 0;
    while (p < s.size()) {
        size_t e = s.find(',', p);
        if (e == std::string::npos) e = s.size();
        const std::string item = s.substr(p, e - p);
        const size_t eq = item.find('=');
        if (eq != std::string::npos && item.substr(0, eq) == chain) return atoi(item.c_str() + eq + 1);
        p = e + 1;
    }
    return 0;
}

void launch_chain(rf_handle h, const std::shared_ptr<TileChain> &cp, int n, cudaStream_t st) {
    TileChain &c = *cp;
    TchArgs a = c.args;
    a.nimg = n;
    a.ntiles = n * a.tiles_per_img;
    a.dbg = h->tile_dbg_dev;
    a.trace = nullptr;
#ifdef RF_TCH_TRACE
    {   // one trace buffer per chain (host-mapped), reset at every launch: holds the LAST launch's timeline of CTA 0
        static std::map<const TileChain *, unsigned long long *> bufs;
        auto it = bufs.find(&c);
        if (it == bufs.end()) {
            unsigned long long *p = nullptr;
            CK(cudaHostAlloc(&p, 8 * 1024, cudaHostAllocMapped));
            it = bufs.emplace(&c, p).first;
        }
        CK(cudaStreamSynchronize(st));
        if (it->second[0]) {
            const unsigned n = (unsigned)std::min<unsigned long long>(it->second[0], 500);
            fprintf(stderr, "TRACE %s:", c.name.c_str());
            for (unsigned i = 0; i < n; i++) fprintf(stderr, " %llu@%llu", it->second[1 + 2 * i], it->second[2 + 2 * i] - it->second[2]);
            fprintf(stderr, "\n");
        }
        memset(it->second, 0, 8 * 1024);
        a.trace = it->second;
    }
#endif
    a.warena = reinterpret_cast<const unsigned char *>(h->d_weights_h);
    a.bias = h->d_weights + c.bias_off;
    TchMaps maps;
    memset(&maps, 0, sizeof maps);
    const TchBuf &b0 = a.buf[0];
    const int bc = std::min(c.in_C, 64);
    if (c.in_s2) maps.in = make_map(h->tptr(c.in_tensor), c.in_C, c.in_W, c.in_H, n, bc, 2 * a.Wl, 2 * b0.nrows, 2);
    else maps.in = make_map(h->tptr(c.in_tensor), c.in_C, c.in_W, c.in_H, n, bc, a.Wl, b0.nrows, 1);
    if (c.merge_tensor >= 0) maps.aux = make_map(h->tptr(c.merge_tensor), 64, c.W / 2, c.H / 2, n, 64, c.W / 2 + 2, a.merge_rows, 1);
    for (int i = 0; i < c.nstores; i++) maps.st[i] = make_map(h->tptr(c.bufs[c.store_buf_of[i]].store_tensor), c.store_C[i], c.W, c.H, n, std::min(c.store_C[i], 64), c.W, 1, 1);
    if (c.level >= 0) {
        a.head.lv = h->lv[c.level];
        a.head.pb = h->pb;
        a.head.params = h->d_params;
        a.head.net_w = h->cfg.net_w; a.head.net_h = h->cfg.net_h;
        a.head.done = h->pb.tile_done;
        a.head.expected = (c.fused_nms && !h->profiling) ? h->tile_expected : 0;    // rf_profile_layers launches single steps: no last-block NMS then
        for (int k = 0; k < 3; k++) a.head.blobs[k] = h->blobs_in_plan ? h->d_blobs[3 * c.level + k] : nullptr;
    }
    const int grid = std::min(a.ntiles, 148);
    CK(launch_k(k_tile_chain<0>, dim3((unsigned)grid), dim3(TCH_THREADS), (size_t)a.smem_bytes, st, maps, a));
}

// adds the step of a finalised chain
void add_chain_step(Builder &B, std::shared_ptr<TileChain> c, int lane, double bytes_per_img) {
    rf_handle h = B.h;
    commit_chain(B, *c);
    h->chains.push_back(c);
    Step s;
    s.name = c->name;
    s.lane = lane;
    s.in = {c->in_tensor};
    if (c->merge_tensor >= 0) s.in.push_back(c->merge_tensor);
    for (int i = 0; i < c->nstores; i++) s.out.push_back(c->bufs[c->store_buf_of[i]].store_tensor);
    double fl = 0;
    for (auto &ls : c->stages) fl += ls.flops * c->W * c->H;
    s.flops_per_img = fl;
    s.bytes_per_img = bytes_per_img;
    s.launch = [h, c](int n, cudaStream_t st) { launch_chain(h, c, n, st); };
    B.step(std::move(s));
}

}  // namespace

// one line per chain: geometry, budgets (for rf_plan_describe and the CPU-side planner tests)
std::string describe_chains(rf_handle h) {
    std::string out;
    for (auto &cp : h->chains) {
        const TileChain &c = *cp;
        const TchArgs &a = c.args;
        out += fmt("%s: %dx%d map, TH=%d HT=%d Wl=%d, %d tiles/image, %d stages, %d MMA tiles/tile, smem %d B (%s weights), TMEM %d x %d cols, stores %d\n",
                   c.name.c_str(), c.W, c.H, a.TH, a.HT, a.Wl, a.tiles_per_img, a.nstages, c.mtiles, a.smem_bytes, a.resident ? "resident" : "streamed", a.nsets,
                   a.set_cols, c.nstores);
    }
    return out;
}

cudaError_t tile_init() {
    if (!g_encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (!fn) return cudaErrorNotSupported;
        g_encode = (EncodeTiledFn)fn;
    }
    return cudaFuncSetAttribute(k_tile_chain<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCH_SMEM_LIMIT);
}

// RF_TILE_MASK bits: which parts of the FP16 plan run as tile chains; the others use the round-1 kernels.
enum { TM_A = 1, TM_B = 2, TM_C = 4, TM_D = 8, TM_E = 16, TM_AGGR = 32, TM_SSH = 64, TM_HEAD = 128, TM_NMS = 256, TM_ALL = 511 };
// Measured on a B200 (profiles/r02_mask_sweep.txt, tools/mask_sweep.py): with ONE execution context (a single forward at a
// time: the blocking / latency mode) the SSH + predictor + NMS chains (and, at batch 1-2, the merge+aggr chains) shorten or
// tie the step; with several contexts overlapping batches (throughput mode) what counts is SM time per step, where the
// per-layer kernels win -- there only the fused decode + NMS tail is taken over.  The backbone chains (depthwise on tensor
// cores) lose to the per-layer kernels since those run their stencils on FHFMA.
constexpr unsigned TM_LATENCY = TM_SSH | TM_HEAD | TM_NMS;
constexpr unsigned TM_LATENCY_SMALL = TM_AGGR | TM_SSH | TM_HEAD | TM_NMS;      // max_batch <= 2
constexpr unsigned TM_THROUGHPUT = 0;

// builds the plan for one mask; false: the predictors could not be fused at all three levels (the caller retries without)
static bool build_tiles_with_mask(rf_handle h, unsigned mask) {
    Builder B{h, h->cfg.net_h, h->cfg.net_w};
    const Model &m = h->model;
    const int H = h->cfg.net_h, W = h->cfg.net_w, mb = h->cfg.max_batch, mf = h->cfg.max_faces;
    if (!(mask & TM_SSH)) mask &= ~(TM_HEAD | TM_NMS);
    if (!(mask & TM_HEAD)) mask &= ~TM_NMS;
    const double es = 2;
    auto pair = [&](int i) -> std::pair<const FoldedConv *, const FoldedConv *> {
        return {&m.conv("mobilenet0_conv" + std::to_string(i) + "_fwd"), &m.conv("mobilenet0_conv" + std::to_string(i + 1) + "_fwd")};
    };
    auto relu_name = [](int i) { return "mobilenet0_relu" + std::to_string(i) + "_fwd"; };

    int cur = plan_stem_tc(B);
    int cur_h = H / 2, cur_w = W / 2, cur_c = 16;

    // a backbone chain over pairs `is` (first may be stride 2) + optionally one trailing 1x1 conv on the last pair's output;
    // falls back to round-1 kernels pair by pair
    int lat1 = -1, lat2 = -1, lat3 = -1;
    auto backbone = [&](const std::string &name, std::vector<int> is, bool enabled, const char *lat_conv, int *lat_out, int lane_lat) {
        const int S = pair(is[0]).first->stride;
        const int oh = cur_h / S, ow = cur_w / S;
        const int Cout = pair(is.back()).second->cout;
        bool done = false;
        if (enabled) {
            auto c = std::make_shared<TileChain>();
            c->name = "tile_" + name;
            c->in_tensor = cur; c->in_C = cur_c; c->in_W = cur_w; c->in_H = cur_h; c->in_s2 = S == 2;
            c->W = ow; c->H = oh;
            int b = add_buf(*c, cur_c);
            c->bufs[b].is_input = true;
            int tout = -1, tlat = -1;
            for (size_t k = 0; k < is.size(); k++) {
                auto pr = pair(is[k]);
                int nb = add_buf(*c, pr.second->cout, k + 1 == is.size());
                add_dwpw(*c, b, *pr.first, *pr.second, nb);
                b = nb;
            }
            if (lat_conv) {
                int lb = add_buf(*c, 64, true);
                add_conv(*c, b, {&m.conv(lat_conv)}, {{lb, 0, 1}});
            }
            // tensors are created only once the chain is known to fit (a failed chain leaves no trace in the plan)
            if (choose_tile(h, *c, mb, mf, forced_th(name))) {
                tout = B.tensor(relu_name(is.back() + 1), oh, ow, Cout);
                c->bufs[(int)is.size()].store_tensor = tout;
                if (lat_conv) { tlat = B.tensor(std::string(lat_conv) + "_relu", oh, ow, 64); c->bufs[(int)is.size() + 1].store_tensor = tlat; }
                add_chain_step(B, c, 0, ((double)cur_h * cur_w * cur_c + (double)oh * ow * Cout + (lat_conv ? (double)oh * ow * 64 : 0.0)) * es);
                cur = tout; cur_h = oh; cur_w = ow; cur_c = Cout;
                if (lat_out) *lat_out = tlat;
                done = true;
            }
        }
        if (!done) {
            for (int i : is) {
                const int S2 = pair(i).first->stride;
                cur = plan_pair_legacy(B, i, cur, cur_h, cur_w);
                cur_h /= S2; cur_w /= S2; cur_c = pair(i).second->cout;
            }
            if (lat_conv) {
                int tl = B.tensor(std::string(lat_conv) + "_relu", cur_h, cur_w, 64);
                plan_conv_legacy(B, std::string(lat_conv) + "_1x1", {&m.conv(lat_conv)}, cur, cur_h, cur_w, tl, 64, 0, 64, 1, -1, 0, 0, 0, lane_lat);
                if (lat_out) *lat_out = tl;
            }
        }
    };
    // RF_TILE_SINGLE=1: one chain per depthwise+pointwise pair (no halo recomputation between layers; more, smaller kernels)
    const bool single = env_int("RF_TILE_SINGLE", 0) != 0;
    if (single) {
        backbone("A3", {3}, mask & TM_A, nullptr, nullptr, 0);
        backbone("A5", {5}, mask & TM_A, nullptr, nullptr, 0);
        backbone("B7", {7}, mask & TM_B, nullptr, nullptr, 0);
        backbone("B9", {9}, mask & TM_B, "rf_c1_red_conv", &lat1, 1);
    } else {
        backbone("A", {3, 5}, mask & TM_A, nullptr, nullptr, 0);
        backbone("B", {7, 9}, mask & TM_B, "rf_c1_red_conv", &lat1, 1);
    }
    const int h8 = cur_h, w8 = cur_w;
    if (single) {
        backbone("C11", {11}, mask & TM_C, nullptr, nullptr, 0);
        backbone("C13", {13}, mask & TM_C, nullptr, nullptr, 0);
        backbone("C15", {15}, mask & TM_C, nullptr, nullptr, 0);
        backbone("D17", {17}, mask & TM_D, nullptr, nullptr, 0);
        backbone("D19", {19}, mask & TM_D, nullptr, nullptr, 0);
        backbone("D21", {21}, mask & TM_D, "rf_c2_lateral", &lat2, 2);
    } else {
        backbone("C", {11, 13, 15}, mask & TM_C, nullptr, nullptr, 0);
        backbone("D", {17, 19, 21}, mask & TM_D, "rf_c2_lateral", &lat2, 2);
    }
    const int h16 = cur_h, w16 = cur_w;
    backbone("E", {23}, mask & TM_E, nullptr, nullptr, 0);
    backbone("F", {25}, false, "rf_c3_lateral", &lat3, 0);
    const int h32 = cur_h, w32 = cur_w;

    // ---- FPN top-down + SSH ---------------------------------------------------------------------------------------------
    // levels: 0 = stride 32 (lat3, no merge), 1 = stride 16, 2 = stride 8
    const char *lvn[3] = {"c3", "c2", "c1"};
    const int fh[3] = {h32, h16, h8}, fw[3] = {w32, w16, w8};
    int feat_in[3] = {lat3, -1, -1};
    int lat[3] = {lat3, lat2, lat1};
    // expected tile count per image (last-block NMS) is known only when all three SSH chains exist
    std::shared_ptr<TileChain> ssh_chain[3];
    auto aggr_level = [&](int l) {
        // merged = lat[l] + upsample(feat_in[l-1]); aggr 3x3 64->64
        const std::string an = std::string("rf_") + lvn[l] + "_aggr";
        int taggr = -1;
        bool done = false;
        if (mask & TM_AGGR) {
            auto c = std::make_shared<TileChain>();
            c->name = std::string("tile_") + lvn[l] + "_merge+aggr";
            c->in_tensor = lat[l]; c->in_C = 64; c->in_W = fw[l]; c->in_H = fh[l];
            c->W = fw[l]; c->H = fh[l];
            c->merge_tensor = feat_in[l - 1];
            c->merge_w.resize(16 * 64);
            for (int ch = 0; ch < 64; ch++)
                for (int t = 0; t < 16; t++) c->merge_w[t * 64 + ch] = __float2half(m.up_w[l - 1][ch * 16 + t]);
            int bi = add_buf(*c, 64);
            int bm = add_buf(*c, 64);
            c->bufs[bm].is_merge = true;
            int bo = add_buf(*c, 64, true);
            add_conv(*c, bi, {&m.conv(an)}, {{bo, 0, 1}});
            if (choose_tile(h, *c, mb, mf, forced_th(std::string("aggr") + lvn[l]))) {
                taggr = B.tensor(an + "_relu", fh[l], fw[l], 64);
                c->bufs[bo].store_tensor = taggr;
                add_chain_step(B, c, 0, ((double)fh[l] * fw[l] * 64 * 2 + (double)(fh[l] / 2) * (fw[l] / 2) * 64) * es);
                done = true;
            }
        }
        if (!done) {
            taggr = B.tensor(an + "_relu", fh[l], fw[l], 64);
            const long tiles = ((long)mb * (fh[l] + 1) * (fw[l] + 2) + 127) / 128;
            if (tiles <= 148) {
                plan_conv_legacy(B, std::string(lvn[l]) + "_upsample+add+aggr_3x3_64to64", {&m.conv(an)}, lat[l], fh[l], fw[l], taggr, 64, 0, 64, 1, -1, 0, 0, 0, 0,
                                 feat_in[l - 1], l - 1);
            } else {
                int plus = plan_fpn_merge_h2(B, l == 1 ? "_plus0" : "_plus1", lat[l], feat_in[l - 1], fh[l], fw[l], l - 1);
                plan_conv_legacy(B, std::string(lvn[l]) + "_aggr_3x3_64to64", {&m.conv(an)}, plus, fh[l], fw[l], taggr, 64, 0, 64, 1, -1, 0, 0, 0);
            }
        }
        feat_in[l] = taggr;
    };
    auto ssh_level = [&](int l, int lane) {
        const std::string p = std::string("rf_") + lvn[l] + "_det";
        const int tin = feat_in[l];
        int cat = B.tensor(p + "_concat_relu", fh[l], fw[l], 64);
        h->feat_tensor[l] = cat;
        bool done = false;
        if (mask & TM_SSH) {
            auto c = std::make_shared<TileChain>();
            c->name = std::string("tile_ssh_") + lvn[l] + ((mask & TM_HEAD) ? "+heads+decode" : "");
            c->in_tensor = tin; c->in_C = 64; c->in_W = fw[l]; c->in_H = fh[l];
            c->W = fw[l]; c->H = fh[l];
            int bi = add_buf(*c, 64);
            int bcat = add_buf(*c, 64, true, cat);
            int bctx1 = add_buf(*c, 16);
            int bctx31 = add_buf(*c, 16);
            // branches that share an input run as ONE stage over the rows the neediest branch wants: an MMA's time is its A-operand
            // read from shared memory (128 rows x 32 bytes whatever N), so the other branch's output columns ride along for free
            add_conv(*c, bi, {&m.conv(p + "_conv1"), &m.conv(p + "_context_conv1")}, {{bcat, 0, 1}, {bctx1, 0, 1}});
            add_conv(*c, bctx1, {&m.conv(p + "_context_conv2"), &m.conv(p + "_context_conv3_1")}, {{bcat, 32, 1}, {bctx31, 0, 1}});
            add_conv(*c, bctx31, {&m.conv(p + "_context_conv3_2")}, {{bcat, 48, 1}});
            if (mask & TM_HEAD) {
                const int strides[3] = {32, 16, 8};
                const std::string sn = "_stride" + std::to_string(strides[l]);
                const FoldedConv *cs[3] = {&m.conv("face_rpn_cls_score" + sn), &m.conv("face_rpn_bbox_pred" + sn), &m.conv("face_rpn_landmark_pred" + sn)};
                add_head(*c, bcat, cs);
                c->level = l;
                c->fused_nms = (mask & TM_NMS) != 0;
            }
            if (choose_tile(h, *c, mb, mf, forced_th(std::string("ssh") + lvn[l]))) {
                add_chain_step(B, c, lane, ((double)fh[l] * fw[l] * 64 * 2) * es);
                ssh_chain[l] = c;
                done = true;
            }
        }
        if (!done) {
            int ctx1 = B.tensor(p + "_context_conv1_relu", fh[l], fw[l], 16);
            int ctx31 = B.tensor(p + "_context_conv3_1_relu", fh[l], fw[l], 16);
            plan_conv_legacy(B, std::string("ssh_") + lvn[l] + "_conv1+ctx1_3x3_64to48", {&m.conv(p + "_conv1"), &m.conv(p + "_context_conv1")}, tin, fh[l], fw[l], cat, 64, 0, 32,
                             1, ctx1, 16, 0, 1, lane);
            plan_conv_legacy(B, std::string("ssh_") + lvn[l] + "_ctx2+ctx3_1_3x3_16to32", {&m.conv(p + "_context_conv2"), &m.conv(p + "_context_conv3_1")}, ctx1, fh[l], fw[l], cat,
                             64, 32, 16, 1, ctx31, 16, 0, 1, lane);
            plan_conv_legacy(B, std::string("ssh_") + lvn[l] + "_ctx3_2_3x3_16to16", {&m.conv(p + "_context_conv3_2")}, ctx31, fh[l], fw[l], cat, 64, 48, 16, 1, -1, 0, 0, 0, lane);
        }
    };
    ssh_level(0, 1);
    aggr_level(1);
    ssh_level(1, 2);
    aggr_level(2);
    ssh_level(2, 0);
    const bool all_heads = ssh_chain[0] && ssh_chain[1] && ssh_chain[2] && (mask & TM_HEAD);
    if (!all_heads) {
        // some level's predictors are not fused: none may be (one decode kernel covers all levels)
        for (auto &c : ssh_chain)
            if (c && c->level >= 0) return false;
        plan_heads_and_nms<__half>(B, true, true);
    } else {
        h->head_step = (int)h->steps.size() - 1;          // the stride-8 SSH chain (last step) emits the last candidates
        h->tile_expected = ssh_chain[0]->args.tiles_per_img + ssh_chain[1]->args.tiles_per_img + ssh_chain[2]->args.tiles_per_img;
        if (!(mask & TM_NMS)) plan_heads_and_nms<__half>(B, false, true);
    }
    h->tile_mask = mask;
    return true;
}

void build_plan_tiles(rf_handle h) {
    // the chain plan pays off for one forward at a time AND small batches (profiles/r02_mask_sweep.txt: one context, batch 1:
    // 111.5 us against 116.7 us per-layer; batch 8: 162.8 against 163.6 us; batch 32: 420 against 387 us -- there every kernel
    // fills the GPU)
    const bool latency_mode = h->cfg.streams == 1 && h->cfg.max_batch <= 16;
    const unsigned dflt = !latency_mode ? TM_THROUGHPUT : (h->cfg.max_batch <= 2 ? TM_LATENCY_SMALL : TM_LATENCY);
    const unsigned mask = (unsigned)env_int("RF_TILE_MASK", (int)dflt);
    for (unsigned m : {mask, mask & ~(unsigned)(TM_HEAD | TM_NMS)}) {
        // a failed attempt leaves no trace
        h->steps.clear(); h->tensors.clear(); h->tensor_by_name.clear(); h->chains.clear();
        h->wstage.clear(); h->wstage_h.clear(); h->wstage_q.clear();
        h->head_step = h->nms_step = -1; h->tile_expected = 0;
        for (int &f : h->feat_tensor) f = -1;
        if (build_tiles_with_mask(h, m)) return;
    }
    throw PlanFail{RF_ERR_UNSUPPORTED, "no tile-chain plan fits this network size; create the handle with RF_FLAG_LEGACY_TC"};
}

}  // namespace rf_eng
