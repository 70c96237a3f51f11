/*
 * oracle/postproc.c -- plain-C restatement of the reference's host post-process.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for the CUDA
 * decode/threshold/NMS kernels and part of the timed CPU baseline.  Never linked into
 * librf_b200.so.
 *
 * Restates, function by function, /root/reference/retinaface/RetinaFace.cpp:
 *   rfo_whctrs / rfo_mkanchors ......... _whctrs, _mkanchors            (:9-33)
 *   rfo_base_anchors ................... _ratio_enum, _scale_enum,
 *                                        generate_anchors(_fpn)          (:35-125)
 *   anchor at (k, ih, iw) .............. anchors_plane                   (:127-154)
 *   rfo_bbox_pred ...................... RetinaFace::bbox_pred           (:378-398)
 *   rfo_clip ........................... clip_boxes (single)             (:179-199)
 *   rfo_landmark_pred .................. RetinaFace::landmark_pred       (:418-432)
 *   rfo_decode_level ................... the per-stride loop of detect   (:666-723)
 *   rfo_nms ............................ RetinaFace::nms + CompareBBox   (:434-492)
 * with the "net3" anchor configuration of the constructor (:245-268).
 *
 * Arithmetic types follow the reference expression by expression: storage is float,
 * sub-expressions containing the literals 0.5 / 1.0 evaluate in double and round to float
 * on assignment, exp() on a float argument is std::exp(float) == expf.
 *
 * The ONE deliberate deviation: the reference sorts with std::sort (unstable; order of
 * equal scores unspecified).  Here ties are broken by emission order (stride 32->16->8,
 * anchor 0->1, row-major j) so both the oracle and the CUDA path are deterministic.
 *
 * Pinned against the reference's own compiled code (oracle/_ref, built by
 * oracle/build_ref.sh) in tests/test_oracle_postproc.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x1, y1, x2, y2; } rfo_box;
typedef struct { float x_ctr, y_ctr, w, h; } rfo_win;

/* Same 15-float record as FaceDetectInfo (RetinaFace.h:37-42). */
typedef struct {
    float score;
    rfo_box rect;
    float px[5];
    float py[5];
} rfo_face;

static rfo_win rfo_whctrs(rfo_box a) {           /* :9-20 */
    rfo_win win;
    win.w = a.x2 - a.x1 + 1;
    win.h = a.y2 - a.y1 + 1;
    win.x_ctr = (float)(a.x1 + 0.5 * (win.w - 1));
    win.y_ctr = (float)(a.y1 + 0.5 * (win.h - 1));
    return win;
}

static rfo_box rfo_mkanchors(rfo_win win) {       /* :22-33 */
    rfo_box a;
    a.x1 = (float)(win.x_ctr - 0.5 * (win.w - 1));
    a.y1 = (float)(win.y_ctr - 0.5 * (win.h - 1));
    a.x2 = (float)(win.x_ctr + 0.5 * (win.w - 1));
    a.y2 = (float)(win.y_ctr + 0.5 * (win.h - 1));
    return a;
}

/* Base anchors of one FPN level: base_size 16, ratios {1.0}, the level's two scales.
 * generate_anchors (:70-104) = _ratio_enum (:35-52) then _scale_enum (:54-68).
 * net3 config (:245-268): stride 32 -> {32,16}; 16 -> {8,4}; 8 -> {2,1}. */
int rfo_base_anchors(int stride, float out[8]) {
    int scales[2];
    if (stride == 32) { scales[0] = 32; scales[1] = 16; }
    else if (stride == 16) { scales[0] = 8; scales[1] = 4; }
    else if (stride == 8) { scales[0] = 2; scales[1] = 1; }
    else return -1;
    const int base_size = 16;
    const float ratio = 1.0f;
    rfo_box base = { 0.f, 0.f, (float)(base_size - 1), (float)(base_size - 1) };
    /* _ratio_enum with the single ratio 1.0 */
    rfo_win win = rfo_whctrs(base);
    float size = win.w * win.h;
    float scale = size / ratio;
    win.w = roundf((float)sqrt(scale));
    win.h = roundf(win.w * ratio);
    rfo_box ratio_anchor = rfo_mkanchors(win);
    /* _scale_enum */
    for (int i = 0; i < 2; i++) {
        rfo_win w2 = rfo_whctrs(ratio_anchor);
        w2.w = w2.w * scales[i];
        w2.h = w2.h * scales[i];
        rfo_box a = rfo_mkanchors(w2);
        out[4 * i + 0] = a.x1; out[4 * i + 1] = a.y1; out[4 * i + 2] = a.x2; out[4 * i + 3] = a.y2;
    }
    return 2;
}

static rfo_box rfo_bbox_pred(rfo_box anchor, const float regress[4]) {   /* :378-398 */
    rfo_box rect;
    float width = anchor.x2 - anchor.x1 + 1;
    float height = anchor.y2 - anchor.y1 + 1;
    float ctr_x = (float)(anchor.x1 + 0.5 * (width - 1.0));
    float ctr_y = (float)(anchor.y1 + 0.5 * (height - 1.0));
    float pred_ctr_x = regress[0] * width + ctr_x;
    float pred_ctr_y = regress[1] * height + ctr_y;
    float pred_w = expf(regress[2]) * width;
    float pred_h = expf(regress[3]) * height;
    rect.x1 = (float)(pred_ctr_x - 0.5 * (pred_w - 1.0));
    rect.y1 = (float)(pred_ctr_y - 0.5 * (pred_h - 1.0));
    rect.x2 = (float)(pred_ctr_x + 0.5 * (pred_w - 1.0));
    rect.y2 = (float)(pred_ctr_y + 0.5 * (pred_h - 1.0));
    return rect;
}

static void rfo_clip(rfo_box *b, int width, int height) {                /* :179-199 */
    if (b->x1 < 0) b->x1 = 0;
    if (b->y1 < 0) b->y1 = 0;
    if (b->x2 > width - 1) b->x2 = (float)(width - 1);
    if (b->y2 > height - 1) b->y2 = (float)(height - 1);
}

static void rfo_landmark_pred(rfo_box anchor, const float dx[5], const float dy[5],
                              float ox[5], float oy[5]) {                /* :418-432 */
    float width = anchor.x2 - anchor.x1 + 1;
    float height = anchor.y2 - anchor.y1 + 1;
    float ctr_x = (float)(anchor.x1 + 0.5 * (width - 1.0));
    float ctr_y = (float)(anchor.y1 + 0.5 * (height - 1.0));
    for (int j = 0; j < 5; j++) {
        ox[j] = dx[j] * width + ctr_x;
        oy[j] = dy[j] * height + ctr_y;
    }
}

/* One FPN level of detect() (:666-723).
 * cls_prob: 4 x h x w (bg0,bg1,face0,face1); bbox: 8 x h x w; lmk: 20 x h x w (NCHW, one image).
 * Appends to out[*n..cap) and emit_idx (global emission index = idx_base + num*h*w + j).
 * Returns number of candidates that did not fit (0 normally). */
int rfo_decode_level(const float *cls_prob, const float *bbox, const float *lmk,
                     int h, int w, int stride, int net_w, int net_h, float thr,
                     rfo_face *out, int32_t *emit_idx, int cap, int *n, int idx_base) {
    float base[8];
    if (rfo_base_anchors(stride, base) != 2) return -1;
    const size_t count = (size_t)h * w;
    const float *score = cls_prob + 2 * count;   /* second half (:671-674) */
    int dropped = 0;
    for (size_t num = 0; num < 2; num++) {
        for (size_t j = 0; j < count; j++) {
            float conf = score[j + count * num];
            if (conf <= thr) continue;                                    /* :693 */
            float regress[4];
            for (int c = 0; c < 4; c++) regress[c] = bbox[j + count * (c + num * 4)];
            /* anchors_plane (:127-154): index k*H*W + ih*W + iw */
            int ih = (int)(j / w), iw = (int)(j % w);
            rfo_box anchor;
            anchor.x1 = base[4 * num + 0] + iw * stride;
            anchor.y1 = base[4 * num + 1] + ih * stride;
            anchor.x2 = base[4 * num + 2] + iw * stride;
            anchor.y2 = base[4 * num + 3] + ih * stride;
            rfo_box rect = rfo_bbox_pred(anchor, regress);
            rfo_clip(&rect, net_w, net_h);
            float dx[5], dy[5];
            for (size_t k = 0; k < 5; k++) {
                dx[k] = lmk[j + count * (num * 10 + k * 2)];
                dy[k] = lmk[j + count * (num * 10 + k * 2 + 1)];
            }
            if (*n >= cap) { dropped++; continue; }
            rfo_face *f = &out[*n];
            f->score = conf;
            f->rect = rect;
            rfo_landmark_pred(anchor, dx, dy, f->px, f->py);
            if (emit_idx) emit_idx[*n] = idx_base + (int32_t)(num * count + j);
            (*n)++;
        }
    }
    return dropped;
}

/* Stable descending sort by score == sort by (score desc, position asc). */
typedef struct { float score; int32_t pos; } rfo_key;
static int rfo_key_cmp(const void *a, const void *b) {
    const rfo_key *x = (const rfo_key *)a, *y = (const rfo_key *)b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return (x->pos > y->pos) - (x->pos < y->pos);
}

/* RetinaFace::nms (:439-492).  in[0..n) in emission order; writes kept faces in score order to
 * out and their positions in `in` to keep_pos (may be NULL).  Returns number kept. */
int rfo_nms(const rfo_face *in, int n, float threshold, rfo_face *out, int32_t *keep_pos) {
    if (n <= 0) return 0;
    rfo_key *keys = (rfo_key *)malloc(sizeof(rfo_key) * (size_t)n);
    uint8_t *merged = (uint8_t *)calloc((size_t)n, 1);
    for (int i = 0; i < n; i++) { keys[i].score = in[i].score; keys[i].pos = i; }
    qsort(keys, (size_t)n, sizeof(rfo_key), rfo_key_cmp);
    int kept = 0;
    int select_idx = 0;
    for (;;) {
        while (select_idx < n && merged[select_idx]) select_idx++;
        if (select_idx == n) break;
        const rfo_face *sel = &in[keys[select_idx].pos];
        out[kept] = *sel;
        if (keep_pos) keep_pos[kept] = keys[select_idx].pos;
        kept++;
        merged[select_idx] = 1;
        rfo_box sb = sel->rect;
        float area1 = (sb.x2 - sb.x1 + 1) * (sb.y2 - sb.y1 + 1);
        float x1 = sb.x1, y1 = sb.y1, x2 = sb.x2, y2 = sb.y2;
        select_idx++;
        for (int i = select_idx; i < n; i++) {
            if (merged[i]) continue;
            const rfo_box *bi = &in[keys[i].pos].rect;
            float x = x1 > bi->x1 ? x1 : bi->x1;          /* std::max<float> */
            float y = y1 > bi->y1 ? y1 : bi->y1;
            float w = (x2 < bi->x2 ? x2 : bi->x2) - x + 1;
            float h = (y2 < bi->y2 ? y2 : bi->y2) - y + 1;
            if (w <= 0 || h <= 0) continue;
            float area2 = (bi->x2 - bi->x1 + 1) * (bi->y2 - bi->y1 + 1);
            float area_intersect = w * h;
            if (area_intersect / (area1 + area2 - area_intersect) > threshold) merged[i] = 1;
        }
    }
    free(keys);
    free(merged);
    return kept;
}

/* Whole post-process of one image: heads[9] in the order of
 * tensorrt/trtretinafacenet.cpp:23-31 (cls_prob, bbox, landmark for stride 32, 16, 8).
 * cand/cand_idx (cap entries) receive the pre-NMS candidates in emission order,
 * out/out_idx the kept faces in score order.  Returns kept count; *n_cand = candidates. */
int rfo_postprocess(const float *const heads[9], int net_h, int net_w, float thr, float nms_thr,
                    rfo_face *cand, int32_t *cand_idx, int cap, int *n_cand,
                    rfo_face *out, int32_t *out_idx) {
    static const int strides[3] = { 32, 16, 8 };
    int n = 0, base = 0;
    for (int l = 0; l < 3; l++) {
        int s = strides[l];
        int h = net_h / s, w = net_w / s;
        rfo_decode_level(heads[3 * l], heads[3 * l + 1], heads[3 * l + 2], h, w, s, net_w, net_h, thr,
                         cand, cand_idx, cap, &n, base);
        base += 2 * h * w;
    }
    *n_cand = n;
    int32_t *pos = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int kept = rfo_nms(cand, n, nms_thr, out, pos);
    if (out_idx) for (int i = 0; i < kept; i++) out_idx[i] = cand_idx ? cand_idx[pos[i]] : pos[i];
    free(pos);
    return kept;
}
