// Per-element arithmetic of the D-FINE criterion (reference: src/d_fine/dfine_criterion.py) shared by the CUDA kernels
// in loss.cu and — compiled for the host by tests/host_harness/loss_host.cpp — by the CPU test that pins this
// arithmetic against the torch restatement (tests/test_loss_math_cpu.py).  Plain C++: no intrinsics, no shuffles.
//
//   VFL      loss_labels_vfl            dfine_criterion.py:92-122
//   boxes    loss_boxes (L1 + GIoU)     dfine_criterion.py:124-143, arch/utils.py:12-73
//   FGL      loss_local, first half     dfine_criterion.py:145-190, bbox2distance / translate_gt arch/utils.py:267-354
//   DDF      loss_local, second half    dfine_criterion.py:192-237
#pragma once
#include <cmath>
#include "../../include/dfine_loss_desc.h"

#if defined(__CUDACC__)
#define LM_HD __host__ __device__ __forceinline__
#else
#define LM_HD inline
#endif

namespace lossmath {

constexpr int NB_MAX = 64;   // reg_max + 1 <= 64

LM_HD float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// cxcywh -> xyxy with w, h clamped at 0 (arch/utils.py:59-67)
LM_HD void to_xyxy(const float* b, float* o) {
    const float hw = 0.5f * fmaxf(b[2], 0.f), hh = 0.5f * fmaxf(b[3], 0.f);
    o[0] = b[0] - hw; o[1] = b[1] - hh; o[2] = b[0] + hw; o[3] = b[1] + hh;
}

struct IouParts {
    float area_a, area_b, wh0, wh1, inter, uni, iou;
    float cw0, cw1, area_c, giou;
};

// IoU / GIoU of two xyxy boxes (arch/utils.py:12-51: box_iou + generalized_box_iou, the matched diagonal)
LM_HD IouParts iou_giou(const float* a, const float* b) {
    IouParts p;
    p.area_a = (a[2] - a[0]) * (a[3] - a[1]);
    p.area_b = (b[2] - b[0]) * (b[3] - b[1]);
    p.wh0 = fmaxf(fminf(a[2], b[2]) - fmaxf(a[0], b[0]), 0.f);
    p.wh1 = fmaxf(fminf(a[3], b[3]) - fmaxf(a[1], b[1]), 0.f);
    p.inter = p.wh0 * p.wh1;
    p.uni = p.area_a + p.area_b - p.inter;
    p.iou = p.inter / p.uni;
    p.cw0 = fmaxf(fmaxf(a[2], b[2]) - fminf(a[0], b[0]), 0.f);
    p.cw1 = fmaxf(fmaxf(a[3], b[3]) - fminf(a[1], b[1]), 0.f);
    p.area_c = p.cw0 * p.cw1;
    p.giou = p.iou - (p.area_c - p.uni) / p.area_c;
    return p;
}

// share of an upstream gradient that torch.minimum / torch.maximum route to `self` (ties split in half)
LM_HD float min_share(float self, float other) { return self < other ? 1.f : (self == other ? 0.5f : 0.f); }
LM_HD float max_share(float self, float other) { return self > other ? 1.f : (self == other ? 0.5f : 0.f); }

// d(giou)/d(src cxcywh) for src box s (cxcywh) against target xyxy t; returns through g[4] (to be scaled by the caller)
LM_HD void giou_grad_cxcywh(const float* s, const float* a, const float* t, const IouParts& p, float* g) {
    // giou = iou - 1 + uni / area_c
    const float d_iou = 1.f;
    float d_uni = 1.f / p.area_c;
    const float d_areac = -p.uni / (p.area_c * p.area_c);
    // iou = inter / uni
    float d_inter = d_iou / p.uni;
    d_uni += -d_iou * p.inter / (p.uni * p.uni);
    // uni = area_a + area_b - inter
    const float d_area_a = d_uni;
    d_inter += -d_uni;
    // inter = wh0 * wh1, wh_k = clamp(min(a_br, t_br) - max(a_tl, t_tl), 0)
    float da[4] = {0.f, 0.f, 0.f, 0.f};
    const float d_wh0 = d_inter * p.wh1, d_wh1 = d_inter * p.wh0;
    const float raw0 = fminf(a[2], t[2]) - fmaxf(a[0], t[0]), raw1 = fminf(a[3], t[3]) - fmaxf(a[1], t[1]);
    if (raw0 >= 0.f) { da[2] += d_wh0 * min_share(a[2], t[2]); da[0] -= d_wh0 * max_share(a[0], t[0]); }
    if (raw1 >= 0.f) { da[3] += d_wh1 * min_share(a[3], t[3]); da[1] -= d_wh1 * max_share(a[1], t[1]); }
    // area_c = cw0 * cw1, cw_k = clamp(max(a_br, t_br) - min(a_tl, t_tl), 0)
    const float d_cw0 = d_areac * p.cw1, d_cw1 = d_areac * p.cw0;
    const float rc0 = fmaxf(a[2], t[2]) - fminf(a[0], t[0]), rc1 = fmaxf(a[3], t[3]) - fminf(a[1], t[1]);
    if (rc0 >= 0.f) { da[2] += d_cw0 * max_share(a[2], t[2]); da[0] -= d_cw0 * min_share(a[0], t[0]); }
    if (rc1 >= 0.f) { da[3] += d_cw1 * max_share(a[3], t[3]); da[1] -= d_cw1 * min_share(a[1], t[1]); }
    // area_a = (a2 - a0) * (a3 - a1)
    const float w = a[2] - a[0], h = a[3] - a[1];
    da[2] += d_area_a * h; da[0] -= d_area_a * h;
    da[3] += d_area_a * w; da[1] -= d_area_a * w;
    // a = (cx - hw, cy - hh, cx + hw, cy + hh), hw = 0.5 * clamp(w, 0)
    g[0] = da[0] + da[2];
    g[1] = da[1] + da[3];
    g[2] = s[2] >= 0.f ? 0.5f * (da[2] - da[0]) : 0.f;
    g[3] = s[3] >= 0.f ? 0.5f * (da[3] - da[1]) : 0.f;
}

// One VFL element (dfine_criterion.py:103-121): target class gets target_score = IoU and weight = IoU, every other
// class weight alpha * sigmoid(x)^gamma (detached) and target 0; loss = weight * BCE-with-logits(x, target_score).
// Returns the loss; *dx = d loss / d x (weight and target detached).
LM_HD float vfl_elem(float x, bool is_target, float iou, float alpha, float gamma, float* dx) {
    const float p = sigmoidf_(x);
    const float ts = is_target ? iou : 0.f;
    const float pg = gamma == 2.f ? p * p : powf(p, gamma);
    const float w = is_target ? ts : alpha * pg;
    // (1 - t) x + max(-x, 0) + log(exp(-max(-x,0)) + exp(-x - max(-x,0)))   [ATen binary_cross_entropy_with_logits]
    const float mv = fmaxf(-x, 0.f);
    const float bce = (1.f - ts) * x + mv + logf(expf(-mv) + expf(-x - mv));
    *dx = w * (p - ts);
    return w * bce;
}

// FGL bin targets of one edge (bbox2distance + translate_gt, arch/utils.py:267-354).  ref = reference box cxcywh, gt =
// target xyxy, edge 0..3 = left, top, right, bottom; wn = W(n) [nb = reg_max + 1].  Outputs the left bin and the two
// interpolation weights.
LM_HD void fgl_target(const float* ref, const float* gt, int edge, const float* wn, int nb, float rs_abs, int* bin,
                      float* w_l, float* w_r) {
    const int reg_max = nb - 1;
    const float sw = (edge & 1) ? ref[3] / rs_abs + 1e-16f : ref[2] / rs_abs + 1e-16f;
    const float c = (edge & 1) ? ref[1] : ref[0];
    const float num = edge < 2 ? c - gt[edge] : gt[edge] - c;
    const float d = num / sw - 0.5f * rs_abs;
    int cnt = 0;
    for (int j = 0; j < nb; ++j) cnt += (wn[j] - d) <= 0.f ? 1 : 0;
    const int left = cnt - 1;                       // in [-1, reg_max]
    const bool valid = left >= 0 && left < reg_max;
    const int li = left < 0 ? 0 : (left > reg_max - 1 ? reg_max - 1 : left);
    const float lv = wn[li], rv = wn[li + 1];
    const float ld = fabsf(d - lv), rd = fabsf(rv - d);
    float wr = valid ? ld / (ld + rd) : 0.f;
    float wl = valid ? 1.f - wr : 0.f;
    const bool below = left < 0, above = left >= reg_max;
    wr = above ? 1.f : (below ? 0.f : wr);
    wl = above ? 0.f : (below ? 1.f : wl);
    float idx = (float)left;
    if (below) idx = 0.f;
    if (above) idx = (float)reg_max - 0.1f;
    idx = fminf(fmaxf(idx, 0.f), (float)reg_max - 0.1f);
    *bin = (int)idx;
    *w_l = wl;
    *w_r = wr;
}

// log-softmax statistics of a row of nb logits scaled by inv_t: returns max and log-sum-exp so that
// log_softmax_j = x_j * inv_t - m - lse
LM_HD void row_lse(const float* x, int nb, float inv_t, float* m_out, float* lse_out) {
    float m = -INFINITY;
    for (int j = 0; j < nb; ++j) m = fmaxf(m, x[j] * inv_t);
    float s = 0.f;
    for (int j = 0; j < nb; ++j) s += expf(x[j] * inv_t - m);
    *m_out = m;
    *lse_out = logf(s);
}

// FGL of one edge row (dfine_criterion.py:837-858 with the weights above): CE to the left bin * w_l + CE to the right
// bin * w_r.  (The IoU weight and the normaliser are applied by the caller.)
LM_HD float fgl_row(const float* x, int nb, int bin, float w_l, float w_r) {
    float m, lse;
    row_lse(x, nb, 1.f, &m, &lse);
    const float ce_l = -(x[bin] - m - lse), ce_r = -(x[bin + 1] - m - lse);
    return ce_l * w_l + ce_r * w_r;
}
// d fgl_row / d x_j, accumulated into g[j] with factor `scale`
LM_HD void fgl_row_grad(const float* x, int nb, int bin, float w_l, float w_r, float scale, float* g) {
    float m, lse;
    row_lse(x, nb, 1.f, &m, &lse);
    const float ws = w_l + w_r;
    for (int j = 0; j < nb; ++j) {
        float v = expf(x[j] - m - lse) * ws;
        if (j == bin) v -= w_l;
        if (j == bin + 1) v -= w_r;
        g[j] += scale * v;
    }
}

// KL(softmax(teacher / T) || softmax(pred / T)) of one edge row (F.kl_div(log_softmax(pred/T), softmax(teacher/T)),
// summed over the bins; dfine_criterion.py:214-219); *same = 1 if the two rows are bit-identical.
LM_HD float ddf_row(const float* pred, const float* teacher, int nb, float inv_t, int* same) {
    float mp, lp, mt, lt;
    row_lse(pred, nb, inv_t, &mp, &lp);
    row_lse(teacher, nb, inv_t, &mt, &lt);
    float kl = 0.f;
    int eq = 1;
    for (int j = 0; j < nb; ++j) {
        const float lsp = pred[j] * inv_t - mp - lp, lst = teacher[j] * inv_t - mt - lt;
        const float tq = expf(lst);
        kl += tq > 0.f ? tq * (lst - lsp) : 0.f;
        eq &= pred[j] == teacher[j] ? 1 : 0;
    }
    *same = eq;
    return kl;
}
// d ddf_row / d pred_j = (softmax(pred/T)_j - softmax(teacher/T)_j) / T, accumulated into g[j] with factor `scale`
LM_HD void ddf_row_grad(const float* pred, const float* teacher, int nb, float inv_t, float scale, float* g) {
    float mp, lp, mt, lt;
    row_lse(pred, nb, inv_t, &mp, &lp);
    row_lse(teacher, nb, inv_t, &mt, &lt);
    for (int j = 0; j < nb; ++j) {
        const float sp = expf(pred[j] * inv_t - mp - lp), st = expf(teacher[j] * inv_t - mt - lt);
        g[j] += scale * (sp - st) * inv_t;
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Launch description of one criterion evaluation (all heads of a train step) and the per-thread work items built on
// the arithmetic above.  The CUDA kernels (loss.cu) and the host harness run exactly these functions.
//
// Heads.  Group A (matching queries, rows [n_dn, n_dn + Q) of the stacked tensors): decoder layers 0..L-1, then the
// `pre` head, then the encoder head (its own [B,Q,*] tensors).  Group DN (denoising queries, rows [0, n_dn)): decoder
// layers 0..L-1, then `dn_pre`.  Index sets come from the criterion's plan table (criterion.IndexPlan): per-head
// Hungarian sets for group A's VFL, the cross-layer GO union for group A's boxes / FGL / DDF, the denoising set for
// everything in group DN.
typedef ::dfine_loss_desc LossDesc;     // include/dfine_loss_desc.h: the C-ABI struct IS the kernels' launch description

LM_HD int n_heads(const LossDesc& d, int g) { return g == 0 ? d.L + 2 : d.L + 1; }
LM_HD int map_go(const LossDesc& d) { return d.L + 2; }
LM_HD int map_dn(const LossDesc& d) { return d.L + 3; }

struct HeadView {
    const float* logits;
    const float* boxes;
    long ldb;        // query rows per image of the tensor
    int q0, nq;      // first row / number of rows of this group inside an image
};
LM_HD HeadView head_view(const LossDesc& d, int g, int h) {
    HeadView v;
    v.ldb = d.Qt; v.q0 = g == 0 ? d.n_dn : 0; v.nq = g == 0 ? d.Q : d.n_dn;
    if (h < d.L) {
        v.logits = d.logits + (long)h * d.B * d.Qt * d.C;
        v.boxes = d.boxes + (long)h * d.B * d.Qt * 4;
    } else if (h == d.L) {
        v.logits = d.pre_logits; v.boxes = d.pre_boxes;
    } else {
        v.logits = d.enc_logits; v.boxes = d.enc_boxes; v.ldb = d.Q; v.q0 = 0;
    }
    return v;
}
LM_HD float norm_vfl(const LossDesc& d, int g) { return g == 0 ? d.counts[1] : d.counts[1] * d.dn_groups; }
LM_HD float norm_box(const LossDesc& d, int g) { return g == 0 ? d.counts[0] : d.counts[1] * d.dn_groups; }

// plan-table column -> (map id, image, query, target); returns false for padded / invalid columns
LM_HD bool table_entry(const LossDesc& d, long j, int* map, int* b, int* q, int* t) {
    if (d.table[3 * d.ncols + j] == 0) return false;
    *b = (int)d.table[j]; *q = (int)d.table[d.ncols + j]; *t = (int)d.table[2 * d.ncols + j];
    const long n_sets = d.L + 2, a_end = n_sets * d.n_layer;
    if (j < a_end) {
        const int s = (int)(j / d.n_layer);               // plan order: main (= last layer), aux_0.., pre, enc
        *map = s == 0 ? d.L - 1 : (s <= d.L - 1 ? s - 1 : s);
    } else if (j < a_end + d.go_cap) {
        *map = map_go(d);
    } else {
        *map = map_dn(d);
    }
    return *q < d.Qm;
}

// VFL: one (group, head, image, query) row; returns the row's loss sum and (if dlogits) writes d loss / d logits * scale
LM_HD float vfl_row(const LossDesc& d, int g, int h, int b, int q, float* drow, float scale) {
    const HeadView v = head_view(d, g, h);
    const int map = g == 0 ? h : map_dn(d);
    const int t = d.maps[((long)map * d.B + b) * d.Qm + q];
    const long row = (long)b * v.ldb + v.q0 + q;
    float iou = 0.f;
    long label = -1;
    if (t >= 0) {
        float a[4], tb[4];
        to_xyxy(v.boxes + row * 4, a);
        to_xyxy(d.tboxes + (long)t * 4, tb);
        iou = iou_giou(a, tb).iou;
        label = d.labels[t];
    }
    const float* x = v.logits + row * d.C;
    float s = 0.f;
    for (int c = 0; c < d.C; ++c) {
        float dx;
        s += vfl_elem(x[c], c == label, iou, d.alpha, d.gamma, &dx);
        if (drow) drow[c] = dx * scale;
    }
    return s;
}

// boxes: one (group, head, set entry); outputs l1 and (1 - giou); if dbox, writes the gradient row (scaled)
LM_HD bool box_entry(const LossDesc& d, int g, int h, long e, float* l1, float* gl, float* dbox, float s_l1, float s_gi,
                     long* row_out) {
    const long n_sets = d.L + 2;
    const long j = g == 0 ? n_sets * d.n_layer + e : n_sets * d.n_layer + d.go_cap + e;
    if (d.table[3 * d.ncols + j] == 0) return false;
    const int b = (int)d.table[j], q = (int)d.table[d.ncols + j], t = (int)d.table[2 * d.ncols + j];
    const HeadView v = head_view(d, g, h);
    const long row = (long)b * v.ldb + v.q0 + q;
    const float* s = v.boxes + row * 4;
    const float* tb = d.tboxes + (long)t * 4;
    float a[4], tx[4];
    to_xyxy(s, a);
    to_xyxy(tb, tx);
    const IouParts p = iou_giou(a, tx);
    *l1 = fabsf(s[0] - tb[0]) + fabsf(s[1] - tb[1]) + fabsf(s[2] - tb[2]) + fabsf(s[3] - tb[3]);
    *gl = 1.f - p.giou;
    if (dbox) {
        float gg[4];
        giou_grad_cxcywh(s, a, tx, p, gg);
        for (int k = 0; k < 4; ++k) {
            const float df = s[k] - tb[k];
            const float sg = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
            dbox[k] = s_l1 * sg - s_gi * gg[k];
        }
    }
    *row_out = row;
    return true;
}

// local losses: one (group, layer, image, query, edge).  Outputs: fgl contribution (already x IoU), ddf `per` value
// (w * T^2 * kl), matched flag, same flag.  Gradient mode: adds into grow[NB] with the given coefficients.
struct LocalOut { float fgl, per; int matched, same; };
LM_HD float teacher_wmax(const LossDesc& d, int b, int qrow) {
    const float* x = d.logits + (((long)(d.L - 1) * d.B + b) * d.Qt + qrow) * d.C;
    float m = x[0];
    for (int c = 1; c < d.C; ++c) m = fmaxf(m, x[c]);
    return sigmoidf_(m);
}
LM_HD LocalOut local_item(const LossDesc& d, int g, int l, int b, int q, int edge, float* grow, float c_fgl, float c_pos,
                          float c_neg) {
    LocalOut o;
    o.fgl = 0.f; o.per = 0.f; o.same = 1;
    const int q0 = g == 0 ? d.n_dn : 0;
    const int qrow = q0 + q;
    const int t = d.maps[((long)(g == 0 ? map_go(d) : map_dn(d)) * d.B + b) * d.Qm + q];
    o.matched = t >= 0;
    const float* pred = d.corners + ((((long)l * d.B + b) * d.Qt + qrow) * 4 + edge) * d.NB;
    float iou = 0.f;
    if (t >= 0) {
        float a[4], tx[4];
        to_xyxy(d.boxes + (((long)l * d.B + b) * d.Qt + qrow) * 4, a);
        to_xyxy(d.tboxes + (long)t * 4, tx);
        iou = iou_giou(a, tx).iou;
        int bin;
        float wl, wr;
        fgl_target(d.ref0 + ((long)b * d.Qt + qrow) * 4, tx, edge, d.project, d.NB, fabsf(d.reg_scale[0]), &bin, &wl, &wr);
        if (grow) fgl_row_grad(pred, d.NB, bin, wl, wr, c_fgl * iou, grow);
        else o.fgl = fgl_row(pred, d.NB, bin, wl, wr) * iou;
    }
    if (l < d.L - 1) {
        const float* teacher = d.corners + ((((long)(d.L - 1) * d.B + b) * d.Qt + qrow) * 4 + edge) * d.NB;
        const float w = t >= 0 ? iou : teacher_wmax(d, b, qrow);
        const float t2 = 1.f / (d.inv_t * d.inv_t);
        if (grow) {
            const float c = t >= 0 ? c_pos : c_neg;
            if (c != 0.f) ddf_row_grad(pred, teacher, d.NB, d.inv_t, c * w * t2, grow);
        } else {
            o.per = w * t2 * ddf_row(pred, teacher, d.NB, d.inv_t, &o.same);
        }
    }
    return o;
}

// Final scalars from the accumulators.  acc layout (double): vfl[2][L+2], l1[2][L+2], gi[2][L+2], fgl[2][L],
// pos[2][L], neg[2][L]; notsame int[2][L].  out layout (float): vfl[2][L+2], l1[2][L+2], gi[2][L+2], fgl[2][L],
// ddf[2][L], then the backward coefficients c_pos[2][L], c_neg[2][L].
LM_HD int acc_off(const LossDesc& d, int which) {
    const int H = d.L + 2;
    return which < 3 ? which * 2 * H : 6 * H + (which - 3) * 2 * d.L;
}
LM_HD int out_count(const LossDesc& d) { return 6 * (d.L + 2) + 8 * d.L; }
LM_HD float nz(float v) { return (v == v && v - v == 0.f) ? v : 0.f; }     // nan / inf -> 0 (torch.nan_to_num(nan=0) + safety)
LM_HD void finalize(const LossDesc& d, const double* acc, const int* notsame, float* out) {
    const int H = d.L + 2, L = d.L;
    float num_pos = 0.f, num_neg = 0.f;
    for (int g = 0; g < 2; ++g) {
        const int nq = g == 0 ? d.Q : d.n_dn;
        const float nv = norm_vfl(d, g), nbx = norm_box(d, g);
        for (int h = 0; h < H; ++h) {
            const bool live = h < n_heads(d, g) && nq > 0;
            out[acc_off(d, 0) + g * H + h] = live ? nz((float)acc[acc_off(d, 0) + g * H + h] / nv) : 0.f;
            out[acc_off(d, 1) + g * H + h] = live ? nz((float)acc[acc_off(d, 1) + g * H + h] / nbx) : 0.f;
            out[acc_off(d, 2) + g * H + h] = live ? nz((float)acc[acc_off(d, 2) + g * H + h] / nbx) : 0.f;
        }
        const float n_pos = 4.f * (float)d.cnt[g], n_neg = 4.f * ((float)d.B * (float)nq - (float)d.cnt[g]);
        if (g == 0) {        // cached from the matching-query pass and reused by the denoising pass (dfine_criterion.py:223-235)
            const float scale = 8.f / (float)d.B;
            num_pos = sqrtf(n_pos * scale);
            num_neg = sqrtf(n_neg * scale);
        }
        for (int l = 0; l < L; ++l) {
            out[acc_off(d, 3) + g * L + l] = nq > 0 ? nz((float)acc[acc_off(d, 3) + g * L + l] / nbx) : 0.f;
            float ddf = 0.f, cp = 0.f, cn = 0.f;
            if (l < L - 1 && nq > 0 && notsame[g * L + l] != 0) {
                const float l_pos = n_pos > 0.f ? (float)acc[acc_off(d, 4) + g * L + l] / n_pos : 0.f;
                const float l_neg = n_neg > 0.f ? (float)acc[acc_off(d, 5) + g * L + l] / n_neg : 0.f;
                const float den = num_pos + num_neg;
                ddf = (l_pos * num_pos + l_neg * num_neg) / den;
                if (ddf == ddf && ddf - ddf == 0.f) {
                    cp = n_pos > 0.f ? num_pos / den / n_pos : 0.f;
                    cn = n_neg > 0.f ? num_neg / den / n_neg : 0.f;
                } else {
                    ddf = 0.f;
                }
            }
            out[acc_off(d, 4) + g * L + l] = ddf;
            out[6 * H + 4 * L + g * L + l] = cp;
            out[6 * H + 6 * L + g * L + l] = cn;
        }
    }
}

}  // namespace lossmath
