// D-FINE criterion kernels: every loss term of a train step in a handful of launches.
//
// Replaces the arithmetic of DFINECriterion.forward (src/d_fine/dfine_criterion.py:609-777) — loss_labels_vfl 92-122,
// loss_boxes 124-143, loss_local (FGL + DDF) 145-237 + 837-858 with bbox2distance / translate_gt (arch/utils.py:267-354)
// — which the reference evaluates head by head (48 loss terms for D-FINE-m, each 5-30 tiny ATen kernels plus host
// syncs) and which round 1 of this repo evaluated layer-batched with ~500 eager torch ops forward and as many
// backward.  Here: prepare (index maps) + one kernel per loss family over ALL heads (main, aux_i, pre, enc_0, dn_i,
// dn_pre) + finalize forward; one kernel per family backward, writing the gradients of the stacked decoder tensors
// in place (every element of d(logits) / d(corners) is written exactly once: no zero fill, no accumulation kernels).
//
// The per-element arithmetic lives in loss_math.cuh (plain C++, also compiled for the host by the CPU test harness);
// the kernels here only map threads to work items and reduce.  All of it is HBM/latency-bound integer + fp32 work:
// ~25 MB of corner logits and ~20 MB of class logits per step at batch 16 — no tensor-core shape anywhere.
#include "common.cuh"
#include "loss_math.cuh"

namespace {
using namespace lossmath;

// workspace layout (bytes): [acc doubles | notsame ints | cnt ints | pad | maps ints]
__host__ __device__ inline int acc_count(int L) { return 6 * (L + 2) + 6 * L; }
struct Ws {
    double* acc;
    int* notsame;
    int* cnt;
    int* maps;
    long zero_bytes, maps_bytes, total;
};
inline Ws carve(void* base, int L, int B, int Q, int n_dn) {
    Ws w;
    const int Qm = Q > n_dn ? Q : n_dn;
    char* p = (char*)base;
    w.acc = (double*)p;
    long off = (long)acc_count(L) * 8;
    w.notsame = (int*)(p + off);
    off += 2L * L * 4;
    w.cnt = (int*)(p + off);
    off += 2 * 4;
    off = (off + 15) / 16 * 16;
    w.zero_bytes = off;
    w.maps = (int*)(p + off);
    w.maps_bytes = (long)(L + 4) * B * Qm * 4;
    w.total = off + w.maps_bytes;
    return w;
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
    v = warp_sum_d(v);
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x + 31) / 32; ++i) r += sh[i];
    return r;      // valid on thread 0
}

// ---------------------------------------------------------------------------------------------- maps
__global__ void loss_maps_kernel(LossDesc d) {
    pdl_entry();
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d.ncols) return;
    int map, b, q, t;
    if (!table_entry(d, j, &map, &b, &q, &t)) return;
    d.maps[((long)map * d.B + b) * d.Qm + q] = t;
    if (map == map_go(d)) atomicAdd(d.cnt, 1);
    if (map == map_dn(d)) atomicAdd(d.cnt + 1, 1);
}

// ---------------------------------------------------------------------------------------------- VFL
// One warp per (group, head, image, query) row; lanes stride the classes.  grid = (row blocks, heads of A + heads of DN).
struct VflOut { float* logits; float* pre; float* enc; const float* gout; };

template <bool BWD>
__global__ void __launch_bounds__(256) loss_vfl_kernel(LossDesc d, double* acc, VflOut o) {
    pdl_entry();
    __shared__ double sh[8];
    const int HA = d.L + 2;
    const int g = blockIdx.y < HA ? 0 : 1, h = blockIdx.y < HA ? blockIdx.y : blockIdx.y - HA;
    const HeadView v = head_view(d, g, h);
    const long rows = (long)d.B * v.nq;
    const long r = (long)blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    double part = 0.0;
    if (r < rows) {
        const int b = (int)(r / v.nq), q = (int)(r % v.nq);
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
        float* drow = nullptr;
        float scale = 0.f;
        if (BWD) {
            float* base = h < d.L ? o.logits + (long)h * d.B * d.Qt * d.C : (h == d.L ? o.pre : o.enc);
            drow = base + row * d.C;
            scale = o.gout[g * HA + h] / norm_vfl(d, g);
        }
        float s = 0.f;
        for (int c = lane; c < d.C; c += 32) {
            float dx;
            s += vfl_elem(x[c], c == label, iou, d.alpha, d.gamma, &dx);
            if (BWD) drow[c] = dx * scale;
        }
        part = (double)s;
    }
    if (!BWD) {
        const double tot = block_sum(part, sh);
        if (threadIdx.x == 0 && tot != 0.0) atomicAdd(acc + acc_off(d, 0) + g * HA + h, tot);
    }
}

// ---------------------------------------------------------------------------------------------- boxes
struct BoxOut { float* boxes; float* pre; float* enc; const float* gout_l1; const float* gout_gi; };

template <bool BWD>
__global__ void __launch_bounds__(128) loss_box_kernel(LossDesc d, double* acc, BoxOut o) {
    pdl_entry();
    __shared__ double sh[4];
    const int HA = d.L + 2;
    const int g = blockIdx.y < HA ? 0 : 1, h = blockIdx.y < HA ? blockIdx.y : blockIdx.y - HA;
    const long n = g == 0 ? d.go_cap : d.n_dn_entries;
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    float l1 = 0.f, gl = 0.f;
    if (e < n) {
        float db[4];
        long row = 0;
        float s1 = 0.f, s2 = 0.f;
        if (BWD) {
            const float nb = norm_box(d, g);
            s1 = o.gout_l1[g * HA + h] / nb;
            s2 = o.gout_gi[g * HA + h] / nb;
        }
        float a = 0.f, b = 0.f;
        if (box_entry(d, g, h, e, &a, &b, BWD ? db : nullptr, s1, s2, &row)) {
            l1 = a; gl = b;
            if (BWD) {
                float* base = h < d.L ? o.boxes + (long)h * d.B * d.Qt * 4 : (h == d.L ? o.pre : o.enc);
                *reinterpret_cast<float4*>(base + row * 4) = make_float4(db[0], db[1], db[2], db[3]);
            }
        }
    }
    if (!BWD) {
        const double t1 = block_sum((double)l1, sh);
        const double t2 = block_sum((double)gl, sh);
        if (threadIdx.x == 0) {
            if (t1 != 0.0) atomicAdd(acc + acc_off(d, 1) + g * HA + h, t1);
            if (t2 != 0.0) atomicAdd(acc + acc_off(d, 2) + g * HA + h, t2);
        }
    }
}

// ---------------------------------------------------------------------------------------------- FGL + DDF
// One thread per (group, layer, image, query, edge); grid = (item blocks, L, 2).
template <bool BWD>
__global__ void __launch_bounds__(128) loss_local_kernel(LossDesc d, double* acc, int* notsame, float* dcorners,
                                                         const float* coef, const float* gout_fgl, const float* gout_ddf) {
    pdl_entry();
    __shared__ double sh[4];
    __shared__ int sh_ns;
    const int l = blockIdx.y, g = blockIdx.z;
    const int nq = g == 0 ? d.Q : d.n_dn;
    const long items = (long)d.B * nq * 4;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x == 0) sh_ns = 0;
    double fgl = 0.0, pos = 0.0, neg = 0.0;
    int ns = 0;
    if (i < items) {
        const int edge = (int)(i % 4);
        const long bq = i / 4;
        const int b = (int)(bq / nq), q = (int)(bq % nq);
        if (BWD) {
            float grow[NB_MAX];
            for (int j = 0; j < d.NB; ++j) grow[j] = 0.f;
            const int H = d.L + 2;
            const float c_fgl = gout_fgl[g * d.L + l] / norm_box(d, g);
            const float gd = gout_ddf[g * d.L + l];
            const float c_pos = coef[6 * H + 4 * d.L + g * d.L + l] * gd, c_neg = coef[6 * H + 6 * d.L + g * d.L + l] * gd;
            local_item(d, g, l, b, q, edge, grow, c_fgl, c_pos, c_neg);
            const int qrow = (g == 0 ? d.n_dn : 0) + q;
            float* dst = dcorners + ((((long)l * d.B + b) * d.Qt + qrow) * 4 + edge) * d.NB;
            for (int j = 0; j < d.NB; ++j) dst[j] = grow[j];
        } else {
            const LocalOut o = local_item(d, g, l, b, q, edge, nullptr, 0.f, 0.f, 0.f);
            fgl = (double)o.fgl;
            if (o.matched) pos = (double)o.per; else neg = (double)o.per;
            ns = o.same ? 0 : 1;
        }
    }
    if (!BWD) {
        const double t0 = block_sum(fgl, sh);
        const double t1 = block_sum(pos, sh);
        const double t2 = block_sum(neg, sh);
        if (ns) sh_ns = 1;            // benign race: every writer stores 1
        __syncthreads();
        if (threadIdx.x == 0) {
            if (t0 != 0.0) atomicAdd(acc + acc_off(d, 3) + g * d.L + l, t0);
            if (t1 != 0.0) atomicAdd(acc + acc_off(d, 4) + g * d.L + l, t1);
            if (t2 != 0.0) atomicAdd(acc + acc_off(d, 5) + g * d.L + l, t2);
            if (sh_ns) atomicOr(notsame + g * d.L + l, 1);
        }
    }
}

__global__ void loss_finalize_kernel(LossDesc d, const double* acc, const int* notsame, float* out) {
    pdl_entry();
    if (blockIdx.x == 0 && threadIdx.x == 0) finalize(d, acc, notsame, out);
}

int check_desc(const dfine_loss_desc* d, const void* ws, const char* who) {
    DFINE_REQUIRE(d != nullptr && ws != nullptr, "%s: null descriptor / workspace", who);
    DFINE_REQUIRE(d->L >= 1 && d->L <= 8 && d->B >= 1 && d->Q >= 1 && d->n_dn >= 0 && d->Qt == d->n_dn + d->Q && d->C >= 1 &&
                      d->NB >= 2 && d->NB <= NB_MAX,
                  "%s: bad dims L=%d B=%d Q=%d n_dn=%d Qt=%d C=%d NB=%d", who, d->L, d->B, d->Q, d->n_dn, d->Qt, d->C, d->NB);
    DFINE_REQUIRE(d->ncols == (long)(d->L + 2) * d->n_layer + d->go_cap + d->n_dn_entries && d->ncols >= 0,
                  "%s: plan table has %ld columns, expected (L+2)*%d + %d + %d", who, d->ncols, d->n_layer, d->go_cap,
                  d->n_dn_entries);
    DFINE_REQUIRE(((uintptr_t)d->boxes % 16) == 0 && ((uintptr_t)d->pre_boxes % 16) == 0 && ((uintptr_t)d->enc_boxes % 16) == 0 &&
                      ((uintptr_t)ws % 16) == 0,
                  "%s: box tensors / workspace must be 16-byte aligned", who);
    return 0;
}

LossDesc with_ws(const dfine_loss_desc* d, void* ws, Ws* w) {
    LossDesc k = *d;
    *w = carve(ws, d->L, d->B, d->Q, d->n_dn);
    k.maps = w->maps;
    k.cnt = w->cnt;
    k.Qm = d->Q > d->n_dn ? d->Q : d->n_dn;
    return k;
}
int heads_total(const LossDesc& d) { return (d.L + 2) + (d.n_dn > 0 ? d.L + 1 : 0); }

}  // namespace

DFINE_API int dfine_loss_desc_size(void) { return (int)sizeof(dfine_loss_desc); }

// Bytes of the caller-provided workspace (accumulators, flags, index maps) of one criterion evaluation; it must stay
// untouched between dfine_loss_prepare and the last dfine_loss_*_bwd of the step.
DFINE_API long dfine_loss_workspace_bytes(int L, int B, int Q, int n_dn) { return carve(nullptr, L, B, Q, n_dn).total; }

// Number of floats written by dfine_loss_finalize: vfl[2][L+2], l1[2][L+2], giou[2][L+2], fgl[2][L], ddf[2][L]
// (group A = matching, group DN = denoising; head order: layers 0..L-1, pre, enc), then 4L backward coefficients.
DFINE_API int dfine_loss_out_count(int L) { return 6 * (L + 2) + 8 * L; }

// Step 1: zero the accumulators and build the (image, query) -> target maps from the plan table.
DFINE_API int dfine_loss_prepare(const dfine_loss_desc* desc, void* workspace, void* stream) {
    if (int rc = check_desc(desc, workspace, "loss_prepare")) return rc;
    Ws w;
    const LossDesc d = with_ws(desc, workspace, &w);
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(workspace, 0, w.zero_bytes, st);
    cudaMemsetAsync(w.maps, 0xFF, w.maps_bytes, st);
    if (d.ncols > 0) launch_k(loss_maps_kernel, ceil_div(d.ncols, 256), 256, 0, st, d);
    DFINE_LAUNCH_CHECK("loss_prepare");
    return 0;
}

// Step 2 (any order): per-family sums over every head into the workspace accumulators.
DFINE_API int dfine_loss_vfl_fwd(const dfine_loss_desc* desc, void* workspace, void* stream) {
    if (int rc = check_desc(desc, workspace, "loss_vfl_fwd")) return rc;
    Ws w;
    const LossDesc d = with_ws(desc, workspace, &w);
    const int nq = d.Q > d.n_dn ? d.Q : d.n_dn;
    dim3 grid(ceil_div((long)d.B * nq, 8), heads_total(d));
    launch_k(loss_vfl_kernel<false>, grid, 256, 0, (cudaStream_t)stream, d, w.acc, VflOut{nullptr, nullptr, nullptr, nullptr});
    DFINE_LAUNCH_CHECK("loss_vfl_fwd");
    return 0;
}
DFINE_API int dfine_loss_box_fwd(const dfine_loss_desc* desc, void* workspace, void* stream) {
    if (int rc = check_desc(desc, workspace, "loss_box_fwd")) return rc;
    Ws w;
    const LossDesc d = with_ws(desc, workspace, &w);
    const long n = d.go_cap > d.n_dn_entries ? d.go_cap : d.n_dn_entries;
    if (n == 0) return 0;
    dim3 grid(ceil_div(n, 128), heads_total(d));
    launch_k(loss_box_kernel<false>, grid, 128, 0, (cudaStream_t)stream, d, w.acc, BoxOut{nullptr, nullptr, nullptr, nullptr, nullptr});
    DFINE_LAUNCH_CHECK("loss_box_fwd");
    return 0;
}
DFINE_API int dfine_loss_fgl_ddf_fwd(const dfine_loss_desc* desc, void* workspace, void* stream) {
    if (int rc = check_desc(desc, workspace, "loss_fgl_ddf_fwd")) return rc;
    Ws w;
    const LossDesc d = with_ws(desc, workspace, &w);
    const int nq = d.Q > d.n_dn ? d.Q : d.n_dn;
    dim3 grid(ceil_div((long)d.B * nq * 4, 128), d.L, d.n_dn > 0 ? 2 : 1);
    launch_k(loss_local_kernel<false>, grid, 128, 0, (cudaStream_t)stream, d, w.acc, w.notsame, nullptr, nullptr, nullptr, nullptr);
    DFINE_LAUNCH_CHECK("loss_fgl_ddf_fwd");
    return 0;
}
// Step 3: accumulators -> the loss scalars (unweighted; nan -> 0) and the DDF coefficients the backward needs.
DFINE_API int dfine_loss_finalize(const dfine_loss_desc* desc, void* workspace, float* out, void* stream) {
    if (int rc = check_desc(desc, workspace, "loss_finalize")) return rc;
    Ws w;
    const LossDesc d = with_ws(desc, workspace, &w);
    launch_k(loss_finalize_kernel, 1, 32, 0, (cudaStream_t)stream, d, w.acc, w.notsame, out);
    DFINE_LAUNCH_CHECK("loss_finalize");
    return 0;
}

// Backward.  gout: upstream gradients in dfine_loss_finalize's layout (first 6(L+2) + 4L floats).  The gradient tensors
// have the shapes of the forward tensors; d(logits), d(pre_logits), d(enc_logits) and d(corners) are written completely,
// the box gradients must be zero-filled by the caller (only matched rows are written).
DFINE_API int dfine_loss_vfl_bwd(const dfine_loss_desc* desc, void* workspace, const float* gout, float* dlogits,
                                 float* dpre_logits, float* denc_logits, void* stream) {
    if (int rc = check_desc(desc, workspace, "loss_vfl_bwd")) return rc;
    Ws w;
    const LossDesc d = with_ws(desc, workspace, &w);
    const int nq = d.Q > d.n_dn ? d.Q : d.n_dn;
    dim3 grid(ceil_div((long)d.B * nq, 8), heads_total(d));
    launch_k(loss_vfl_kernel<true>, grid, 256, 0, (cudaStream_t)stream, d, w.acc, VflOut{dlogits, dpre_logits, denc_logits, gout});
    DFINE_LAUNCH_CHECK("loss_vfl_bwd");
    return 0;
}
DFINE_API int dfine_loss_box_bwd(const dfine_loss_desc* desc, void* workspace, const float* gout, float* dboxes,
                                 float* dpre_boxes, float* denc_boxes, void* stream) {
    if (int rc = check_desc(desc, workspace, "loss_box_bwd")) return rc;
    DFINE_REQUIRE(((uintptr_t)dboxes % 16) == 0 && ((uintptr_t)dpre_boxes % 16) == 0 && ((uintptr_t)denc_boxes % 16) == 0,
                  "loss_box_bwd: gradient tensors must be 16-byte aligned");
    Ws w;
    const LossDesc d = with_ws(desc, workspace, &w);
    const long n = d.go_cap > d.n_dn_entries ? d.go_cap : d.n_dn_entries;
    if (n == 0) return 0;
    const int H = d.L + 2;
    dim3 grid(ceil_div(n, 128), heads_total(d));
    launch_k(loss_box_kernel<true>, grid, 128, 0, (cudaStream_t)stream, 
        d, w.acc, BoxOut{dboxes, dpre_boxes, denc_boxes, gout + 2 * H, gout + 4 * H});
    DFINE_LAUNCH_CHECK("loss_box_bwd");
    return 0;
}
DFINE_API int dfine_loss_fgl_ddf_bwd(const dfine_loss_desc* desc, void* workspace, const float* out, const float* gout,
                                     float* dcorners, void* stream) {
    if (int rc = check_desc(desc, workspace, "loss_fgl_ddf_bwd")) return rc;
    Ws w;
    const LossDesc d = with_ws(desc, workspace, &w);
    const int nq = d.Q > d.n_dn ? d.Q : d.n_dn;
    const int H = d.L + 2;
    dim3 grid(ceil_div((long)d.B * nq * 4, 128), d.L, d.n_dn > 0 ? 2 : 1);
    launch_k(loss_local_kernel<true>, grid, 128, 0, (cudaStream_t)stream, d, w.acc, w.notsame, dcorners, out, gout + 6 * H,
                                                                   gout + 6 * H + 2 * d.L);
    DFINE_LAUNCH_CHECK("loss_fgl_ddf_bwd");
    return 0;
}
