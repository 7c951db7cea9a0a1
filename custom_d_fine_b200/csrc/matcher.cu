// Hungarian matcher for sm_100a: block-diagonal cost matrices + rectangular LSAP, all
// (decoder layer, image) problems of a training step in ONE launch, no host round trip.
//
// Replaces HungarianMatcher.forward (reference src/d_fine/matcher.py:110-257): the reference builds
// the full [B*Q, sum T] cost matrix with ~25 ATen kernels, copies it to the host (device sync) and
// calls SciPy's C++ rectangular_lsap per image; this kernel computes only the per-image [Q, T_b]
// blocks (matcher.py:135-172 arithmetic, fp32, same operation order, FMA contraction disabled for
// this file) and solves each with the same shortest-augmenting-path algorithm and the same
// tie-breaking as SciPy (Crouse 2016; see oracle/lsap.c for the CPU restatement used as checker):
//   * costs promoted to double, transpose solved when n_rows > n_cols,
//   * rows inserted in order; candidates scanned in the order of SciPy's `remaining` list
//     (reverse-filled, swap-with-last removal); ties prefer an unassigned column (the last such
//     one in scan order), otherwise the first in scan order,
//   * pairs emitted sorted by query index.
// One warp per problem: the scan over the remaining columns is lane-strided over list positions and
// merged with a shuffle reduction whose comparator reproduces the sequential rule exactly.
//
// This translation unit is compiled with -fmad=false.
#include "common.cuh"

namespace {

struct Key {
    double v;
    int it;   // position in `remaining`
    int un;   // 1 if the column is unassigned
};

// true if a beats b under the sequential scan rule (see header comment)
__device__ __forceinline__ bool better(const Key& a, const Key& b) {
    if (a.it < 0) return false;
    if (b.it < 0) return true;
    if (a.v < b.v) return true;
    if (a.v > b.v) return false;
    if (a.un != b.un) return a.un > b.un;
    return a.un ? (a.it > b.it) : (a.it < b.it);
}

__device__ __forceinline__ Key shfl_key(const Key& k, int o) {
    Key r;
    r.v = __shfl_xor_sync(0xffffffffu, k.v, o);
    r.it = __shfl_xor_sync(0xffffffffu, k.it, o);
    r.un = __shfl_xor_sync(0xffffffffu, k.un, o);
    return r;
}

// A label outside [0, C) would be an out-of-bounds read of the logits row (the reference raises an index error there,
// matcher.py:150); labels are validated on the host where they are host tensors (train.DevicePrefetcher) and clamped
// here so that a bad device-resident label can never read outside the tensor.
__device__ __forceinline__ long clamp_label(long l, int C) { return l < 0 ? 0 : (l >= C ? C - 1 : l); }

__device__ __forceinline__ float cost_entry(const float* __restrict__ lg, const float* __restrict__ bx,
                                            long label, const float* __restrict__ tb, float alpha, float gamma,
                                            float w_class, float w_bbox, float w_giou, float extra = 0.f,
                                            bool has_extra = false) {
    // class term (matcher.py:150-158)
    const float x = lg[label];
    const float p = 1.f / (1.f + expf(-x));
    const float pg = gamma == 2.f ? p * p : powf(p, gamma);
    const float qg = gamma == 2.f ? (1.f - p) * (1.f - p) : powf(1.f - p, gamma);
    const float neg = (1.f - alpha) * pg * (-logf(1.f - p + 1e-8f));
    const float pos = alpha * qg * (-logf(p + 1e-8f));
    const float c_cls = pos - neg;
    // L1 on cxcywh (matcher.py:163)
    const float c_box = fabsf(bx[0] - tb[0]) + fabsf(bx[1] - tb[1]) + fabsf(bx[2] - tb[2]) + fabsf(bx[3] - tb[3]);
    // -GIoU on xyxy with w,h clamped at 0 (arch/utils.py:28-67)
    const float w1 = fmaxf(bx[2], 0.f), h1 = fmaxf(bx[3], 0.f), w2 = fmaxf(tb[2], 0.f), h2 = fmaxf(tb[3], 0.f);
    const float ax0 = bx[0] - 0.5f * w1, ay0 = bx[1] - 0.5f * h1, ax1 = bx[0] + 0.5f * w1, ay1 = bx[1] + 0.5f * h1;
    const float bx0 = tb[0] - 0.5f * w2, by0 = tb[1] - 0.5f * h2, bx1 = tb[0] + 0.5f * w2, by1 = tb[1] + 0.5f * h2;
    const float a1 = (ax1 - ax0) * (ay1 - ay0), a2 = (bx1 - bx0) * (by1 - by0);
    const float iw = fmaxf(fminf(ax1, bx1) - fmaxf(ax0, bx0), 0.f), ih = fmaxf(fminf(ay1, by1) - fmaxf(ay0, by0), 0.f);
    const float inter = iw * ih;
    const float uni = a1 + a2 - inter;
    const float iou = inter / uni;
    const float cw = fmaxf(fmaxf(ax1, bx1) - fminf(ax0, bx0), 0.f), ch = fmaxf(fmaxf(ay1, by1) - fminf(ay0, by0), 0.f);
    const float area = cw * ch;
    const float giou = iou - (area - uni) / area;
    float c = w_bbox * c_box + w_class * c_cls + w_giou * (-giou);
    if (has_extra) c = c + extra;     // segmentation: mask cost block added before the NaN guard (matcher.py:233-242)
    // torch.nan_to_num(C, nan=1.0) (matcher.py:242)
    if (isnan(c)) c = 1.0f;
    else if (isinf(c)) c = c > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
    return c;
}

// shared-memory carve-up for one problem with solve dims R (rows) x Cc (cols)
struct Work {
    float* cost;  // [R][Cc]
    double *u, *v, *spc;
    int *path, *row4col, *col4row, *remaining;
    unsigned char *SR, *SC;
};

__global__ void __launch_bounds__(32) matcher_kernel(
    const float* __restrict__ logits,  // [NL,B,Q,C]
    const float* __restrict__ boxes,   // [NL,B,Q,4]
    const long* __restrict__ labels,   // [sumT]
    const float* __restrict__ tboxes,  // [sumT,4]
    const int* __restrict__ toff,      // [B+1]
    long* __restrict__ out_q, long* __restrict__ out_t,  // [NL,sumT]
    float* __restrict__ cost_out,      // optional [NL, Q*sumT] (block b at Q*toff[b], [Q,T_b] row-major) or null
    float* __restrict__ workspace,     // used for the cost block when it does not fit in smem
    const float* __restrict__ extra,   // optional additive cost, same layout as cost_out (mask cost), or null
    int NL, int B, int Q, int C, int sumT, int Tmax, int cost_in_smem, float alpha, float gamma,
    float w_class, float w_bbox, float w_giou) {
    pdl_entry();
    extern __shared__ __align__(16) unsigned char smem[];
    const int prob = blockIdx.x;
    const int layer = prob / B, b = prob % B;
    const int lane = threadIdx.x;
    const int t0 = toff[b], T = toff[b + 1] - t0;
    if (T <= 0) return;
    const bool transposed = T < Q;           // SciPy solves the transpose when n_rows > n_cols
    const int R = transposed ? T : Q;        // solve rows
    const int Cc = transposed ? Q : T;       // solve cols
    const int Rmax = Tmax < Q ? Tmax : Q, Cmax = Tmax < Q ? Q : Tmax;

    // carve
    Work w;
    unsigned char* sp = smem;
    w.u = (double*)sp; sp += sizeof(double) * Rmax;
    w.v = (double*)sp; sp += sizeof(double) * Cmax;
    w.spc = (double*)sp; sp += sizeof(double) * Cmax;
    w.path = (int*)sp; sp += sizeof(int) * Cmax;
    w.row4col = (int*)sp; sp += sizeof(int) * Cmax;
    w.remaining = (int*)sp; sp += sizeof(int) * Cmax;
    w.col4row = (int*)sp; sp += sizeof(int) * Rmax;
    w.SR = sp; sp += (Rmax + 3) / 4 * 4;
    w.SC = sp; sp += (Cmax + 3) / 4 * 4;
    sp = (unsigned char*)(((uintptr_t)sp + 15) & ~(uintptr_t)15);
    w.cost = cost_in_smem ? (float*)sp : workspace + (long)prob * Q * Tmax;

    // ---- cost block, stored in solve orientation cost[i*Cc + j] ----
    const float* lg = logits + ((long)layer * B + b) * Q * C;
    const float* bx = boxes + ((long)layer * B + b) * Q * 4;
    for (int e = lane; e < Q * T; e += 32) {
        const int q = e / T, t = e % T;
        const float ex = extra ? __ldg(extra + (long)layer * Q * sumT + (long)Q * t0 + e) : 0.f;
        const float c = cost_entry(lg + (long)q * C, bx + q * 4, clamp_label(labels[t0 + t], C), tboxes + (long)(t0 + t) * 4, alpha,
                                   gamma, w_class, w_bbox, w_giou, ex, extra != nullptr);
        if (cost_out) cost_out[(long)layer * Q * sumT + (long)Q * t0 + e] = c;
        if (transposed) w.cost[t * Cc + q] = c; else w.cost[q * Cc + t] = c;
    }
    for (int i = lane; i < R; i += 32) { w.u[i] = 0.0; w.col4row[i] = -1; }
    for (int j = lane; j < Cc; j += 32) { w.v[j] = 0.0; w.path[j] = -1; w.row4col[j] = -1; }
    __syncwarp();

    for (int cur = 0; cur < R; ++cur) {
        // ---- augmenting path search from row `cur` ----
        for (int i = lane; i < R; i += 32) w.SR[i] = 0;
        for (int j = lane; j < Cc; j += 32) { w.SC[j] = 0; w.spc[j] = INFINITY; w.remaining[j] = Cc - j - 1; }
        __syncwarp();
        int n_rem = Cc, i = cur, sink = -1;
        double min_val = 0.0;
        while (sink == -1) {
            if (lane == 0) w.SR[i] = 1;
            const double ui = w.u[i];
            const float* crow = w.cost + (long)i * Cc;
            Key best; best.v = INFINITY; best.it = -1; best.un = 0;
            for (int it = lane; it < n_rem; it += 32) {
                const int j = w.remaining[it];
                const double r = min_val + (double)crow[j] - ui - w.v[j];
                double s = w.spc[j];
                if (r < s) { w.path[j] = i; w.spc[j] = r; s = r; }
                Key k; k.v = s; k.it = it; k.un = (w.row4col[j] == -1);
                if (better(k, best)) best = k;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const Key other = shfl_key(best, o);
                if (better(other, best)) best = other;
            }
            min_val = best.v;
            if (best.it < 0 || min_val == INFINITY) { sink = -2; break; }  // infeasible (cannot happen after nan_to_num)
            const int j = w.remaining[best.it];
            const int r4c = w.row4col[j];
            if (r4c == -1) sink = j; else i = r4c;
            __syncwarp();
            if (lane == 0) { w.SC[j] = 1; w.remaining[best.it] = w.remaining[n_rem - 1]; }
            --n_rem;
            __syncwarp();
        }
        if (sink < 0) break;
        // ---- dual updates ----
        for (int r = lane; r < R; r += 32) {
            if (r == cur) w.u[r] += min_val;
            else if (w.SR[r]) w.u[r] += min_val - w.spc[w.col4row[r]];
        }
        for (int j = lane; j < Cc; j += 32)
            if (w.SC[j]) w.v[j] -= min_val - w.spc[j];
        __syncwarp();
        // ---- augment ----
        if (lane == 0) {
            int j = sink;
            for (;;) {
                const int r = w.path[j];
                w.row4col[j] = r;
                const int t = w.col4row[r];
                w.col4row[r] = j;
                j = t;
                if (r == cur) break;
            }
        }
        __syncwarp();
    }

    // ---- emit pairs sorted by query index ----
    long* oq = out_q + (long)layer * sumT + t0;
    long* ot = out_t + (long)layer * sumT + t0;
    if (transposed) {
        for (int i = lane; i < R; i += 32) {
            const int q = w.col4row[i];
            int rank = 0;
            for (int k = 0; k < R; ++k) rank += (w.col4row[k] < q);
            oq[rank] = q;
            ot[rank] = i;
        }
    } else {
        for (int i = lane; i < R; i += 32) { oq[i] = i; ot[i] = w.col4row[i]; }
    }
}

}  // namespace

DFINE_API long dfine_matcher_workspace_bytes(int NL, int B, int Q, int Tmax) {
    const int Rmax = Tmax < Q ? Tmax : Q, Cmax = Tmax < Q ? Q : Tmax;
    const long fixed = 8L * Rmax + 16L * Cmax + 12L * Cmax + 4L * Rmax + (Rmax + 3) / 4 * 4 + (Cmax + 3) / 4 * 4 + 16;
    const long cost = 4L * Q * Tmax;
    if (fixed + cost <= 200 * 1024) return 0;
    return (long)NL * B * cost;
}

// logits [NL,B,Q,C], boxes [NL,B,Q,4] (cxcywh), labels int64 [sumT], tboxes [sumT,4], toff int32 [B+1]
// (device).  out_q/out_t int64 [NL,sumT]: for image b of layer l the min(Q,T_b) matched pairs, sorted
// by query index, start at l*sumT + toff[b].  cost_out (optional) receives the fp32 cost blocks.
namespace {
int matcher_impl(const float* logits, const float* boxes, const long* labels, const float* tboxes, const int* toff,
                 long* out_q, long* out_t, float* cost_out, float* workspace, const float* extra, int NL, int B, int Q,
                 int C, int sumT, int Tmax, float alpha, float gamma, float w_class, float w_bbox, float w_giou,
                 void* stream) {
    if (NL * B == 0 || sumT == 0) return 0;
    DFINE_REQUIRE(Tmax > 0 && Q > 0 && C > 0, "matcher: bad dims");
    const int Rmax = Tmax < Q ? Tmax : Q, Cmax = Tmax < Q ? Q : Tmax;
    const long fixed = 8L * Rmax + 16L * Cmax + 12L * Cmax + 4L * Rmax + (Rmax + 3) / 4 * 4 + (Cmax + 3) / 4 * 4 + 16;
    const long cost = 4L * Q * Tmax;
    const int in_smem = fixed + cost <= 200 * 1024;
    DFINE_REQUIRE(fixed <= 200 * 1024, "matcher: problem too large (Q=%d, Tmax=%d)", Q, Tmax);
    DFINE_REQUIRE(in_smem || workspace != nullptr, "matcher: workspace required for Q=%d Tmax=%d", Q, Tmax);
    const long smem = fixed + (in_smem ? cost : 0);
    DFINE_SET_SMEM_ONCE(matcher_kernel, 220 * 1024, "matcher");
    launch_k(matcher_kernel, NL * B, 32, smem, (cudaStream_t)stream, logits, boxes, labels, tboxes, toff, out_q, out_t,
                                                              cost_out, workspace, extra, NL, B, Q, C, sumT, Tmax,
                                                              in_smem, alpha, gamma, w_class, w_bbox, w_giou);
    DFINE_LAUNCH_CHECK("matcher");
    return 0;
}
}  // namespace

DFINE_API int dfine_matcher(const float* logits, const float* boxes, const long* labels, const float* tboxes,
                            const int* toff, long* out_q, long* out_t, float* cost_out, float* workspace, int NL,
                            int B, int Q, int C, int sumT, int Tmax, float alpha, float gamma, float w_class,
                            float w_bbox, float w_giou, void* stream) {
    return matcher_impl(logits, boxes, labels, tboxes, toff, out_q, out_t, cost_out, workspace, nullptr, NL, B, Q, C,
                        sumT, Tmax, alpha, gamma, w_class, w_bbox, w_giou, stream);
}

// The same with an additive cost `extra` [NL, Q*sumT] in cost_out's layout (image b's [Q, T_b] block row-major at
// Q*toff[b]), added before the NaN guard: the mask term of the segmentation matcher (matcher.py:175-237).
DFINE_API int dfine_matcher_extra(const float* logits, const float* boxes, const long* labels, const float* tboxes,
                                  const int* toff, const float* extra, long* out_q, long* out_t, float* cost_out,
                                  float* workspace, int NL, int B, int Q, int C, int sumT, int Tmax, float alpha,
                                  float gamma, float w_class, float w_bbox, float w_giou, void* stream) {
    DFINE_REQUIRE(extra != nullptr, "matcher_extra: null extra cost");
    return matcher_impl(logits, boxes, labels, tboxes, toff, out_q, out_t, cost_out, workspace, extra, NL, B, Q, C, sumT,
                        Tmax, alpha, gamma, w_class, w_bbox, w_giou, stream);
}
