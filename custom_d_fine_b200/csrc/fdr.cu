// FDR head of the D-FINE decoder, forward + backward, one kernel each.
//
// Replaces, per decoder layer, the reference's Integral (softmax over the reg_max+1 bins of each of the 4 box
// edges, expectation against the weighting function W(n), dfine_decoder.py:291-295), distance2bbox
// (arch/utils.py:119-142) and the LQE statistics (softmax again, top-k probabilities + their mean,
// dfine_decoder.py:307-311) — ~25 tiny ATen kernels forward (softmax x2, a gemv, topk, cat, mean, ~15
// elementwise) and as many backward — with ONE pass over pred_corners.
//
// Work decomposition: one warp per query row (4 edges x NB bins contiguous), 8 lanes per edge; lane j of an edge
// group holds bins j, j+8, ..., so every global access of the row is coalesced.  Softmax, expectation and the
// top-k search are xor-shuffle reductions inside the 8-lane group.  NB <= 64 (reg_max 32 -> 33 bins).
//
// Gradients: pred_corners only.  The reference points entering distance2bbox are detached
// (dfine_decoder.py:493), `project` is a function of the non-trainable up / reg_scale parameters.
#include "common.cuh"

namespace {

constexpr int MAXB = 8;    // bins per lane (NB <= 64)
constexpr int MAXK = 8;

template <typename T>
__device__ __forceinline__ T grp_xor(T v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ float grp_sum(float v) {
    v += grp_xor(v, 1); v += grp_xor(v, 2); v += grp_xor(v, 4);
    return v;
}
__device__ __forceinline__ float grp_max(float v) {
    v = fmaxf(v, grp_xor(v, 1)); v = fmaxf(v, grp_xor(v, 2)); v = fmaxf(v, grp_xor(v, 4));
    return v;
}

// softmax of this lane's bins of one edge (prob[i] for bin j + 8 i); returns nothing else
__device__ __forceinline__ void edge_softmax(const float* __restrict__ crow, int NB, int j, float (&prob)[MAXB]) {
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < MAXB; ++i) {
        const int b = j + 8 * i;
        prob[i] = b < NB ? __ldg(crow + b) : -INFINITY;
        m = fmaxf(m, prob[i]);
    }
    m = grp_max(m);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXB; ++i) {
        prob[i] = (j + 8 * i) < NB ? expf(prob[i] - m) : 0.f;
        s += prob[i];
    }
    s = grp_sum(s);
    const float inv = 1.f / s;
#pragma unroll
    for (int i = 0; i < MAXB; ++i) prob[i] *= inv;
}

// the k largest probabilities of the edge, descending, ties to the lower bin; every lane of the group gets them
__device__ __forceinline__ void edge_topk(const float (&prob)[MAXB], int NB, int j, int k, float (&topv)[MAXK],
                                          int (&topi)[MAXK]) {
    unsigned taken = 0;   // bit i: this lane's bin j + 8 i was selected
    for (int r = 0; r < k; ++r) {
        float bv = -1.f;
        int bi = 1 << 30;
#pragma unroll
        for (int i = 0; i < MAXB; ++i) {
            const int b = j + 8 * i;
            if (b < NB && !((taken >> i) & 1u) && prob[i] > bv) { bv = prob[i]; bi = b; }   // ascending b: lowest bin wins ties
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const float ov = grp_xor(bv, o);
            const int oi = grp_xor(bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        topv[r] = bv;
        topi[r] = bi;
        if ((bi & 7) == j && bi < NB) taken |= 1u << (bi >> 3);
    }
}

__global__ void __launch_bounds__(256) fdr_head_fwd_kernel(const float* __restrict__ corners, const float* __restrict__ ref,
                                                           const float* __restrict__ project,
                                                           const float* __restrict__ reg_scale, float* __restrict__ box,
                                                           float* __restrict__ stat, long rows, int NB, int k) {
    pdl_entry();
    const long row = (long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (row >= rows) return;
    const int lane = threadIdx.x % 32, e = lane / 8, j = lane % 8;
    float prob[MAXB];
    edge_softmax(corners + (row * 4 + e) * NB, NB, j, prob);
    if (box) {
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < MAXB; ++i)
            if (j + 8 * i < NB) d += prob[i] * __ldg(project + j + 8 * i);
        d = grp_sum(d);
        const float d0 = __shfl_sync(0xffffffffu, d, 0), d1 = __shfl_sync(0xffffffffu, d, 8);
        const float d2 = __shfl_sync(0xffffffffu, d, 16), d3 = __shfl_sync(0xffffffffu, d, 24);
        if (lane == 0) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(ref + row * 4));
            const float rs = fabsf(__ldg(reg_scale));
            const float sw = r.z / rs, sh = r.w / rs;
            const float x1 = r.x - (0.5f * rs + d0) * sw, y1 = r.y - (0.5f * rs + d1) * sh;
            const float x2 = r.x + (0.5f * rs + d2) * sw, y2 = r.y + (0.5f * rs + d3) * sh;
            *reinterpret_cast<float4*>(box + row * 4) = make_float4((x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1);
        }
    }
    if (stat) {
        float topv[MAXK];
        int topi[MAXK];
        edge_topk(prob, NB, j, k, topv, topi);
        if (j == 0) {
            float* o = stat + (row * 4 + e) * (k + 1);
            float s = 0.f;
            for (int r = 0; r < k; ++r) { o[r] = topv[r]; s += topv[r]; }
            o[k] = s / (float)k;
        }
    }
}

__global__ void __launch_bounds__(256) fdr_head_bwd_kernel(const float* __restrict__ corners, const float* __restrict__ ref,
                                                           const float* __restrict__ project,
                                                           const float* __restrict__ reg_scale,
                                                           const float* __restrict__ dbox, const float* __restrict__ dstat,
                                                           float* __restrict__ dcorners, long rows, int NB, int k) {
    pdl_entry();
    const long row = (long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (row >= rows) return;
    const int lane = threadIdx.x % 32, e = lane / 8, j = lane % 8;
    float prob[MAXB], dprob[MAXB];
    edge_softmax(corners + (row * 4 + e) * NB, NB, j, prob);
#pragma unroll
    for (int i = 0; i < MAXB; ++i) dprob[i] = 0.f;
    if (dbox) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(ref + row * 4));
        const float4 gb = __ldg(reinterpret_cast<const float4*>(dbox + row * 4));
        const float rs = fabsf(__ldg(reg_scale));
        const float sw = r.z / rs, sh = r.w / rs;
        // box = ((x1+x2)/2, (y1+y2)/2, x2-x1, y2-y1); x1 = rx-(rs/2+d0) sw, y1 = ry-(rs/2+d1) sh, x2 = rx+(rs/2+d2) sw, ...
        const float dd = e == 0 ? -sw * (0.5f * gb.x - gb.z) : e == 1 ? -sh * (0.5f * gb.y - gb.w)
                       : e == 2 ? sw * (0.5f * gb.x + gb.z) : sh * (0.5f * gb.y + gb.w);
#pragma unroll
        for (int i = 0; i < MAXB; ++i)
            if (j + 8 * i < NB) dprob[i] = dd * __ldg(project + j + 8 * i);
    }
    if (dstat) {
        float topv[MAXK];
        int topi[MAXK];
        edge_topk(prob, NB, j, k, topv, topi);
        const float* g = dstat + (row * 4 + e) * (k + 1);
        const float gm = __ldg(g + k) / (float)k;
        for (int r = 0; r < k; ++r) {
            const int b = topi[r];
            if ((b & 7) == j && b < NB) {
                const float gv = __ldg(g + r) + gm;
#pragma unroll
                for (int i = 0; i < MAXB; ++i)
                    if (i == (b >> 3)) dprob[i] += gv;
            }
        }
    }
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < MAXB; ++i) dot += prob[i] * dprob[i];
    dot = grp_sum(dot);
    float* o = dcorners + (row * 4 + e) * NB;
#pragma unroll
    for (int i = 0; i < MAXB; ++i)
        if (j + 8 * i < NB) o[j + 8 * i] = prob[i] * (dprob[i] - dot);
}

}  // namespace

// corners [rows, 4*NB] (rows = B*Lq), ref [rows, 4] cxcywh, project [NB] = W(n), reg_scale: device scalar.
// box [rows, 4] (nullable) = distance2bbox(ref, Integral(corners)); stat [rows, 4*(k+1)] (nullable) = per edge the
// k largest bin probabilities (descending) and their mean.
DFINE_API int dfine_fdr_head_fwd(const float* corners, const float* ref, const float* project, const float* reg_scale,
                                 float* box, float* stat, long rows, int NB, int k, void* stream) {
    DFINE_REQUIRE(NB >= 1 && NB <= 8 * MAXB && k >= 0 && k <= MAXK && k <= NB, "fdr_head_fwd: NB=%d k=%d unsupported", NB, k);
    DFINE_REQUIRE(!box || (((uintptr_t)ref % 16) == 0 && ((uintptr_t)box % 16) == 0), "fdr_head_fwd: alignment");
    if (rows == 0) return 0;
    launch_k(fdr_head_fwd_kernel, ceil_div(rows, 8), 256, 0, (cudaStream_t)stream, corners, ref, project, reg_scale, box,
                                                                             stat, rows, NB, k);
    DFINE_LAUNCH_CHECK("fdr_head_fwd");
    return 0;
}

// dcorners [rows, 4*NB] (fully overwritten) from dbox [rows, 4] and / or dstat [rows, 4*(k+1)] (either nullable).
DFINE_API int dfine_fdr_head_bwd(const float* corners, const float* ref, const float* project, const float* reg_scale,
                                 const float* dbox, const float* dstat, float* dcorners, long rows, int NB, int k,
                                 void* stream) {
    DFINE_REQUIRE(NB >= 1 && NB <= 8 * MAXB && k >= 0 && k <= MAXK && k <= NB, "fdr_head_bwd: NB=%d k=%d unsupported", NB, k);
    DFINE_REQUIRE(!dbox || (((uintptr_t)ref % 16) == 0 && ((uintptr_t)dbox % 16) == 0), "fdr_head_bwd: alignment");
    if (rows == 0) return 0;
    launch_k(fdr_head_bwd_kernel, ceil_div(rows, 8), 256, 0, (cudaStream_t)stream, corners, ref, project, reg_scale, dbox,
                                                                             dstat, dcorners, rows, NB, k);
    DFINE_LAUNCH_CHECK("fdr_head_bwd");
    return 0;
}
