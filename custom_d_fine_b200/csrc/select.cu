// Query selection and decoder gate — the remaining "glue" ops of the D-FINE decoder as kernels.
//
//   dfine_topk_rowmax    DFINETransformer._select_topk (dfine_decoder.py:875-910, "default" method): the top-k memory
//                        tokens of every image by their best class logit — row max over C classes + top-k of L tokens
//                        (8 400 at 640x640, 33 600 at 1280x1280) + descending sort, one CTA per image, no score tensor
//                        in HBM (replaces max + torch.topk's gather / sort kernels).
//   dfine_gate_mix_*     Gate.forward's mixing (dfine_decoder.py:267-271): sigmoid(g[:, :D]) * x1 + sigmoid(g[:, D:]) * x2
//                        forward and backward (the LayerNorm that follows is the library's layernorm kernel).
//
// Both are HBM/latency-bound: 2.7 MB of logits per image for the selection, 3 x 2 MB per decoder layer for the gate.
#include "common.cuh"

namespace {

// order-preserving float -> uint key (larger float <-> larger key); NaN sorts above +inf like torch.topk (NaN = largest)
__device__ __forceinline__ uint32_t fkey(float f) {
    if (f != f) return 0xFFFFFFFFu;
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int TK_THREADS = 1024;

// keys[L] in dynamic shared memory.  Radix select (4 x 8-bit passes) of the k-th largest key, then the selected
// (key, index) pairs — every key above the threshold plus the lowest-index keys equal to it — are sorted descending
// (ties: lower index first) by a bitonic network over KP = next power of two >= k slots.
// scores[b, t] = max_c logits[b, t, c] (NaN if any class is NaN): one warp per token over the whole grid
__global__ void __launch_bounds__(256) rowmax_kernel(const float* __restrict__ logits, float* __restrict__ scores, long rows,
                                                     int C) {
    pdl_entry();
    const long t = (long)blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (t >= rows) return;
    float m = -INFINITY;
    bool nan = false;
    for (int c = lane; c < C; c += 32) {
        const float v = __ldg(logits + t * C + c);
        nan |= v != v;
        m = fmaxf(m, v);
    }
    m = warp_max(m);
    nan = __any_sync(0xffffffffu, nan);
    if (lane == 0) scores[t] = nan ? __int_as_float(0x7fc00000) : m;
}

template <int KP>
__global__ void __launch_bounds__(TK_THREADS) topk_rowmax_kernel(const float* __restrict__ scores, long* __restrict__ out,
                                                                 int L, int k) {
    pdl_entry();
    extern __shared__ uint32_t keys[];
    __shared__ uint32_t hist[256];
    __shared__ uint32_t sel_prefix, sel_remaining, n_above, n_equal_taken;
    __shared__ unsigned long long slots[KP];       // (key << 32) | (0xFFFFFFFF - index): descending sort = key desc, index asc
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    // phase 1: the image's scores as order-preserving keys
    for (int t = tid; t < L; t += TK_THREADS) keys[t] = fkey(__ldg(scores + (long)b * L + t));
    if (tid == 0) { sel_prefix = 0; sel_remaining = (uint32_t)k; }
    __syncthreads();
    // phase 2: radix select from the most significant byte down; afterwards sel_prefix = the k-th largest key
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        const uint32_t hi_mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const uint32_t prefix = sel_prefix;
        for (int t = tid; t < L; t += TK_THREADS) {
            const uint32_t key = keys[t];
            if ((key & hi_mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t rem = sel_remaining;
            int d = 255;
            for (; d > 0; --d) {
                if (hist[d] >= rem) break;
                rem -= hist[d];
            }
            sel_prefix = prefix | ((uint32_t)d << shift);
            sel_remaining = rem;          // how many keys with this digit (and, after the last pass, equal to the threshold) are needed
        }
        __syncthreads();
    }
    const uint32_t thr = sel_prefix;
    const uint32_t need_equal = sel_remaining;
    for (int i = tid; i < KP; i += TK_THREADS) slots[i] = 0ull;      // padding sorts last
    if (tid == 0) { n_above = 0; n_equal_taken = 0; }
    __syncthreads();
    // phase 3a: keys strictly above the threshold (fewer than k of them), any order
    for (int t = tid; t < L; t += TK_THREADS) {
        const uint32_t key = keys[t];
        if (key > thr) {
            const uint32_t pos = atomicAdd(&n_above, 1u);
            slots[pos] = ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)t);
        }
    }
    __syncthreads();
    // phase 3b: the lowest-index keys EQUAL to the threshold, in index order (deterministic tie handling): chunks of
    // TK_THREADS tokens, ballot-based ordered compaction
    {
        __shared__ uint32_t warp_cnt[TK_THREADS / 32];
        __shared__ uint32_t chunk_base;
        if (tid == 0) chunk_base = 0;
        __syncthreads();
        for (int t0 = 0; t0 < L; t0 += TK_THREADS) {
            const int t = t0 + tid;
            const bool eq = t < L && keys[t] == thr;
            const uint32_t bal = __ballot_sync(0xffffffffu, eq);
            if (lane == 0) warp_cnt[warp] = __popc(bal);
            __syncthreads();
            uint32_t before = chunk_base;
            for (int w = 0; w < warp; ++w) before += warp_cnt[w];
            const uint32_t rank = before + __popc(bal & ((1u << lane) - 1u));
            if (eq && rank < need_equal)
                slots[n_above + rank] = ((unsigned long long)thr << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)t);
            __syncthreads();
            if (tid == 0) {
                uint32_t tot = 0;
                for (int w = 0; w < TK_THREADS / 32; ++w) tot += warp_cnt[w];
                chunk_base += tot;
            }
            __syncthreads();
            if (chunk_base >= need_equal) break;
        }
    }
    __syncthreads();
    // phase 4: bitonic sort, descending
    for (int size = 2; size <= KP; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < KP / 2; i += TK_THREADS) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long a = slots[lo], c = slots[hi];
                if ((a < c) == desc) { slots[lo] = c; slots[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < k; i += TK_THREADS)
        out[(long)b * k + i] = (long)(0xFFFFFFFFu - (uint32_t)(slots[i] & 0xFFFFFFFFull));
}

__global__ void gate_mix_fwd_kernel(const float* __restrict__ g, const float* __restrict__ x1, const float* __restrict__ x2,
                                    float* __restrict__ out, long rows, int D) {
    pdl_entry();
    const long n4 = rows * (D / 4);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const long r = i / (D / 4);
        const int c = (int)(i % (D / 4)) * 4;
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(g + r * 2 * D + c));
        const float4 g2 = __ldg(reinterpret_cast<const float4*>(g + r * 2 * D + D + c));
        const float4 a = __ldg(reinterpret_cast<const float4*>(x1 + r * D + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(x2 + r * D + c));
        float4 o;
        o.x = a.x / (1.f + expf(-g1.x)) + b.x / (1.f + expf(-g2.x));
        o.y = a.y / (1.f + expf(-g1.y)) + b.y / (1.f + expf(-g2.y));
        o.z = a.z / (1.f + expf(-g1.z)) + b.z / (1.f + expf(-g2.z));
        o.w = a.w / (1.f + expf(-g1.w)) + b.w / (1.f + expf(-g2.w));
        *reinterpret_cast<float4*>(out + r * D + c) = o;
    }
}

__device__ __forceinline__ void gate_bwd1(float go, float g, float x, float& dg, float& dx) {
    const float s = 1.f / (1.f + expf(-g));
    dx = go * s;
    dg = go * x * s * (1.f - s);
}
__global__ void gate_mix_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ g, const float* __restrict__ x1,
                                    const float* __restrict__ x2, float* __restrict__ dg, float* __restrict__ dx1,
                                    float* __restrict__ dx2, long rows, int D) {
    pdl_entry();
    const long n4 = rows * (D / 4);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const long r = i / (D / 4);
        const int c = (int)(i % (D / 4)) * 4;
        const float4 go = __ldg(reinterpret_cast<const float4*>(dout + r * D + c));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(g + r * 2 * D + c));
        const float4 g2 = __ldg(reinterpret_cast<const float4*>(g + r * 2 * D + D + c));
        const float4 a = __ldg(reinterpret_cast<const float4*>(x1 + r * D + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(x2 + r * D + c));
        float4 d1, d2, da, db;
        gate_bwd1(go.x, g1.x, a.x, d1.x, da.x); gate_bwd1(go.y, g1.y, a.y, d1.y, da.y);
        gate_bwd1(go.z, g1.z, a.z, d1.z, da.z); gate_bwd1(go.w, g1.w, a.w, d1.w, da.w);
        gate_bwd1(go.x, g2.x, b.x, d2.x, db.x); gate_bwd1(go.y, g2.y, b.y, d2.y, db.y);
        gate_bwd1(go.z, g2.z, b.z, d2.z, db.z); gate_bwd1(go.w, g2.w, b.w, d2.w, db.w);
        *reinterpret_cast<float4*>(dg + r * 2 * D + c) = d1;
        *reinterpret_cast<float4*>(dg + r * 2 * D + D + c) = d2;
        *reinterpret_cast<float4*>(dx1 + r * D + c) = da;
        *reinterpret_cast<float4*>(dx2 + r * D + c) = db;
    }
}

}  // namespace

// idx[b, i] (int64 [B, k]) = the token with the i-th largest row maximum of logits[b] ([B, L, C] contiguous), descending,
// ties broken towards the lower token index.  `scores` = caller-provided scratch [B, L] floats (the row maxima, written
// by a grid-wide pass); the selection runs one CTA per image with the image's L keys in shared memory.
DFINE_API int dfine_topk_rowmax(const float* logits, float* scores, long* idx, int B, int L, int C, int k, void* stream) {
    DFINE_REQUIRE(B >= 0 && L >= 1 && C >= 1 && k >= 1 && k <= L && k <= 1024, "topk_rowmax: B=%d L=%d C=%d k=%d", B, L, C, k);
    DFINE_REQUIRE((long)L * 4 <= 200 * 1024, "topk_rowmax: %d tokens exceed the shared-memory score buffer", L);
    if (B == 0) return 0;
    const int smem = L * 4;
    cudaStream_t st = (cudaStream_t)stream;
    launch_k(rowmax_kernel, ceil_div((long)B * L, 8), 256, 0, st, logits, scores, (long)B * L, C);
    if (k <= 512) {
        DFINE_SET_SMEM_ONCE((topk_rowmax_kernel<512>), 200 * 1024, "topk_rowmax");
        launch_k(topk_rowmax_kernel<512>, B, TK_THREADS, smem, st, scores, idx, L, k);
    } else {
        DFINE_SET_SMEM_ONCE((topk_rowmax_kernel<1024>), 200 * 1024, "topk_rowmax");
        launch_k(topk_rowmax_kernel<1024>, B, TK_THREADS, smem, st, scores, idx, L, k);
    }
    DFINE_LAUNCH_CHECK("topk_rowmax");
    return 0;
}

// out[r, :] = sigmoid(g[r, :D]) * x1[r, :] + sigmoid(g[r, D:]) * x2[r, :]; all row-major contiguous, D % 4 == 0.
DFINE_API int dfine_gate_mix_fwd(const float* g, const float* x1, const float* x2, float* out, long rows, int D, void* stream) {
    DFINE_REQUIRE(D % 4 == 0 && D > 0, "gate_mix: D=%d must be a multiple of 4", D);
    DFINE_REQUIRE(((uintptr_t)g % 16) == 0 && ((uintptr_t)x1 % 16) == 0 && ((uintptr_t)x2 % 16) == 0 && ((uintptr_t)out % 16) == 0,
                  "gate_mix: pointers must be 16-byte aligned");
    if (rows == 0) return 0;
    long blocks = (rows * (D / 4) + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_k(gate_mix_fwd_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, g, x1, x2, out, rows, D);
    DFINE_LAUNCH_CHECK("gate_mix_fwd");
    return 0;
}
DFINE_API int dfine_gate_mix_bwd(const float* dout, const float* g, const float* x1, const float* x2, float* dg, float* dx1,
                                 float* dx2, long rows, int D, void* stream) {
    DFINE_REQUIRE(D % 4 == 0 && D > 0, "gate_mix_bwd: D=%d must be a multiple of 4", D);
    if (rows == 0) return 0;
    long blocks = (rows * (D / 4) + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_k(gate_mix_bwd_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, dout, g, x1, x2, dg, dx1, dx2, rows, D);
    DFINE_LAUNCH_CHECK("gate_mix_bwd");
    return 0;
}
