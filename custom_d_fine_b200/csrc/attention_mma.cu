// Flash-style multi-head attention core on the warp-level tensor-core path (mma.sync m16n8k8, tf32 operands with
// 3xTF32 error compensation, fp32 accumulate), forward + backward.
//
// Replaces nn.MultiheadAttention's explicit-softmax path (hybrid_encoder.py:256,277 AIFI; dfine_decoder.py:200,239
// decoder self-attention with the CDN block mask).  The sequences are short (400 / 500 tokens, head_dim 32) and
// the work tiny next to the convolutions (2.4 GFLOP/img forward), so the kernels favour simplicity: 64-query x
// 64-key tiles, 4 warps x 16 rows, K/V (or Q/dO) tiles staged in shared memory, online softmax in registers.
// Every product a*b is evaluated as a_lo*b_hi + a_hi*b_lo + a_hi*b_hi with hi = round-to-nearest tf32, so the
// result has fp32-class accuracy (the 1e-3 parity bar of the decoder outputs leaves no room for plain tf32
// logits).  tcgen05 is reserved for the GEMM-shaped bulk of the network (gemm_tc.cu); at M = 16 rows per warp
// and K = 32 a TMEM round trip per 64-key tile would cost more than the math.
//
// Fragment layout of mma.m16n8k8 (g = lane / 4, t = lane % 4):
//   A[16x8]  a0 = (g, t)  a1 = (g+8, t)  a2 = (g, t+4)  a3 = (g+8, t+4)
//   B[8x8]   b0 = (k = t, n = g)          b1 = (k = t+4, n = g)
//   C[16x8]  c0 = (g, 2t) c1 = (g, 2t+1)  c2 = (g+8, 2t) c3 = (g+8, 2t+1)
#include "common.cuh"

namespace attn_mma {

constexpr int NT = 128;   // 4 warps
constexpr int BQ = 64;    // rows owned by a CTA (16 per warp)
constexpr int BT = 64;    // streamed tile (keys in fwd / dq, queries in dkv)
constexpr int PLD = BT + 4;

__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    lo = __float_as_uint(x - __uint_as_float(hi));   // consumed truncated to tf32: 2^-21 relative in total
}
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += a * b with both operands split (small terms first)
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], float b0, float b1) {
    uint32_t bh0, bl0, bh1, bl1;
    split(b0, bh0, bl0);
    split(b1, bh1, bl1);
    mma(c, al, bh0, bh1);
    mma(c, ah, bl0, bl1);
    mma(c, ah, bh0, bh1);
}

// c += a * b with the B operand already split (tiles staged as hi / lo planes)
__device__ __forceinline__ void mma3s(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                      uint32_t bh1, uint32_t bl0, uint32_t bl1) {
    mma(c, al, bh0, bh1);
    mma(c, ah, bl0, bl1);
    mma(c, ah, bh0, bh1);
}

// rows r0 = row0 + g and r0 + 8 of a [S, ld] matrix as A fragments for every k-step (8 columns each), scaled
template <int HD>
__device__ __forceinline__ void load_a_frags(const float* __restrict__ base, long ld, int row0, int S, float scale, int g,
                                             int t, uint32_t (&hi)[HD / 8][4], uint32_t (&lo)[HD / 8][4]) {
    const int r0 = row0 + g, r1 = r0 + 8;
#pragma unroll
    for (int kk = 0; kk < HD / 8; ++kk) {
        const float v0 = r0 < S ? __ldg(base + (long)r0 * ld + 8 * kk + t) * scale : 0.f;
        const float v1 = r1 < S ? __ldg(base + (long)r1 * ld + 8 * kk + t) * scale : 0.f;
        const float v2 = r0 < S ? __ldg(base + (long)r0 * ld + 8 * kk + t + 4) * scale : 0.f;
        const float v3 = r1 < S ? __ldg(base + (long)r1 * ld + 8 * kk + t + 4) * scale : 0.f;
        split(v0, hi[kk][0], lo[kk][0]); split(v1, hi[kk][1], lo[kk][1]);
        split(v2, hi[kk][2], lo[kk][2]); split(v3, hi[kk][3], lo[kk][3]);
    }
}

// stage rows [row0, row0 + BT) of a [S, ld] matrix (HD columns) into smem as two planes (tf32 hi | lo remainder)
// with row pitch LD; zero past S.  Splitting once per tile (not once per warp and use) halves the instruction
// count of the MMA loops.
template <int HD, int LD>
__device__ __forceinline__ void stage(uint32_t* hi, uint32_t* lo, const float* __restrict__ src, long ld, int row0, int S) {
    for (int e = threadIdx.x; e < BT * (HD / 4); e += NT) {
        const int r = e / (HD / 4), c4 = e % (HD / 4);
        float4 v = make_float4(0, 0, 0, 0);
        if (row0 + r < S) v = __ldg(reinterpret_cast<const float4*>(src + (long)(row0 + r) * ld + c4 * 4));
        uint4 h, l;
        split(v.x, h.x, l.x); split(v.y, h.y, l.y); split(v.z, h.z, l.z); split(v.w, h.w, l.w);
        *reinterpret_cast<uint4*>(hi + r * LD + c4 * 4) = h;
        *reinterpret_cast<uint4*>(lo + r * LD + c4 * 4) = l;
    }
}

// acc[nt] (16 x 8 per n-tile, NTILES tiles) += A(regs, HD wide) * Bs^T where Bs is [rows = n][HD] in smem (pitch LD)
template <int HD, int LD, int NTILES>
__device__ __forceinline__ void gemm_abt(float (&acc)[NTILES][4], const uint32_t (&ah)[HD / 8][4],
                                         const uint32_t (&al)[HD / 8][4], const uint32_t* Bh, const uint32_t* Bl, int g,
                                         int t) {
#pragma unroll
    for (int nt = 0; nt < NTILES; ++nt)
#pragma unroll
        for (int kk = 0; kk < HD / 8; ++kk) {
            const int i0 = (8 * nt + g) * LD + 8 * kk + t;
            mma3s(acc[nt], ah[kk], al[kk], Bh[i0], Bh[i0 + 4], Bl[i0], Bl[i0 + 4]);
        }
}

// acc[nt] (16 x 8 per n-tile over HD columns) += P(16 x BT, smem pitch PLD) * Bs where Bs is [rows = k][HD] (pitch LD)
template <int HD, int LD>
__device__ __forceinline__ void gemm_pb(float (&acc)[HD / 8][4], const float* Ps, const uint32_t* Bh, const uint32_t* Bl,
                                        int g, int t) {
#pragma unroll
    for (int kk = 0; kk < BT / 8; ++kk) {
        uint32_t ah[4], al[4];
        split(Ps[g * PLD + 8 * kk + t], ah[0], al[0]);
        split(Ps[(g + 8) * PLD + 8 * kk + t], ah[1], al[1]);
        split(Ps[g * PLD + 8 * kk + t + 4], ah[2], al[2]);
        split(Ps[(g + 8) * PLD + 8 * kk + t + 4], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt) {
            const int i0 = (8 * kk + t) * LD + 8 * nt + g, i1 = (8 * kk + t + 4) * LD + 8 * nt + g;
            mma3s(acc[nt], ah, al, Bh[i0], Bh[i1], Bl[i0], Bl[i1]);
        }
    }
}

__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// store a C-layout 16 x BT tile into the warp's smem buffer (pitch PLD)
__device__ __forceinline__ void store_c(float* Ps, const float (&c)[BT / 8][4], int g, int t) {
#pragma unroll
    for (int nt = 0; nt < BT / 8; ++nt) {
        *reinterpret_cast<float2*>(Ps + g * PLD + 8 * nt + 2 * t) = make_float2(c[nt][0], c[nt][1]);
        *reinterpret_cast<float2*>(Ps + (g + 8) * PLD + 8 * nt + 2 * t) = make_float2(c[nt][2], c[nt][3]);
    }
}

// ------------------------------------------------------------------------------------------- forward
template <int HD>
struct FwdSmem {
    uint32_t k[BT * (HD + 4)], kl[BT * (HD + 4)];
    uint32_t v[BT * (HD + 8)], vl[BT * (HD + 8)];
    float p[NT / 32][16 * PLD];
};

template <int HD>
__global__ void __launch_bounds__(NT) fwd_kernel(const float* __restrict__ q, long ldq, const float* __restrict__ k, long ldk,
                                                 const float* __restrict__ v, long ldv,
                                                 const unsigned char* __restrict__ mask, float* __restrict__ o, long ldo,
                                                 float* __restrict__ lse, int S, int H, float scale) {
    extern __shared__ __align__(16) unsigned char raw[];
    FwdSmem<HD>& sm = *reinterpret_cast<FwdSmem<HD>*>(raw);
    const int b = blockIdx.z, h = blockIdx.y, warp = threadIdx.x / 32, lane = threadIdx.x % 32, g = lane / 4, t = lane % 4;
    const int row0 = blockIdx.x * BQ + 16 * warp, r0 = row0 + g, r1 = r0 + 8;
    const float* qb = q + (long)b * S * ldq + h * HD;
    const float* kb = k + (long)b * S * ldk + h * HD;
    const float* vb = v + (long)b * S * ldv + h * HD;
    uint32_t qh[HD / 8][4], ql[HD / 8][4];
    load_a_frags<HD>(qb, ldq, row0, S, scale, g, t, qh, ql);
    float oacc[HD / 8][4] = {};
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float* Ps = sm.p[warp];
    for (int t0 = 0; t0 < S; t0 += BT) {
        __syncthreads();
        stage<HD, HD + 4>(sm.k, sm.kl, kb, ldk, t0, S);
        stage<HD, HD + 8>(sm.v, sm.vl, vb, ldv, t0, S);
        __syncthreads();
        float s[BT / 8][4] = {};
        gemm_abt<HD, HD + 4, BT / 8>(s, qh, ql, sm.k, sm.kl, g, t);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < BT / 8; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int key = t0 + 8 * nt + 2 * t + j;
                const bool in = key < S;
                if (!in || r0 >= S || (mask && mask[(long)r0 * S + key])) s[nt][j] = -INFINITY;
                if (!in || r1 >= S || (mask && mask[(long)r1 * S + key])) s[nt][2 + j] = -INFINITY;
                mx0 = fmaxf(mx0, s[nt][j]);
                mx1 = fmaxf(mx1, s[nt][2 + j]);
            }
        mx0 = quad_max(mx0);
        mx1 = quad_max(mx1);
        const float n0 = fmaxf(m0, mx0), n1 = fmaxf(m1, mx1);
        const float c0 = (n0 == -INFINITY) ? 1.f : expf(m0 - n0), c1 = (n1 == -INFINITY) ? 1.f : expf(m1 - n1);
        float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < BT / 8; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                s[nt][j] = (s[nt][j] == -INFINITY) ? 0.f : expf(s[nt][j] - n0);
                s[nt][2 + j] = (s[nt][2 + j] == -INFINITY) ? 0.f : expf(s[nt][2 + j] - n1);
                ps0 += s[nt][j];
                ps1 += s[nt][2 + j];
            }
        l0 = l0 * c0 + quad_sum(ps0);
        l1 = l1 * c1 + quad_sum(ps1);
        m0 = n0;
        m1 = n1;
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt) { oacc[nt][0] *= c0; oacc[nt][1] *= c0; oacc[nt][2] *= c1; oacc[nt][3] *= c1; }
        __syncwarp();
        store_c(Ps, s, g, t);
        __syncwarp();
        gemm_pb<HD, HD + 8>(oacc, Ps, sm.v, sm.vl, g, t);
    }
    const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
    float* ob = o + (long)b * S * ldo + h * HD;
#pragma unroll
    for (int nt = 0; nt < HD / 8; ++nt) {
        if (r0 < S) *reinterpret_cast<float2*>(ob + (long)r0 * ldo + 8 * nt + 2 * t) = make_float2(oacc[nt][0] * i0, oacc[nt][1] * i0);
        if (r1 < S) *reinterpret_cast<float2*>(ob + (long)r1 * ldo + 8 * nt + 2 * t) = make_float2(oacc[nt][2] * i1, oacc[nt][3] * i1);
    }
    if (t == 0) {
        if (r0 < S) lse[((long)b * H + h) * S + r0] = m0 + logf(l0);
        if (r1 < S) lse[((long)b * H + h) * S + r1] = m1 + logf(l1);
    }
}

// ------------------------------------------------------------------------------------------- dQ (+ D = rowsum(dO*O))
template <int HD>
struct DqSmem {
    uint32_t k[BT * (HD + 4)], kl[BT * (HD + 4)];
    uint32_t v[BT * (HD + 4)], vl[BT * (HD + 4)];
    float p[NT / 32][16 * PLD];
};

template <int HD>
__global__ void __launch_bounds__(NT) dq_kernel(const float* __restrict__ q, long ldq, const float* __restrict__ k, long ldk,
                                                const float* __restrict__ v, long ldv,
                                                const unsigned char* __restrict__ mask, const float* __restrict__ o,
                                                long ldo, const float* __restrict__ dout, long ldd,
                                                const float* __restrict__ lse, float* __restrict__ dsum,
                                                float* __restrict__ dq, long lddq, int S, int H, float scale) {
    extern __shared__ __align__(16) unsigned char raw[];
    DqSmem<HD>& sm = *reinterpret_cast<DqSmem<HD>*>(raw);
    const int b = blockIdx.z, h = blockIdx.y, warp = threadIdx.x / 32, lane = threadIdx.x % 32, g = lane / 4, t = lane % 4;
    const int row0 = blockIdx.x * BQ + 16 * warp, r0 = row0 + g, r1 = r0 + 8;
    const float* qb = q + (long)b * S * ldq + h * HD;
    const float* kb = k + (long)b * S * ldk + h * HD;
    const float* vb = v + (long)b * S * ldv + h * HD;
    const float* ob = o + (long)b * S * ldo + h * HD;
    const float* db = dout + (long)b * S * ldd + h * HD;
    uint32_t qh[HD / 8][4], ql[HD / 8][4], dh[HD / 8][4], dl[HD / 8][4];
    load_a_frags<HD>(qb, ldq, row0, S, scale, g, t, qh, ql);
    load_a_frags<HD>(db, ldd, row0, S, 1.f, g, t, dh, dl);
    // D_i = sum_d dO_i[d] * O_i[d]: this lane's columns (8kk + t, 8kk + t + 4), then the quad
    float D0 = 0.f, D1 = 0.f;
#pragma unroll
    for (int kk = 0; kk < HD / 8; ++kk)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = 8 * kk + t + 4 * j;
            if (r0 < S) D0 += __ldg(db + (long)r0 * ldd + c) * __ldg(ob + (long)r0 * ldo + c);
            if (r1 < S) D1 += __ldg(db + (long)r1 * ldd + c) * __ldg(ob + (long)r1 * ldo + c);
        }
    D0 = quad_sum(D0);
    D1 = quad_sum(D1);
    if (t == 0) {
        if (r0 < S) dsum[((long)b * H + h) * S + r0] = D0;
        if (r1 < S) dsum[((long)b * H + h) * S + r1] = D1;
    }
    const float L0 = r0 < S ? __ldg(lse + ((long)b * H + h) * S + r0) : 0.f;
    const float L1 = r1 < S ? __ldg(lse + ((long)b * H + h) * S + r1) : 0.f;
    float acc[HD / 8][4] = {};
    float* Ps = sm.p[warp];
    for (int t0 = 0; t0 < S; t0 += BT) {
        __syncthreads();
        stage<HD, HD + 4>(sm.k, sm.kl, kb, ldk, t0, S);
        stage<HD, HD + 4>(sm.v, sm.vl, vb, ldv, t0, S);
        __syncthreads();
        float s[BT / 8][4] = {}, dp[BT / 8][4] = {};
        gemm_abt<HD, HD + 4, BT / 8>(s, qh, ql, sm.k, sm.kl, g, t);
        gemm_abt<HD, HD + 4, BT / 8>(dp, dh, dl, sm.v, sm.vl, g, t);
#pragma unroll
        for (int nt = 0; nt < BT / 8; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int key = t0 + 8 * nt + 2 * t + j;
                const bool in = key < S;
                const bool ok0 = in && r0 < S && !(mask && mask[(long)r0 * S + key]);
                const bool ok1 = in && r1 < S && !(mask && mask[(long)r1 * S + key]);
                s[nt][j] = ok0 ? expf(s[nt][j] - L0) * (dp[nt][j] - D0) : 0.f;
                s[nt][2 + j] = ok1 ? expf(s[nt][2 + j] - L1) * (dp[nt][2 + j] - D1) : 0.f;
            }
        __syncwarp();
        store_c(Ps, s, g, t);
        __syncwarp();
        gemm_pb<HD, HD + 4>(acc, Ps, sm.k, sm.kl, g, t);
    }
    float* qo = dq + (long)b * S * lddq + h * HD;
#pragma unroll
    for (int nt = 0; nt < HD / 8; ++nt) {
        if (r0 < S) *reinterpret_cast<float2*>(qo + (long)r0 * lddq + 8 * nt + 2 * t) = make_float2(acc[nt][0] * scale, acc[nt][1] * scale);
        if (r1 < S) *reinterpret_cast<float2*>(qo + (long)r1 * lddq + 8 * nt + 2 * t) = make_float2(acc[nt][2] * scale, acc[nt][3] * scale);
    }
}

// ------------------------------------------------------------------------------------------- dK, dV
template <int HD>
struct DkvSmem {
    uint32_t q[BT * (HD + 4)], ql[BT * (HD + 4)];
    uint32_t d[BT * (HD + 4)], dl[BT * (HD + 4)];
    float lse[BT], dsum[BT];
    float p[NT / 32][16 * PLD];
    float ds[NT / 32][16 * PLD];
};

template <int HD>
__global__ void __launch_bounds__(NT) dkv_kernel(const float* __restrict__ q, long ldq, const float* __restrict__ k, long ldk,
                                                 const float* __restrict__ v, long ldv,
                                                 const unsigned char* __restrict__ mask,
                                                 const float* __restrict__ dout, long ldd, const float* __restrict__ lse,
                                                 const float* __restrict__ dsum, float* __restrict__ dk, long lddk,
                                                 float* __restrict__ dv, long lddv, int S, int H, float scale) {
    extern __shared__ __align__(16) unsigned char raw[];
    DkvSmem<HD>& sm = *reinterpret_cast<DkvSmem<HD>*>(raw);
    const int b = blockIdx.z, h = blockIdx.y, warp = threadIdx.x / 32, lane = threadIdx.x % 32, g = lane / 4, t = lane % 4;
    const int row0 = blockIdx.x * BQ + 16 * warp, r0 = row0 + g, r1 = r0 + 8;   // key rows
    const float* qb = q + (long)b * S * ldq + h * HD;
    const float* kb = k + (long)b * S * ldk + h * HD;
    const float* vb = v + (long)b * S * ldv + h * HD;
    const float* db = dout + (long)b * S * ldd + h * HD;
    uint32_t kh[HD / 8][4], kl[HD / 8][4], vh[HD / 8][4], vl[HD / 8][4];
    load_a_frags<HD>(kb, ldk, row0, S, scale, g, t, kh, kl);
    load_a_frags<HD>(vb, ldv, row0, S, 1.f, g, t, vh, vl);
    float av[HD / 8][4] = {}, ak[HD / 8][4] = {};
    float* Ps = sm.p[warp];
    float* Ds = sm.ds[warp];
    for (int t0 = 0; t0 < S; t0 += BT) {
        __syncthreads();
        stage<HD, HD + 4>(sm.q, sm.ql, qb, ldq, t0, S);
        stage<HD, HD + 4>(sm.d, sm.dl, db, ldd, t0, S);
        for (int e = threadIdx.x; e < BT; e += NT) {
            const bool ok = t0 + e < S;
            sm.lse[e] = ok ? __ldg(lse + ((long)b * H + h) * S + t0 + e) : 0.f;
            sm.dsum[e] = ok ? __ldg(dsum + ((long)b * H + h) * S + t0 + e) : 0.f;
        }
        __syncthreads();
        float s[BT / 8][4] = {}, dp[BT / 8][4] = {};
        gemm_abt<HD, HD + 4, BT / 8>(s, kh, kl, sm.q, sm.ql, g, t);     // S^T[key, query]
        gemm_abt<HD, HD + 4, BT / 8>(dp, vh, vl, sm.d, sm.dl, g, t);    // dP^T[key, query] = V dO^T
#pragma unroll
        for (int nt = 0; nt < BT / 8; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int qi_l = 8 * nt + 2 * t + j, qi = t0 + qi_l;
                const bool in = qi < S;
                const bool ok0 = in && r0 < S && !(mask && mask[(long)qi * S + r0]);
                const bool ok1 = in && r1 < S && !(mask && mask[(long)qi * S + r1]);
                const float p0 = ok0 ? expf(s[nt][j] - sm.lse[qi_l]) : 0.f;
                const float p1 = ok1 ? expf(s[nt][2 + j] - sm.lse[qi_l]) : 0.f;
                s[nt][j] = p0;
                s[nt][2 + j] = p1;
                dp[nt][j] = p0 * (dp[nt][j] - sm.dsum[qi_l]);
                dp[nt][2 + j] = p1 * (dp[nt][2 + j] - sm.dsum[qi_l]);
            }
        __syncwarp();
        store_c(Ps, s, g, t);
        store_c(Ds, dp, g, t);
        __syncwarp();
        gemm_pb<HD, HD + 4>(av, Ps, sm.d, sm.dl, g, t);   // dV += P^T dO
        gemm_pb<HD, HD + 4>(ak, Ds, sm.q, sm.ql, g, t);   // dK += dS^T Q
    }
    float* ko = dk + (long)b * S * lddk + h * HD;
    float* vo = dv + (long)b * S * lddv + h * HD;
#pragma unroll
    for (int nt = 0; nt < HD / 8; ++nt) {
        if (r0 < S) {
            *reinterpret_cast<float2*>(vo + (long)r0 * lddv + 8 * nt + 2 * t) = make_float2(av[nt][0], av[nt][1]);
            *reinterpret_cast<float2*>(ko + (long)r0 * lddk + 8 * nt + 2 * t) = make_float2(ak[nt][0] * scale, ak[nt][1] * scale);
        }
        if (r1 < S) {
            *reinterpret_cast<float2*>(vo + (long)r1 * lddv + 8 * nt + 2 * t) = make_float2(av[nt][2], av[nt][3]);
            *reinterpret_cast<float2*>(ko + (long)r1 * lddk + 8 * nt + 2 * t) = make_float2(ak[nt][2] * scale, ak[nt][3] * scale);
        }
    }
}

template <typename K>
int set_smem(K fn, int bytes, bool& done) {
    if (done) return 0;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) { dfine_set_error("attention(mma): smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    done = true;
    return 0;
}

template <int HD>
int launch_fwd(const float* q, long ldq, const float* k, long ldk, const float* v, long ldv, const unsigned char* mask,
               float* o, long ldo, float* lse, int B, int S, int H, float scale, cudaStream_t st) {
    static bool done = false;
    int rc = set_smem(fwd_kernel<HD>, (int)sizeof(FwdSmem<HD>), done);
    if (rc) return rc;
    fwd_kernel<HD><<<dim3(ceil_div(S, BQ), H, B), NT, sizeof(FwdSmem<HD>), st>>>(q, ldq, k, ldk, v, ldv, mask, o, ldo, lse,
                                                                                S, H, scale);
    return 0;
}

template <int HD>
int launch_bwd(const float* q, long ldq, const float* k, long ldk, const float* v, long ldv, const unsigned char* mask,
               const float* o, long ldo, const float* dout, long ldd, const float* lse, float* dsum, float* dq, long lddq,
               float* dk, long lddk, float* dv, long lddv, int B, int S, int H, float scale, cudaStream_t st) {
    static bool d1 = false, d2 = false;
    int rc = set_smem(dq_kernel<HD>, (int)sizeof(DqSmem<HD>), d1);
    if (rc) return rc;
    rc = set_smem(dkv_kernel<HD>, (int)sizeof(DkvSmem<HD>), d2);
    if (rc) return rc;
    const dim3 grid(ceil_div(S, BQ), H, B);
    dq_kernel<HD><<<grid, NT, sizeof(DqSmem<HD>), st>>>(q, ldq, k, ldk, v, ldv, mask, o, ldo, dout, ldd, lse, dsum, dq, lddq,
                                                       S, H, scale);
    dkv_kernel<HD><<<grid, NT, sizeof(DkvSmem<HD>), st>>>(q, ldq, k, ldk, v, ldv, mask, dout, ldd, lse, dsum, dk, lddk, dv,
                                                         lddv, S, H, scale);
    return 0;
}

}  // namespace attn_mma

// Internal (not part of the C ABI): called by dfine_attn_fwd / dfine_attn_bwd for head_dim 16 / 32 / 48 / 64.
int attn_mma_fwd(const float* q, long ldq, const float* k, long ldk, const float* v, long ldv, const unsigned char* mask,
                 float* o, long ldo, float* lse, int B, int S, int H, int head_dim, float scale, cudaStream_t st) {
    switch (head_dim) {
        case 16: return attn_mma::launch_fwd<16>(q, ldq, k, ldk, v, ldv, mask, o, ldo, lse, B, S, H, scale, st);
        case 32: return attn_mma::launch_fwd<32>(q, ldq, k, ldk, v, ldv, mask, o, ldo, lse, B, S, H, scale, st);
        case 48: return attn_mma::launch_fwd<48>(q, ldq, k, ldk, v, ldv, mask, o, ldo, lse, B, S, H, scale, st);
        case 64: return attn_mma::launch_fwd<64>(q, ldq, k, ldk, v, ldv, mask, o, ldo, lse, B, S, H, scale, st);
    }
    return -1;
}

int attn_mma_bwd(const float* q, long ldq, const float* k, long ldk, const float* v, long ldv, const unsigned char* mask,
                 const float* o, long ldo, const float* dout, long ldd, const float* lse, float* dsum, float* dq, long lddq,
                 float* dk, long lddk, float* dv, long lddv, int B, int S, int H, int head_dim, float scale,
                 cudaStream_t st) {
    switch (head_dim) {
        case 16: return attn_mma::launch_bwd<16>(q, ldq, k, ldk, v, ldv, mask, o, ldo, dout, ldd, lse, dsum, dq, lddq, dk, lddk, dv, lddv, B, S, H, scale, st);
        case 32: return attn_mma::launch_bwd<32>(q, ldq, k, ldk, v, ldv, mask, o, ldo, dout, ldd, lse, dsum, dq, lddq, dk, lddk, dv, lddv, B, S, H, scale, st);
        case 48: return attn_mma::launch_bwd<48>(q, ldq, k, ldk, v, ldv, mask, o, ldo, dout, ldd, lse, dsum, dq, lddq, dk, lddk, dv, lddv, B, S, H, scale, st);
        case 64: return attn_mma::launch_bwd<64>(q, ldq, k, ldk, v, ldv, mask, o, ldo, dout, ldd, lse, dsum, dq, lddq, dk, lddk, dv, lddv, B, S, H, scale, st);
    }
    return -1;
}
