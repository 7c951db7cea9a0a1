// Flash-style multi-head attention core on the warp-level tensor-core path (mma.sync m16n8k8, tf32 operands,
// fp32 accumulate), forward + backward.
//
// Replaces nn.MultiheadAttention's explicit-softmax path (hybrid_encoder.py:256,277 AIFI; dfine_decoder.py:200,239
// decoder self-attention with the CDN block mask).  The sequences are short (400 / 500 tokens, head_dim 32):
// tcgen05 is reserved for the GEMM-shaped bulk of the network (gemm_tc.cu); at K = 32 a TMEM round trip per
// 64-key tile would cost more than the math.
//
// Precision.  Everything that decides the forward output is error-compensated 3xTF32
// (a*b = a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, hi = round-to-nearest tf32: fp32-class logits and P*V — the 1e-3 parity
// bar of the decoder outputs leaves no room for plain tf32 logits).  The backward recomputes the logits the same
// way (P must equal the forward's) and evaluates the four gradient products (dP, dQ, dK, dV) with single
// round-to-nearest tf32 MMAs — the operand precision of every other gradient GEMM of the library.
//
// Structure.  CTA = 4 warps x 32 rows (two m16 tiles per warp, so every B fragment read from shared memory feeds
// two MMAs chains); the streamed operand (keys in fwd / dq, queries in dkv) is staged 64 rows at a time as
// pre-split tf32 hi / lo planes, the next tile's global loads are in flight in registers while the current one
// is consumed.
//
// Fragment layout of mma.m16n8k8 (g = lane / 4, t = lane % 4):
//   A[16x8]  a0 = (g, t)  a1 = (g+8, t)  a2 = (g, t+4)  a3 = (g+8, t+4)
//   B[8x8]   b0 = (k = t, n = g)          b1 = (k = t+4, n = g)
//   C[16x8]  c0 = (g, 2t) c1 = (g, 2t+1)  c2 = (g+8, 2t) c3 = (g+8, 2t+1)
// The contraction index of a k-step is permuted (slot t <-> element 2t, slot t+4 <-> element 2t+1) on BOTH
// operands, which leaves the product unchanged and buys two things: the B fragment of an A*B^T product is one
// 8-byte shared load per plane (elements 2t, 2t+1 of a row), and a C tile IS the A fragment of the next product
// (a0 = c0, a1 = c2, a2 = c1, a3 = c3) — P never goes through shared memory.
#include "common.cuh"

namespace attn_mma {

constexpr int NT = 128;   // 4 warps
// m16 tiles per warp: two while the fragments fit the register file (head_dim <= 32)
template <int HD> struct Cfg { static constexpr int MI = HD <= 32 ? 2 : 1, WR = 16 * MI, BQ = 4 * WR; };
constexpr int BT = 64;    // streamed tile
constexpr int HT = 32;    // processed in two halves (bounds the live C fragments)

__device__ __forceinline__ uint32_t rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
    hi = rn(x);
    lo = __float_as_uint(x - __uint_as_float(hi));   // consumed truncated to tf32: 2^-21 relative in total
}
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += a * b, both operands split (small terms first)
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                     uint32_t bh1, uint32_t bl0, uint32_t bl1) {
    mma(c, al, bh0, bh1);
    mma(c, ah, bl0, bl1);
    mma(c, ah, bh0, bh1);
}

// ---- global -> registers -> shared staging of a [BT x HD] tile --------------------------------------------
template <int HD>
struct TileRegs {
    static constexpr int N = BT * (HD / 4) / NT;
    float4 v[N];
};
template <int HD>
__device__ __forceinline__ void tile_load(TileRegs<HD>& r, const float* __restrict__ src, long ld, int row0, int S) {
#pragma unroll
    for (int i = 0; i < TileRegs<HD>::N; ++i) {
        const int e = threadIdx.x + i * NT, row = e / (HD / 4), c4 = e % (HD / 4);
        r.v[i] = (row0 + row < S) ? __ldg(reinterpret_cast<const float4*>(src + (long)(row0 + row) * ld + c4 * 4))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
// hi plane (pitch LD), optionally the lo plane (same pitch) and a second copy of the hi plane with pitch LD2
template <int HD, int LD, int LD2>
__device__ __forceinline__ void tile_store(const TileRegs<HD>& r, uint32_t* hi, uint32_t* lo, uint32_t* hi2) {
#pragma unroll
    for (int i = 0; i < TileRegs<HD>::N; ++i) {
        const int e = threadIdx.x + i * NT, row = e / (HD / 4), c4 = e % (HD / 4);
        uint4 h, l;
        split(r.v[i].x, h.x, l.x); split(r.v[i].y, h.y, l.y); split(r.v[i].z, h.z, l.z); split(r.v[i].w, h.w, l.w);
        *reinterpret_cast<uint4*>(hi + row * LD + c4 * 4) = h;
        if (lo) *reinterpret_cast<uint4*>(lo + row * LD + c4 * 4) = l;
        if (hi2) *reinterpret_cast<uint4*>(hi2 + row * LD2 + c4 * 4) = h;
    }
}

// A fragments (with the permuted contraction index) of rows row0+g, row0+g+8 of a [S, ld] matrix, scaled
template <int HD, bool LO>
__device__ __forceinline__ void load_a(const float* __restrict__ base, long ld, int row0, int S, float scale, int g, int t,
                                       uint32_t (&hi)[HD / 8][4], uint32_t (&lo)[HD / 8][4]) {
    const int r0 = row0 + g, r1 = r0 + 8;
#pragma unroll
    for (int kk = 0; kk < HD / 8; ++kk) {
        float2 u = make_float2(0.f, 0.f), w = u;
        if (r0 < S) u = __ldg(reinterpret_cast<const float2*>(base + (long)r0 * ld + 8 * kk + 2 * t));
        if (r1 < S) w = __ldg(reinterpret_cast<const float2*>(base + (long)r1 * ld + 8 * kk + 2 * t));
        if (LO) {
            split(u.x * scale, hi[kk][0], lo[kk][0]); split(w.x * scale, hi[kk][1], lo[kk][1]);
            split(u.y * scale, hi[kk][2], lo[kk][2]); split(w.y * scale, hi[kk][3], lo[kk][3]);
        } else {
            hi[kk][0] = rn(u.x * scale); hi[kk][1] = rn(w.x * scale);
            hi[kk][2] = rn(u.y * scale); hi[kk][3] = rn(w.y * scale);
        }
    }
}

// acc[mi][nt] (nt over HT/8 streamed rows starting at kb) += A * Bs^T, Bs = [rows][HD] hi / lo planes, pitch LD
template <int HD, int LD, int MI>
__device__ __forceinline__ void gemm_abt3(float (&acc)[MI][HT / 8][4], const uint32_t (&ah)[MI][HD / 8][4],
                                          const uint32_t (&al)[MI][HD / 8][4], const uint32_t* Bh, const uint32_t* Bl,
                                          int kb, int g, int t) {
    // term-major issue order: consecutive MMAs accumulate into different C tiles (no back-to-back dependent
    // tensor instructions; a dependent chain costs the full MMA latency per link)
#pragma unroll
    for (int kk = 0; kk < HD / 8; ++kk) {
        uint2 bh[HT / 8], bl[HT / 8];
#pragma unroll
        for (int nt = 0; nt < HT / 8; ++nt) {
            const int i0 = (kb + 8 * nt + g) * LD + 8 * kk + 2 * t;
            bh[nt] = *reinterpret_cast<const uint2*>(Bh + i0);
            bl[nt] = *reinterpret_cast<const uint2*>(Bl + i0);
        }
#pragma unroll
        for (int nt = 0; nt < HT / 8; ++nt)
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) mma(acc[mi][nt], al[mi][kk], bh[nt].x, bh[nt].y);
#pragma unroll
        for (int nt = 0; nt < HT / 8; ++nt)
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) mma(acc[mi][nt], ah[mi][kk], bl[nt].x, bl[nt].y);
#pragma unroll
        for (int nt = 0; nt < HT / 8; ++nt)
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) mma(acc[mi][nt], ah[mi][kk], bh[nt].x, bh[nt].y);
    }
}
// the same with single tf32 MMAs
template <int HD, int LD, int MI>
__device__ __forceinline__ void gemm_abt1(float (&acc)[MI][HT / 8][4], const uint32_t (&ah)[MI][HD / 8][4],
                                          const uint32_t* Bh, int kb, int g, int t) {
#pragma unroll
    for (int kk = 0; kk < HD / 8; ++kk)
#pragma unroll
        for (int nt = 0; nt < HT / 8; ++nt) {
            const uint2 bh = *reinterpret_cast<const uint2*>(Bh + (kb + 8 * nt + g) * LD + 8 * kk + 2 * t);
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) mma(acc[mi][nt], ah[mi][kk], bh.x, bh.y);
        }
}
// acc[mi][nt] (nt over HD/8 columns) += P * Bs, P = the C tiles p[mi][kk] (HT streamed rows from kb),
// Bs = [rows][HD] hi / lo planes with pitch LD (LD % 16 == 4: conflict-free)
template <int HD, int LD, int MI>
__device__ __forceinline__ void gemm_pb3(float (&acc)[MI][HD / 8][4], const float (&p)[MI][HT / 8][4], const uint32_t* Bh,
                                         const uint32_t* Bl, int kb, int g, int t) {
#pragma unroll
    for (int kk = 0; kk < HT / 8; ++kk) {
        uint32_t ah[MI][4], al[MI][4];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
            split(p[mi][kk][0], ah[mi][0], al[mi][0]); split(p[mi][kk][2], ah[mi][1], al[mi][1]);
            split(p[mi][kk][1], ah[mi][2], al[mi][2]); split(p[mi][kk][3], ah[mi][3], al[mi][3]);
        }
        uint32_t h0[HD / 8], h1[HD / 8], l0[HD / 8], l1[HD / 8];
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt) {
            const int i0 = (kb + 8 * kk + 2 * t) * LD + 8 * nt + g;
            h0[nt] = Bh[i0]; h1[nt] = Bh[i0 + LD]; l0[nt] = Bl[i0]; l1[nt] = Bl[i0 + LD];
        }
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt)
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) mma(acc[mi][nt], al[mi], h0[nt], h1[nt]);
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt)
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) mma(acc[mi][nt], ah[mi], l0[nt], l1[nt]);
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt)
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) mma(acc[mi][nt], ah[mi], h0[nt], h1[nt]);
    }
}
template <int HD, int LD, int MI>
__device__ __forceinline__ void gemm_pb1(float (&acc)[MI][HD / 8][4], const float (&p)[MI][HT / 8][4], const uint32_t* Bh,
                                         int kb, int g, int t) {
#pragma unroll
    for (int kk = 0; kk < HT / 8; ++kk) {
        uint32_t ah[MI][4];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
            ah[mi][0] = rn(p[mi][kk][0]); ah[mi][1] = rn(p[mi][kk][2]);
            ah[mi][2] = rn(p[mi][kk][1]); ah[mi][3] = rn(p[mi][kk][3]);
        }
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt) {
            const int i0 = (kb + 8 * kk + 2 * t) * LD + 8 * nt + g;
            const uint32_t h0 = Bh[i0], h1 = Bh[i0 + LD];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) mma(acc[mi][nt], ah[mi], h0, h1);
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
// mask[row][col], mask[row][col + 1] (col even); S even -> one 2-byte load
__device__ __forceinline__ void mask2(const unsigned char* __restrict__ mask, long row, int col, int S, bool even, bool& m0,
                                      bool& m1) {
    if (even) {
        const unsigned short v = __ldg(reinterpret_cast<const unsigned short*>(mask + row * S + col));
        m0 = v & 0xff;
        m1 = v >> 8;
    } else {
        m0 = __ldg(mask + row * S + col);
        m1 = (col + 1 < S) ? __ldg(mask + row * S + col + 1) : true;
    }
}

// ------------------------------------------------------------------------------------------- forward
template <int HD>
struct FwdSmem {
    uint32_t kh[BT * (HD + 8)], kl[BT * (HD + 8)];
    uint32_t vh[BT * (HD + 4)], vl[BT * (HD + 4)];
};

template <int HD>
__global__ void __launch_bounds__(NT) fwd_kernel(const float* __restrict__ q, long ldq, const float* __restrict__ k, long ldk,
                                                 const float* __restrict__ v, long ldv,
                                                 const unsigned char* __restrict__ mask, float* __restrict__ o, long ldo,
                                                 float* __restrict__ lse, int S, int H, float scale) {
    pdl_entry();
    extern __shared__ __align__(16) unsigned char raw[];
    FwdSmem<HD>& sm = *reinterpret_cast<FwdSmem<HD>*>(raw);
    const int b = blockIdx.z, h = blockIdx.y, warp = threadIdx.x / 32, lane = threadIdx.x % 32, g = lane / 4, t = lane % 4;
    constexpr int MI = Cfg<HD>::MI, WR = Cfg<HD>::WR, BQ = Cfg<HD>::BQ;
    const int row0 = blockIdx.x * BQ + WR * warp;
    const float* qb = q + (long)b * S * ldq + h * HD;
    const float* kb_ = k + (long)b * S * ldk + h * HD;
    const float* vb = v + (long)b * S * ldv + h * HD;
    const bool even = (S % 2) == 0 && (reinterpret_cast<uintptr_t>(mask) % 2) == 0;
    uint32_t qh[MI][HD / 8][4], ql[MI][HD / 8][4];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) load_a<HD, true>(qb, ldq, row0 + 16 * mi, S, scale, g, t, qh[mi], ql[mi]);
    float oacc[MI][HD / 8][4] = {};
    float mrun[MI][2], lrun[MI][2];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) { mrun[mi][0] = mrun[mi][1] = -INFINITY; lrun[mi][0] = lrun[mi][1] = 0.f; }
    TileRegs<HD> rk, rv;
    tile_load<HD>(rk, kb_, ldk, 0, S);
    tile_load<HD>(rv, vb, ldv, 0, S);
    const bool active = row0 < S;
    for (int t0 = 0; t0 < S; t0 += BT) {
        __syncthreads();
        tile_store<HD, HD + 8, 0>(rk, sm.kh, sm.kl, nullptr);
        tile_store<HD, HD + 4, 0>(rv, sm.vh, sm.vl, nullptr);
        __syncthreads();
        if (t0 + BT < S) {
            tile_load<HD>(rk, kb_, ldk, t0 + BT, S);
            tile_load<HD>(rv, vb, ldv, t0 + BT, S);
        }
        if (!active) continue;
#pragma unroll
        for (int half = 0; half < BT / HT; ++half) {
            const int kb = HT * half;
            if (t0 + kb >= S) break;
            float s[MI][HT / 8][4] = {};
            gemm_abt3<HD, HD + 8, MI>(s, qh, ql, sm.kh, sm.kl, kb, g, t);
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
                const int r0 = row0 + 16 * mi + g, r1 = r0 + 8;
                float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < HT / 8; ++nt) {
                    const int key = t0 + kb + 8 * nt + 2 * t;
                    bool a0 = false, a1 = false, b0 = false, b1 = false;
                    if (mask && key < S) {
                        if (r0 < S) mask2(mask, r0, key, S, even, a0, a1);
                        if (r1 < S) mask2(mask, r1, key, S, even, b0, b1);
                    }
                    if (key >= S || r0 >= S || a0) s[mi][nt][0] = -INFINITY;
                    if (key + 1 >= S || r0 >= S || a1) s[mi][nt][1] = -INFINITY;
                    if (key >= S || r1 >= S || b0) s[mi][nt][2] = -INFINITY;
                    if (key + 1 >= S || r1 >= S || b1) s[mi][nt][3] = -INFINITY;
                    mx0 = fmaxf(mx0, fmaxf(s[mi][nt][0], s[mi][nt][1]));
                    mx1 = fmaxf(mx1, fmaxf(s[mi][nt][2], s[mi][nt][3]));
                }
                mx0 = quad_max(mx0);
                mx1 = quad_max(mx1);
                const float n0 = fmaxf(mrun[mi][0], mx0), n1 = fmaxf(mrun[mi][1], mx1);
                // __expf = ex2.approx(x * log2 e): 2 ulp, one MUFU op instead of expf's ~20 instructions (the softmax,
                // not the MMAs, was the larger half of this kernel's instruction stream).  A row that has seen only
                // masked keys has n = -inf: subtract 0 instead, every exponent is then exp(-inf) = 0.
                const float z0 = (n0 == -INFINITY) ? 0.f : n0, z1 = (n1 == -INFINITY) ? 0.f : n1;
                const float c0 = __expf(mrun[mi][0] - z0), c1 = __expf(mrun[mi][1] - z1);
                float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
                for (int nt = 0; nt < HT / 8; ++nt)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        s[mi][nt][j] = __expf(s[mi][nt][j] - z0);
                        s[mi][nt][2 + j] = __expf(s[mi][nt][2 + j] - z1);
                        ps0 += s[mi][nt][j];
                        ps1 += s[mi][nt][2 + j];
                    }
                lrun[mi][0] = lrun[mi][0] * c0 + quad_sum(ps0);
                lrun[mi][1] = lrun[mi][1] * c1 + quad_sum(ps1);
                mrun[mi][0] = n0;
                mrun[mi][1] = n1;
#pragma unroll
                for (int nt = 0; nt < HD / 8; ++nt) {
                    oacc[mi][nt][0] *= c0; oacc[mi][nt][1] *= c0; oacc[mi][nt][2] *= c1; oacc[mi][nt][3] *= c1;
                }
            }
            gemm_pb3<HD, HD + 4, MI>(oacc, s, sm.vh, sm.vl, kb, g, t);
        }
    }
    if (!active) return;
    float* ob = o + (long)b * S * ldo + h * HD;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
        const int r0 = row0 + 16 * mi + g, r1 = r0 + 8;
        const float l0 = lrun[mi][0], l1 = lrun[mi][1];
        const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt) {
            if (r0 < S) *reinterpret_cast<float2*>(ob + (long)r0 * ldo + 8 * nt + 2 * t) = make_float2(oacc[mi][nt][0] * i0, oacc[mi][nt][1] * i0);
            if (r1 < S) *reinterpret_cast<float2*>(ob + (long)r1 * ldo + 8 * nt + 2 * t) = make_float2(oacc[mi][nt][2] * i1, oacc[mi][nt][3] * i1);
        }
        if (t == 0) {
            if (r0 < S) lse[((long)b * H + h) * S + r0] = mrun[mi][0] + logf(l0);
            if (r1 < S) lse[((long)b * H + h) * S + r1] = mrun[mi][1] + logf(l1);
        }
    }
}

// ------------------------------------------------------------------------------------------- dQ (+ D = rowsum(dO*O))
template <int HD>
struct DqSmem {
    uint32_t kh[BT * (HD + 8)], kl[BT * (HD + 8)], kh2[BT * (HD + 4)];
    uint32_t vh[BT * (HD + 8)];
};

template <int HD>
__global__ void __launch_bounds__(NT) dq_kernel(const float* __restrict__ q, long ldq, const float* __restrict__ k, long ldk,
                                                const float* __restrict__ v, long ldv,
                                                const unsigned char* __restrict__ mask, const float* __restrict__ o,
                                                long ldo, const float* __restrict__ dout, long ldd,
                                                const float* __restrict__ lse, float* __restrict__ dsum,
                                                float* __restrict__ dq, long lddq, int S, int H, float scale) {
    pdl_entry();
    extern __shared__ __align__(16) unsigned char raw[];
    DqSmem<HD>& sm = *reinterpret_cast<DqSmem<HD>*>(raw);
    const int b = blockIdx.z, h = blockIdx.y, warp = threadIdx.x / 32, lane = threadIdx.x % 32, g = lane / 4, t = lane % 4;
    constexpr int MI = Cfg<HD>::MI, WR = Cfg<HD>::WR, BQ = Cfg<HD>::BQ;
    const int row0 = blockIdx.x * BQ + WR * warp;
    const float* qb = q + (long)b * S * ldq + h * HD;
    const float* kb_ = k + (long)b * S * ldk + h * HD;
    const float* vb = v + (long)b * S * ldv + h * HD;
    const float* ob = o + (long)b * S * ldo + h * HD;
    const float* db = dout + (long)b * S * ldd + h * HD;
    const bool even = (S % 2) == 0 && (reinterpret_cast<uintptr_t>(mask) % 2) == 0;
    uint32_t qh[MI][HD / 8][4], ql[MI][HD / 8][4], dh[MI][HD / 8][4];
    float Lr[MI][2], Dr[MI][2];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
        load_a<HD, true>(qb, ldq, row0 + 16 * mi, S, scale, g, t, qh[mi], ql[mi]);
        load_a<HD, false>(db, ldd, row0 + 16 * mi, S, 1.f, g, t, dh[mi], dh[mi]);
        // D_i = sum_d dO_i[d] * O_i[d]: this lane's columns (8kk + 2t, 8kk + 2t + 1), then the quad
        const int r0 = row0 + 16 * mi + g, r1 = r0 + 8;
        float D0 = 0.f, D1 = 0.f;
#pragma unroll
        for (int kk = 0; kk < HD / 8; ++kk) {
            const int c = 8 * kk + 2 * t;
            if (r0 < S) {
                const float2 a = __ldg(reinterpret_cast<const float2*>(db + (long)r0 * ldd + c));
                const float2 w = __ldg(reinterpret_cast<const float2*>(ob + (long)r0 * ldo + c));
                D0 += a.x * w.x + a.y * w.y;
            }
            if (r1 < S) {
                const float2 a = __ldg(reinterpret_cast<const float2*>(db + (long)r1 * ldd + c));
                const float2 w = __ldg(reinterpret_cast<const float2*>(ob + (long)r1 * ldo + c));
                D1 += a.x * w.x + a.y * w.y;
            }
        }
        D0 = quad_sum(D0);
        D1 = quad_sum(D1);
        if (t == 0) {
            if (r0 < S) dsum[((long)b * H + h) * S + r0] = D0;
            if (r1 < S) dsum[((long)b * H + h) * S + r1] = D1;
        }
        Dr[mi][0] = D0;
        Dr[mi][1] = D1;
        Lr[mi][0] = r0 < S ? __ldg(lse + ((long)b * H + h) * S + r0) : 0.f;
        Lr[mi][1] = r1 < S ? __ldg(lse + ((long)b * H + h) * S + r1) : 0.f;
    }
    float acc[MI][HD / 8][4] = {};
    TileRegs<HD> rk, rv;
    tile_load<HD>(rk, kb_, ldk, 0, S);
    tile_load<HD>(rv, vb, ldv, 0, S);
    const bool active = row0 < S;
    for (int t0 = 0; t0 < S; t0 += BT) {
        __syncthreads();
        tile_store<HD, HD + 8, HD + 4>(rk, sm.kh, sm.kl, sm.kh2);
        tile_store<HD, HD + 8, 0>(rv, sm.vh, nullptr, nullptr);
        __syncthreads();
        if (t0 + BT < S) {
            tile_load<HD>(rk, kb_, ldk, t0 + BT, S);
            tile_load<HD>(rv, vb, ldv, t0 + BT, S);
        }
        if (!active) continue;
#pragma unroll
        for (int half = 0; half < BT / HT; ++half) {
            const int kb = HT * half;
            if (t0 + kb >= S) break;
            float s[MI][HT / 8][4] = {}, dp[MI][HT / 8][4] = {};
            gemm_abt3<HD, HD + 8, MI>(s, qh, ql, sm.kh, sm.kl, kb, g, t);
            gemm_abt1<HD, HD + 8, MI>(dp, dh, sm.vh, kb, g, t);
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
                const int r0 = row0 + 16 * mi + g, r1 = r0 + 8;
#pragma unroll
                for (int nt = 0; nt < HT / 8; ++nt) {
                    const int key = t0 + kb + 8 * nt + 2 * t;
                    bool a0 = false, a1 = false, b0 = false, b1 = false;
                    if (mask && key < S) {
                        if (r0 < S) mask2(mask, r0, key, S, even, a0, a1);
                        if (r1 < S) mask2(mask, r1, key, S, even, b0, b1);
                    }
                    const bool k0 = key < S, k1 = key + 1 < S;
                    s[mi][nt][0] = (k0 && r0 < S && !a0) ? __expf(s[mi][nt][0] - Lr[mi][0]) * (dp[mi][nt][0] - Dr[mi][0]) : 0.f;
                    s[mi][nt][1] = (k1 && r0 < S && !a1) ? __expf(s[mi][nt][1] - Lr[mi][0]) * (dp[mi][nt][1] - Dr[mi][0]) : 0.f;
                    s[mi][nt][2] = (k0 && r1 < S && !b0) ? __expf(s[mi][nt][2] - Lr[mi][1]) * (dp[mi][nt][2] - Dr[mi][1]) : 0.f;
                    s[mi][nt][3] = (k1 && r1 < S && !b1) ? __expf(s[mi][nt][3] - Lr[mi][1]) * (dp[mi][nt][3] - Dr[mi][1]) : 0.f;
                }
            }
            gemm_pb1<HD, HD + 4, MI>(acc, s, sm.kh2, kb, g, t);
        }
    }
    if (!active) return;
    float* qo = dq + (long)b * S * lddq + h * HD;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
        const int r0 = row0 + 16 * mi + g, r1 = r0 + 8;
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt) {
            if (r0 < S) *reinterpret_cast<float2*>(qo + (long)r0 * lddq + 8 * nt + 2 * t) = make_float2(acc[mi][nt][0] * scale, acc[mi][nt][1] * scale);
            if (r1 < S) *reinterpret_cast<float2*>(qo + (long)r1 * lddq + 8 * nt + 2 * t) = make_float2(acc[mi][nt][2] * scale, acc[mi][nt][3] * scale);
        }
    }
}

// ------------------------------------------------------------------------------------------- dK, dV
template <int HD>
struct DkvSmem {
    uint32_t qh[BT * (HD + 8)], ql[BT * (HD + 8)], qh2[BT * (HD + 4)];
    uint32_t dh[BT * (HD + 8)], dh2[BT * (HD + 4)];
    float lse[BT], dsum[BT];
};

template <int HD>
__global__ void __launch_bounds__(NT) dkv_kernel(const float* __restrict__ q, long ldq, const float* __restrict__ k, long ldk,
                                                 const float* __restrict__ v, long ldv,
                                                 const unsigned char* __restrict__ mask,
                                                 const float* __restrict__ dout, long ldd, const float* __restrict__ lse,
                                                 const float* __restrict__ dsum, float* __restrict__ dk, long lddk,
                                                 float* __restrict__ dv, long lddv, int S, int H, float scale) {
    pdl_entry();
    extern __shared__ __align__(16) unsigned char raw[];
    DkvSmem<HD>& sm = *reinterpret_cast<DkvSmem<HD>*>(raw);
    const int b = blockIdx.z, h = blockIdx.y, warp = threadIdx.x / 32, lane = threadIdx.x % 32, g = lane / 4, t = lane % 4;
    constexpr int MI = Cfg<HD>::MI, WR = Cfg<HD>::WR, BQ = Cfg<HD>::BQ;
    const int row0 = blockIdx.x * BQ + WR * warp;   // key rows
    const float* qb = q + (long)b * S * ldq + h * HD;
    const float* kb_ = k + (long)b * S * ldk + h * HD;
    const float* vb = v + (long)b * S * ldv + h * HD;
    const float* db = dout + (long)b * S * ldd + h * HD;
    uint32_t kh[MI][HD / 8][4], kl[MI][HD / 8][4], vh[MI][HD / 8][4];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
        load_a<HD, true>(kb_, ldk, row0 + 16 * mi, S, scale, g, t, kh[mi], kl[mi]);
        load_a<HD, false>(vb, ldv, row0 + 16 * mi, S, 1.f, g, t, vh[mi], vh[mi]);
    }
    float av[MI][HD / 8][4] = {}, ak[MI][HD / 8][4] = {};
    const bool active = row0 < S;
    for (int t0 = 0; t0 < S; t0 += BT) {
        __syncthreads();
        {   // no register prefetch here: the accumulators of both gradients leave no room (the co-resident CTA
            // covers the staging latency)
            TileRegs<HD> rq, rd;
            tile_load<HD>(rq, qb, ldq, t0, S);
            tile_load<HD>(rd, db, ldd, t0, S);
            tile_store<HD, HD + 8, HD + 4>(rq, sm.qh, sm.ql, sm.qh2);
            tile_store<HD, HD + 8, HD + 4>(rd, sm.dh, nullptr, sm.dh2);
        }
        for (int e = threadIdx.x; e < BT; e += NT) {
            const bool ok = t0 + e < S;
            sm.lse[e] = ok ? __ldg(lse + ((long)b * H + h) * S + t0 + e) : 0.f;
            sm.dsum[e] = ok ? __ldg(dsum + ((long)b * H + h) * S + t0 + e) : 0.f;
        }
        __syncthreads();
        if (!active) continue;
#pragma unroll
        for (int half = 0; half < BT / HT; ++half) {
            const int qb0 = HT * half;
            if (t0 + qb0 >= S) break;
            float s[MI][HT / 8][4] = {}, dp[MI][HT / 8][4] = {};
            gemm_abt3<HD, HD + 8, MI>(s, kh, kl, sm.qh, sm.ql, qb0, g, t);     // S^T[key, query]
            gemm_abt1<HD, HD + 8, MI>(dp, vh, sm.dh, qb0, g, t);               // dP^T[key, query] = V dO^T
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
                const int r0 = row0 + 16 * mi + g, r1 = r0 + 8;
#pragma unroll
                for (int nt = 0; nt < HT / 8; ++nt)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int qi_l = qb0 + 8 * nt + 2 * t + j, qi = t0 + qi_l;
                        const bool in = qi < S;
                        const bool ok0 = in && r0 < S && !(mask && __ldg(mask + (long)qi * S + r0));
                        const bool ok1 = in && r1 < S && !(mask && __ldg(mask + (long)qi * S + r1));
                        const float Lq = sm.lse[qi_l], Dq = sm.dsum[qi_l];
                        const float p0 = ok0 ? __expf(s[mi][nt][j] - Lq) : 0.f;
                        const float p1 = ok1 ? __expf(s[mi][nt][2 + j] - Lq) : 0.f;
                        s[mi][nt][j] = p0;
                        s[mi][nt][2 + j] = p1;
                        dp[mi][nt][j] = p0 * (dp[mi][nt][j] - Dq);
                        dp[mi][nt][2 + j] = p1 * (dp[mi][nt][2 + j] - Dq);
                    }
            }
            gemm_pb1<HD, HD + 4, MI>(av, s, sm.dh2, qb0, g, t);    // dV += P^T dO
            gemm_pb1<HD, HD + 4, MI>(ak, dp, sm.qh2, qb0, g, t);   // dK += dS^T Q
        }
    }
    if (!active) return;
    float* ko = dk + (long)b * S * lddk + h * HD;
    float* vo = dv + (long)b * S * lddv + h * HD;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
        const int r0 = row0 + 16 * mi + g, r1 = r0 + 8;
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt) {
            if (r0 < S) {
                *reinterpret_cast<float2*>(vo + (long)r0 * lddv + 8 * nt + 2 * t) = make_float2(av[mi][nt][0], av[mi][nt][1]);
                *reinterpret_cast<float2*>(ko + (long)r0 * lddk + 8 * nt + 2 * t) = make_float2(ak[mi][nt][0] * scale, ak[mi][nt][1] * scale);
            }
            if (r1 < S) {
                *reinterpret_cast<float2*>(vo + (long)r1 * lddv + 8 * nt + 2 * t) = make_float2(av[mi][nt][2], av[mi][nt][3]);
                *reinterpret_cast<float2*>(ko + (long)r1 * lddk + 8 * nt + 2 * t) = make_float2(ak[mi][nt][2] * scale, ak[mi][nt][3] * scale);
            }
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
    launch_k(fwd_kernel<HD>, dim3(ceil_div(S, Cfg<HD>::BQ), H, B), NT, sizeof(FwdSmem<HD>), st, q, ldq, k, ldk, v, ldv, mask, o, ldo, lse,
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
    const dim3 grid(ceil_div(S, Cfg<HD>::BQ), H, B);
    launch_k(dq_kernel<HD>, grid, NT, sizeof(DqSmem<HD>), st, q, ldq, k, ldk, v, ldv, mask, o, ldo, dout, ldd, lse, dsum, dq, lddq,
                                                       S, H, scale);
    launch_k(dkv_kernel<HD>, grid, NT, sizeof(DkvSmem<HD>), st, q, ldq, k, ldk, v, ldv, mask, dout, ldd, lse, dsum, dk, lddk, dv,
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
