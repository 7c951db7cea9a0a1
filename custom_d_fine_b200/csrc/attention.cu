// Small-sequence multi-head attention core (softmax(q k^T * scale + mask) v), forward + backward.
//
// Replaces the explicit-softmax path nn.MultiheadAttention takes in the reference
// (hybrid_encoder.py:256,277 AIFI; dfine_decoder.py:200,239 decoder self-attention with the CDN
// block mask): no [B,h,S,S] probability tensor, no head-averaged weights.  S <= ~1600, head_dim
// 16..64, so the op is small next to the convolutions; the kernels stream K/V (or Q/dO) tiles of
// 128 rows through shared memory with an online softmax; one warp owns one query (key) at a time,
// lanes run across the tile's keys for the dot products and across head_dim for the P.V products.
//   q/k/v/o rows are addressed as  base + (b*S + s)*ld + h*HD  (strided views of the packed
//   in-projection output are consumed in place).  mask: uint8 [S,S], non-zero = blocked, or null.
#include "common.cuh"
#include <cstdlib>

namespace {

constexpr int NT = 128;          // 4 warps
constexpr int RPW = 8;           // rows (queries or keys) owned per warp
constexpr int ROWS = (NT / 32) * RPW;  // 32 rows per CTA
constexpr int KT = 128;          // streamed tile

template <int HD>
struct Smem {
    float a[KT][HD + 1];   // K   (fwd, dq)  |  Q   (dkv)
    float b[KT][HD + 1];   // V   (fwd, dq)  |  dO  (dkv)
    float acc[ROWS][HD];   // per-row output accumulators
    float acc2[ROWS][HD];  // dkv only: second accumulator (dK)
    float m[ROWS], l[ROWS];
    float p[NT / 32][KT];  // per-warp probabilities / ds
    float p2[NT / 32][KT];
    float lse[KT], dsum[KT];  // dkv: per-query stats of the streamed tile
};

template <int HD>
__device__ __forceinline__ void load_tile(float (*dst)[HD + 1], const float* __restrict__ src, long ld, int row0,
                                          int rows_total) {
    for (int e = threadIdx.x; e < KT * (HD / 4); e += NT) {
        const int r = e / (HD / 4), c4 = e % (HD / 4);
        float4 v = make_float4(0, 0, 0, 0);
        if (row0 + r < rows_total) v = __ldg(reinterpret_cast<const float4*>(src + (long)(row0 + r) * ld + c4 * 4));
        dst[r][c4 * 4 + 0] = v.x; dst[r][c4 * 4 + 1] = v.y; dst[r][c4 * 4 + 2] = v.z; dst[r][c4 * 4 + 3] = v.w;
    }
}

template <int HD>
__device__ __forceinline__ float dot_row(const float* reg, const float* row) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) s += reg[d] * row[d];
    return s;
}

// ------------------------------------------------------------------------------------------- forward
template <int HD>
__global__ void __launch_bounds__(NT) attn_fwd_kernel(const float* __restrict__ q, long ldq, const float* __restrict__ k,
                                                      long ldk, const float* __restrict__ v, long ldv,
                                                      const unsigned char* __restrict__ mask, float* __restrict__ o,
                                                      long ldo, float* __restrict__ lse, int S, int H, float scale) {
    pdl_entry();
    extern __shared__ __align__(16) unsigned char raw[];
    Smem<HD>& sm = *reinterpret_cast<Smem<HD>*>(raw);
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ROWS;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const float* qb = q + (long)b * S * ldq + h * HD;
    const float* kb = k + (long)b * S * ldk + h * HD;
    const float* vb = v + (long)b * S * ldv + h * HD;
    for (int e = threadIdx.x; e < ROWS * HD; e += NT) sm.acc[e / HD][e % HD] = 0.f;
    for (int e = threadIdx.x; e < ROWS; e += NT) { sm.m[e] = -INFINITY; sm.l[e] = 0.f; }
    for (int t0 = 0; t0 < S; t0 += KT) {
        __syncthreads();
        load_tile<HD>(sm.a, kb, ldk, t0, S);
        load_tile<HD>(sm.b, vb, ldv, t0, S);
        __syncthreads();
        for (int r = 0; r < RPW; ++r) {
            const int row = warp * RPW + r, qi = q0 + row;
            if (qi >= S) break;
            float qr[HD];
#pragma unroll
            for (int d = 0; d < HD; ++d) qr[d] = __ldg(qb + (long)qi * ldq + d) * scale;
            float sc[KT / 32], mx = -INFINITY;
#pragma unroll
            for (int t = 0; t < KT / 32; ++t) {
                const int j = lane + 32 * t, kj = t0 + j;
                float s = -INFINITY;
                if (kj < S && !(mask && mask[(long)qi * S + kj])) s = dot_row<HD>(qr, sm.a[j]);
                sc[t] = s;
                mx = fmaxf(mx, s);
            }
            mx = warp_max(mx);
            const float m_old = sm.m[row], m_new = fmaxf(m_old, mx);
            const float corr = (m_new == -INFINITY) ? 1.f : expf(m_old - m_new);
            float ps = 0.f;
#pragma unroll
            for (int t = 0; t < KT / 32; ++t) {
                const float p = (sc[t] == -INFINITY) ? 0.f : expf(sc[t] - m_new);
                sm.p[warp][lane + 32 * t] = p;
                ps += p;
            }
            ps = warp_sum(ps);
            __syncwarp();
            for (int d = lane; d < HD; d += 32) {
                float a = sm.acc[row][d] * corr;
                for (int j = 0; j < KT; ++j) a += sm.p[warp][j] * sm.b[j][d];
                sm.acc[row][d] = a;
            }
            if (lane == 0) { sm.m[row] = m_new; sm.l[row] = sm.l[row] * corr + ps; }
            __syncwarp();
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ROWS * HD; e += NT) {
        const int row = e / HD, d = e % HD, qi = q0 + row;
        if (qi < S) o[((long)b * S + qi) * ldo + h * HD + d] = sm.acc[row][d] / sm.l[row];
    }
    for (int e = threadIdx.x; e < ROWS; e += NT)
        if (q0 + e < S) lse[((long)b * H + h) * S + q0 + e] = sm.m[e] + logf(sm.l[e]);
}

// ------------------------------------------------------------------------------------------- dQ (+ D = rowsum(dO*O))
template <int HD>
__global__ void __launch_bounds__(NT) attn_bwd_dq_kernel(const float* __restrict__ q, long ldq,
                                                         const float* __restrict__ k, long ldk,
                                                         const float* __restrict__ v, long ldv,
                                                         const unsigned char* __restrict__ mask,
                                                         const float* __restrict__ o, long ldo,
                                                         const float* __restrict__ dout, long ldd,
                                                         const float* __restrict__ lse, float* __restrict__ dsum,
                                                         float* __restrict__ dq, long lddq, int S, int H, float scale) {
    pdl_entry();
    extern __shared__ __align__(16) unsigned char raw[];
    Smem<HD>& sm = *reinterpret_cast<Smem<HD>*>(raw);
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ROWS;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const float* qb = q + (long)b * S * ldq + h * HD;
    const float* kb = k + (long)b * S * ldk + h * HD;
    const float* vb = v + (long)b * S * ldv + h * HD;
    const float* ob = o + (long)b * S * ldo + h * HD;
    const float* db = dout + (long)b * S * ldd + h * HD;
    for (int e = threadIdx.x; e < ROWS * HD; e += NT) sm.acc[e / HD][e % HD] = 0.f;
    // D_i = sum_d dO_i[d] * O_i[d]
    for (int r = 0; r < RPW; ++r) {
        const int row = warp * RPW + r, qi = q0 + row;
        float s = 0.f;
        if (qi < S)
            for (int d = lane; d < HD; d += 32) s += __ldg(db + (long)qi * ldd + d) * __ldg(ob + (long)qi * ldo + d);
        s = warp_sum(s);
        if (lane == 0) {
            sm.m[row] = s;
            if (qi < S) dsum[((long)b * H + h) * S + qi] = s;
        }
    }
    for (int t0 = 0; t0 < S; t0 += KT) {
        __syncthreads();
        load_tile<HD>(sm.a, kb, ldk, t0, S);
        load_tile<HD>(sm.b, vb, ldv, t0, S);
        __syncthreads();
        for (int r = 0; r < RPW; ++r) {
            const int row = warp * RPW + r, qi = q0 + row;
            if (qi >= S) break;
            float qr[HD], dr[HD];
#pragma unroll
            for (int d = 0; d < HD; ++d) {
                qr[d] = __ldg(qb + (long)qi * ldq + d) * scale;
                dr[d] = __ldg(db + (long)qi * ldd + d);
            }
            const float L = __ldg(lse + ((long)b * H + h) * S + qi), Di = sm.m[row];
#pragma unroll
            for (int t = 0; t < KT / 32; ++t) {
                const int j = lane + 32 * t, kj = t0 + j;
                float ds = 0.f;
                if (kj < S && !(mask && mask[(long)qi * S + kj])) {
                    const float p = expf(dot_row<HD>(qr, sm.a[j]) - L);
                    ds = p * (dot_row<HD>(dr, sm.b[j]) - Di);
                }
                sm.p[warp][j] = ds;
            }
            __syncwarp();
            for (int d = lane; d < HD; d += 32) {
                float a = sm.acc[row][d];
                for (int j = 0; j < KT; ++j) a += sm.p[warp][j] * sm.a[j][d];
                sm.acc[row][d] = a;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ROWS * HD; e += NT) {
        const int row = e / HD, d = e % HD, qi = q0 + row;
        if (qi < S) dq[((long)b * S + qi) * lddq + h * HD + d] = sm.acc[row][d] * scale;
    }
}

// ------------------------------------------------------------------------------------------- dK, dV
template <int HD>
__global__ void __launch_bounds__(NT) attn_bwd_dkv_kernel(const float* __restrict__ q, long ldq,
                                                          const float* __restrict__ k, long ldk,
                                                          const float* __restrict__ v, long ldv,
                                                          const unsigned char* __restrict__ mask,
                                                          const float* __restrict__ dout, long ldd,
                                                          const float* __restrict__ lse,
                                                          const float* __restrict__ dsum, float* __restrict__ dk,
                                                          long lddk, float* __restrict__ dv, long lddv, int S, int H,
                                                          float scale) {
    pdl_entry();
    extern __shared__ __align__(16) unsigned char raw[];
    Smem<HD>& sm = *reinterpret_cast<Smem<HD>*>(raw);
    const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * ROWS;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const float* qb = q + (long)b * S * ldq + h * HD;
    const float* kb = k + (long)b * S * ldk + h * HD;
    const float* vb = v + (long)b * S * ldv + h * HD;
    const float* db = dout + (long)b * S * ldd + h * HD;
    for (int e = threadIdx.x; e < ROWS * HD; e += NT) { sm.acc[e / HD][e % HD] = 0.f; sm.acc2[e / HD][e % HD] = 0.f; }
    for (int t0 = 0; t0 < S; t0 += KT) {
        __syncthreads();
        load_tile<HD>(sm.a, qb, ldq, t0, S);
        load_tile<HD>(sm.b, db, ldd, t0, S);
        for (int e = threadIdx.x; e < KT; e += NT) {
            const bool ok = t0 + e < S;
            sm.lse[e] = ok ? __ldg(lse + ((long)b * H + h) * S + t0 + e) : 0.f;
            sm.dsum[e] = ok ? __ldg(dsum + ((long)b * H + h) * S + t0 + e) : 0.f;
        }
        __syncthreads();
        for (int r = 0; r < RPW; ++r) {
            const int row = warp * RPW + r, kj = k0 + row;
            if (kj >= S) break;
            float kr[HD], vr[HD];
#pragma unroll
            for (int d = 0; d < HD; ++d) {
                kr[d] = __ldg(kb + (long)kj * ldk + d) * scale;
                vr[d] = __ldg(vb + (long)kj * ldv + d);
            }
#pragma unroll
            for (int t = 0; t < KT / 32; ++t) {
                const int i = lane + 32 * t, qi = t0 + i;
                float p = 0.f, ds = 0.f;
                if (qi < S && !(mask && mask[(long)qi * S + kj])) {
                    p = expf(dot_row<HD>(kr, sm.a[i]) - sm.lse[i]);
                    ds = p * (dot_row<HD>(vr, sm.b[i]) - sm.dsum[i]);
                }
                sm.p[warp][i] = p;
                sm.p2[warp][i] = ds;
            }
            __syncwarp();
            for (int d = lane; d < HD; d += 32) {
                float a = sm.acc[row][d], a2 = sm.acc2[row][d];
                for (int i = 0; i < KT; ++i) {
                    a += sm.p[warp][i] * sm.b[i][d];
                    a2 += sm.p2[warp][i] * sm.a[i][d];
                }
                sm.acc[row][d] = a;
                sm.acc2[row][d] = a2;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ROWS * HD; e += NT) {
        const int row = e / HD, d = e % HD, kj = k0 + row;
        if (kj < S) {
            dv[((long)b * S + kj) * lddv + h * HD + d] = sm.acc[row][d];
            dk[((long)b * S + kj) * lddk + h * HD + d] = sm.acc2[row][d] * scale;
        }
    }
}

template <int HD, int WHICH>
int set_smem(const void* fn) {
    static bool done = false;   // once per kernel instantiation (also keeps the call out of graph captures)
    if (done) return 0;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<HD>));
    if (e != cudaSuccess) { dfine_set_error("attention: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    done = true;
    return 0;
}

#define ATTN_DISPATCH(HD_, ...)                                       \
    switch (HD_) {                                                    \
        case 16: { constexpr int HD = 16; __VA_ARGS__; break; }              \
        case 32: { constexpr int HD = 32; __VA_ARGS__; break; }              \
        case 48: { constexpr int HD = 48; __VA_ARGS__; break; }              \
        case 64: { constexpr int HD = 64; __VA_ARGS__; break; }              \
        default: dfine_set_error("attention: head_dim %d unsupported (16|32|48|64)", HD_); return -1; \
    }

bool use_mma() {
    static const bool on = [] { const char* e = getenv("DFINE_ATTN"); return !(e && e[0] == 's'); }();   // "simt" = CUDA cores
    return on;
}

}  // namespace

// tensor-core (mma.sync, 3xTF32) kernels of attention_mma.cu
int attn_mma_fwd(const float* q, long ldq, const float* k, long ldk, const float* v, long ldv, const unsigned char* mask,
                 float* o, long ldo, float* lse, int B, int S, int H, int head_dim, float scale, cudaStream_t st);
int attn_mma_bwd(const float* q, long ldq, const float* k, long ldk, const float* v, long ldv, const unsigned char* mask,
                 const float* o, long ldo, const float* dout, long ldd, const float* lse, float* dsum, float* dq, long lddq,
                 float* dk, long lddk, float* dv, long lddv, int B, int S, int H, int head_dim, float scale,
                 cudaStream_t st);

// o [B,S,*] (row stride ldo), lse [B,H,S].  scale = 1/sqrt(head_dim).  All row strides in elements,
// multiples of 4; base pointers 16-byte aligned.
DFINE_API int dfine_attn_fwd(const float* q, long ldq, const float* k, long ldk, const float* v, long ldv,
                             const unsigned char* mask, float* o, long ldo, float* lse, int B, int S, int H,
                             int head_dim, float scale, void* stream) {
    if (B * S * H == 0) return 0;
    DFINE_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 2 == 0,
                  "attn_fwd: row strides must be multiples of 4");
    if (use_mma() && (head_dim == 16 || head_dim == 32 || head_dim == 48 || head_dim == 64)) {
        int rc = attn_mma_fwd(q, ldq, k, ldk, v, ldv, mask, o, ldo, lse, B, S, H, head_dim, scale, (cudaStream_t)stream);
        if (rc) return rc;
        DFINE_LAUNCH_CHECK("attn_fwd(mma)");
        return 0;
    }
    dim3 grid(ceil_div(S, ROWS), H, B);
    ATTN_DISPATCH(head_dim, {
        int rc = set_smem<HD, 0>((const void*)attn_fwd_kernel<HD>);
        if (rc) return rc;
        launch_k(attn_fwd_kernel<HD>, grid, NT, sizeof(Smem<HD>), (cudaStream_t)stream, q, ldq, k, ldk, v, ldv, mask, o, ldo,
                                                                                  lse, S, H, scale);
    });
    DFINE_LAUNCH_CHECK("attn_fwd");
    return 0;
}

// dsum [B,H,S] scratch (written by the dQ kernel, read by the dK/dV kernel).
DFINE_API int dfine_attn_bwd(const float* q, long ldq, const float* k, long ldk, const float* v, long ldv,
                             const unsigned char* mask, const float* o, long ldo, const float* dout, long ldd,
                             const float* lse, float* dsum, float* dq, long lddq, float* dk, long lddk, float* dv,
                             long lddv, int B, int S, int H, int head_dim, float scale, void* stream) {
    if (B * S * H == 0) return 0;
    DFINE_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldd % 4 == 0,
                  "attn_bwd: row strides must be multiples of 4");
    if (use_mma() && (head_dim == 16 || head_dim == 32 || head_dim == 48 || head_dim == 64) && lddq % 2 == 0 &&
        lddk % 2 == 0 && lddv % 2 == 0) {
        int rc = attn_mma_bwd(q, ldq, k, ldk, v, ldv, mask, o, ldo, dout, ldd, lse, dsum, dq, lddq, dk, lddk, dv, lddv, B, S,
                              H, head_dim, scale, (cudaStream_t)stream);
        if (rc) return rc;
        DFINE_LAUNCH_CHECK("attn_bwd(mma)");
        return 0;
    }
    dim3 grid(ceil_div(S, ROWS), H, B);
    ATTN_DISPATCH(head_dim, {
        int rc = set_smem<HD, 1>((const void*)attn_bwd_dq_kernel<HD>);
        if (rc) return rc;
        rc = set_smem<HD, 2>((const void*)attn_bwd_dkv_kernel<HD>);
        if (rc) return rc;
        launch_k(attn_bwd_dq_kernel<HD>, grid, NT, sizeof(Smem<HD>), (cudaStream_t)stream, 
            q, ldq, k, ldk, v, ldv, mask, o, ldo, dout, ldd, lse, dsum, dq, lddq, S, H, scale);
        launch_k(attn_bwd_dkv_kernel<HD>, grid, NT, sizeof(Smem<HD>), (cudaStream_t)stream, 
            q, ldq, k, ldk, v, ldv, mask, dout, ldd, lse, dsum, dk, lddk, dv, lddv, S, H, scale);
    });
    DFINE_LAUNCH_CHECK("attn_bwd");
    return 0;
}
