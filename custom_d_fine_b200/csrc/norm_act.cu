// BatchNorm(+add)+activation(+LAB)(+residual) and LayerNorm kernels, NHWC / token-major fp32.
//
// Replaces the per-op ATen chains of the reference's ConvBNAct / ConvNormLayer blocks
// (hgnetv2.py:35-80 incl. LearnableAffineBlock 25-32, hybrid_encoder.py:22-45,106-121) and the
// nn.LayerNorm calls of AIFI / decoder layers.  All kernels are HBM-bound: 16-byte vector
// accesses along the contiguous channel dimension, per-channel reductions accumulated in fp32
// per thread over a short row run and merged in double (shared, then one global atomic per CTA
// and channel).
#include "common.cuh"

namespace {

constexpr int NT = 256;

struct RowMap {  // how the 256 threads of a CTA tile a [rows, C/4] slab of float4
    int VC;      // float4 per row
    int LPR;     // lanes per row  = min(VC, NT)
    int RPP;     // rows per pass  = NT / LPR
};
__host__ __device__ inline RowMap make_rowmap(int C) {
    RowMap m;
    m.VC = C / 4;
    m.LPR = m.VC < NT ? m.VC : NT;
    m.RPP = NT / m.LPR;
    return m;
}

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ---------------------------------------------------------------------------------------------
// BatchNorm statistics: stats[0:C] += sum(x), stats[C:2C] += sum(x^2)      (double accumulators)
// ---------------------------------------------------------------------------------------------
// Partial sums are merged without atomics: thread (row lane, channel group) owns the slot part[row lane][channel]
// of a shared fp32 table, the CTA then sums the <= 256/LPR row lanes in double and issues ONE global fp64 atomic
// per channel.  (Shared-memory fp64 atomicAdd is a CAS loop on sm_100: with 256 threads hitting 2*C addresses it
// cost more than the HBM pass itself.)
__global__ void __launch_bounds__(NT) bn_stats_kernel(const float* __restrict__ x, double* __restrict__ stats,
                                                      long M, int C, long rows_per_cta) {
    pdl_entry();
    extern __shared__ float part[];  // [RPP][2*C]
    const RowMap m = make_rowmap(C);
    const long r0 = (long)blockIdx.x * rows_per_cta;
    const long r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
    const int lane_r = threadIdx.x / m.LPR, lane_c = threadIdx.x % m.LPR;
    if (lane_r < m.RPP) {
        float* mine = part + (size_t)lane_r * 2 * C;
        for (int cv = lane_c; cv < m.VC; cv += m.LPR) {
            float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
            for (long r = r0 + lane_r; r < r1; r += m.RPP) {
                const float4 v = ld4(x + r * C + cv * 4);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                q.x += v.x * v.x; q.y += v.y * v.y; q.z += v.z * v.z; q.w += v.w * v.w;
            }
            st4(mine + cv * 4, s);
            st4(mine + C + cv * 4, q);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += NT) {
        double a = 0.0;
        for (int r = 0; r < m.RPP; ++r) a += (double)part[(size_t)r * 2 * C + i];
        atomicAdd(&stats[i], a);
    }
}

// mean / invstd / folded scale+shift / running-stat update (momentum, unbiased var) — C threads.
__global__ void bn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ weight,
                                   const float* __restrict__ bias, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out, float* __restrict__ scale_out,
                                   float* __restrict__ shift_out, long M, int C, float momentum, float eps) {
    pdl_entry();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    bn_finalize_channel(stats[c], stats[C + c], c, weight, bias, running_mean, running_var, mean_out, invstd_out, scale_out,
                        shift_out, M, momentum, eps);
}

// scale/shift from running statistics (eval BatchNorm / FrozenBatchNorm2d, common.py:58-67)
__global__ void bn_fold_kernel(const float* __restrict__ weight, const float* __restrict__ bias,
                               const float* __restrict__ running_mean, const float* __restrict__ running_var,
                               float* __restrict__ scale_out, float* __restrict__ shift_out, int C, float eps) {
    pdl_entry();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float sc = weight[c] * rsqrtf(running_var[c] + eps);
    scale_out[c] = sc;
    shift_out[c] = bias[c] - running_mean[c] * sc;
}

// y = lab_s * act(x*scale[c] + shift[c] + pre_add) + lab_b + post_add
__global__ void __launch_bounds__(NT) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                      const float* __restrict__ shift,
                                                      const float* __restrict__ pre_add,
                                                      const float* __restrict__ post_add,
                                                      const float* __restrict__ lab, const float* __restrict__ lab_b,
                                                      float* __restrict__ y, long n4, int VC, int act, long ldy,
                                                      long ld_post) {
    pdl_entry();
    const float ls = lab ? __ldg(lab) : 1.f, lb = lab ? __ldg(lab_b) : 0.f;
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < n4; i += (long)gridDim.x * NT) {
        const int c = (int)(i % VC) * 4;
        const float4 v = ld4(x + i * 4), sc = ld4(scale + c), sh = ld4(shift + c);
        float4 z = make_float4(v.x * sc.x + sh.x, v.y * sc.y + sh.y, v.z * sc.z + sh.z, v.w * sc.w + sh.w);
        if (pre_add) { const float4 a = ld4(pre_add + i * 4); z.x += a.x; z.y += a.y; z.z += a.z; z.w += a.w; }
        float4 o = make_float4(act_fwd(z.x, act), act_fwd(z.y, act), act_fwd(z.z, act), act_fwd(z.w, act));
        if (lab) { o.x = ls * o.x + lb; o.y = ls * o.y + lb; o.z = ls * o.z + lb; o.w = ls * o.w + lb; }
        if (post_add) { const float4 a = ld4(post_add + (i / VC) * ld_post + c); o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
        st4(y + (i / VC) * ldy + c, o);
    }
}

// Register budget of the reduce kernel: four rows in flight per thread cost 97 registers = 2 CTAs per SM, and the
// 4-CTAs-per-SM grid of pick_rows_per_cta then ran in TWO half-occupied waves (ncu: 12-25 % warps active).  Two rows per
// thread fit 64 registers: four resident CTAs per SM, the grid is one wave.
#ifndef BN_RED_UNROLL
#define BN_RED_UNROLL 2
#endif
#ifndef BN_RED_CTAS
#define BN_RED_CTAS 4
#endif
// Backward pass 1: per-channel sums of dz and dz*xhat (double), LAB scalar grads.
//   dz = dy * lab_s * act'(z),  z = x*scale+shift (+pre_add),  xhat = (x-mean)*invstd
// red[0:C] += sum dz ; red[C:2C] += sum dz*xhat ; red[2C] += sum dy*act(z) ; red[2C+1] += sum dy
__global__ void __launch_bounds__(NT, BN_RED_CTAS) bn_bwd_reduce_kernel(
    const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ scale,
    const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
    const float* __restrict__ pre_add, const float* __restrict__ lab, double* __restrict__ red, long M, int C,
    long rows_per_cta, int act, long ld_dy) {
    pdl_entry();
    extern __shared__ float part[];  // [RPP][2*C] partial sums + [NT/32][2] LAB partials (see bn_stats_kernel)
    const float ls = lab ? __ldg(lab) : 1.f;
    const RowMap m = make_rowmap(C);
    float* labp = part + (size_t)m.RPP * 2 * C;
    const long r0 = (long)blockIdx.x * rows_per_cta;
    const long r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
    const int lane_r = threadIdx.x / m.LPR, lane_c = threadIdx.x % m.LPR;
    float lab_gs = 0.f, lab_gb = 0.f;
    if (lane_r < m.RPP) {
        float* mine = part + (size_t)lane_r * 2 * C;
        for (int cv = lane_c; cv < m.VC; cv += m.LPR) {
            const int c = cv * 4;
            const float4 sc = ld4(scale + c), sf = ld4(shift + c), mu = ld4(mean + c), is = ld4(invstd + c);
            float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
            // rows are walked four at a time so that eight independent 16-byte loads are in flight per thread
            auto body = [&](const float4 g, const float4 v, const long o) {
                float4 z = make_float4(v.x * sc.x + sf.x, v.y * sc.y + sf.y, v.z * sc.z + sf.z, v.w * sc.w + sf.w);
                if (pre_add) { const float4 a = ld4(pre_add + o); z.x += a.x; z.y += a.y; z.z += a.z; z.w += a.w; }
                if (lab) {
                    lab_gs += g.x * act_fwd(z.x, act) + g.y * act_fwd(z.y, act) + g.z * act_fwd(z.z, act) +
                              g.w * act_fwd(z.w, act);
                    lab_gb += g.x + g.y + g.z + g.w;
                }
                const float4 dz = make_float4(g.x * ls * act_bwd(z.x, act), g.y * ls * act_bwd(z.y, act),
                                              g.z * ls * act_bwd(z.z, act), g.w * ls * act_bwd(z.w, act));
                s.x += dz.x; s.y += dz.y; s.z += dz.z; s.w += dz.w;
                q.x += dz.x * (v.x - mu.x) * is.x; q.y += dz.y * (v.y - mu.y) * is.y;
                q.z += dz.z * (v.z - mu.z) * is.z; q.w += dz.w * (v.w - mu.w) * is.w;
            };
            long r = r0 + lane_r;
            constexpr int U = BN_RED_UNROLL;
            for (; r + (long)(U - 1) * m.RPP < r1; r += (long)U * m.RPP) {
                float4 g[U], v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    g[u] = ld4(dy + (r + (long)u * m.RPP) * ld_dy + c);
                    v[u] = ld4(x + (r + (long)u * m.RPP) * C + c);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) body(g[u], v[u], (r + (long)u * m.RPP) * C + c);
            }
            for (; r < r1; r += m.RPP) body(ld4(dy + r * ld_dy + c), ld4(x + r * C + c), r * C + c);
            st4(mine + c, s);
            st4(mine + C + c, q);
        }
    }
    if (lab) {
        lab_gs = warp_sum(lab_gs);
        lab_gb = warp_sum(lab_gb);
        if ((threadIdx.x & 31) == 0) { labp[2 * (threadIdx.x / 32)] = lab_gs; labp[2 * (threadIdx.x / 32) + 1] = lab_gb; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += NT) {
        double a = 0.0;
        for (int r = 0; r < m.RPP; ++r) a += (double)part[(size_t)r * 2 * C + i];
        atomicAdd(&red[i], a);
    }
    if (lab && threadIdx.x < 2) {
        double a = 0.0;
        for (int w = 0; w < NT / 32; ++w) a += (double)labp[2 * w + threadIdx.x];
        atomicAdd(&red[2 * C + threadIdx.x], a);
    }
}

// Backward of a frozen / eval-mode BatchNorm + ReLU that was folded into the conv epilogue (y = act(conv * scale + shift),
// no pre-activation kept): dconv = dy * act'(y) * scale, with act' read off the OUTPUT (ReLU: y > 0; none: 1).
__global__ void __launch_bounds__(NT) frozen_bn_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                           const float* __restrict__ scale, float* __restrict__ dx, long n4,
                                                           int VC, int relu, long ld_dy, long ld_y, long ld_dx) {
    pdl_entry();
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < n4; i += (long)gridDim.x * NT) {
        const int c = (int)(i % VC) * 4;
        const long r = i / VC;
        const float4 g = ld4(dy + r * ld_dy + c), sc = ld4(scale + c);
        float4 o = make_float4(g.x * sc.x, g.y * sc.y, g.z * sc.z, g.w * sc.w);
        if (relu) {
            const float4 v = ld4(y + r * ld_y + c);
            o.x = v.x > 0.f ? o.x : 0.f; o.y = v.y > 0.f ? o.y : 0.f; o.z = v.z > 0.f ? o.z : 0.f; o.w = v.w > 0.f ? o.w : 0.f;
        }
        st4(dx + r * ld_dx + c, o);
    }
}

// Backward pass 2.  training: dx = scale*(dz - sum_dz/M - xhat*sum_dzx/M) ; else dx = scale*dz.
// Optionally also writes dz (= gradient of pre_add).
__global__ void __launch_bounds__(NT, 8) bn_bwd_apply_kernel(
    const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ scale,
    const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
    const float* __restrict__ pre_add, const float* __restrict__ lab, const double* __restrict__ red,
    float* __restrict__ dx, float* __restrict__ dpre, long n4, int VC, long M, int act, int training,
    float* __restrict__ g_w, float* __restrict__ g_b, float* __restrict__ g_lab_s, float* __restrict__ g_lab_b,
    long ld_dy, long ld_dx) {
    pdl_entry();
    const float ls = lab ? __ldg(lab) : 1.f;
    const int C = VC * 4;
    const double invM = 1.0 / (double)M;
    extern __shared__ float mean_terms[];   // [2*C]: sum(dz)/M and sum(dz*xhat)/M as floats (one fp64 multiply per channel)
    if (training) {
        for (int c = threadIdx.x; c < 2 * C; c += NT) mean_terms[c] = (float)(red[c] * invM);
        __syncthreads();
    }
    if (blockIdx.x == 0) {
        // parameter gradients straight into the (flat-arena) .grad tensors: d(bn.weight) = sum dz*xhat,
        // d(bn.bias) = sum dz, LAB scalars; `red` is complete (written by the previous launch)
        if (g_w)
            for (int c = threadIdx.x; c < C; c += NT) { g_w[c] += (float)red[C + c]; g_b[c] += (float)red[c]; }
        if (g_lab_s && threadIdx.x == 0) { g_lab_s[0] += (float)red[2 * C]; g_lab_b[0] += (float)red[2 * C + 1]; }
    }
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < n4; i += (long)gridDim.x * NT) {
        const int c = (int)(i % VC) * 4;
        const float4 g = ld4(dy + (i / VC) * ld_dy + c), v = ld4(x + i * 4), sc = ld4(scale + c), sf = ld4(shift + c);
        float z[4] = {v.x * sc.x + sf.x, v.y * sc.y + sf.y, v.z * sc.z + sf.z, v.w * sc.w + sf.w};
        if (pre_add) { const float4 a = ld4(pre_add + i * 4); z[0] += a.x; z[1] += a.y; z[2] += a.z; z[3] += a.w; }
        const float gg[4] = {g.x, g.y, g.z, g.w}, vv[4] = {v.x, v.y, v.z, v.w}, ss[4] = {sc.x, sc.y, sc.z, sc.w};
        float dz[4], o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) dz[k] = gg[k] * ls * act_bwd(z[k], act);
        if (training) {
            const float4 mu = ld4(mean + c), is = ld4(invstd + c);
            const float mm[4] = {mu.x, mu.y, mu.z, mu.w}, ii[4] = {is.x, is.y, is.z, is.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float m1 = mean_terms[c + k], m2 = mean_terms[C + c + k];
                o[k] = ss[k] * (dz[k] - m1 - (vv[k] - mm[k]) * ii[k] * m2);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = ss[k] * dz[k];
        }
        st4(dx + (i / VC) * ld_dx + c, make_float4(o[0], o[1], o[2], o[3]));
        if (dpre) st4(dpre + i * 4, make_float4(dz[0], dz[1], dz[2], dz[3]));
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the last dim D (multiple of 4, <= 1024): one warp per row.
// ---------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 8;  // float4 per lane -> D <= 1024

__global__ void __launch_bounds__(NT) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                           const float* __restrict__ w, const float* __restrict__ b,
                                                           float* __restrict__ y, float* __restrict__ mean_out,
                                                           float* __restrict__ rstd_out, long rows, int D, float eps) {
    pdl_entry();
    const long row = (long)blockIdx.x * (NT / 32) + threadIdx.x / 32;
    if (row >= rows) return;
    const int lane = threadIdx.x & 31, VD = D / 4;
    float4 v[LN_MAXV];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAXV; ++k) {
        const int i = lane + k * 32;
        if (i < VD) {
            v[k] = ld4(x + row * D + i * 4);
            if (res) { const float4 r = ld4(res + row * D + i * 4); v[k].x += r.x; v[k].y += r.y; v[k].z += r.z; v[k].w += r.w; }
            s += v[k].x + v[k].y + v[k].z + v[k].w;
        }
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAXV; ++k) {
        const int i = lane + k * 32;
        if (i < VD) {
            const float a = v[k].x - mean, bb = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
            q += a * a + bb * bb + c * c + d * d;
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
#pragma unroll
    for (int k = 0; k < LN_MAXV; ++k) {
        const int i = lane + k * 32;
        if (i < VD) {
            const float4 ww = ld4(w + i * 4), bb = ld4(b + i * 4);
            float4 o;
            o.x = (v[k].x - mean) * rstd * ww.x + bb.x; o.y = (v[k].y - mean) * rstd * ww.y + bb.y;
            o.z = (v[k].z - mean) * rstd * ww.z + bb.z; o.w = (v[k].w - mean) * rstd * ww.w + bb.w;
            st4(y + row * D + i * 4, o);
        }
    }
}

// dx = rstd*(g - mean(g) - xhat*mean(g*xhat)), g = dy*w ; dw += dy*xhat ; db += dy (fp32 atomics,
// pre-reduced over the CTA's 8 rows x ROWS_PER_WARP in shared memory).  x here is the LN *input*
// (x + res when the forward fused a residual; the caller passes that sum's two terms again).
constexpr int LN_ROWS_PER_WARP = 8;
template <int MAXV>
__global__ void __launch_bounds__(NT) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           const float* __restrict__ res,
                                                           const float* __restrict__ w, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, float* __restrict__ dx,
                                                           float* __restrict__ dw, float* __restrict__ db, long rows,
                                                           int D) {
    pdl_entry();
    extern __shared__ float shf[];  // [2*D]
    for (int i = threadIdx.x; i < 2 * D; i += NT) shf[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x / 32, VD = D / 4;
    float4 aw[MAXV], ab[MAXV];
#pragma unroll
    for (int k = 0; k < MAXV; ++k) { aw[k] = make_float4(0, 0, 0, 0); ab[k] = make_float4(0, 0, 0, 0); }
    // grid-stride over slabs of (NT/32) x LN_ROWS_PER_WARP rows: the grid is capped (a few CTAs per SM), so the
    // parameter-gradient atomics below run once per CTA, not once per 64 rows (the 134 400-row encoder LayerNorm
    // issued 1.07 M global atomics on 512 addresses: 282 us for a 100 MB pass)
    const long n_slabs = (rows + (NT / 32) * LN_ROWS_PER_WARP - 1) / ((NT / 32) * LN_ROWS_PER_WARP);
    for (long slab = blockIdx.x; slab < n_slabs; slab += gridDim.x)
    for (int rr = 0; rr < LN_ROWS_PER_WARP; ++rr) {
        const long row = (slab * (NT / 32) + warp) * LN_ROWS_PER_WARP + rr;
        if (row >= rows) break;
        const float mu = mean[row], rs = rstd[row];
        float4 xh[MAXV], g[MAXV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int i = lane + k * 32;
            if (i < VD) {
                float4 v = ld4(x + row * D + i * 4);
                if (res) { const float4 r = ld4(res + row * D + i * 4); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
                const float4 d = ld4(dy + row * D + i * 4), ww = ld4(w + i * 4);
                xh[k] = make_float4((v.x - mu) * rs, (v.y - mu) * rs, (v.z - mu) * rs, (v.w - mu) * rs);
                g[k] = make_float4(d.x * ww.x, d.y * ww.y, d.z * ww.z, d.w * ww.w);
                s1 += g[k].x + g[k].y + g[k].z + g[k].w;
                s2 += g[k].x * xh[k].x + g[k].y * xh[k].y + g[k].z * xh[k].z + g[k].w * xh[k].w;
                aw[k].x += d.x * xh[k].x; aw[k].y += d.y * xh[k].y; aw[k].z += d.z * xh[k].z; aw[k].w += d.w * xh[k].w;
                ab[k].x += d.x; ab[k].y += d.y; ab[k].z += d.z; ab[k].w += d.w;
            }
        }
        s1 = warp_sum(s1) / (float)D;
        s2 = warp_sum(s2) / (float)D;
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int i = lane + k * 32;
            if (i < VD) {
                float4 o;
                o.x = rs * (g[k].x - s1 - xh[k].x * s2); o.y = rs * (g[k].y - s1 - xh[k].y * s2);
                o.z = rs * (g[k].z - s1 - xh[k].z * s2); o.w = rs * (g[k].w - s1 - xh[k].w * s2);
                st4(dx + row * D + i * 4, o);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int i = lane + k * 32;
        if (i < VD) {
            atomicAdd(&shf[i * 4 + 0], aw[k].x); atomicAdd(&shf[i * 4 + 1], aw[k].y);
            atomicAdd(&shf[i * 4 + 2], aw[k].z); atomicAdd(&shf[i * 4 + 3], aw[k].w);
            atomicAdd(&shf[D + i * 4 + 0], ab[k].x); atomicAdd(&shf[D + i * 4 + 1], ab[k].y);
            atomicAdd(&shf[D + i * 4 + 2], ab[k].z); atomicAdd(&shf[D + i * 4 + 3], ab[k].w);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += NT) { atomicAdd(&dw[i], shf[i]); atomicAdd(&db[i], shf[D + i]); }
}

inline long pick_rows_per_cta(long M, int C) {
    // ~4 CTAs per SM worth of slabs, at least 8 rows per thread (every CTA ends with one global fp64 atomic per channel
    // sum).  Measured on the D-FINE-m step: 2 per SM makes the reduce kernels 7 % faster in isolation and the step 2 %
    // slower (the weight-gradient stream fills the idle SMs); DFINE_BN_CTAS_PER_SM overrides.
    static const long per_sm = [] { const char* e = getenv("DFINE_BN_CTAS_PER_SM"); const long v = e ? atol(e) : 4; return v > 0 ? v : 4; }();
    const RowMap m = make_rowmap(C);
    long target = 148L * per_sm;
    long rpc = (M + target - 1) / target;
    if (rpc < m.RPP * 8) rpc = m.RPP * 8;
    return rpc;
}
}  // namespace

// stats: double [2*C], zero-initialised by the caller.
DFINE_API int dfine_bn_stats(const float* x, double* stats, long M, int C, void* stream) {
    DFINE_REQUIRE(C % 4 == 0 && C > 0 && C <= 3064, "bn_stats: C=%d must be a multiple of 4 (<=4096)", C);
    if (M == 0) return 0;
    const long rpc = pick_rows_per_cta(M, C);
    const RowMap rm = make_rowmap(C);
    launch_k(bn_stats_kernel, ceil_div(M, rpc), NT, (size_t)rm.RPP * 2 * C * sizeof(float), (cudaStream_t)stream, x, stats, M, C, rpc);
    DFINE_LAUNCH_CHECK("bn_stats");
    return 0;
}

DFINE_API int dfine_bn_finalize(const double* stats, const float* weight, const float* bias, float* running_mean,
                                float* running_var, float* mean, float* invstd, float* scale, float* shift, long M,
                                int C, float momentum, float eps, void* stream) {
    launch_k(bn_finalize_kernel, ceil_div(C, 128), 128, 0, (cudaStream_t)stream, stats, weight, bias, running_mean,
                                                                          running_var, mean, invstd, scale, shift, M,
                                                                          C, momentum, eps);
    DFINE_LAUNCH_CHECK("bn_finalize");
    return 0;
}

DFINE_API int dfine_bn_fold(const float* weight, const float* bias, const float* running_mean,
                            const float* running_var, float* scale, float* shift, int C, float eps, void* stream) {
    launch_k(bn_fold_kernel, ceil_div(C, 128), 128, 0, (cudaStream_t)stream, weight, bias, running_mean, running_var, scale,
                                                                      shift, C, eps);
    DFINE_LAUNCH_CHECK("bn_fold");
    return 0;
}

// lab / lab_b: device pointers to the scalar LAB scale and bias (both or neither).  pre_add/post_add optional [M,C].
DFINE_API int dfine_bn_apply(const float* x, const float* scale, const float* shift, const float* pre_add,
                             const float* post_add, const float* lab, const float* lab_b, float* y, long M, int C,
                             int act, long ldy, long ld_post, void* stream) {
    DFINE_REQUIRE(C % 4 == 0, "bn_apply: C=%d must be a multiple of 4", C);
    DFINE_REQUIRE(ldy >= C && ldy % 4 == 0 && ((uintptr_t)y % 16) == 0, "bn_apply: output row stride %ld", ldy);
    DFINE_REQUIRE(!post_add || (ld_post >= C && ld_post % 4 == 0 && ((uintptr_t)post_add % 16) == 0),
                  "bn_apply: post_add row stride %ld", ld_post);
    const long n4 = M * C / 4;
    if (n4 == 0) return 0;
    launch_k(bn_apply_kernel, ew_grid_k(bn_apply_kernel, n4, NT, 0), NT, 0, (cudaStream_t)stream, x, scale, shift, pre_add, post_add, lab, lab_b, y,
                                                                 n4, C / 4, act, ldy, post_add ? ld_post : (long)C);
    DFINE_LAUNCH_CHECK("bn_apply");
    return 0;
}

// red: double [2*C+2], zero-initialised by the caller.
DFINE_API int dfine_bn_bwd_reduce(const float* dy, const float* x, const float* scale, const float* shift,
                                  const float* mean, const float* invstd, const float* pre_add, const float* lab,
                                  double* red, long M, int C, int act, long ld_dy, void* stream) {
    DFINE_REQUIRE(C % 4 == 0 && C > 0 && C <= 3064, "bn_bwd_reduce: C=%d", C);
    DFINE_REQUIRE(ld_dy >= C && ld_dy % 4 == 0 && ((uintptr_t)dy % 16) == 0, "bn_bwd_reduce: dy row stride %ld", ld_dy);
    if (M == 0) return 0;
    const long rpc = pick_rows_per_cta(M, C);
    const RowMap rm = make_rowmap(C);
    const size_t smem = ((size_t)rm.RPP * 2 * C + 2 * (NT / 32)) * sizeof(float);
    launch_k(bn_bwd_reduce_kernel, ceil_div(M, rpc), NT, smem, (cudaStream_t)stream, 
        dy, x, scale, shift, mean, invstd, pre_add, lab, red, M, C, rpc, act, ld_dy);
    DFINE_LAUNCH_CHECK("bn_bwd_reduce");
    return 0;
}

DFINE_API int dfine_bn_bwd_apply(const float* dy, const float* x, const float* scale, const float* shift,
                                 const float* mean, const float* invstd, const float* pre_add, const float* lab,
                                 const double* red, float* dx, float* dpre, long M, int C, int act, int training,
                                 float* g_w, float* g_b, float* g_lab_s, float* g_lab_b, long ld_dy, long ld_dx,
                                 void* stream) {
    DFINE_REQUIRE(C % 4 == 0, "bn_bwd_apply: C=%d", C);
    DFINE_REQUIRE(ld_dx >= C && ld_dx % 4 == 0 && ((uintptr_t)dx % 16) == 0, "bn_bwd_apply: dx row stride %ld", ld_dx);
    DFINE_REQUIRE(ld_dy >= C && ld_dy % 4 == 0 && ((uintptr_t)dy % 16) == 0, "bn_bwd_apply: dy row stride %ld", ld_dy);
    DFINE_REQUIRE((g_w == nullptr) == (g_b == nullptr) && (g_lab_s == nullptr) == (g_lab_b == nullptr),
                  "bn_bwd_apply: gradient outputs come in pairs");
    const long n4 = M * C / 4;
    if (n4 == 0) return 0;
    launch_k(bn_bwd_apply_kernel, ew_grid_k(bn_bwd_apply_kernel, n4, NT, 2 * C * sizeof(float)), NT, 2 * C * sizeof(float), (cudaStream_t)stream, dy, x, scale, shift, mean, invstd, pre_add, lab,
                                                                     red, dx, dpre, n4, C / 4, M, act, training, g_w,
                                                                     g_b, g_lab_s, g_lab_b, ld_dy, ld_dx);
    DFINE_LAUNCH_CHECK("bn_bwd_apply");
    return 0;
}

// dx[M,C] (row stride ld_dx) = dy * act'(y) * scale[c] for a BatchNorm with fixed statistics folded into the producing conv's
// epilogue (dfine_conv_tc_f16x3's ch_scale): act = DFINE_ACT_NONE or DFINE_ACT_RELU (the derivative is read off the output).
DFINE_API int dfine_frozen_bn_bwd(const float* dy, const float* y, const float* scale, float* dx, long M, int C, int act,
                                  long ld_dy, long ld_y, long ld_dx, void* stream) {
    DFINE_REQUIRE(C % 4 == 0 && (act == 0 || act == 1), "frozen_bn_bwd: C=%d act=%d", C, act);
    DFINE_REQUIRE(ld_dy >= C && ld_dy % 4 == 0 && ld_y >= C && ld_y % 4 == 0 && ld_dx >= C && ld_dx % 4 == 0 &&
                      ((uintptr_t)dy % 16) == 0 && ((uintptr_t)y % 16) == 0 && ((uintptr_t)dx % 16) == 0,
                  "frozen_bn_bwd: strides / alignment");
    const long n4 = M * C / 4;
    if (n4 == 0) return 0;
    launch_k(frozen_bn_bwd_kernel, ew_grid_k(frozen_bn_bwd_kernel, n4, NT, 0), NT, 0, (cudaStream_t)stream, dy, y, scale, dx, n4, C / 4, act == 1, ld_dy, ld_y, ld_dx);
    DFINE_LAUNCH_CHECK("frozen_bn_bwd");
    return 0;
}

// y = LayerNorm(x (+ res)) * w + b ; saves mean / rstd per row.
DFINE_API int dfine_layernorm_fwd(const float* x, const float* res, const float* w, const float* b, float* y,
                                  float* mean, float* rstd, long rows, int D, float eps, void* stream) {
    DFINE_REQUIRE(D % 4 == 0 && D <= 1024, "layernorm: D=%d must be a multiple of 4 and <= 1024", D);
    if (rows == 0) return 0;
    launch_k(layernorm_fwd_kernel, ceil_div(rows, NT / 32), NT, 0, (cudaStream_t)stream, x, res, w, b, y, mean, rstd, rows,
                                                                                  D, eps);
    DFINE_LAUNCH_CHECK("layernorm_fwd");
    return 0;
}

// dw/db fp32 [D], zero-initialised by the caller.  dx is the gradient of (x + res).
DFINE_API int dfine_layernorm_bwd(const float* dy, const float* x, const float* res, const float* w,
                                  const float* mean, const float* rstd, float* dx, float* dw, float* db, long rows,
                                  int D, void* stream) {
    DFINE_REQUIRE(D % 4 == 0 && D <= 1024, "layernorm_bwd: D=%d", D);
    if (rows == 0) return 0;
    const long per_cta = (long)(NT / 32) * LN_ROWS_PER_WARP;
    const long slabs = (rows + per_cta - 1) / per_cta;
    // float4 per lane as a template parameter: at D = 256 the register arrays shrink 4x (170 -> ~60 registers) and
    // four times as many rows are in flight per SM (the 134 400-row encoder LayerNorm ran at 1.6 TB/s)
    const int grid = (int)(slabs < 148L * 8 ? slabs : 148L * 8);
    if (D <= 256)
        launch_k(layernorm_bwd_kernel<2>, grid, NT, 2 * D * sizeof(float), (cudaStream_t)stream, dy, x, res, w, mean, rstd, dx,
                                                                                          dw, db, rows, D);
    else
        launch_k(layernorm_bwd_kernel<LN_MAXV>, grid, NT, 2 * D * sizeof(float), (cudaStream_t)stream, dy, x, res, w, mean, rstd,
                                                                                                dx, dw, db, rows, D);
    DFINE_LAUNCH_CHECK("layernorm_bwd");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Stand-alone activation (used where the pre-activation must be kept for the backward: GELU FFN)
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(NT) act_fwd_kernel(const float* __restrict__ z, float* __restrict__ y, long n4, int act) {
    pdl_entry();
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < n4; i += (long)gridDim.x * NT) {
        const float4 v = ld4(z + i * 4);
        st4(y + i * 4, make_float4(act_fwd(v.x, act), act_fwd(v.y, act), act_fwd(v.z, act), act_fwd(v.w, act)));
    }
}
// dz = dy * act'(z).  For ReLU `z` may be the activation output (same sign test).
__global__ void __launch_bounds__(NT) act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                     float* __restrict__ dz, long n4, int act) {
    pdl_entry();
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < n4; i += (long)gridDim.x * NT) {
        const float4 g = ld4(dy + i * 4), v = ld4(z + i * 4);
        st4(dz + i * 4, make_float4(g.x * act_bwd(v.x, act), g.y * act_bwd(v.y, act), g.z * act_bwd(v.z, act),
                                    g.w * act_bwd(v.w, act)));
    }
}
}  // namespace

DFINE_API int dfine_act_fwd(const float* z, float* y, long n, int act, void* stream) {
    DFINE_REQUIRE(n % 4 == 0, "act_fwd: n=%ld must be a multiple of 4", n);
    if (n == 0) return 0;
    launch_k(act_fwd_kernel, ew_grid_k(act_fwd_kernel, n / 4, NT, 0), NT, 0, (cudaStream_t)stream, z, y, n / 4, act);
    DFINE_LAUNCH_CHECK("act_fwd");
    return 0;
}
DFINE_API int dfine_act_bwd(const float* dy, const float* z, float* dz, long n, int act, void* stream) {
    DFINE_REQUIRE(n % 4 == 0, "act_bwd: n=%ld must be a multiple of 4", n);
    if (n == 0) return 0;
    launch_k(act_bwd_kernel, ew_grid_k(act_bwd_kernel, n / 4, NT, 0), NT, 0, (cudaStream_t)stream, dy, z, dz, n / 4, act);
    DFINE_LAUNCH_CHECK("act_bwd");
    return 0;
}
