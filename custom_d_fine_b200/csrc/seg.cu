// Segmentation-head kernels (NHWC fp32): GroupNorm (+ReLU) forward / backward and the backward of the bilinear resize.
//
// Replaces, in MaskDecoder.forward (src/d_fine/arch/dfine_decoder.py:353-370), nn.GroupNorm(32, C) after the lateral /
// fusion / up convolutions (with the ReLU that follows two of them) and F.interpolate(mode="bilinear",
// align_corners=False) on feature maps — which on the NHWC graph cost a permute + contiguous copy each way around ATen's
// NCHW kernels.  (The forward resize is dfine_resize_bilinear_f32 in io.cu; the mask product runs on the tcgen05
// implicit-GEMM kernel with per-image weights.)  HBM-bound: one read + one write of a [B,H,W,256] map per pass
// (210 MB at 160x160, batch 8).
#include "common.cuh"

namespace {

// ---- GroupNorm ---------------------------------------------------------------------------------------------------
// stats[b][g] = {sum, sum of squares} (double, zeroed by the caller).  One thread = one pixel x 4 channels; a group's
// channels are contiguous (cpg = C / G, cpg % 4 == 0), so a float4 belongs to one group.  grid = (chunks, B).
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, double* __restrict__ stats, long HW,
                                                       int C, int G) {
    pdl_entry();
    extern __shared__ double sh[];        // [2 * G]
    const int b = blockIdx.y, c4 = C / 4, cpg4 = C / G / 4;
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sh[i] = 0.0;
    __syncthreads();
    const long per = (HW + gridDim.x - 1) / gridDim.x;
    const long p0 = (long)blockIdx.x * per, p1 = p0 + per < HW ? p0 + per : HW;
    // thread t handles channel quad (t % c4) of pixels p0 + t / c4, stepping by blockDim / c4 pixels
    const int q = threadIdx.x % c4, pstep = blockDim.x / c4;
    float s = 0.f, ss = 0.f;
    if (threadIdx.x < pstep * c4) {
        for (long p = p0 + threadIdx.x / c4; p < p1; p += pstep) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(x + ((long)b * HW + p) * C + 4 * q));
            s += v.x + v.y + v.z + v.w;
            ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        const int g = q / cpg4;
        atomicAdd(&sh[g], (double)s);
        atomicAdd(&sh[G + g], (double)ss);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x)
        if (sh[i] != 0.0) atomicAdd(stats + ((long)b * G + i % G) * 2 + i / G, sh[i]);
}

__global__ void gn_finalize_kernel(const double* __restrict__ stats, float* __restrict__ mean, float* __restrict__ rstd,
                                   int BG, double n, float eps) {
    pdl_entry();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BG) return;
    const double m = stats[2 * i] / n;
    double var = stats[2 * i + 1] / n - m * m;
    var = var < 0.0 ? 0.0 : var;
    mean[i] = (float)m;
    rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float* __restrict__ y, long HW,
                                                       int C, int G, int B, int relu) {
    pdl_entry();
    const int c4 = C / 4, cpg4 = C / G / 4;
    const long n = (long)B * HW * c4;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int q = (int)(i % c4);
        const long bp = i / c4;
        const int b = (int)(bp / HW), g = q / cpg4;
        const float m = mean[b * G + g], r = rstd[b * G + g];
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + bp * C + 4 * q));
        const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + 4 * q));
        const float4 be = __ldg(reinterpret_cast<const float4*>(beta + 4 * q));
        float4 o;
        o.x = (v.x - m) * r * ga.x + be.x; o.y = (v.y - m) * r * ga.y + be.y;
        o.z = (v.z - m) * r * ga.z + be.z; o.w = (v.w - m) * r * ga.w + be.w;
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4*>(y + bp * C + 4 * q) = o;
    }
}

// Backward reductions: red[b][g] = {sum gamma*dz, sum gamma*dz*xhat} (double), dgamma[c] += sum dz*xhat, dbeta[c] += sum dz
// with dz = dy masked by the ReLU (y > 0 <=> xhat*gamma + beta > 0).
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            double* __restrict__ red, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, long HW, int C, int G, int relu) {
    pdl_entry();
    extern __shared__ double sh[];        // [2 * G]
    const int b = blockIdx.y, c4 = C / 4, cpg4 = C / G / 4;
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sh[i] = 0.0;
    __syncthreads();
    const long per = (HW + gridDim.x - 1) / gridDim.x;
    const long p0 = (long)blockIdx.x * per, p1 = p0 + per < HW ? p0 + per : HW;
    const int q = threadIdx.x % c4, pstep = blockDim.x / c4;
    if (threadIdx.x < pstep * c4) {
        const int g = q / cpg4;
        const float m = mean[b * G + g], r = rstd[b * G + g];
        const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + 4 * q));
        const float4 be = __ldg(reinterpret_cast<const float4*>(beta + 4 * q));
        const float gam[4] = {ga.x, ga.y, ga.z, ga.w}, bet[4] = {be.x, be.y, be.z, be.w};
        float s1 = 0.f, s2 = 0.f, dg[4] = {0, 0, 0, 0}, db[4] = {0, 0, 0, 0};
        for (long p = p0 + threadIdx.x / c4; p < p1; p += pstep) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(x + ((long)b * HW + p) * C + 4 * q));
            const float4 dv = __ldg(reinterpret_cast<const float4*>(dy + ((long)b * HW + p) * C + 4 * q));
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float xh = (xs[j] - m) * r;
                const float dz = (relu && xh * gam[j] + bet[j] <= 0.f) ? 0.f : ds[j];
                s1 += gam[j] * dz;
                s2 += gam[j] * dz * xh;
                dg[j] += dz * xh;
                db[j] += dz;
            }
        }
        atomicAdd(&sh[g], (double)s1);
        atomicAdd(&sh[G + g], (double)s2);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(dgamma + 4 * q + j, dg[j]);
            atomicAdd(dbeta + 4 * q + j, db[j]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x)
        if (sh[i] != 0.0) atomicAdd(red + ((long)b * G + i % G) * 2 + i / G, sh[i]);
}

// dx = rstd * (gamma*dz - mean_g(gamma*dz) - xhat * mean_g(gamma*dz*xhat))
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const double* __restrict__ red, float* __restrict__ dx, long HW,
                                                           int C, int G, int B, int relu, double inv_n) {
    pdl_entry();
    const int c4 = C / 4, cpg4 = C / G / 4;
    const long n = (long)B * HW * c4;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int q = (int)(i % c4);
        const long bp = i / c4;
        const int b = (int)(bp / HW), g = q / cpg4;
        const float m = mean[b * G + g], r = rstd[b * G + g];
        const float a1 = (float)(red[((long)b * G + g) * 2] * inv_n), a2 = (float)(red[((long)b * G + g) * 2 + 1] * inv_n);
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + bp * C + 4 * q));
        const float4 dv = __ldg(reinterpret_cast<const float4*>(dy + bp * C + 4 * q));
        const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + 4 * q));
        const float4 be = __ldg(reinterpret_cast<const float4*>(beta + 4 * q));
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
        const float gam[4] = {ga.x, ga.y, ga.z, ga.w}, bet[4] = {be.x, be.y, be.z, be.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float xh = (xs[j] - m) * r;
            const float dz = (relu && xh * gam[j] + bet[j] <= 0.f) ? 0.f : ds[j];
            o[j] = r * (gam[j] * dz - a1 - xh * a2);
        }
        *reinterpret_cast<float4*>(dx + bp * C + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ---- bilinear resize backward (gather form: deterministic, no atomics) ---------------------------------------------
__device__ __forceinline__ void src_index(int d, float scale, int in_size, int* i0, int* i1, float* w1) {
    float s = ((float)d + 0.5f) * scale - 0.5f;
    s = s < 0.f ? 0.f : s;
    const int a = (int)s;
    *i0 = a < in_size - 1 ? a : in_size - 1;
    *i1 = a < in_size - 1 ? a + 1 : in_size - 1;
    *w1 = s - (float)a;
}
// total weight with which destination index d reads source index s (0 if it does not)
__device__ __forceinline__ float tap_weight(int d, int s, float scale, int in_size) {
    int i0, i1;
    float w1;
    src_index(d, scale, in_size, &i0, &i1, &w1);
    float w = 0.f;
    if (i0 == s) w += 1.f - w1;
    if (i1 == s) w += w1;
    return w;
}
// dx[b, ys, xs, :] = sum over destination pixels of dy * weight; x: [B,Hs,Ws,C] source grid, dy: [B,H,W,C]
__global__ void __launch_bounds__(256) resize_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int Hs,
                                                         int Ws, int H, int W, int C) {
    pdl_entry();
    const int c4 = C / 4;
    const long n = (long)B * Hs * Ws * c4;
    const float sy = (float)Hs / (float)H, sx = (float)Ws / (float)W;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int q = (int)(i % c4);
        long r = i / c4;
        const int xs = (int)(r % Ws); r /= Ws;
        const int ys = (int)(r % Hs);
        const int b = (int)(r / Hs);
        // destination rows / columns whose source coordinate lies in (s - 1, s + 1)
        int ylo = (int)floorf(((float)ys - 0.5f) / sy - 0.5f) - 1, yhi = (int)ceilf(((float)ys + 1.5f) / sy - 0.5f) + 1;
        int xlo = (int)floorf(((float)xs - 0.5f) / sx - 0.5f) - 1, xhi = (int)ceilf(((float)xs + 1.5f) / sx - 0.5f) + 1;
        ylo = ylo < 0 ? 0 : ylo; xlo = xlo < 0 ? 0 : xlo;
        yhi = yhi > H - 1 ? H - 1 : yhi; xhi = xhi > W - 1 ? W - 1 : xhi;
        if (ys == Hs - 1) yhi = H - 1;      // the clamped border source row / column also feeds every destination beyond it
        if (xs == Ws - 1) xhi = W - 1;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int yd = ylo; yd <= yhi; ++yd) {
            const float wy = tap_weight(yd, ys, sy, Hs);
            if (wy == 0.f) continue;
            for (int xd = xlo; xd <= xhi; ++xd) {
                const float w = wy * tap_weight(xd, xs, sx, Ws);
                if (w == 0.f) continue;
                const float4 g = __ldg(reinterpret_cast<const float4*>(dy + (((long)b * H + yd) * W + xd) * C + 4 * q));
                acc.x += w * g.x; acc.y += w * g.y; acc.z += w * g.z; acc.w += w * g.w;
            }
        }
        *reinterpret_cast<float4*>(dx + (((long)b * Hs + ys) * Ws + xs) * C + 4 * q) = acc;
    }
}

// ---- mask matching cost (matcher.py:19-71, 175-237) ---------------------------------------------------------------
// logits [B, HW, Q] (pixel-major: the layout the mask product writes), gt [sumT, HW] (GT masks resized to the prediction
// size), toff [B+1].  One pass over the logits accumulates, per (query, target of the same image),
//   S1 = sum_p prob * gt,  S2 = sum_p (pos - neg) * gt   and per query  P = sum_p prob,  N = sum_p neg
// with prob = sigmoid(logit), pos = alpha (1-prob)^gamma (-log(prob + 1e-8)), neg = (1-alpha) prob^gamma (-log(1-prob+1e-8)).
// acc layout (float, zeroed by the caller): per image b a block of Q * (2*T_b + 2) floats at 2*Q*toff[b] + 2*Q*b:
// [q][0..T_b) = S1, [q][T_b..2T_b) = S2, [q][2T_b] = P, [q][2T_b+1] = N.
constexpr int MC_T = 16;         // targets per register pass
__global__ void __launch_bounds__(512) mask_cost_acc_kernel(const float* __restrict__ logits, const float* __restrict__ gt,
                                                            const int* __restrict__ toff, float* __restrict__ acc, long HW,
                                                            int Q, float alpha, float gamma, int chunk) {
    pdl_entry();
    extern __shared__ float gts[];              // [MC_T][chunk]
    const int b = blockIdx.y;
    const int t0 = toff[b], T = toff[b + 1] - t0;
    if (T == 0) return;
    const long p0 = (long)blockIdx.x * chunk;
    const int np = (int)(p0 + chunk <= HW ? chunk : HW - p0);
    if (np <= 0) return;
    float* out = acc + 2L * Q * t0 + 2L * Q * b;
    const int stride = 2 * T + 2;
    for (int tb = 0; tb < T; tb += MC_T) {
        const int nt = T - tb < MC_T ? T - tb : MC_T;
        __syncthreads();
        for (int i = threadIdx.x; i < nt * np; i += blockDim.x)
            gts[(i / np) * chunk + i % np] = __ldg(gt + (long)(t0 + tb + i / np) * HW + p0 + i % np);
        __syncthreads();
        for (int q = threadIdx.x; q < Q; q += blockDim.x) {
            float s1[MC_T], s2[MC_T], P = 0.f, N = 0.f;
#pragma unroll
            for (int t = 0; t < MC_T; ++t) { s1[t] = 0.f; s2[t] = 0.f; }
            const float* lp = logits + ((long)b * HW + p0) * Q + q;
            // eight pixels per iteration, their (row-strided) loads all issued before the first use: with one load in
            // flight per thread the kernel ran at 0.4 TB/s (612 us per layer on the 160x160 masks of D-FINE-l-seg)
            constexpr int U = 8;
            for (int pb = 0; pb < np; pb += U) {
                float xs[U];
#pragma unroll
                for (int u = 0; u < U; ++u) xs[u] = pb + u < np ? __ldg(lp + (long)(pb + u) * Q) : 0.f;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int p = pb + u;
                    if (p < np) {
                        const float x = xs[u];
                        const float prob = 1.f / (1.f + expf(-x));
                        const float pg = gamma == 2.f ? prob * prob : powf(prob, gamma);
                        const float qg = gamma == 2.f ? (1.f - prob) * (1.f - prob) : powf(1.f - prob, gamma);
                        const float neg = (1.f - alpha) * pg * (-logf(1.f - prob + 1e-8f));
                        const float pos = alpha * qg * (-logf(prob + 1e-8f));
                        const float d = pos - neg;
                        P += prob; N += neg;
#pragma unroll
                        for (int t = 0; t < MC_T; ++t)
                            if (t < nt) { const float g = gts[t * chunk + p]; s1[t] += prob * g; s2[t] += d * g; }
                    }
                }
            }
            float* o = out + (long)q * stride;
#pragma unroll
            for (int t = 0; t < MC_T; ++t)
                if (t < nt) { atomicAdd(o + tb + t, s1[t]); atomicAdd(o + T + tb + t, s2[t]); }
            if (tb == 0) { atomicAdd(o + 2 * T, P); atomicAdd(o + 2 * T + 1, N); }
        }
    }
}
// extra[Q*toff[b] + q*T_b + t] (+)= w_dice * (1 - (2 S1 + 1e-6) / (P + G_t + 1e-6)) + w_mask * (S2 + N) / HW
__global__ void mask_cost_final_kernel(const float* __restrict__ acc, const float* __restrict__ gsum, const int* __restrict__ toff,
                                       float* __restrict__ extra, int B, int Q, long HW, float w_dice, float w_mask) {
    pdl_entry();
    const int b = blockIdx.y;
    const int t0 = toff[b], T = toff[b + 1] - t0;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q * T) return;
    const int q = i / T, t = i % T;
    const float* a = acc + 2L * Q * t0 + 2L * Q * b + (long)q * (2 * T + 2);
    const float S1 = a[t], S2 = a[T + t], P = a[2 * T], N = a[2 * T + 1];
    float c = 0.f;
    if (w_dice > 0.f) c += w_dice * (1.f - (2.f * S1 + 1e-6f) / (P + gsum[t0 + t] + 1e-6f));
    if (w_mask > 0.f) c += w_mask * ((S2 + N) / (float)HW);
    extra[(long)Q * t0 + (long)q * T + t] = c;
}

// ---- mask losses (dfine_criterion.py:335-386 cropped BCE, 404-450 cropped Dice, 504-556) ---------------------------
// pred [M, HW] matched mask logits, gt [sumT, HW] resized GT masks in [0,1], t_idx [M] (int64) the target of every row,
// tboxes [sumT,4] normalised cxcywh.  Both losses are evaluated INSIDE the GT box only:
//   bce_row  = sum_inside BCE-with-logits(pred, gt) / max(box area in mask pixels, 1)
//   dice_row = 1 - (2 sum p*t + 1e-6) / (sum p + sum t + 1e-6),  p = sigmoid(pred) inside the box, t = gt inside the box
// One CTA per row; sums[m] = {sum p*t, sum p, sum t, area} is kept for the backward.
struct BoxPx { float x1, y1, x2, y2; };
__device__ __forceinline__ BoxPx box_px(const float* b, int Hm, int Wm) {
    BoxPx r;
    r.x1 = fminf(fmaxf((b[0] - b[2] / 2) * Wm, 0.f), (float)(Wm - 1));
    r.y1 = fminf(fmaxf((b[1] - b[3] / 2) * Hm, 0.f), (float)(Hm - 1));
    r.x2 = fminf(fmaxf((b[0] + b[2] / 2) * Wm, 1.f), (float)Wm);
    r.y2 = fminf(fmaxf((b[1] + b[3] / 2) * Hm, 1.f), (float)Hm);
    return r;
}
__device__ __forceinline__ float block_sum_f(float v, float* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x / 32] = v;
    __syncthreads();
    float r = 0.f;
    for (int i = 0; i < (int)(blockDim.x + 31) / 32; ++i) r += sh[i];
    return r;
}
__global__ void __launch_bounds__(256) mask_loss_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                            const long* __restrict__ t_idx, const float* __restrict__ tboxes,
                                                            float* __restrict__ bce_row, float* __restrict__ dice_row,
                                                            float* __restrict__ sums, int Hm, int Wm) {
    pdl_entry();
    __shared__ float sh[8];
    const int m = blockIdx.x;
    const long t = t_idx[m];
    const BoxPx bx = box_px(tboxes + t * 4, Hm, Wm);
    const float* pr = pred + (long)m * Hm * Wm;
    const float* g = gt + t * (long)Hm * Wm;
    // only the rows / columns that can be inside the box are visited
    const int ya = (int)ceilf(bx.y1), yb = (int)ceilf(bx.y2);     // y in [ya, yb): y >= y1 and y < y2
    const int xa = (int)ceilf(bx.x1), xb = (int)ceilf(bx.x2);
    const int bw = xb - xa, n = (yb - ya) * (bw > 0 ? bw : 0);
    float bce = 0.f, spt = 0.f, sp = 0.f, st = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int y = ya + i / bw, x = xa + i % bw;
        const float xv = __ldg(pr + (long)y * Wm + x), tv = __ldg(g + (long)y * Wm + x);
        const float mv = fmaxf(-xv, 0.f);
        bce += (1.f - tv) * xv + mv + logf(expf(-mv) + expf(-xv - mv));
        const float pv = 1.f / (1.f + expf(-xv));
        spt += pv * tv; sp += pv; st += tv;
    }
    bce = block_sum_f(bce, sh); spt = block_sum_f(spt, sh); sp = block_sum_f(sp, sh); st = block_sum_f(st, sh);
    if (threadIdx.x == 0) {
        const float area = fmaxf((bx.x2 - bx.x1) * (bx.y2 - bx.y1), 1.f);
        bce_row[m] = bce / area;
        dice_row[m] = 1.f - (2.f * spt + 1e-6f) / (sp + st + 1e-6f);
        sums[4 * m] = spt; sums[4 * m + 1] = sp; sums[4 * m + 2] = st; sums[4 * m + 3] = area;
    }
}
// dpred [M, HW] (written completely: zero outside the box)
__global__ void __launch_bounds__(256) mask_loss_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                            const long* __restrict__ t_idx, const float* __restrict__ tboxes,
                                                            const float* __restrict__ sums, const float* __restrict__ g_bce,
                                                            const float* __restrict__ g_dice, float* __restrict__ dpred,
                                                            int Hm, int Wm) {
    pdl_entry();
    const int m = blockIdx.x;
    const long t = t_idx[m];
    const BoxPx bx = box_px(tboxes + t * 4, Hm, Wm);
    const float* pr = pred + (long)m * Hm * Wm;
    const float* g = gt + t * (long)Hm * Wm;
    float* d = dpred + (long)m * Hm * Wm;
    const float spt = sums[4 * m], sp = sums[4 * m + 1], st = sums[4 * m + 2], area = sums[4 * m + 3];
    const float N = 2.f * spt + 1e-6f, D = sp + st + 1e-6f;
    const float cb = g_bce[m] / area, cd = g_dice[m];
    const int HW = Hm * Wm;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        const int y = i / Wm, x = i % Wm;
        float o = 0.f;
        if ((float)x >= bx.x1 && (float)x < bx.x2 && (float)y >= bx.y1 && (float)y < bx.y2) {
            const float xv = __ldg(pr + i), tv = __ldg(g + i);
            const float pv = 1.f / (1.f + expf(-xv));
            // d dice / d p = -(2 t D - N) / D^2 ; d p / d x = p (1 - p)
            o = cb * (pv - tv) + cd * (-(2.f * tv * D - N) / (D * D)) * pv * (1.f - pv);
        }
        d[i] = o;
    }
}

// ---- the same losses on PIXEL-MAJOR logits ------------------------------------------------------------------------
// pred [B, HW, R]: row r of image b is the column pred[(b*HW + px)*R + r] — the layout the tcgen05 mask product writes
// (pixels = GEMM rows), so the matched logits never go through a transposing copy or a library sgemm.  Lane = row
// (32 consecutive r: 128-byte coalesced accesses), a warp walks pixel rows y, y+8, ... of its CTA's slab; only the union
// of the 32 lanes' boxes is visited.  sums [B*R, 4] = {bce sum, sum p*t, sum p, sum t}, zeroed by the caller.
constexpr int PM_YROWS = 32;       // pixel rows per CTA slab
__global__ void __launch_bounds__(256) mask_loss_pm_acc_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                               const long* __restrict__ t_idx, const float* __restrict__ tboxes,
                                                               float* __restrict__ sums, int R, int Hm, int Wm) {
    pdl_entry();
    __shared__ float red[8][32][4];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int b = blockIdx.y, r = blockIdx.x * 32 + lane;
    const int ybeg = blockIdx.z * PM_YROWS, yend = min(Hm, ybeg + PM_YROWS);
    const long t = r < R ? t_idx[(long)b * R + r] : -1;
    const bool valid = t >= 0;
    int ya = Hm, yb = 0, xa = Wm, xb = 0;
    if (valid) {
        const BoxPx bx = box_px(tboxes + t * 4, Hm, Wm);
        ya = (int)ceilf(bx.y1); yb = min((int)ceilf(bx.y2), Hm);
        xa = (int)ceilf(bx.x1); xb = min((int)ceilf(bx.x2), Wm);
    }
    int uya = ya, uyb = yb, uxa = xa, uxb = xb;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        uya = min(uya, __shfl_xor_sync(0xffffffffu, uya, o)); uyb = max(uyb, __shfl_xor_sync(0xffffffffu, uyb, o));
        uxa = min(uxa, __shfl_xor_sync(0xffffffffu, uxa, o)); uxb = max(uxb, __shfl_xor_sync(0xffffffffu, uxb, o));
    }
    float bce = 0.f, spt = 0.f, sp = 0.f, st = 0.f;
    const float* g = gt + (valid ? t : 0) * (long)Hm * Wm;
    for (int y = max(ybeg, uya) + warp; y < min(yend, uyb); y += 8) {
        const bool yin = valid && y >= ya && y < yb;
        const float* prow = pred + ((long)b * Hm * Wm + (long)y * Wm) * R + r;
        const float* grow = g + (long)y * Wm;
#pragma unroll 4
        for (int x = uxa; x < uxb; ++x) {
            if (yin && x >= xa && x < xb) {
                const float xv = __ldg(prow + (long)x * R), tv = __ldg(grow + x);
                const float mv = fmaxf(-xv, 0.f);
                bce += (1.f - tv) * xv + mv + logf(expf(-mv) + expf(-xv - mv));
                const float pv = 1.f / (1.f + expf(-xv));
                spt += pv * tv; sp += pv; st += tv;
            }
        }
    }
    red[warp][lane][0] = bce; red[warp][lane][1] = spt; red[warp][lane][2] = sp; red[warp][lane][3] = st;
    __syncthreads();
    if (warp < 4) {        // warp j sums component j of the 8 partials of every lane's row
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w][lane][warp];
        if (valid && v != 0.f) atomicAdd(sums + ((long)b * R + r) * 4 + warp, v);
    }
}
__global__ void mask_loss_pm_final_kernel(const float* __restrict__ sums, const long* __restrict__ t_idx,
                                          const float* __restrict__ tboxes, float* __restrict__ bce_row,
                                          float* __restrict__ dice_row, long n, int Hm, int Wm) {
    pdl_entry();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long t = t_idx[i];
    if (t < 0) { bce_row[i] = 0.f; dice_row[i] = 0.f; return; }
    const BoxPx bx = box_px(tboxes + t * 4, Hm, Wm);
    const float area = fmaxf((bx.x2 - bx.x1) * (bx.y2 - bx.y1), 1.f);
    bce_row[i] = sums[4 * i] / area;
    dice_row[i] = 1.f - (2.f * sums[4 * i + 1] + 1e-6f) / (sums[4 * i + 2] + sums[4 * i + 3] + 1e-6f);
}
// dpred [B, HW, R], written completely (zero outside the boxes and on the padding rows)
__global__ void __launch_bounds__(256) mask_loss_pm_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                               const long* __restrict__ t_idx, const float* __restrict__ tboxes,
                                                               const float* __restrict__ sums, const float* __restrict__ g_bce,
                                                               const float* __restrict__ g_dice, float* __restrict__ dpred,
                                                               int R, int Hm, int Wm) {
    pdl_entry();
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int b = blockIdx.y, r = blockIdx.x * 32 + lane;
    if (r >= R) return;
    const int ybeg = blockIdx.z * PM_YROWS, yend = min(Hm, ybeg + PM_YROWS);
    const long row = (long)b * R + r;
    const long t = t_idx[row];
    const bool valid = t >= 0;
    int ya = Hm, yb = 0, xa = Wm, xb = 0;
    float cb = 0.f, cd = 0.f, N = 0.f, D = 1.f;
    if (valid) {
        const BoxPx bx = box_px(tboxes + t * 4, Hm, Wm);
        ya = (int)ceilf(bx.y1); yb = min((int)ceilf(bx.y2), Hm);
        xa = (int)ceilf(bx.x1); xb = min((int)ceilf(bx.x2), Wm);
        const float area = fmaxf((bx.x2 - bx.x1) * (bx.y2 - bx.y1), 1.f);
        cb = g_bce[row] / area; cd = g_dice[row];
        N = 2.f * sums[4 * row + 1] + 1e-6f; D = sums[4 * row + 2] + sums[4 * row + 3] + 1e-6f;
    }
    const float* g = gt + (valid ? t : 0) * (long)Hm * Wm;
    const float invD2 = 1.f / (D * D);
    for (int y = ybeg + warp; y < yend; y += 8) {
        const bool yin = valid && y >= ya && y < yb;
        const long base = ((long)b * Hm * Wm + (long)y * Wm) * R + r;
        const float* grow = g + (long)y * Wm;
#pragma unroll 4
        for (int x = 0; x < Wm; ++x) {
            float o = 0.f;
            if (yin && x >= xa && x < xb) {
                const float xv = __ldg(pred + base + (long)x * R), tv = __ldg(grow + x);
                const float pv = 1.f / (1.f + expf(-xv));
                o = cb * (pv - tv) - cd * (2.f * tv * D - N) * invD2 * pv * (1.f - pv);
            }
            dpred[base + (long)x * R] = o;
        }
    }
}

}  // namespace

// GroupNorm forward over an NHWC tensor x [B,HW,C]: y = (x - mean_g) * rstd_g * gamma + beta (+ReLU).  stats: caller-zeroed
// double [B*G*2] scratch; mean / rstd [B*G] are written for the backward.  C % (4*G) == 0.
DFINE_API int dfine_groupnorm_fwd(const float* x, const float* gamma, const float* beta, float* y, double* stats, float* mean,
                                  float* rstd, int B, long HW, int C, int G, float eps, int relu, void* stream) {
    DFINE_REQUIRE(B >= 0 && HW > 0 && C > 0 && G > 0 && C % G == 0 && (C / G) % 4 == 0 && C / 4 <= 256,
                  "groupnorm: C=%d G=%d (channels per group must be a multiple of 4, C <= 1024)", C, G);
    DFINE_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0 && ((uintptr_t)gamma % 16) == 0 && ((uintptr_t)beta % 16) == 0,
                  "groupnorm: pointers must be 16-byte aligned");
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int chunks = (int)((HW + 255) / 256);
    if (chunks > 148 * 4 / (B > 0 ? B : 1) + 1) chunks = 148 * 4 / B + 1;
    launch_k(gn_stats_kernel, dim3(chunks, B), 256, 2 * G * sizeof(double), st, x, stats, HW, C, G);
    launch_k(gn_finalize_kernel, ceil_div((long)B * G, 128), 128, 0, st, stats, mean, rstd, B * G, (double)HW * (C / G), eps);
    launch_k(gn_apply_kernel, ew_grid_k(gn_apply_kernel, (long)B * HW * (C / 4), 256, 0), 256, 0, st, x, mean, rstd, gamma, beta, y, HW, C, G, B, relu);
    DFINE_LAUNCH_CHECK("groupnorm_fwd");
    return 0;
}

// GroupNorm backward: dx [B,HW,C]; dgamma / dbeta [C] are ACCUMULATED (caller zero-fills or passes the gradient arena);
// red: caller-zeroed double [B*G*2] scratch.
DFINE_API int dfine_groupnorm_bwd(const float* dy, const float* x, const float* gamma, const float* beta, const float* mean,
                                  const float* rstd, float* dx, float* dgamma, float* dbeta, double* red, int B, long HW,
                                  int C, int G, int relu, void* stream) {
    DFINE_REQUIRE(B >= 0 && HW > 0 && C > 0 && G > 0 && C % G == 0 && (C / G) % 4 == 0 && C / 4 <= 256, "groupnorm_bwd: C=%d G=%d", C, G);
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int chunks = (int)((HW + 255) / 256);
    if (chunks > 148 * 4 / (B > 0 ? B : 1) + 1) chunks = 148 * 4 / B + 1;
    launch_k(gn_bwd_reduce_kernel, dim3(chunks, B), 256, 2 * G * sizeof(double), st, dy, x, mean, rstd, gamma, beta, red, dgamma,
                                                                               dbeta, HW, C, G, relu);
    launch_k(gn_bwd_apply_kernel, ew_grid_k(gn_bwd_apply_kernel, (long)B * HW * (C / 4), 256, 0), 256, 0, st, dy, x, mean, rstd, gamma, beta, red, dx, HW, C, G, B,
                                                                         relu, 1.0 / ((double)HW * (C / G)));
    DFINE_LAUNCH_CHECK("groupnorm_bwd");
    return 0;
}

// Mask term of the Hungarian cost for one prediction layer (matcher.py:175-237): logits [B,HW,Q] pixel-major, gt [sumT,HW]
// float (resized GT masks), gsum [sumT] = their pixel sums, toff int32 [B+1] (device).  extra [Q*sumT] receives, for image b,
// the [Q, T_b] block (row-major) at Q*toff[b] — the additive input of dfine_matcher_extra.  workspace: 2*Q*(sumT + B) floats,
// zeroed here.  Tmax = largest T_b (host).
DFINE_API int dfine_mask_cost(const float* logits, const float* gt, const float* gsum, const int* toff, float* extra,
                              float* workspace, int B, long HW, int Q, int sumT, int Tmax, float alpha, float gamma,
                              float w_dice, float w_mask, void* stream) {
    DFINE_REQUIRE(B >= 0 && HW > 0 && Q > 0 && sumT >= 0, "mask_cost: bad dims");
    if (B == 0 || sumT == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(workspace, 0, sizeof(float) * 2L * Q * ((long)sumT + B), st);
    const int chunk = 512;                                   // pixels per CTA: 16 x 512 floats of GT in shared memory
    dim3 grid(ceil_div(HW, chunk), B);
    const int threads = Q >= 512 ? 512 : (Q + 31) / 32 * 32;      // one thread per query where they fit
    launch_k(mask_cost_acc_kernel, grid, threads, MC_T * chunk * sizeof(float), st, logits, gt, toff, workspace, HW, Q, alpha, gamma, chunk);
    dim3 g2(ceil_div((long)Q * Tmax, 128), B);
    launch_k(mask_cost_final_kernel, g2, 128, 0, st, workspace, gsum, toff, extra, B, Q, HW, w_dice, w_mask);
    DFINE_LAUNCH_CHECK("mask_cost");
    return 0;
}

// Cropped BCE + cropped Dice of M matched mask logits (rows) against their GT masks: bce_row / dice_row [M] (the caller
// averages them), sums [M,4] scratch kept for the backward.
DFINE_API int dfine_mask_loss_fwd(const float* pred, const float* gt, const long* t_idx, const float* tboxes, float* bce_row,
                                  float* dice_row, float* sums, int M, int Hm, int Wm, void* stream) {
    DFINE_REQUIRE(M >= 0 && Hm > 0 && Wm > 0, "mask_loss: bad dims");
    if (M == 0) return 0;
    launch_k(mask_loss_fwd_kernel, M, 256, 0, (cudaStream_t)stream, pred, gt, t_idx, tboxes, bce_row, dice_row, sums, Hm, Wm);
    DFINE_LAUNCH_CHECK("mask_loss_fwd");
    return 0;
}
DFINE_API int dfine_mask_loss_bwd(const float* pred, const float* gt, const long* t_idx, const float* tboxes, const float* sums,
                                  const float* g_bce, const float* g_dice, float* dpred, int M, int Hm, int Wm, void* stream) {
    DFINE_REQUIRE(M >= 0 && Hm > 0 && Wm > 0, "mask_loss_bwd: bad dims");
    if (M == 0) return 0;
    launch_k(mask_loss_bwd_kernel, M, 256, 0, (cudaStream_t)stream, pred, gt, t_idx, tboxes, sums, g_bce, g_dice, dpred, Hm, Wm);
    DFINE_LAUNCH_CHECK("mask_loss_bwd");
    return 0;
}

// The same losses on pixel-major logits pred [B, Hm*Wm, R] (the layout of dfine_conv_tc's mask product): t_idx [B*R] int64,
// negative = padding row (loss 0, zero gradient); sums [B*R, 4] scratch, ZEROED by the caller, kept for the backward.
DFINE_API int dfine_mask_loss_pm_fwd(const float* pred, const float* gt, const long* t_idx, const float* tboxes, float* bce_row,
                                     float* dice_row, float* sums, int B, int R, int Hm, int Wm, void* stream) {
    DFINE_REQUIRE(B >= 0 && R >= 0 && Hm > 0 && Wm > 0, "mask_loss_pm: bad dims");
    if ((long)B * R == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(ceil_div(R, 32), B, ceil_div(Hm, PM_YROWS));
    launch_k(mask_loss_pm_acc_kernel, grid, 256, 0, st, pred, gt, t_idx, tboxes, sums, R, Hm, Wm);
    DFINE_LAUNCH_CHECK("mask_loss_pm_acc");
    launch_k(mask_loss_pm_final_kernel, ceil_div((long)B * R, 256), 256, 0, st, sums, t_idx, tboxes, bce_row, dice_row, (long)B * R, Hm, Wm);
    DFINE_LAUNCH_CHECK("mask_loss_pm_final");
    return 0;
}
DFINE_API int dfine_mask_loss_pm_bwd(const float* pred, const float* gt, const long* t_idx, const float* tboxes,
                                     const float* sums, const float* g_bce, const float* g_dice, float* dpred, int B, int R, int Hm,
                                     int Wm, void* stream) {
    DFINE_REQUIRE(B >= 0 && R >= 0 && Hm > 0 && Wm > 0, "mask_loss_pm_bwd: bad dims");
    if ((long)B * R == 0) return 0;
    dim3 grid(ceil_div(R, 32), B, ceil_div(Hm, PM_YROWS));
    launch_k(mask_loss_pm_bwd_kernel, grid, 256, 0, (cudaStream_t)stream, pred, gt, t_idx, tboxes, sums, g_bce, g_dice, dpred, R, Hm, Wm);
    DFINE_LAUNCH_CHECK("mask_loss_pm_bwd");
    return 0;
}

// Backward of dfine_resize_bilinear_f32: dx [B,Hs,Ws,C] from dy [B,H,W,C] (C % 4 == 0), gather form.
DFINE_API int dfine_resize_bilinear_bwd(const float* dy, float* dx, int B, int Hs, int Ws, int H, int W, int C, void* stream) {
    DFINE_REQUIRE(B >= 0 && Hs > 0 && Ws > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "resize_bilinear_bwd: bad dims (C %% 4)");
    if (B == 0) return 0;
    launch_k(resize_bwd_kernel, ew_grid_k(resize_bwd_kernel, (long)B * Hs * Ws * (C / 4), 256, 0), 256, 0, (cudaStream_t)stream, dy, dx, B, Hs, Ws, H, W, C);
    DFINE_LAUNCH_CHECK("resize_bilinear_bwd");
    return 0;
}
