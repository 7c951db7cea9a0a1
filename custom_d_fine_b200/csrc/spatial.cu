// Depthwise convolutions and the small spatial ops of the backbone / encoder, NHWC fp32.
//
//  * depthwise k x k (k = 3 stride 2 "downsample", k = 5 stride 1 LightConv, k = 3 stride 2 SCDown):
//    hgnetv2.py:83-112,295-304, hybrid_encoder.py:96-103.  Pure HBM-bound stencils: one thread per
//    (pixel, 4 channels), 16-byte accesses along C, neighbouring taps served by L1/L2.
//  * stem max-pool (F.pad(0,1,0,1) + MaxPool2d(2, stride 1, ceil_mode) hgnetv2.py:154-162).
//  * nearest x2 upsample of the FPN top-down path (hybrid_encoder.py:472).
#include "common.cuh"

namespace {

constexpr int NT = 256;
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void fma4(float4& a, const float4& x, const float4& w) {
    a.x += x.x * w.x; a.y += x.y * w.y; a.z += x.z * w.z; a.w += x.w * w.w;
}

// weights here are [k*k, C] (tap-major) — the host re-lays the [C,1,k,k] parameter once per step.
__global__ void __launch_bounds__(NT) dwconv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        float* __restrict__ y, int B, int H, int W, int C, int OH,
                                                        int OW, int k, int stride, int pad) {
    pdl_entry();
    const int VC = C / 4;
    const long total = (long)B * OH * OW * VC;
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < total; i += (long)gridDim.x * NT) {
        const int cv = (int)(i % VC);
        long p = i / VC;
        const int ow = (int)(p % OW); p /= OW;
        const int oh = (int)(p % OH);
        const int b = (int)(p / OH);
        float4 acc = make_float4(0, 0, 0, 0);
        for (int kh = 0; kh < k; ++kh) {
            const int ih = oh * stride + kh - pad;
            if (ih < 0 || ih >= H) continue;
            for (int kw = 0; kw < k; ++kw) {
                const int iw = ow * stride + kw - pad;
                if (iw < 0 || iw >= W) continue;
                fma4(acc, ld4(x + (((long)b * H + ih) * W + iw) * C + cv * 4), ld4(w + (long)(kh * k + kw) * C + cv * 4));
            }
        }
        st4(y + i * 4, acc);
    }
}

__global__ void __launch_bounds__(NT) dwconv_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                             float* __restrict__ dx, int B, int H, int W, int C,
                                                             int OH, int OW, int k, int stride, int pad) {
    pdl_entry();
    const int VC = C / 4;
    const long total = (long)B * H * W * VC;
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < total; i += (long)gridDim.x * NT) {
        const int cv = (int)(i % VC);
        long p = i / VC;
        const int iw = (int)(p % W); p /= W;
        const int ih = (int)(p % H);
        const int b = (int)(p / H);
        float4 acc = make_float4(0, 0, 0, 0);
        for (int kh = 0; kh < k; ++kh) {
            const int th = ih + pad - kh;
            if (th < 0 || th % stride) continue;
            const int oh = th / stride;
            if (oh >= OH) continue;
            for (int kw = 0; kw < k; ++kw) {
                const int tw = iw + pad - kw;
                if (tw < 0 || tw % stride) continue;
                const int ow = tw / stride;
                if (ow >= OW) continue;
                fma4(acc, ld4(dy + (((long)b * OH + oh) * OW + ow) * C + cv * 4),
                     ld4(w + (long)(kh * k + kw) * C + cv * 4));
            }
        }
        st4(dx + i * 4, acc);
    }
}

// ---- register-blocked variants: one thread = PX consecutive output pixels along W x 4 channels -------------
// A K x K depthwise stencil issues K*K activation + K*K weight loads per output in the simple kernels above
// (LSU-bound: 36 us for the 13 MB 5x5 layers).  Blocking PX = 4 outputs per thread loads each input row
// segment once ((PX-1)*S + K float4) and every weight once per thread: 16 instead of 50 loads per output.
constexpr int PX = 4;

template <int K, int S>
__global__ void __launch_bounds__(NT) dwconv_fwd_blocked(const float* __restrict__ x, const float* __restrict__ w,
                                                         float* __restrict__ y, int B, int H, int W, int C, int OH,
                                                         int OW, int pad) {
    pdl_entry();
    constexpr int SPAN = (PX - 1) * S + K;
    const int VC = C / 4, WB = (OW + PX - 1) / PX;
    const long total = (long)B * OH * WB * VC;
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < total; i += (long)gridDim.x * NT) {
        const int cv = (int)(i % VC);
        long p = i / VC;
        const int ow0 = (int)(p % WB) * PX; p /= WB;
        const int oh = (int)(p % OH);
        const int b = (int)(p / OH);
        float4 acc[PX];
#pragma unroll
        for (int q = 0; q < PX; ++q) acc[q] = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int kh = 0; kh < K; ++kh) {
            const int ih = oh * S + kh - pad;
            if (ih < 0 || ih >= H) continue;
            const float* row = x + ((long)b * H + ih) * W * C + cv * 4;
            float4 xin[SPAN];
#pragma unroll
            for (int s = 0; s < SPAN; ++s) {
                const int iw = ow0 * S + s - pad;
                xin[s] = (iw >= 0 && iw < W) ? ld4(row + (long)iw * C) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int kw = 0; kw < K; ++kw) {
                const float4 wv = ld4(w + (long)(kh * K + kw) * C + cv * 4);
#pragma unroll
                for (int q = 0; q < PX; ++q) fma4(acc[q], xin[q * S + kw], wv);
            }
        }
#pragma unroll
        for (int q = 0; q < PX; ++q)
            if (ow0 + q < OW) st4(y + ((((long)b * OH + oh) * OW + ow0 + q) * VC + cv) * 4, acc[q]);
    }
}

// stride-1 data gradient: dx[ih, iw] = sum_{kh,kw} dy[ih + pad - kh, iw + pad - kw] * w[kh, kw]
template <int K>
__global__ void __launch_bounds__(NT) dwconv_bwd_data_blocked(const float* __restrict__ dy, const float* __restrict__ w,
                                                              float* __restrict__ dx, int B, int H, int W, int C,
                                                              int OH, int OW, int pad) {
    pdl_entry();
    constexpr int SPAN = PX - 1 + K;
    const int VC = C / 4, WB = (W + PX - 1) / PX;
    const long total = (long)B * H * WB * VC;
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < total; i += (long)gridDim.x * NT) {
        const int cv = (int)(i % VC);
        long p = i / VC;
        const int iw0 = (int)(p % WB) * PX; p /= WB;
        const int ih = (int)(p % H);
        const int b = (int)(p / H);
        float4 acc[PX];
#pragma unroll
        for (int q = 0; q < PX; ++q) acc[q] = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int kh = 0; kh < K; ++kh) {
            const int oh = ih + pad - kh;
            if (oh < 0 || oh >= OH) continue;
            const float* row = dy + ((long)b * OH + oh) * OW * C + cv * 4;
            float4 g[SPAN];   // g[s] = dy[oh, iw0 + pad - (K-1) + s]
#pragma unroll
            for (int s = 0; s < SPAN; ++s) {
                const int ow = iw0 + pad - (K - 1) + s;
                g[s] = (ow >= 0 && ow < OW) ? ld4(row + (long)ow * C) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int kw = 0; kw < K; ++kw) {
                const float4 wv = ld4(w + (long)(kh * K + kw) * C + cv * 4);
#pragma unroll
                for (int q = 0; q < PX; ++q) fma4(acc[q], g[q + (K - 1 - kw)], wv);   // ow = iw0 + q + pad - kw
            }
        }
#pragma unroll
        for (int q = 0; q < PX; ++q)
            if (iw0 + q < W) st4(dx + ((((long)b * H + ih) * W + iw0 + q) * VC + cv) * 4, acc[q]);
    }
}

// 3x3 / stride 2 / pad 1 data gradient (HG_Stage.downsample, SCDown): one thread = a 2x2 block of input pixels
// (even row, even column origin) x 4 channels.  With stride 2 an input pixel receives 1, 2, 2 or 4 taps depending on
// its row / column parity, all from the four outputs (oh, ow), (oh, ow+1), (oh+1, ow), (oh+1, ow+1), oh = ih0 / 2:
// four dy loads feed four stores (the generic kernel walked 9 taps with modulo tests per input pixel).
__global__ void __launch_bounds__(NT) dwconv3x3s2_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                                  float* __restrict__ dx, int B, int H, int W, int C,
                                                                  int OH, int OW) {
    pdl_entry();
    const int VC = C / 4, HB = (H + 1) / 2, WB = (W + 1) / 2;
    const long total = (long)B * HB * WB * VC;
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < total; i += (long)gridDim.x * NT) {
        const int cv = (int)(i % VC);
        long p = i / VC;
        const int ow = (int)(p % WB); p /= WB;
        const int oh = (int)(p % HB);
        const int b = (int)(p / HB);
        const int ih0 = 2 * oh, iw0 = 2 * ow;
        const float4 z = make_float4(0, 0, 0, 0);
        const float* base = dy + ((long)b * OH * OW) * C + cv * 4;
        const bool r0 = oh < OH, r1 = oh + 1 < OH, c0 = ow < OW, c1 = ow + 1 < OW;
        const float4 g00 = (r0 && c0) ? ld4(base + ((long)oh * OW + ow) * C) : z;
        const float4 g01 = (r0 && c1) ? ld4(base + ((long)oh * OW + ow + 1) * C) : z;
        const float4 g10 = (r1 && c0) ? ld4(base + ((long)(oh + 1) * OW + ow) * C) : z;
        const float4 g11 = (r1 && c1) ? ld4(base + ((long)(oh + 1) * OW + ow + 1) * C) : z;
        const float* wc = w + cv * 4;
        float4 a;
        // (even row, even col): ih = 2 oh + kh - 1 -> kh = 1; iw likewise kw = 1
        a = z; fma4(a, g00, ld4(wc + (long)(1 * 3 + 1) * C));
        st4(dx + ((((long)b * H + ih0) * W + iw0) * VC + cv) * 4, a);
        if (iw0 + 1 < W) {   // (even, odd): kw = 0 from output ow + 1, kw = 2 from output ow
            a = z; fma4(a, g01, ld4(wc + (long)(1 * 3 + 0) * C)); fma4(a, g00, ld4(wc + (long)(1 * 3 + 2) * C));
            st4(dx + ((((long)b * H + ih0) * W + iw0 + 1) * VC + cv) * 4, a);
        }
        if (ih0 + 1 < H) {   // (odd, even): kh = 0 from output oh + 1, kh = 2 from output oh
            a = z; fma4(a, g10, ld4(wc + (long)(0 * 3 + 1) * C)); fma4(a, g00, ld4(wc + (long)(2 * 3 + 1) * C));
            st4(dx + ((((long)b * H + ih0 + 1) * W + iw0) * VC + cv) * 4, a);
            if (iw0 + 1 < W) {
                a = z;
                fma4(a, g11, ld4(wc + (long)(0 * 3 + 0) * C)); fma4(a, g10, ld4(wc + (long)(0 * 3 + 2) * C));
                fma4(a, g01, ld4(wc + (long)(2 * 3 + 0) * C)); fma4(a, g00, ld4(wc + (long)(2 * 3 + 2) * C));
                st4(dx + ((((long)b * H + ih0 + 1) * W + iw0 + 1) * VC + cv) * 4, a);
            }
        }
    }
}

// dw[tap, c] += sum_pixels dy * x_shifted.  CTA = slab of output pixels; lanes tile [pixels, C/4];
// per-thread register accumulators for all taps (K*K float4), merged through shared atomics.
template <int K, int S>
__global__ void __launch_bounds__(NT) dwconv_bwd_weight_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                               float* __restrict__ dw, int B, int H, int W, int C,
                                                               int OH, int OW, int pad, long pix_per_cta) {
    pdl_entry();
    constexpr int stride = S;
    extern __shared__ float shw[];  // [K*K*C]
    for (int i = threadIdx.x; i < K * K * C; i += NT) shw[i] = 0.f;
    __syncthreads();
    const int VC = C / 4;
    const int LPR = VC < NT ? VC : NT, RPP = NT / LPR;
    const int lane_r = threadIdx.x / LPR, lane_c = threadIdx.x % LPR;
    const long P = (long)B * OH * ((OW + PX - 1) / PX);      // pixel blocks
    const long p0 = (long)blockIdx.x * pix_per_cta, p1 = p0 + pix_per_cta < P ? p0 + pix_per_cta : P;
    if (lane_r < RPP) {
        for (int cv = lane_c; cv < VC; cv += LPR) {
            float4 acc[K * K];
#pragma unroll
            for (int t = 0; t < K * K; ++t) acc[t] = make_float4(0, 0, 0, 0);
            // pixel blocks of PX outputs along W: dy loaded once per block, each input row segment once per kh
            constexpr int SPANW = (PX - 1) * S + K;
            const int WB = (OW + PX - 1) / PX;
            for (long p = p0 + lane_r; p < p1; p += RPP) {
                const int ow0 = (int)(p % WB) * PX;
                const long q = p / WB;
                const int oh = (int)(q % OH), b = (int)(q / OH);
                float4 g[PX];
#pragma unroll
                for (int j = 0; j < PX; ++j)
                    g[j] = (ow0 + j < OW) ? ld4(dy + ((((long)b * OH + oh) * OW + ow0 + j) * C) + cv * 4)
                                          : make_float4(0, 0, 0, 0);
#pragma unroll
                for (int kh = 0; kh < K; ++kh) {
                    const int ih = oh * stride + kh - pad;
                    if (ih < 0 || ih >= H) continue;
                    const float* row = x + ((long)b * H + ih) * W * C + cv * 4;
                    float4 xin[SPANW];
#pragma unroll
                    for (int sI = 0; sI < SPANW; ++sI) {
                        const int iw = ow0 * stride + sI - pad;
                        xin[sI] = (iw >= 0 && iw < W) ? ld4(row + (long)iw * C) : make_float4(0, 0, 0, 0);
                    }
#pragma unroll
                    for (int kw = 0; kw < K; ++kw)
#pragma unroll
                        for (int j = 0; j < PX; ++j)
                            fma4(acc[kh * K + kw], g[j], xin[S * j + kw]);
                }
            }
#pragma unroll
            for (int t = 0; t < K * K; ++t) {
                float* s = shw + (long)t * C + cv * 4;
                atomicAdd(s + 0, acc[t].x); atomicAdd(s + 1, acc[t].y); atomicAdd(s + 2, acc[t].z); atomicAdd(s + 3, acc[t].w);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * K * C; i += NT) atomicAdd(&dw[i], shw[i]);
}

// out[h,w] = max over the 2x2 window of the right/bottom zero-padded map; first max wins on ties.
__device__ __forceinline__ float padded(const float* __restrict__ x, int b, int h, int w, int c, int H, int W, int C) {
    return (h < H && w < W) ? __ldg(x + (((long)b * H + h) * W + w) * C + c) : 0.f;
}
__device__ __forceinline__ int window_argmax(const float* __restrict__ x, int b, int h, int w, int c, int H, int W,
                                             int C, float* mx) {
    float best = padded(x, b, h, w, c, H, W, C);
    int arg = 0;
    float v = padded(x, b, h, w + 1, c, H, W, C);
    if (v > best || isnan(v)) { best = v; arg = 1; }
    v = padded(x, b, h + 1, w, c, H, W, C);
    if (v > best || isnan(v)) { best = v; arg = 2; }
    v = padded(x, b, h + 1, w + 1, c, H, W, C);
    if (v > best || isnan(v)) { best = v; arg = 3; }
    *mx = best;
    return arg;
}
__global__ void __launch_bounds__(NT) maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B,
                                                         int H, int W, int C) {
    pdl_entry();
    const long total = (long)B * H * W * C;
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < total; i += (long)gridDim.x * NT) {
        const int c = (int)(i % C);
        long p = i / C;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const int b = (int)(p / H);
        float mx;
        window_argmax(x, b, h, w, c, H, W, C, &mx);
        y[i] = mx;
    }
}
// gather form of the backward: input (h,w) receives dy of each of its <=4 windows whose argmax it is.
__global__ void __launch_bounds__(NT) maxpool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                         float* __restrict__ dx, int B, int H, int W, int C) {
    pdl_entry();
    const long total = (long)B * H * W * C;
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < total; i += (long)gridDim.x * NT) {
        const int c = (int)(i % C);
        long p = i / C;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const int b = (int)(p / H);
        float acc = 0.f, mx;
        // window origin (h-dh, w-dw) sees this element at slot dh*2+dw
#pragma unroll
        for (int dh = 0; dh < 2; ++dh)
#pragma unroll
            for (int dw = 0; dw < 2; ++dw) {
                const int oh = h - dh, ow = w - dw;
                if (oh < 0 || ow < 0) continue;
                if (window_argmax(x, b, oh, ow, c, H, W, C, &mx) == dh * 2 + dw)
                    acc += __ldg(dy + (((long)b * H + oh) * W + ow) * C + c);
            }
        dx[i] = acc;
    }
}

// ---- float4 variants of the stem max-pool (C % 4 == 0): one thread = one pixel x 4 channels ----------------
__device__ __forceinline__ float4 padded4(const float* __restrict__ x, int b, int h, int w, int cv, int H, int W, long ldx) {
    return (h < H && w < W) ? ld4(x + (((long)b * H + h) * W + w) * ldx + cv * 4) : make_float4(0, 0, 0, 0);
}
// argmax slot (0..3, first max wins, NaN propagates) of each of the 4 channels, packed in one int
__device__ __forceinline__ int window_argmax4(const float* __restrict__ x, int b, int h, int w, int cv, int H, int W,
                                              long ldx, float4* mx) {
    const float4 v[4] = {padded4(x, b, h, w, cv, H, W, ldx), padded4(x, b, h, w + 1, cv, H, W, ldx),
                         padded4(x, b, h + 1, w, cv, H, W, ldx), padded4(x, b, h + 1, w + 1, cv, H, W, ldx)};
    float best[4] = {v[0].x, v[0].y, v[0].z, v[0].w};
    int arg[4] = {0, 0, 0, 0};
#pragma unroll
    for (int s = 1; s < 4; ++s) {
        const float e[4] = {v[s].x, v[s].y, v[s].z, v[s].w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (e[c] > best[c] || isnan(e[c])) { best[c] = e[c]; arg[c] = s; }
    }
    *mx = make_float4(best[0], best[1], best[2], best[3]);
    return arg[0] | (arg[1] << 2) | (arg[2] << 4) | (arg[3] << 6);
}
__global__ void __launch_bounds__(NT) maxpool_fwd4_kernel(const float* __restrict__ x, float* __restrict__ y, int B,
                                                          int H, int W, int VC, long ldx) {
    pdl_entry();
    const long total = (long)B * H * W * VC;
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < total; i += (long)gridDim.x * NT) {
        const int cv = (int)(i % VC);
        long p = i / VC;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const int b = (int)(p / H);
        float4 mx;
        window_argmax4(x, b, h, w, cv, H, W, ldx, &mx);
        st4(y + i * 4, mx);
    }
}
__global__ void __launch_bounds__(NT) maxpool_bwd4_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          float* __restrict__ dx, int B, int H, int W, int VC, long ldx) {
    pdl_entry();
    const long total = (long)B * H * W * VC;
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < total; i += (long)gridDim.x * NT) {
        const int cv = (int)(i % VC);
        long p = i / VC;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const int b = (int)(p / H);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        float4 mx;
#pragma unroll
        for (int dh = 0; dh < 2; ++dh)
#pragma unroll
            for (int dw = 0; dw < 2; ++dw) {
                const int oh = h - dh, ow = w - dw;
                if (oh < 0 || ow < 0) continue;
                const int arg = window_argmax4(x, b, oh, ow, cv, H, W, ldx, &mx);
                const float4 g = ld4(dy + ((((long)b * H + oh) * W + ow) * VC + cv) * 4);
                const float ge[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (((arg >> (2 * c)) & 3) == dh * 2 + dw) acc[c] += ge[c];
            }
        st4(dx + i * 4, make_float4(acc[0], acc[1], acc[2], acc[3]));
    }
}

__global__ void __launch_bounds__(NT) upsample2x_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B,
                                                            int H, int W, int VC) {
    pdl_entry();
    const long total = (long)B * (2 * H) * (2 * W) * VC;
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < total; i += (long)gridDim.x * NT) {
        const int cv = (int)(i % VC);
        long p = i / VC;
        const int ow = (int)(p % (2 * W)); p /= (2 * W);
        const int oh = (int)(p % (2 * H));
        const int b = (int)(p / (2 * H));
        st4(y + i * 4, ld4(x + ((((long)b * H + oh / 2) * W + ow / 2) * VC + cv) * 4));
    }
}
__global__ void __launch_bounds__(NT) upsample2x_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                                                            int B, int H, int W, int VC) {
    pdl_entry();
    const long total = (long)B * H * W * VC;
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < total; i += (long)gridDim.x * NT) {
        const int cv = (int)(i % VC);
        long p = i / VC;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const int b = (int)(p / H);
        const float* base = dy + ((((long)b * 2 * H + 2 * h) * 2 * W + 2 * w) * VC + cv) * 4;
        const long rs = (long)2 * W * VC * 4;
        const float4 a = ld4(base), bb = ld4(base + VC * 4), c = ld4(base + rs), d = ld4(base + rs + VC * 4);
        st4(dx + i * 4, make_float4(a.x + bb.x + c.x + d.x, a.y + bb.y + c.y + d.y, a.z + bb.z + c.z + d.z,
                                    a.w + bb.w + c.w + d.w));
    }
}

}  // namespace

// x [B,H,W,C], w [k*k,C] tap-major, y [B,OH,OW,C]; symmetric padding `pad`.
DFINE_API int dfine_dwconv_fwd(const float* x, const float* w, float* y, int B, int H, int W, int C, int k,
                               int stride, int pad, void* stream) {
    DFINE_REQUIRE(C % 4 == 0, "dwconv_fwd: C=%d must be a multiple of 4", C);
    const int OH = (H + 2 * pad - k) / stride + 1, OW = (W + 2 * pad - k) / stride + 1;
    const long total = (long)B * OH * OW * (C / 4);
    if (total == 0) return 0;
    const long blocked = (long)B * OH * ((OW + PX - 1) / PX) * (C / 4);
    cudaStream_t st = (cudaStream_t)stream;
    if (k == 5 && stride == 1)
        launch_k(dwconv_fwd_blocked<5, 1>, ew_grid_k(dwconv_fwd_blocked<5, 1>, blocked, NT, 0), NT, 0, st, x, w, y, B, H, W, C, OH, OW, pad);
    else if (k == 3 && stride == 2)
        launch_k(dwconv_fwd_blocked<3, 2>, ew_grid_k(dwconv_fwd_blocked<3, 2>, blocked, NT, 0), NT, 0, st, x, w, y, B, H, W, C, OH, OW, pad);
    else if (k == 3 && stride == 1)
        launch_k(dwconv_fwd_blocked<3, 1>, ew_grid_k(dwconv_fwd_blocked<3, 1>, blocked, NT, 0), NT, 0, st, x, w, y, B, H, W, C, OH, OW, pad);
    else
        launch_k(dwconv_fwd_kernel, ew_grid_k(dwconv_fwd_kernel, total, NT, 0), NT, 0, st, x, w, y, B, H, W, C, OH, OW, k, stride, pad);
    DFINE_LAUNCH_CHECK("dwconv_fwd");
    return 0;
}

DFINE_API int dfine_dwconv_bwd_data(const float* dy, const float* w, float* dx, int B, int H, int W, int C, int k,
                                    int stride, int pad, void* stream) {
    DFINE_REQUIRE(C % 4 == 0, "dwconv_bwd_data: C=%d", C);
    const int OH = (H + 2 * pad - k) / stride + 1, OW = (W + 2 * pad - k) / stride + 1;
    const long total = (long)B * H * W * (C / 4);
    if (total == 0) return 0;
    const long blocked = (long)B * H * ((W + PX - 1) / PX) * (C / 4);
    cudaStream_t st = (cudaStream_t)stream;
    if (k == 5 && stride == 1)
        launch_k(dwconv_bwd_data_blocked<5>, ew_grid_k(dwconv_bwd_data_blocked<5>, blocked, NT, 0), NT, 0, st, dy, w, dx, B, H, W, C, OH, OW, pad);
    else if (k == 3 && stride == 1)
        launch_k(dwconv_bwd_data_blocked<3>, ew_grid_k(dwconv_bwd_data_blocked<3>, blocked, NT, 0), NT, 0, st, dy, w, dx, B, H, W, C, OH, OW, pad);
    else if (k == 3 && stride == 2 && pad == 1)
        launch_k(dwconv3x3s2_bwd_data_kernel, ew_grid_k(dwconv3x3s2_bwd_data_kernel, (long)B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4), NT, 0), NT, 0, st, 
            dy, w, dx, B, H, W, C, OH, OW);
    else
        launch_k(dwconv_bwd_data_kernel, ew_grid_k(dwconv_bwd_data_kernel, total, NT, 0), NT, 0, st, dy, w, dx, B, H, W, C, OH, OW, k, stride, pad);
    DFINE_LAUNCH_CHECK("dwconv_bwd_data");
    return 0;
}

// dw [k*k,C] tap-major, zero-initialised by the caller.
DFINE_API int dfine_dwconv_bwd_weight(const float* dy, const float* x, float* dw, int B, int H, int W, int C, int k,
                                      int stride, int pad, void* stream) {
    DFINE_REQUIRE(C % 4 == 0 && (k == 3 || k == 5), "dwconv_bwd_weight: C=%d k=%d", C, k);
    DFINE_REQUIRE((long)k * k * C * 4 <= 200 * 1024, "dwconv_bwd_weight: k*k*C too large for shared memory");
    const int OH = (H + 2 * pad - k) / stride + 1, OW = (W + 2 * pad - k) / stride + 1;
    DFINE_REQUIRE(stride == 1 || stride == 2, "dwconv_bwd_weight: stride %d", stride);
    const long P = (long)B * OH * ((OW + PX - 1) / PX);      // blocks of PX output pixels along W
    if (P == 0) return 0;
    // one CTA per SM: every CTA ends with k*k*C global atomics on the same k*k*C addresses, which bounded the small
    // 5x5 layers (400 CTAs x 3200 atomics for 13 MB of operands: 40 us) — fewer, longer CTAs
    long ppc = (P + 148L - 1) / 148L;
    if (ppc < 16) ppc = 16;
    const size_t smem = (size_t)k * k * C * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ceil_div(P, ppc);
    // wide 5x5 layers (D-FINE-x: 512 channels = 51 KB of per-CTA partial sums) need the opt-in shared-memory limit
    auto launch = [&](auto kern) -> int {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { dfine_set_error("dwconv_bwd_weight: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
        }
        launch_k(kern, grid, NT, smem, st, dy, x, dw, B, H, W, C, OH, OW, pad, ppc);
        return 0;
    };
    int rc;
    if (k == 3 && stride == 1) rc = launch(dwconv_bwd_weight_kernel<3, 1>);
    else if (k == 3) rc = launch(dwconv_bwd_weight_kernel<3, 2>);
    else if (stride == 1) rc = launch(dwconv_bwd_weight_kernel<5, 1>);
    else rc = launch(dwconv_bwd_weight_kernel<5, 2>);
    if (rc) return rc;
    DFINE_LAUNCH_CHECK("dwconv_bwd_weight");
    return 0;
}

// x may carry a padded pixel stride ldx >= C (elements); y / dy / dx are dense [B,H,W,C].
DFINE_API int dfine_maxpool2x2_fwd(const float* x, long ldx, float* y, int B, int H, int W, int C, void* stream) {
    const long total = (long)B * H * W * C;
    if (total == 0) return 0;
    DFINE_REQUIRE(ldx == C || (C % 4 == 0 && ldx % 4 == 0 && ldx > C), "maxpool_fwd: pixel stride %ld", ldx);
    if (C % 4 == 0)
        launch_k(maxpool_fwd4_kernel, ew_grid_k(maxpool_fwd4_kernel, total / 4, NT, 0), NT, 0, (cudaStream_t)stream, x, y, B, H, W, C / 4, ldx);
    else
        launch_k(maxpool_fwd_kernel, ew_grid_k(maxpool_fwd_kernel, total, NT, 0), NT, 0, (cudaStream_t)stream, x, y, B, H, W, C);
    DFINE_LAUNCH_CHECK("maxpool_fwd");
    return 0;
}
DFINE_API int dfine_maxpool2x2_bwd(const float* x, long ldx, const float* dy, float* dx, int B, int H, int W, int C,
                                   void* stream) {
    const long total = (long)B * H * W * C;
    if (total == 0) return 0;
    DFINE_REQUIRE(ldx == C || (C % 4 == 0 && ldx % 4 == 0 && ldx > C), "maxpool_bwd: pixel stride %ld", ldx);
    if (C % 4 == 0)
        launch_k(maxpool_bwd4_kernel, ew_grid_k(maxpool_bwd4_kernel, total / 4, NT, 0), NT, 0, (cudaStream_t)stream, x, dy, dx, B, H, W, C / 4, ldx);
    else
        launch_k(maxpool_bwd_kernel, ew_grid_k(maxpool_bwd_kernel, total, NT, 0), NT, 0, (cudaStream_t)stream, x, dy, dx, B, H, W, C);
    DFINE_LAUNCH_CHECK("maxpool_bwd");
    return 0;
}
DFINE_API int dfine_upsample2x_fwd(const float* x, float* y, int B, int H, int W, int C, void* stream) {
    DFINE_REQUIRE(C % 4 == 0, "upsample2x: C=%d", C);
    const long total = (long)B * 4 * H * W * (C / 4);
    if (total == 0) return 0;
    launch_k(upsample2x_fwd_kernel, ew_grid_k(upsample2x_fwd_kernel, total, NT, 0), NT, 0, (cudaStream_t)stream, x, y, B, H, W, C / 4);
    DFINE_LAUNCH_CHECK("upsample2x_fwd");
    return 0;
}
DFINE_API int dfine_upsample2x_bwd(const float* dy, float* dx, int B, int H, int W, int C, void* stream) {
    DFINE_REQUIRE(C % 4 == 0, "upsample2x: C=%d", C);
    const long total = (long)B * H * W * (C / 4);
    if (total == 0) return 0;
    launch_k(upsample2x_bwd_kernel, ew_grid_k(upsample2x_bwd_kernel, total, NT, 0), NT, 0, (cudaStream_t)stream, dy, dx, B, H, W, C / 4);
    DFINE_LAUNCH_CHECK("upsample2x_bwd");
    return 0;
}
