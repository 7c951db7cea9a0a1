// Direct kernels for the image convolution of the HGNetv2 stem (hgnetv2.py:117-124 `stem1`: 3 -> 16/24/32
// channels, 3x3, stride 2, pad 1, on the 640x640 input).  K = 27 and N <= 32 waste a GEMM tiling (the generic
// 64x64x16 CUDA-core kernel took 553 us forward / 771 us weight-gradient per step for a layer whose compulsory
// traffic is 235 MB = 36 us); the tensor-core path cannot take 3-channel (12-byte) pixels through TMA.
//
//   forward : one output pixel x all Cout channels per thread; the 27 x Cout weights sit in shared memory as
//             float4 over channels (warp-uniform broadcast reads), the 27 inputs in registers.
//   wgrad   : warp w of a CTA owns output channels 4w..4w+3, lanes stream over pixels; 4 x 27 accumulators per
//             thread, one shuffle reduction and 108 atomics per warp at the end (grid capped to a few CTAs per SM).
#include "common.cuh"

namespace {

__device__ __forceinline__ void load_patch(const float* __restrict__ x, long ldx, int b, int oh, int ow, int H, int W,
                                           float (&xin)[27]) {
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int ih = 2 * oh - 1 + kh;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int iw = 2 * ow - 1 + kw;
            const bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
            const float* p = x + (((long)b * H + (ok ? ih : 0)) * W + (ok ? iw : 0)) * ldx;
#pragma unroll
            for (int c = 0; c < 3; ++c) xin[(kh * 3 + kw) * 3 + c] = ok ? __ldg(p + c) : 0.f;
        }
    }
}

template <int CO4>
__global__ void __launch_bounds__(256) stem_fwd_kernel(const float* __restrict__ x, long ldx, const float* __restrict__ wr,
                                                       float* __restrict__ y, long ldy, int B, int H, int W, int OH,
                                                       int OW) {
    pdl_entry();
    __shared__ float4 ws[27][CO4];
    for (int i = threadIdx.x; i < 27 * CO4; i += blockDim.x) {
        const int k = i / CO4, c4 = i % CO4;
        ws[k][c4] = make_float4(__ldg(wr + (4 * c4 + 0) * 27 + k), __ldg(wr + (4 * c4 + 1) * 27 + k),
                                __ldg(wr + (4 * c4 + 2) * 27 + k), __ldg(wr + (4 * c4 + 3) * 27 + k));
    }
    __syncthreads();
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long)B * OH * OW) return;
    const int ow = (int)(p % OW), oh = (int)((p / OW) % OH), b = (int)(p / ((long)OW * OH));
    float xin[27];
    load_patch(x, ldx, b, oh, ow, H, W, xin);
    float4 acc[CO4];
#pragma unroll
    for (int c4 = 0; c4 < CO4; ++c4) acc[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 27; ++k)
#pragma unroll
        for (int c4 = 0; c4 < CO4; ++c4) {
            const float4 w = ws[k][c4];
            acc[c4].x += xin[k] * w.x; acc[c4].y += xin[k] * w.y; acc[c4].z += xin[k] * w.z; acc[c4].w += xin[k] * w.w;
        }
    float* o = y + p * ldy;
#pragma unroll
    for (int c4 = 0; c4 < CO4; ++c4) *reinterpret_cast<float4*>(o + 4 * c4) = acc[c4];
}

template <int CO4>
__global__ void __launch_bounds__(32 * CO4) stem_wgrad_kernel(const float* __restrict__ dy, long ldy,
                                                              const float* __restrict__ x, long ldx,
                                                              float* __restrict__ dwr, int B, int H, int W, int OH, int OW) {
    pdl_entry();
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const long P = (long)B * OH * OW;
    float acc[4][27];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 27; ++k) acc[i][k] = 0.f;
    for (long p = (long)blockIdx.x * 32 + lane; p < P; p += (long)gridDim.x * 32) {
        const int ow = (int)(p % OW), oh = (int)((p / OW) % OH), b = (int)(p / ((long)OW * OH));
        float xin[27];
        load_patch(x, ldx, b, oh, ow, H, W, xin);
        const float4 g = __ldg(reinterpret_cast<const float4*>(dy + p * ldy + 4 * warp));
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            acc[0][k] += g.x * xin[k]; acc[1][k] += g.y * xin[k]; acc[2][k] += g.z * xin[k]; acc[3][k] += g.w * xin[k];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            const float v = warp_sum(acc[i][k]);
            if (lane == 0) atomicAdd(dwr + (4 * warp + i) * 27 + k, v);
        }
}

}  // namespace

// 1 if the direct stem kernels take this geometry.
DFINE_API int dfine_stem_conv_supported(int Cin, int Cout, int KH, int KW, int stride, int pad_t, int pad_l, int pad_b,
                                        int pad_r, long ldy) {
    return Cin == 3 && KH == 3 && KW == 3 && stride == 2 && pad_t == 1 && pad_l == 1 && pad_b == 1 && pad_r == 1 &&
           (Cout == 16 || Cout == 24 || Cout == 32) && ldy % 4 == 0;
}

// y[B,OH,OW,Cout] (pixel stride ldy) = conv3x3 s2 p1 of x[B,H,W,3] (pixel stride ldx) with wr[Cout][3][3][3].
DFINE_API int dfine_stem_conv_fwd(const float* x, long ldx, const float* wr, float* y, long ldy, int B, int H, int W,
                                  int Cout, void* stream) {
    DFINE_REQUIRE(dfine_stem_conv_supported(3, Cout, 3, 3, 2, 1, 1, 1, 1, ldy) && ((uintptr_t)y % 16) == 0,
                  "stem_conv_fwd: Cout=%d ldy=%ld unsupported", Cout, ldy);
    const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
    const long P = (long)B * OH * OW;
    if (P == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ceil_div(P, 256);
    if (Cout == 16) launch_k(stem_fwd_kernel<4>, grid, 256, 0, st, x, ldx, wr, y, ldy, B, H, W, OH, OW);
    else if (Cout == 24) launch_k(stem_fwd_kernel<6>, grid, 256, 0, st, x, ldx, wr, y, ldy, B, H, W, OH, OW);
    else launch_k(stem_fwd_kernel<8>, grid, 256, 0, st, x, ldx, wr, y, ldy, B, H, W, OH, OW);
    DFINE_LAUNCH_CHECK("stem_conv_fwd");
    return 0;
}

// dwr[Cout][3][3][3] += sum over output pixels dy[p, co] * x[patch(p), k]; dwr zeroed or holding a running gradient.
DFINE_API int dfine_stem_conv_wgrad(const float* dy, long ldy, const float* x, long ldx, float* dwr, int B, int H, int W,
                                    int Cout, void* stream) {
    DFINE_REQUIRE(dfine_stem_conv_supported(3, Cout, 3, 3, 2, 1, 1, 1, 1, ldy) && ((uintptr_t)dy % 16) == 0,
                  "stem_conv_wgrad: Cout=%d ldy=%ld unsupported", Cout, ldy);
    const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
    const long P = (long)B * OH * OW;
    if (P == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    long g = (P + 31) / 32;
    if (g > 148 * 4) g = 148 * 4;
    const int grid = (int)g;
    if (Cout == 16) launch_k(stem_wgrad_kernel<4>, grid, 128, 0, st, dy, ldy, x, ldx, dwr, B, H, W, OH, OW);
    else if (Cout == 24) launch_k(stem_wgrad_kernel<6>, grid, 192, 0, st, dy, ldy, x, ldx, dwr, B, H, W, OH, OW);
    else launch_k(stem_wgrad_kernel<8>, grid, 256, 0, st, dy, ldy, x, ldx, dwr, B, H, W, OH, OW);
    DFINE_LAUNCH_CHECK("stem_conv_wgrad");
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// The two 2x2 convolutions of the stem (hgnetv2.py:125-140 `stem2a` C1 -> C1/2 and `stem2b` C1/2 -> C1, kernel 2,
// stride 1, on the input padded by one pixel at the bottom / right): 8..32 channels on the 320x320 map.  Through the
// tensor-core path these layers ran at 0.1-0.2 of the HBM roofline (250 us forward, 170-270 us data gradient, 365 us
// weight gradient per layer at batch 16: channel counts below the 32-channel k-block, N = 12 / 24 tiles of 128-row MMAs,
// every pixel fetched once per tap).  Direct fp32 kernel (158 us forward, 157 us data gradient):
//   conv2x2_kernel   forward and data gradient (the data gradient is the same correlation on dy with the taps flipped and
//                    the zero border at the top / left: the host passes flipped, transposed weights and origin = -1).
//                    A CTA stages a (8+1) x (64+1) pixel tile in shared memory (pixel pitch CI + 4 floats), every thread
//                    owns a 2 x 2 output block x all CO channels: 16 FMAs per 16-byte broadcast weight read.  Optional
//                    fused BatchNorm statistics (sum | sum of squares per channel, double), flushed once per CTA.
// The weight gradient stays on the tensor-core kernel (a direct version measured 384 us against its 365 us).
namespace {

constexpr int C2_TH = 8, C2_TW = 64, C2_THREADS = 128;

template <int CI, int CO>
__global__ void __launch_bounds__(C2_THREADS) conv2x2_kernel(const float* __restrict__ x, long ldx, const float* __restrict__ wt,
                                                             float* __restrict__ y, long ldy, double* __restrict__ stats, int B,
                                                             int H, int W, int origin, int tiles_w, int tiles_h) {
    pdl_entry();
    constexpr int PITCH = CI + 4;                       // floats per staged pixel: conflict-free float4 reads across lanes
    extern __shared__ float4 c2_smem4[];
    float* xs = reinterpret_cast<float*>(c2_smem4);                                   // [(TH+1)*(TW+1)][PITCH]
    float* ws = xs + (C2_TH + 1) * (C2_TW + 1) * PITCH;                               // [4*CI][CO]
    double* st = reinterpret_cast<double*>(ws + 4 * CI * CO);                         // [4 warps][2*CO]
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    for (int i = threadIdx.x; i < 4 * CI * CO / 4; i += C2_THREADS)
        reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(wt) + i);
    if (stats)
        for (int i = threadIdx.x; i < 4 * 2 * CO; i += C2_THREADS) st[i] = 0.0;
    const int total = B * tiles_h * tiles_w;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int b = t / (tiles_h * tiles_w), tt = t % (tiles_h * tiles_w);
        const int h0 = (tt / tiles_w) * C2_TH, w0 = (tt % tiles_w) * C2_TW;
        __syncthreads();                                // the previous tile's readers are done (and ws / st are written)
        // staged U float4 per thread at a time, all loads issued before the first shared-memory store (one load in
        // flight per thread left the kernel latency-bound at 1.2 TB/s)
        constexpr int NV = (C2_TH + 1) * (C2_TW + 1) * (CI / 4), U = 8;
        for (int base = 0; base < NV; base += C2_THREADS * U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * C2_THREADS + threadIdx.x;
                const int c4 = i % (CI / 4), px = i / (CI / 4);
                const int r = px / (C2_TW + 1), c = px % (C2_TW + 1);
                const int ih = h0 + origin + r, iw = w0 + origin + c;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < NV && ih >= 0 && ih < H && iw >= 0 && iw < W)
                    v[u] = __ldg(reinterpret_cast<const float4*>(x + (((long)b * H + ih) * W + iw) * ldx + 4 * c4));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * C2_THREADS + threadIdx.x;
                if (i < NV) *reinterpret_cast<float4*>(xs + (i / (CI / 4)) * PITCH + 4 * (i % (CI / 4))) = v[u];
            }
        }
        __syncthreads();
        float acc[2][2][CO];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int bb = 0; bb < 2; ++bb)
#pragma unroll
                for (int c = 0; c < CO; ++c) acc[a][bb][c] = 0.f;
        const float* xb = xs + ((2 * warp) * (C2_TW + 1) + 2 * lane) * PITCH;
#pragma unroll 1
        for (int c4 = 0; c4 < CI / 4; ++c4) {
            float4 xv[3][3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    xv[i][j] = *reinterpret_cast<const float4*>(xb + (i * (C2_TW + 1) + j) * PITCH + 4 * c4);
#pragma unroll
            for (int tap = 0; tap < 4; ++tap) {
                const int dh = tap / 2, dw = tap % 2;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const float* wrow = ws + ((tap * CI) + 4 * c4 + cc) * CO;
#pragma unroll
                    for (int q = 0; q < CO / 4; ++q) {
                        const float4 w = *reinterpret_cast<const float4*>(wrow + 4 * q);
#pragma unroll
                        for (int a = 0; a < 2; ++a)
#pragma unroll
                            for (int bb = 0; bb < 2; ++bb) {
                                const float4 xq = xv[a + dh][bb + dw];
                                const float xe = cc == 0 ? xq.x : (cc == 1 ? xq.y : (cc == 2 ? xq.z : xq.w));
                                acc[a][bb][4 * q] += xe * w.x; acc[a][bb][4 * q + 1] += xe * w.y;
                                acc[a][bb][4 * q + 2] += xe * w.z; acc[a][bb][4 * q + 3] += xe * w.w;
                            }
                    }
                }
            }
        }
        float s1[CO], s2[CO];
#pragma unroll
        for (int c = 0; c < CO; ++c) s1[c] = s2[c] = 0.f;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int bb = 0; bb < 2; ++bb) {
                const int oh = h0 + 2 * warp + a, ow = w0 + 2 * lane + bb;
                if (oh < H && ow < W) {
                    float* o = y + (((long)b * H + oh) * W + ow) * ldy;
#pragma unroll
                    for (int q = 0; q < CO / 4; ++q)
                        *reinterpret_cast<float4*>(o + 4 * q) =
                            make_float4(acc[a][bb][4 * q], acc[a][bb][4 * q + 1], acc[a][bb][4 * q + 2], acc[a][bb][4 * q + 3]);
                    if (stats) {
#pragma unroll
                        for (int c = 0; c < CO; ++c) { s1[c] += acc[a][bb][c]; s2[c] += acc[a][bb][c] * acc[a][bb][c]; }
                    }
                }
            }
        if (stats) {
#pragma unroll
            for (int c = 0; c < CO; ++c) {
                const float v1 = warp_sum(s1[c]), v2 = warp_sum(s2[c]);
                if (lane == 0) { st[warp * 2 * CO + c] += (double)v1; st[warp * 2 * CO + CO + c] += (double)v2; }
            }
        }
    }
    if (stats) {
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * CO; i += C2_THREADS) {
            const double v = st[i] + st[2 * CO + i] + st[4 * CO + i] + st[6 * CO + i];
            if (v != 0.0) atomicAdd(stats + i, v);
        }
    }
}

template <int CI, int CO>
int launch_conv2x2(const float* x, long ldx, const float* wt, float* y, long ldy, double* stats, int B, int H, int W, int origin,
                   cudaStream_t st) {
    const int smem = ((C2_TH + 1) * (C2_TW + 1) * (CI + 4) + 4 * CI * CO) * 4 + 4 * 2 * CO * 8;
    DFINE_SET_SMEM_ONCE((conv2x2_kernel<CI, CO>), smem, "conv2x2");
    const int tiles_w = ceil_div(W, C2_TW), tiles_h = ceil_div(H, C2_TH);
    const long total = (long)B * tiles_w * tiles_h;
    const int per_sm = smem > 110 * 1024 ? 1 : (smem > 72 * 1024 ? 2 : 3);
    const long cap = 148L * per_sm;
    launch_k(conv2x2_kernel<CI, CO>, (int)(total < cap ? total : cap), C2_THREADS, smem, st, x, ldx, wt, y, ldy, stats, B, H, W, origin,
                                                                                     tiles_w, tiles_h);
    return 0;
}

#define C2_DISPATCH(CI_, CO_, CALL)                                         \
    if (CI_ == 24 && CO_ == 12) return CALL(24, 12);                        \
    if (CI_ == 12 && CO_ == 24) return CALL(12, 24);                        \
    if (CI_ == 32 && CO_ == 16) return CALL(32, 16);                        \
    if (CI_ == 16 && CO_ == 32) return CALL(16, 32);                        \
    if (CI_ == 16 && CO_ == 8) return CALL(16, 8);                          \
    if (CI_ == 8 && CO_ == 16) return CALL(8, 16);

}  // namespace

// 1 if the direct 2x2 kernels take this geometry (kernel 2, stride 1, one pixel of zero padding at the bottom / right).
DFINE_API int dfine_conv2x2_supported(int Cin, int Cout, int KH, int KW, int stride, int pad_t, int pad_l, int pad_b, int pad_r,
                                      long ldx, long ldy) {
    const bool pair = (Cin == 24 && Cout == 12) || (Cin == 12 && Cout == 24) || (Cin == 32 && Cout == 16) ||
                      (Cin == 16 && Cout == 32) || (Cin == 16 && Cout == 8) || (Cin == 8 && Cout == 16);
    return pair && KH == 2 && KW == 2 && stride == 1 && pad_t == 0 && pad_l == 0 && pad_b == 1 && pad_r == 1 && ldx % 4 == 0 &&
           ldy % 4 == 0;
}

// y[B,H,W,Cout] (pixel stride ldy) = sum over the 4 taps (i, j) of x[b, h + origin + i, w + origin + j, :] . wt[(2i + j)][Cin][Cout]
// with zeros outside the image.  origin = 0 with wt[tap][ci][co] = w[co][ci][i][j]: the forward conv; origin = -1 on dy with
// wt[(2i + j)][co][ci] = w[co][ci][1 - i][1 - j]: its data gradient (Cin / Cout then name dy's / dx's channels).
// stats (optional, double [2*Cout], accumulated): per-channel sum | sum of squares of y.
DFINE_API int dfine_conv2x2(const float* x, long ldx, const float* wt, float* y, long ldy, double* stats, int B, int H, int W,
                            int Cin, int Cout, int origin, void* stream) {
    DFINE_REQUIRE(dfine_conv2x2_supported(Cin, Cout, 2, 2, 1, 0, 0, 1, 1, ldx, ldy) && (origin == 0 || origin == -1) &&
                      ((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0 && ((uintptr_t)wt % 16) == 0,
                  "conv2x2: %d -> %d, strides %ld / %ld unsupported", Cin, Cout, ldx, ldy);
    if ((long)B * H * W == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
#define C2_FWD(CI_, CO_) launch_conv2x2<CI_, CO_>(x, ldx, wt, y, ldy, stats, B, H, W, origin, st)
    auto run = [&]() -> int { C2_DISPATCH(Cin, Cout, C2_FWD) return -1; };
    const int rc = run();
#undef C2_FWD
    if (rc) return rc;
    DFINE_LAUNCH_CHECK("conv2x2");
    return 0;
}
