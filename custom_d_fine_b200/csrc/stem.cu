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
    if (Cout == 16) stem_fwd_kernel<4><<<grid, 256, 0, st>>>(x, ldx, wr, y, ldy, B, H, W, OH, OW);
    else if (Cout == 24) stem_fwd_kernel<6><<<grid, 256, 0, st>>>(x, ldx, wr, y, ldy, B, H, W, OH, OW);
    else stem_fwd_kernel<8><<<grid, 256, 0, st>>>(x, ldx, wr, y, ldy, B, H, W, OH, OW);
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
    if (Cout == 16) stem_wgrad_kernel<4><<<grid, 128, 0, st>>>(dy, ldy, x, ldx, dwr, B, H, W, OH, OW);
    else if (Cout == 24) stem_wgrad_kernel<6><<<grid, 192, 0, st>>>(dy, ldy, x, ldx, dwr, B, H, W, OH, OW);
    else stem_wgrad_kernel<8><<<grid, 256, 0, st>>>(dy, ldy, x, ldx, dwr, B, H, W, OH, OW);
    DFINE_LAUNCH_CHECK("stem_conv_wgrad");
    return 0;
}
