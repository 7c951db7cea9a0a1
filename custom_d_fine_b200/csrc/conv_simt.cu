// CUDA-core implicit-GEMM convolution / linear kernels (fp32, NHWC), forward, data-gradient and
// weight-gradient.  They cover every dense conv and nn.Linear shape of the model and are the path
// for the shapes the tcgen05 kernels (gemm_tc.cu) do not take: tiny-channel stem convs
// (Cin 3..48, hgnetv2.py:115-166), strided / asymmetric-padded convs and ragged tails.
//
// One 64x64x16 tiling, 256 threads, 4x4 register micro-tile, operands staged k-major in shared
// memory; the three entry points differ only in how tile elements are gathered:
//   fwd   : C[m=(b,oh,ow), n=co]      = sum_{k=(kh,kw,ci)} x[b,oh*s+kh-pt,ow*s+kw-pl,ci] * Wr[co,kh,kw,ci]
//   dgrad : C[m=(b,ih,iw), n=ci]      = sum_{k=(kh,kw,co)} dy[b,(ih+pt-kh)/s,(iw+pl-kw)/s,co] * Wr[co,kh,kw,ci]
//   wgrad : C[m=co, n=(kh,kw,ci)]    += sum_{k=pixel}      dy[pixel,co] * x[pixel shifted by tap, ci]   (split-K, atomics)
// Wr is the conv weight re-laid as [Cout, KH, KW, Cin] (an nn.Linear weight [N,K] already is that).
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct ConvGeom {
    int B, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad_t, pad_l;
    long ldx, ldy;  // pixel (row) strides of the input and output tensors, in elements
};

__device__ __forceinline__ void mma_tile(const float (*As)[BM + 4], const float (*Bs)[BN + 4], float acc[4][4], int ty,
                                         int tx) {
#pragma unroll
    for (int k = 0; k < BK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
    }
}

// ---- forward / data-gradient: A is a spatial gather, B is the weight ------------------------------
// DGRAD=false: src = x  (geometry as declared).  DGRAD=true: src = dy, output pixels are input pixels.
template <bool DGRAD>
__global__ void __launch_bounds__(NT) conv_gemm_kernel(const float* __restrict__ src, const float* __restrict__ wr,
                                                       const float* __restrict__ bias, float* __restrict__ dst,
                                                       ConvGeom g, int act) {
    pdl_entry();
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
    const int KK = g.KH * g.KW;
    const int CK = DGRAD ? g.Cout : g.Cin;         // channels of the gathered tensor
    const int N = DGRAD ? g.Cin : g.Cout;
    const int K = KK * CK;
    const int PH = DGRAD ? g.H : g.OH, PW = DGRAD ? g.W : g.OW;  // output-pixel grid of this GEMM
    const long M = (long)g.B * PH * PW;
    const long m0 = (long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // A loads: element e = tid + i*256 -> kk = e % 16, mm = e / 16  (threads run along k: contiguous channels)
    const int a_kk = tid % BK;
    int a_b[4], a_h[4], a_w[4];
    bool a_ok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long m = m0 + tid / BK + i * (NT / BK);
        a_ok[i] = m < M;
        const long mm = a_ok[i] ? m : 0;
        a_w[i] = (int)(mm % PW);
        const long t = mm / PW;
        a_h[i] = (int)(t % PH);
        a_b[i] = (int)(t / PH);
    }
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += BK) {
        {
            const int k = k0 + a_kk;
            const bool kok = k < K;
            const int c = kok ? k % CK : 0, tap = kok ? k / CK : 0;
            const int kh = tap / g.KW, kw = tap % g.KW;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float v = 0.f;
                if (kok && a_ok[i]) {
                    if (!DGRAD) {
                        const int ih = a_h[i] * g.stride + kh - g.pad_t, iw = a_w[i] * g.stride + kw - g.pad_l;
                        if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
                            v = __ldg(src + (((long)a_b[i] * g.H + ih) * g.W + iw) * g.ldx + c);
                    } else {
                        const int th = a_h[i] + g.pad_t - kh, tw = a_w[i] + g.pad_l - kw;
                        if (th >= 0 && tw >= 0 && th % g.stride == 0 && tw % g.stride == 0) {
                            const int oh = th / g.stride, ow = tw / g.stride;
                            if (oh < g.OH && ow < g.OW)
                                v = __ldg(src + (((long)a_b[i] * g.OH + oh) * g.OW + ow) * g.ldy + c);
                        }
                    }
                }
                As[a_kk][tid / BK + i * (NT / BK)] = v;
            }
        }
        if (!DGRAD) {
            // B(k,n) = wr[n*K + k]: threads along k
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int kk = tid % BK, nn = tid / BK + i * (NT / BK);
                const int k = k0 + kk, n = n0 + nn;
                Bs[kk][nn] = (k < K && n < N) ? __ldg(wr + (long)n * K + k) : 0.f;
            }
        } else {
            // B(k=(tap,co), n=ci) = wr[(co*KK + tap)*Cin + ci]: threads along n
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int nn = tid % BN, kk = tid / BN + i * (NT / BN);
                const int k = k0 + kk, n = n0 + nn;
                float v = 0.f;
                if (k < K && n < N) {
                    const int co = k % CK, tap = k / CK;
                    v = __ldg(wr + ((long)co * KK + tap) * g.Cin + n);
                }
                Bs[kk][nn] = v;
            }
        }
        __syncthreads();
        mma_tile(As, Bs, acc, ty, tx);
        __syncthreads();
    }
    const long ldo = DGRAD ? g.ldx : g.ldy;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (!DGRAD) {
                if (bias) v += __ldg(bias + n);
                v = act_fwd(v, act);
            }
            dst[m * ldo + n] = v;
        }
    }
}

// ---- weight gradient: reduction over pixels, split across blockIdx.z ------------------------------
__global__ void __launch_bounds__(NT) conv_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                        float* __restrict__ dwr, ConvGeom g, long pix_per_split) {
    pdl_entry();
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
    const int KK = g.KH * g.KW;
    const int Mo = g.Cout, No = KK * g.Cin;
    const long P = (long)g.B * g.OH * g.OW;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const long p_begin = (long)blockIdx.z * pix_per_split;
    const long p_end = p_begin + pix_per_split < P ? p_begin + pix_per_split : P;
    // both operands are contiguous along the tile's non-reduction dim: threads run along m / n
    const int mm = tid % BM, nn = tid % BN, kq = tid / BM;  // kq in [0,4)
    const int co = m0 + mm, nidx = n0 + nn;
    const bool m_ok = co < Mo, n_ok = nidx < No;
    const int ci = n_ok ? nidx % g.Cin : 0, tap = n_ok ? nidx / g.Cin : 0;
    const int kh = tap / g.KW, kw = tap % g.KW;
    float acc[4][4] = {};
    for (long p0 = p_begin; p0 < p_end; p0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int kk = kq + i * (NT / BM);
            const long p = p0 + kk;
            float a = 0.f, b = 0.f;
            if (p < p_end) {
                if (m_ok) a = __ldg(dy + p * g.ldy + co);
                if (n_ok) {
                    const int ow = (int)(p % g.OW);
                    const long t = p / g.OW;
                    const int oh = (int)(t % g.OH), bi = (int)(t / g.OH);
                    const int ih = oh * g.stride + kh - g.pad_t, iw = ow * g.stride + kw - g.pad_l;
                    if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
                        b = __ldg(x + (((long)bi * g.H + ih) * g.W + iw) * g.ldx + ci);
                }
            }
            As[kk][mm] = a;
            Bs[kk][nn] = b;
        }
        __syncthreads();
        mma_tile(As, Bs, acc, ty, tx);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= Mo) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < No) atomicAdd(dwr + (long)m * No + n, acc[i][j]);
        }
    }
}

// out[c] += sum_rows x[row*ld + c]   (bias gradients)
__global__ void __launch_bounds__(NT) colsum_kernel(const float* __restrict__ x, float* __restrict__ out, long M,
                                                    int C, long ld, long rows_per_cta) {
    pdl_entry();
    const long r0 = (long)blockIdx.x * rows_per_cta;
    const long r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
    for (int c = threadIdx.x; c < C; c += NT) {
        float s = 0.f;
        for (long r = r0; r < r1; ++r) s += __ldg(x + r * ld + c);
        atomicAdd(out + c, s);
    }
}

int check_geom(const ConvGeom& g, const char* who) {
    if (g.B < 0 || g.H <= 0 || g.W <= 0 || g.Cin <= 0 || g.Cout <= 0 || g.KH <= 0 || g.KW <= 0 || g.stride <= 0 ||
        g.OH <= 0 || g.OW <= 0 || g.ldx < g.Cin || g.ldy < g.Cout) {
        dfine_set_error("%s: bad geometry", who);
        return -1;
    }
    return 0;
}

}  // namespace

#define GEOM_ARGS                                                                                                  \
    int B, int H, int W, int Cin, int OH, int OW, int Cout, int KH, int KW, int stride, int pad_t, int pad_l,       \
        long ldx, long ldy
#define GEOM_INIT {B, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad_t, pad_l, ldx, ldy}

// y[b,oh,ow,co] = act(conv(x, Wr) + bias).  x pixel stride ldx (>= Cin), y pixel stride ldy (>= Cout).
// nn.Linear: B=rows, H=W=OH=OW=KH=KW=1.
DFINE_API int dfine_conv_fwd_simt(const float* x, const float* wr, const float* bias, float* y, GEOM_ARGS, int act,
                                  void* stream) {
    ConvGeom g = GEOM_INIT;
    if (check_geom(g, "conv_fwd_simt")) return -1;
    const long M = (long)B * OH * OW;
    if (M == 0) return 0;
    dim3 grid(ceil_div(M, BM), ceil_div(Cout, BN));
    launch_k(conv_gemm_kernel<false>, grid, NT, 0, (cudaStream_t)stream, x, wr, bias, y, g, act);
    DFINE_LAUNCH_CHECK("conv_fwd_simt");
    return 0;
}

// dx[b,ih,iw,ci] (pixel stride ldx) from dy (pixel stride ldy); fully overwrites dx's Cin channels.
DFINE_API int dfine_conv_dgrad_simt(const float* dy, const float* wr, float* dx, GEOM_ARGS, void* stream) {
    ConvGeom g = GEOM_INIT;
    if (check_geom(g, "conv_dgrad_simt")) return -1;
    const long M = (long)B * H * W;
    if (M == 0) return 0;
    dim3 grid(ceil_div(M, BM), ceil_div(Cin, BN));
    launch_k(conv_gemm_kernel<true>, grid, NT, 0, (cudaStream_t)stream, dy, wr, nullptr, dx, g, 0);
    DFINE_LAUNCH_CHECK("conv_dgrad_simt");
    return 0;
}

// dwr [Cout,KH,KW,Cin] += ... ; zero-initialised by the caller.
DFINE_API int dfine_conv_wgrad_simt(const float* dy, const float* x, float* dwr, GEOM_ARGS, void* stream) {
    ConvGeom g = GEOM_INIT;
    if (check_geom(g, "conv_wgrad_simt")) return -1;
    const long P = (long)B * OH * OW;
    if (P == 0) return 0;
    const int gx = ceil_div(Cout, BM), gy = ceil_div((long)KH * KW * Cin, BN);
    long splits = (148L * 4) / ((long)gx * gy);
    if (splits < 1) splits = 1;
    long pps = (P + splits - 1) / splits;
    pps = (pps + BK - 1) / BK * BK;
    if (pps < 8 * BK) pps = 8 * BK;
    dim3 grid(gx, gy, ceil_div(P, pps));
    launch_k(conv_wgrad_kernel, grid, NT, 0, (cudaStream_t)stream, dy, x, dwr, g, pps);
    DFINE_LAUNCH_CHECK("conv_wgrad_simt");
    return 0;
}

// out [C] += column sums of x [M, C] with row stride ld; zero-initialised by the caller.
DFINE_API int dfine_colsum(const float* x, float* out, long M, int C, long ld, void* stream) {
    if (M == 0 || C == 0) return 0;
    long rpc = (M + 148L * 2 - 1) / (148L * 2);
    if (rpc < 32) rpc = 32;
    launch_k(colsum_kernel, ceil_div(M, rpc), NT, 0, (cudaStream_t)stream, x, out, M, C, ld, rpc);
    DFINE_LAUNCH_CHECK("colsum");
    return 0;
}
