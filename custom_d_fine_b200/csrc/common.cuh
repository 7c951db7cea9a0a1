// Shared helpers for libdfine_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cmath>

#define DFINE_API extern "C" __attribute__((visibility("default")))

void dfine_set_error(const char* fmt, ...);

// argument check: negative return codes are caller errors (see include/dfine_sm100.h)
#define DFINE_REQUIRE(cond, ...)                     \
    do {                                             \
        if (!(cond)) {                               \
            dfine_set_error(__VA_ARGS__);            \
            return -1;                               \
        }                                            \
    } while (0)

#define DFINE_LAUNCH_CHECK(name)                                                         \
    do {                                                                                 \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) {                                                        \
            dfine_set_error("%s: %s", name, cudaGetErrorString(e__));                    \
            return (int)e__;                                                             \
        }                                                                                \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (call site, device): a process that later switches to
// another GPU must set the attribute there too.  `kernel` may be a parenthesised template-id.
#define DFINE_SET_SMEM_ONCE(kernel, bytes, what)                                                                   \
    do {                                                                                                           \
        static unsigned long long done__ = 0ull;                                                                   \
        int dev__ = 0;                                                                                             \
        cudaGetDevice(&dev__);                                                                                     \
        if (!((done__ >> (dev__ & 63)) & 1ull)) {                                                                  \
            cudaError_t e__ = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
            if (e__ != cudaSuccess) {                                                                              \
                dfine_set_error("%s: smem attribute (%d B): %s", what, (int)(bytes), cudaGetErrorString(e__));     \
                return (int)e__;                                                                                   \
            }                                                                                                      \
            done__ |= 1ull << (dev__ & 63);                                                                        \
        }                                                                                                          \
    } while (0)

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch ---------------------------------------------------------------------------------
// A train step is ~2000 short kernels in one stream order; between two dependent kernels the GPU idles for the launch
// latency of the second one.  Every kernel of the library is launched with the programmatic-stream-serialization
// attribute (also recorded by CUDA-graph capture as a programmatic edge) and starts with pdl_entry(): its CTAs may
// become resident while the previous kernel drains, wait for that kernel's completion + memory flush
// (griddepcontrol.wait) BEFORE touching global memory, then allow their own successor to be scheduled.  A kernel
// launched without the attribute (or after a non-kernel stream item) executes both instructions as no-ops.
// DFINE_PDL=0 launches without the attribute (plain stream serialization).
__device__ __forceinline__ void pdl_entry() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
static inline bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("DFINE_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
// Grid of a grid-strided elementwise kernel: enough CTAs for n work items, capped at TWO FULL WAVES of what is resident
// with this kernel's registers / shared memory (occupancy x SM count).  Every CTA of a capped grid does the same amount of
// work, so a cap that is not a multiple of the resident count leaves the last round partly empty: the old fixed cap of
// 148 x 16 CTAs cost the 46-register BatchNorm-backward kernel (5 CTAs per SM) a fourth round for 3.2 rounds of work.
template <typename K>
static inline int ew_grid_k(K kernel, long n, int block, size_t smem) {
    int occ = 0, dev = 0, sms = 148;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, smem) != cudaSuccess || occ < 1) occ = 1;
    static int sm_cache[64] = {0};
    cudaGetDevice(&dev);
    if (sm_cache[dev & 63] == 0 &&
        (cudaDeviceGetAttribute(&sm_cache[dev & 63], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_cache[dev & 63] <= 0))
        sm_cache[dev & 63] = 148;
    sms = sm_cache[dev & 63];
    const long g = (n + block - 1) / block, cap = 2L * occ * sms;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                   Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SILU = 2, ACT_GELU = 3 };

__device__ __forceinline__ float act_fwd(float z, int act) {
    switch (act) {
        case ACT_RELU: return z > 0.f ? z : 0.f;
        case ACT_SILU: return __fdividef(z, 1.f + __expf(-z));   // fast-math intrinsics: 2 ulp, HBM-bound callers
        case ACT_GELU: return 0.5f * z * (1.f + erff(z * 0.70710678118654752f));
        default: return z;
    }
}
// d act(z) / dz
__device__ __forceinline__ float act_bwd(float z, int act) {
    switch (act) {
        case ACT_RELU: return z > 0.f ? 1.f : 0.f;
        case ACT_SILU: {
            float s = __fdividef(1.f, 1.f + __expf(-z));
            return s * (1.f + z * (1.f - s));
        }
        case ACT_GELU: {
            float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752f));
            float pdf = 0.39894228040143268f * expf(-0.5f * z * z);
            return cdf + z * pdf;
        }
        default: return 1.f;
    }
}

// Train-mode BatchNorm finalize of ONE channel from its complete fp64 sums (sum x, sum x^2 over M pixels): mean / invstd,
// the folded scale / shift, and the running-statistics update (momentum, unbiased variance) — torch.nn.BatchNorm2d's
// arithmetic (hgnetv2.py:65 / hybrid_encoder.py:36).  Shared by bn_finalize_kernel and the conv kernel's last-CTA tail.
__device__ __forceinline__ void bn_finalize_channel(double sum, double sumsq, int c, const float* __restrict__ weight,
                                                    const float* __restrict__ bias, float* __restrict__ running_mean,
                                                    float* __restrict__ running_var, float* __restrict__ mean_out,
                                                    float* __restrict__ invstd_out, float* __restrict__ scale_out,
                                                    float* __restrict__ shift_out, long M, float momentum, float eps) {
    const double mean = sum / (double)M;
    double var = sumsq / (double)M - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float w = weight ? weight[c] : 1.f, b = bias ? bias[c] : 0.f;
    mean_out[c] = (float)mean;
    invstd_out[c] = invstd;
    const float sc = w * invstd;
    scale_out[c] = sc;
    shift_out[c] = b - (float)mean * sc;
    if (running_mean) {
        const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
