// Fused optimizer step over flat parameter arenas: gradient-norm clip + AdamW + EMA blend + grad zeroing.
//
// Replaces the device part of the reference's optimizer_step (train.py:512-535:
// clip_grad_norm_(0.1) -> AdamW.step -> zero_grad) and ModelEMA.update (train.py:62-73), which the
// reference issues as ~2 k tiny kernels per step (per-tensor Python loop for the EMA, foreach kernels for
// AdamW).  Parameters, gradients, both Adam moments and the EMA copy live in five flat fp32 arenas with the
// same element order, so one pass reads p, g, m, v, ema and writes p, m, v, ema, g(=0): 36 bytes per
// parameter, HBM-bound (19.6 M parameters of D-FINE-m = 0.7 GB per step).
//
// Every scalar that changes between steps (learning rate, weight decay, step count, EMA momentum, the
// squared gradient norm) is read from device memory, so the launches can be replayed from a CUDA graph.
#include "common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int NT = 256;

__global__ void __launch_bounds__(NT) sumsq_kernel(const float4* __restrict__ g, long n4, double* __restrict__ out) {
    pdl_entry();
    float s = 0.f;
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < n4; i += (long)gridDim.x * NT) {
        const float4 v = __ldg(g + i);
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    __shared__ double part[NT / 32];
    double d = warp_sum_d((double)s);
    if (threadIdx.x % 32 == 0) part[threadIdx.x / 32] = d;
    __syncthreads();
    if (threadIdx.x < 32) {
        d = threadIdx.x < NT / 32 ? part[threadIdx.x] : 0.0;
        d = warp_sum_d(d);
        if (threadIdx.x == 0) atomicAdd(out, d);
    }
}

__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// four weights x scale -> fp16 hi (RN) and lo (RN of the remainder), packed 4 halves per uint2
__device__ __forceinline__ void split_f16x4(const float4& w, float scale, uint2& hi, uint2& lo) {
    const float v[4] = {w.x * scale, w.y * scale, w.z * scale, w.w * scale};
    unsigned short h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __half hh = __float2half_rn(v[j]);
        h[j] = __half_as_ushort(hh);
        l[j] = __half_as_ushort(__float2half_rn(v[j] - __half2float(hh)));
    }
    hi = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
    lo = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
}

__global__ void __launch_bounds__(NT) f16_split_flat_kernel(const float4* __restrict__ w, uint2* __restrict__ hi,
                                                            uint2* __restrict__ lo, long n4, float scale) {
    pdl_entry();
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < n4; i += (long)gridDim.x * NT) {
        uint2 h, l;
        split_f16x4(__ldg(w + i), scale, h, l);
        hi[i] = h;
        lo[i] = l;
    }
}

// hyper (device, 4 floats): lr, weight_decay, step (1-based, already incremented), ema momentum
__global__ void __launch_bounds__(NT) adamw_ema_kernel(float4* __restrict__ p, float4* __restrict__ g,
                                                       float4* __restrict__ m, float4* __restrict__ v,
                                                       float4* __restrict__ ema, long n4,
                                                       const float* __restrict__ hyper,
                                                       const double* __restrict__ gnorm_sq, float max_norm,
                                                       float beta1, float beta2, float eps, int zero_grad,
                                                       float4* __restrict__ hi, float4* __restrict__ lo, int plane_mode,
                                                       float plane_scale) {
    pdl_entry();
    const float lr = __ldg(hyper + 0), wd = __ldg(hyper + 1), step = __ldg(hyper + 2), em = __ldg(hyper + 3);
    // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
    float coef = 1.f;
    if (gnorm_sq != nullptr && max_norm > 0.f) {
        const float total = (float)sqrt(*gnorm_sq);
        coef = fminf(max_norm / (total + 1e-6f), 1.f);
    }
    // torch.optim.AdamW (non-amsgrad): bias corrections from the step count
    const float bc1 = 1.f - powf(beta1, step), bc2 = 1.f - powf(beta2, step);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    const float decay = 1.f - lr * wd;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < n4; i += (long)gridDim.x * NT) {
        float4 pv = p[i], gv = g[i], mv = m[i], vv = v[i];
        float* pp = reinterpret_cast<float*>(&pv);
        float* gp = reinterpret_cast<float*>(&gv);
        float* mp = reinterpret_cast<float*>(&mv);
        float* vp = reinterpret_cast<float*>(&vv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gr = gp[j] * coef;
            float x = pp[j] * decay;
            mp[j] = beta1 * mp[j] + (1.f - beta1) * gr;
            vp[j] = beta2 * vp[j] + (1.f - beta2) * gr * gr;
            const float denom = sqrtf(vp[j]) * inv_sqrt_bc2 + eps;
            pp[j] = x - step_size * (mp[j] / denom);
        }
        p[i] = pv; m[i] = mv; v[i] = vv;
        if (zero_grad) g[i] = zero;
        if (hi != nullptr && plane_mode == 1) {     // 3xFP16 operand planes of the new weights x plane_scale (fp16 arenas)
            uint2 h16, l16;
            split_f16x4(pv, plane_scale, h16, l16);
            reinterpret_cast<uint2*>(hi)[i] = h16;
            reinterpret_cast<uint2*>(lo)[i] = l16;
        } else if (hi != nullptr) {     // 3xTF32 operand planes of the new weights (hi = RN tf32, lo = exact remainder)
            float4 h;
            h.x = tf32_rn(pv.x); h.y = tf32_rn(pv.y); h.z = tf32_rn(pv.z); h.w = tf32_rn(pv.w);
            hi[i] = h;
            lo[i] = make_float4(pv.x - h.x, pv.y - h.y, pv.z - h.z, pv.w - h.w);
        }
        if (ema != nullptr) {
            float4 ev = ema[i];
            ev.x = em * ev.x + (1.f - em) * pv.x; ev.y = em * ev.y + (1.f - em) * pv.y;
            ev.z = em * ev.z + (1.f - em) * pv.z; ev.w = em * ev.w + (1.f - em) * pv.w;
            ema[i] = ev;
        }
    }
}

// ema = m * ema + (1 - m) * src   (floating-point buffers: BatchNorm running statistics etc., train.py:70-73)
__global__ void __launch_bounds__(NT) ema_blend_kernel(float4* __restrict__ ema, const float4* __restrict__ src, long n4,
                                                       const float* __restrict__ momentum) {
    pdl_entry();
    const float em = __ldg(momentum);
    for (long i = (long)blockIdx.x * NT + threadIdx.x; i < n4; i += (long)gridDim.x * NT) {
        float4 e = ema[i];
        const float4 s = __ldg(src + i);
        e.x = em * e.x + (1.f - em) * s.x; e.y = em * e.y + (1.f - em) * s.y;
        e.z = em * e.z + (1.f - em) * s.z; e.w = em * e.w + (1.f - em) * s.w;
        ema[i] = e;
    }
}

int grid_for(long n4) {
    long g = (n4 + NT - 1) / NT;
    const long cap = 148L * 8;   // 8 resident 256-thread CTAs per SM, grid-stride beyond that
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace

// out[0] += sum(g[i]^2); n must be a multiple of 4 (arenas are padded), out zeroed by the caller.
DFINE_API int dfine_sumsq(const float* g, long n, double* out, void* stream) {
    DFINE_REQUIRE(n % 4 == 0 && ((uintptr_t)g % 16) == 0, "sumsq: arena must be 16-byte aligned and padded to 4 floats");
    if (n == 0) return 0;
    launch_k(sumsq_kernel, grid_for(n / 4), NT, 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(g), n / 4, out);
    DFINE_LAUNCH_CHECK("sumsq");
    return 0;
}

// One AdamW (+clip, +EMA, +zero_grad) pass over a parameter group laid out flat.
//   hyper     device float[4] = {lr, weight_decay, step, ema_momentum}
//   gnorm_sq  device double, squared global gradient norm (null: no clipping)
//   ema       null: no EMA blend
//   hi, lo    null, or arenas receiving the operand split of the updated parameters (same element order): the forward
//             GEMMs of the next step read their weight planes from there, no per-layer split launches.
//             plane_mode 0: fp32 arenas, 3xTF32 split (hi = RN tf32, lo = remainder); plane_mode 1: fp16 arenas
//             (n halves each), 3xFP16 split of w * plane_scale
DFINE_API int dfine_adamw_ema(float* p, float* g, float* m, float* v, float* ema, long n, const float* hyper,
                              const double* gnorm_sq, float max_norm, float beta1, float beta2, float eps,
                              int zero_grad, void* hi, void* lo, int plane_mode, float plane_scale, void* stream) {
    DFINE_REQUIRE(n % 4 == 0, "adamw_ema: arena length must be a multiple of 4");
    DFINE_REQUIRE(((uintptr_t)p % 16) == 0 && ((uintptr_t)g % 16) == 0 && ((uintptr_t)m % 16) == 0 &&
                      ((uintptr_t)v % 16) == 0 && ((uintptr_t)ema % 16) == 0 && ((uintptr_t)hi % 16) == 0 &&
                      ((uintptr_t)lo % 16) == 0 && ((hi == nullptr) == (lo == nullptr)),
                  "adamw_ema: arenas must be 16-byte aligned");
    if (n == 0) return 0;
    launch_k(adamw_ema_kernel, grid_for(n / 4), NT, 0, (cudaStream_t)stream, 
        reinterpret_cast<float4*>(p), reinterpret_cast<float4*>(g), reinterpret_cast<float4*>(m),
        reinterpret_cast<float4*>(v), reinterpret_cast<float4*>(ema), n / 4, hyper, gnorm_sq, max_norm, beta1, beta2,
        eps, zero_grad, reinterpret_cast<float4*>(hi), reinterpret_cast<float4*>(lo), plane_mode, plane_scale);
    DFINE_LAUNCH_CHECK("adamw_ema");
    return 0;
}

// hi16 / lo16 (n halves each) = the 3xFP16 split of w * scale over a flat arena (n % 4 == 0).
DFINE_API int dfine_f16_split_flat(const float* w, void* hi16, void* lo16, long n, float scale, void* stream) {
    DFINE_REQUIRE(n % 4 == 0 && ((uintptr_t)w % 16) == 0 && ((uintptr_t)hi16 % 8) == 0 && ((uintptr_t)lo16 % 8) == 0 && scale > 0.f,
                  "f16_split_flat: alignment / n %% 4");
    if (n == 0) return 0;
    launch_k(f16_split_flat_kernel, grid_for(n / 4), NT, 0, (cudaStream_t)stream, 
        reinterpret_cast<const float4*>(w), reinterpret_cast<uint2*>(hi16), reinterpret_cast<uint2*>(lo16), n / 4, scale);
    DFINE_LAUNCH_CHECK("f16_split_flat");
    return 0;
}

DFINE_API int dfine_ema_blend(float* ema, const float* src, long n, const float* momentum, void* stream) {
    DFINE_REQUIRE(n % 4 == 0 && ((uintptr_t)ema % 16) == 0 && ((uintptr_t)src % 16) == 0, "ema_blend: alignment");
    if (n == 0) return 0;
    launch_k(ema_blend_kernel, grid_for(n / 4), NT, 0, (cudaStream_t)stream, reinterpret_cast<float4*>(ema),
                                                                       reinterpret_cast<const float4*>(src), n / 4,
                                                                       momentum);
    DFINE_LAUNCH_CHECK("ema_blend");
    return 0;
}
