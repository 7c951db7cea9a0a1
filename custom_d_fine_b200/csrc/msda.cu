// Multi-scale deformable attention, forward + backward, for sm_100a.
//
// Replaces the reference's MSDeformableAttention core: softmax over the P sampling points,
// sampling-location arithmetic (dfine_decoder.py:137-166) and deformable_attention_core_func_v2
// (arch/utils.py:191-264: per-level NCHW copy + grid_sample(bilinear, zeros, align_corners=False)
// + concat + weighted sum) with ONE kernel that gathers straight from the token-major memory
// tensor [B, L, heads, head_dim].  A (token, head) row is head_dim*4 bytes (128 B for head_dim 32),
// so every bilinear corner is one fully used cache line, fetched as 16-byte vectors by
// head_dim/4 adjacent lanes.
//
// Work decomposition: one "group" of TPG = head_dim/4 lanes per (batch, query, head); 256-thread
// CTAs = 256/TPG groups.  Phase 1: the group's lanes split the P points, compute the softmax
// (xor-shuffles inside the group), pixel coordinates and level geometry, and park them in shared
// memory.  Phase 2: every lane walks the P points and accumulates its float4 of channels.
// Backward re-runs phase 1, scatters d(value) with vector red.global.add.v4.f32 and reduces
// d(weight)/d(x)/d(y) across the group with shuffles; phase 3 applies the softmax/location chain rule.
#include "common.cuh"

namespace {

constexpr int MAX_LEVELS = 4;
constexpr int MAX_POINTS = 16;

struct MsdaGeom {
    int n_levels;
    int H[MAX_LEVELS], W[MAX_LEVELS], start[MAX_LEVELS], pts[MAX_LEVELS];
    int P;  // total points
};

struct PointSlot {  // one sampling point of one (b,q,h), shared-memory resident
    float x, y, w;  // pixel coords (align_corners=False convention) and softmax weight
    int start, W, H;
    float dscale_x, dscale_y;  // d(pixel)/d(offset) for the backward chain rule
};

struct FwdSlot {  // forward only: the four bilinear corners resolved once per point instead of once per lane
    int tok[4];   // token index of each corner (0 where the corner is outside the map: weight 0, load still legal)
    float w[4];   // softmax weight x bilinear weight (0 outside)
};

template <int TPG>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = TPG / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int TPG>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = TPG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Phase 1 shared by fwd and bwd.  `lane` = lane inside the group.
__device__ __forceinline__ void put_slot(PointSlot* slots, int p, const PointSlot& s) { slots[p] = s; }
__device__ __forceinline__ void put_slot(FwdSlot* slots, int p, const PointSlot& s) {
    const float xf = floorf(s.x), yf = floorf(s.y);
    const int x0 = (int)xf, y0 = (int)yf;
    const float fx = s.x - xf, fy = s.y - yf;
    const bool vx0 = x0 >= 0 && x0 < s.W, vx1 = x0 + 1 >= 0 && x0 + 1 < s.W;
    const bool vy0 = y0 >= 0 && y0 < s.H, vy1 = y0 + 1 >= 0 && y0 + 1 < s.H;
    const int t00 = s.start + y0 * s.W + x0;
    FwdSlot f;
    f.tok[0] = (vx0 && vy0) ? t00 : 0;
    f.tok[1] = (vx1 && vy0) ? t00 + 1 : 0;
    f.tok[2] = (vx0 && vy1) ? t00 + s.W : 0;
    f.tok[3] = (vx1 && vy1) ? t00 + s.W + 1 : 0;
    f.w[0] = (vx0 && vy0) ? s.w * (1.f - fx) * (1.f - fy) : 0.f;
    f.w[1] = (vx1 && vy0) ? s.w * fx * (1.f - fy) : 0.f;
    f.w[2] = (vx0 && vy1) ? s.w * (1.f - fx) * fy : 0.f;
    f.w[3] = (vx1 && vy1) ? s.w * fx * fy : 0.f;
    slots[p] = f;
}

template <int TPG, typename SlotT>
__device__ __forceinline__ void msda_prepare(const MsdaGeom& g, const float* __restrict__ off_row,
                                             const float* __restrict__ logit_row,
                                             const float* __restrict__ ref4,
                                             const float* __restrict__ pscale, float offset_scale,
                                             int lane, SlotT* slots) {
    constexpr int NS = MAX_POINTS / TPG;
    float lg[NS];
    float m = -INFINITY;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        int p = s * TPG + lane;
        lg[s] = p < g.P ? __ldg(logit_row + p) : -INFINITY;
        m = fmaxf(m, lg[s]);
    }
    m = group_max<TPG>(m);
    float sum = 0.f;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        lg[s] = (s * TPG + lane) < g.P ? expf(lg[s] - m) : 0.f;
        sum += lg[s];
    }
    sum = group_sum<TPG>(sum);
    const float inv = 1.f / sum;
    const float rx = __ldg(ref4 + 0), ry = __ldg(ref4 + 1), rw = __ldg(ref4 + 2), rh = __ldg(ref4 + 3);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        int p = s * TPG + lane;
        if (p < g.P) {
            int l = 0, acc = g.pts[0];
            while (p >= acc && l + 1 < g.n_levels) { ++l; acc += g.pts[l]; }
            const float2 o = *reinterpret_cast<const float2*>(off_row + 2 * p);
            const float ps = __ldg(pscale + p);
            // loc = ref_xy + off * (1/P_l) * ref_wh * offset_scale      (dfine_decoder.py:160-163)
            const float sx = ps * rw * offset_scale, sy = ps * rh * offset_scale;
            const float lx = rx + o.x * sx, ly = ry + o.y * sy;
            // grid = 2*loc-1; pixel = ((grid+1)*size-1)/2               (utils.py:226, ATen unnormalize)
            const float gx = 2.f * lx - 1.f, gy = 2.f * ly - 1.f;
            PointSlot ps_out;
            // clamp far-outside / NaN coordinates to "just outside" so the int conversion is defined
            ps_out.x = fminf(fmaxf(((gx + 1.f) * (float)g.W[l] - 1.f) * 0.5f, -2.f), (float)g.W[l] + 1.f);
            ps_out.y = fminf(fmaxf(((gy + 1.f) * (float)g.H[l] - 1.f) * 0.5f, -2.f), (float)g.H[l] + 1.f);
            ps_out.w = lg[s] * inv;
            ps_out.start = g.start[l];
            ps_out.W = g.W[l];
            ps_out.H = g.H[l];
            ps_out.dscale_x = sx * (float)g.W[l];
            ps_out.dscale_y = sy * (float)g.H[l];
            put_slot(slots, p, ps_out);
        }
    }
}

template <int TPG>
__global__ void __launch_bounds__(256) msda_fwd_kernel(
    const float* __restrict__ value, const float* __restrict__ offsets, long off_ld,
    const float* __restrict__ logits, long logit_ld, const float* __restrict__ ref,
    const float* __restrict__ pscale, float* __restrict__ out, MsdaGeom g, int B, int Q, int heads,
    int L, float offset_scale) {
    pdl_entry();
    constexpr int D = TPG * 4;
    constexpr int GROUPS = 256 / TPG;
    __shared__ __align__(16) FwdSlot slots[GROUPS][MAX_POINTS];
    const int grp = threadIdx.x / TPG, lane = threadIdx.x % TPG;
    const long gid = (long)blockIdx.x * GROUPS + grp;  // (b*Q+q)*heads + h
    const long total = (long)B * Q * heads;
    const bool live = gid < total;
    const long bq = live ? gid / heads : 0;
    const int h = live ? (int)(gid % heads) : 0;
    const int b = (int)(bq / Q);
    // dead groups (tail CTA) run phase 1 on row 0 so that the full-mask shuffles stay convergent
    msda_prepare<TPG, FwdSlot>(g, offsets + bq * off_ld + (long)h * g.P * 2, logits + bq * logit_ld + (long)h * g.P,
                      ref + bq * 4, pscale, offset_scale, lane, slots[grp]);
    __syncwarp();
    if (!live) return;
    const float* vbase = value + ((long)b * L * heads + h) * D + lane * 4;
    const long tok_stride = (long)heads * D;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // 2 x LDS.128 + 4 x LDG.128 + 16 FMA per point; two points (8 gathers) in flight per iteration
#pragma unroll 2
    for (int p = 0; p < g.P; ++p) {
        const int4 tk = *reinterpret_cast<const int4*>(slots[grp][p].tok);
        const float4 w = *reinterpret_cast<const float4*>(slots[grp][p].w);
        const float4 v00 = __ldg(reinterpret_cast<const float4*>(vbase + (long)tk.x * tok_stride));
        const float4 v01 = __ldg(reinterpret_cast<const float4*>(vbase + (long)tk.y * tok_stride));
        const float4 v10 = __ldg(reinterpret_cast<const float4*>(vbase + (long)tk.z * tok_stride));
        const float4 v11 = __ldg(reinterpret_cast<const float4*>(vbase + (long)tk.w * tok_stride));
        acc.x += w.x * v00.x + w.y * v01.x + w.z * v10.x + w.w * v11.x;
        acc.y += w.x * v00.y + w.y * v01.y + w.z * v10.y + w.w * v11.y;
        acc.z += w.x * v00.z + w.y * v01.z + w.z * v10.z + w.w * v11.z;
        acc.w += w.x * v00.w + w.y * v01.w + w.z * v10.w + w.w * v11.w;
    }
    *reinterpret_cast<float4*>(out + gid * D + lane * 4) = acc;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}

template <int TPG>
__global__ void __launch_bounds__(256) msda_bwd_kernel(
    const float* __restrict__ value, const float* __restrict__ offsets, long off_ld,
    const float* __restrict__ logits, long logit_ld, const float* __restrict__ ref,
    const float* __restrict__ pscale, const float* __restrict__ gout, float* __restrict__ gvalue,
    float* __restrict__ goff, long goff_ld, float* __restrict__ glogit, long glogit_ld, MsdaGeom g,
    int B, int Q, int heads, int L, float offset_scale) {
    pdl_entry();
    constexpr int D = TPG * 4;
    constexpr int GROUPS = 256 / TPG;
    __shared__ PointSlot slots[GROUPS][MAX_POINTS];
    __shared__ float gw_s[GROUPS][MAX_POINTS], gx_s[GROUPS][MAX_POINTS], gy_s[GROUPS][MAX_POINTS];
    const int grp = threadIdx.x / TPG, lane = threadIdx.x % TPG;
    const long gid = (long)blockIdx.x * GROUPS + grp;
    const long total = (long)B * Q * heads;
    const bool live = gid < total;
    const long bq = live ? gid / heads : 0;
    const int h = live ? (int)(gid % heads) : 0;
    const int b = (int)(bq / Q);
    // dead groups (tail CTA) run phase 1 on row 0 so that the full-mask shuffles stay convergent
    msda_prepare<TPG, PointSlot>(g, offsets + bq * off_ld + (long)h * g.P * 2, logits + bq * logit_ld + (long)h * g.P,
                      ref + bq * 4, pscale, offset_scale, lane, slots[grp]);
    __syncwarp();
    // NB: no early return before the shuffles below — dead groups run with zero gradients.
    const long voff = ((long)b * L * heads + h) * D + lane * 4;
    const long tok_stride = (long)heads * D;
    const float4 go = live ? __ldg(reinterpret_cast<const float4*>(gout + gid * D + lane * 4))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
    const int P = g.P;
    for (int p = 0; p < P; ++p) {
        float tw = 0.f, tx = 0.f, ty = 0.f;
        if (live) {
            const PointSlot s = slots[grp][p];
            const float xf = floorf(s.x), yf = floorf(s.y);
            const int x0 = (int)xf, y0 = (int)yf;
            const float fx = s.x - xf, fy = s.y - yf;
            const bool vx0 = x0 >= 0 && x0 < s.W, vx1 = x0 + 1 >= 0 && x0 + 1 < s.W;
            const bool vy0 = y0 >= 0 && y0 < s.H, vy1 = y0 + 1 >= 0 && y0 + 1 < s.H;
            const long o00 = voff + (long)(s.start + y0 * s.W + x0) * tok_stride;
            const long o01 = o00 + tok_stride, o10 = o00 + (long)s.W * tok_stride, o11 = o10 + tok_stride;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 v00 = (vx0 && vy0) ? __ldg(reinterpret_cast<const float4*>(value + o00)) : z;
            const float4 v01 = (vx1 && vy0) ? __ldg(reinterpret_cast<const float4*>(value + o01)) : z;
            const float4 v10 = (vx0 && vy1) ? __ldg(reinterpret_cast<const float4*>(value + o10)) : z;
            const float4 v11 = (vx1 && vy1) ? __ldg(reinterpret_cast<const float4*>(value + o11)) : z;
            const float b00 = (1.f - fx) * (1.f - fy), b01 = fx * (1.f - fy), b10 = (1.f - fx) * fy, b11 = fx * fy;
            if (vx0 && vy0) { float c = s.w * b00; red_add_v4(gvalue + o00, c * go.x, c * go.y, c * go.z, c * go.w); }
            if (vx1 && vy0) { float c = s.w * b01; red_add_v4(gvalue + o01, c * go.x, c * go.y, c * go.z, c * go.w); }
            if (vx0 && vy1) { float c = s.w * b10; red_add_v4(gvalue + o10, c * go.x, c * go.y, c * go.z, c * go.w); }
            if (vx1 && vy1) { float c = s.w * b11; red_add_v4(gvalue + o11, c * go.x, c * go.y, c * go.z, c * go.w); }
            const float d00 = dot4(v00, go), d01 = dot4(v01, go), d10 = dot4(v10, go), d11 = dot4(v11, go);
            tw = b00 * d00 + b01 * d01 + b10 * d10 + b11 * d11;
            tx = (1.f - fy) * (d01 - d00) + fy * (d11 - d10);
            ty = (1.f - fx) * (d10 - d00) + fx * (d11 - d01);
        }
        tw = group_sum<TPG>(tw);
        tx = group_sum<TPG>(tx);
        ty = group_sum<TPG>(ty);
        if (lane == 0) { gw_s[grp][p] = tw; gx_s[grp][p] = tx; gy_s[grp][p] = ty; }
    }
    __syncwarp();
    // Phase 3: softmax backward over the P points + location chain rule.
    constexpr int NS = MAX_POINTS / TPG;
    float part = 0.f;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        int p = s * TPG + lane;
        if (live && p < P) part += slots[grp][p].w * gw_s[grp][p];
    }
    const float wsum = group_sum<TPG>(part);
    if (!live) return;
    float* go_row = goff + bq * goff_ld + (long)h * P * 2;
    float* gl_row = glogit + bq * glogit_ld + (long)h * P;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        int p = s * TPG + lane;
        if (p < P) {
            const PointSlot sl = slots[grp][p];
            gl_row[p] = sl.w * (gw_s[grp][p] - wsum);
            float2 r;
            r.x = sl.w * gx_s[grp][p] * sl.dscale_x;
            r.y = sl.w * gy_s[grp][p] * sl.dscale_y;
            *reinterpret_cast<float2*>(go_row + 2 * p) = r;
        }
    }
}

int fill_geom(MsdaGeom& g, int n_levels, const int* hw, const int* pts, int L) {
    if (n_levels < 1 || n_levels > MAX_LEVELS) return -1;
    g.n_levels = n_levels;
    int start = 0, P = 0;
    for (int l = 0; l < MAX_LEVELS; ++l) {
        if (l < n_levels) {
            g.H[l] = hw[2 * l];
            g.W[l] = hw[2 * l + 1];
            g.start[l] = start;
            g.pts[l] = pts[l];
            start += g.H[l] * g.W[l];
            P += pts[l];
        } else {
            g.H[l] = g.W[l] = g.start[l] = g.pts[l] = 0;
        }
    }
    g.P = P;
    if (P > MAX_POINTS || start != L) return -1;
    return 0;
}

}  // namespace

// value [B,L,heads,head_dim]; offsets rows of heads*P*2 floats with row stride off_ld (elements);
// logits rows of heads*P floats with stride logit_ld; ref [B,Q,4]; pscale [P]; out [B,Q,heads*head_dim].
DFINE_API int dfine_msda_fwd(const float* value, const float* offsets, long off_ld, const float* logits,
                             long logit_ld, const float* ref, const float* pscale, float* out, int B, int Q,
                             int L, int heads, int head_dim, int n_levels, const int* level_hw,
                             const int* level_points, float offset_scale, void* stream) {
    MsdaGeom g;
    DFINE_REQUIRE(fill_geom(g, n_levels, level_hw, level_points, L) == 0, "msda_fwd: bad level geometry");
    DFINE_REQUIRE(head_dim == 32 || head_dim == 16, "msda_fwd: head_dim %d unsupported (16|32)", head_dim);
    DFINE_REQUIRE(off_ld % 2 == 0 && ((uintptr_t)offsets % 8) == 0 && ((uintptr_t)value % 16) == 0 &&
                      ((uintptr_t)out % 16) == 0,
                  "msda_fwd: alignment");
    const long groups = (long)B * Q * heads;
    if (groups == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (head_dim == 32) {
        launch_k(msda_fwd_kernel<8>, ceil_div(groups, 32), 256, 0, st, value, offsets, off_ld, logits, logit_ld, ref,
                                                                pscale, out, g, B, Q, heads, L, offset_scale);
    } else {
        launch_k(msda_fwd_kernel<4>, ceil_div(groups, 64), 256, 0, st, value, offsets, off_ld, logits, logit_ld, ref,
                                                                pscale, out, g, B, Q, heads, L, offset_scale);
    }
    DFINE_LAUNCH_CHECK("msda_fwd");
    return 0;
}

// gvalue [B,L,heads,head_dim] must be zero-initialised by the caller (accumulated with red.add);
// goff / glogit rows use strides goff_ld / glogit_ld and are fully overwritten.
DFINE_API int dfine_msda_bwd(const float* value, const float* offsets, long off_ld, const float* logits,
                             long logit_ld, const float* ref, const float* pscale, const float* gout,
                             float* gvalue, float* goff, long goff_ld, float* glogit, long glogit_ld, int B,
                             int Q, int L, int heads, int head_dim, int n_levels, const int* level_hw,
                             const int* level_points, float offset_scale, void* stream) {
    MsdaGeom g;
    DFINE_REQUIRE(fill_geom(g, n_levels, level_hw, level_points, L) == 0, "msda_bwd: bad level geometry");
    DFINE_REQUIRE(head_dim == 32 || head_dim == 16, "msda_bwd: head_dim %d unsupported (16|32)", head_dim);
    DFINE_REQUIRE(off_ld % 2 == 0 && goff_ld % 2 == 0 && ((uintptr_t)offsets % 8) == 0 &&
                      ((uintptr_t)goff % 8) == 0 && ((uintptr_t)value % 16) == 0 &&
                      ((uintptr_t)gvalue % 16) == 0 && ((uintptr_t)gout % 16) == 0,
                  "msda_bwd: alignment");
    const long groups = (long)B * Q * heads;
    if (groups == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (head_dim == 32) {
        launch_k(msda_bwd_kernel<8>, ceil_div(groups, 32), 256, 0, st, value, offsets, off_ld, logits, logit_ld, ref,
                                                                pscale, gout, gvalue, goff, goff_ld, glogit,
                                                                glogit_ld, g, B, Q, heads, L, offset_scale);
    } else {
        launch_k(msda_bwd_kernel<4>, ceil_div(groups, 64), 256, 0, st, value, offsets, off_ld, logits, logit_ld, ref,
                                                                pscale, gout, gvalue, goff, goff_ld, glogit,
                                                                glogit_ld, g, B, Q, heads, L, offset_scale);
    }
    DFINE_LAUNCH_CHECK("msda_bwd");
    return 0;
}
