// Input side and detection post-processing on the device — the callers either side of the hot path.
//
//   dfine_preprocess_u8      uint8 HWC images -> float32 NHWC in [0,1] with optional BGR->RGB swap and bilinear resize in
//                            ONE pass: Torch_model._prepare_inputs' device half (infer/torch_model.py:283-292: uint8 H2D
//                            then .float().div_(255)) fused with the resize that cv2 does on the host (245-248) and with the
//                            train loader's multiscale F.interpolate(bilinear, align_corners=False) (dl/dataset.py:675-683).
//                            The result is the NHWC tensor the first conv kernel reads (no NCHW round trip).
//   dfine_resize_bilinear_f32  the same resize for float NHWC tensors (multiscale on an already normalised batch).
//   dfine_postprocess        DFINEPostProcessor.forward (dl/export.py:59-100) / Torch_model._preds_postprocess
//                            (infer/torch_model.py:153-227), focal-loss branch: scores = sigmoid(logits), top-K over the
//                            Q*C scores (dfine_topk_rowmax with C = 1 on the logits — sigmoid is monotonic), labels = idx % C,
//                            query = idx / C, boxes cxcywh (normalised) -> xyxy in input pixels with the reference's
//                            floor / ceil / clamp rounding.
// HBM-bound byte / fp32 work: 1.2 MB of uint8 per 640x640 image in, 4.9 MB of fp32 out.
#include "common.cuh"

namespace {

// PyTorch's area_pixel_compute_source_index for bilinear, align_corners = False: src = (dst + 0.5) * scale - 0.5, clamped at 0
__device__ __forceinline__ void src_index(int d, float scale, int in_size, int* i0, int* i1, float* w1) {
    float s = ((float)d + 0.5f) * scale - 0.5f;
    s = s < 0.f ? 0.f : s;
    const int a = (int)s;
    *i0 = a < in_size - 1 ? a : in_size - 1;
    *i1 = a < in_size - 1 ? a + 1 : in_size - 1;
    *w1 = s - (float)a;
}

template <typename T>
__global__ void resize_kernel(const T* __restrict__ src, float* __restrict__ dst, int B, int Hs, int Ws, int H, int W, int C,
                              float mul, int swap_rb) {
    pdl_entry();
    const long n = (long)B * H * W;
    const float sy = (float)Hs / (float)H, sx = (float)Ws / (float)W;
    const bool same = Hs == H && Ws == W;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H), b = (int)(i / ((long)W * H));
        const T* img = src + (long)b * Hs * Ws * C;
        float* o = dst + i * C;
        if (same) {
            for (int c = 0; c < C; ++c) o[swap_rb && C == 3 ? 2 - c : c] = (float)img[((long)y * Ws + x) * C + c] * mul;
            continue;
        }
        int y0, y1, x0, x1;
        float wy, wx;
        src_index(y, sy, Hs, &y0, &y1, &wy);
        src_index(x, sx, Ws, &x0, &x1, &wx);
        for (int c = 0; c < C; ++c) {
            const float v00 = (float)img[((long)y0 * Ws + x0) * C + c], v01 = (float)img[((long)y0 * Ws + x1) * C + c];
            const float v10 = (float)img[((long)y1 * Ws + x0) * C + c], v11 = (float)img[((long)y1 * Ws + x1) * C + c];
            const float top = v00 + wx * (v01 - v00), bot = v10 + wx * (v11 - v10);
            o[swap_rb && C == 3 ? 2 - c : c] = (top + wy * (bot - top)) * mul;
        }
    }
}

// float NHWC, C % 4 == 0: one thread = one destination pixel x 4 channels (coalesced float4 accesses)
__global__ void __launch_bounds__(256) resize_f32x4_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int Hs,
                                                           int Ws, int H, int W, int C) {
    pdl_entry();
    const int c4 = C / 4;
    const long n = (long)B * H * W * c4;
    const float sy = (float)Hs / (float)H, sx = (float)Ws / (float)W;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int q = (int)(i % c4);
        long r = i / c4;
        const int x = (int)(r % W); r /= W;
        const int y = (int)(r % H);
        const int b = (int)(r / H);
        int y0, y1, x0, x1;
        float wy, wx;
        src_index(y, sy, Hs, &y0, &y1, &wy);
        src_index(x, sx, Ws, &x0, &x1, &wx);
        const float* img = src + (long)b * Hs * Ws * C + 4 * q;
        const float4 v00 = __ldg(reinterpret_cast<const float4*>(img + ((long)y0 * Ws + x0) * C));
        const float4 v01 = __ldg(reinterpret_cast<const float4*>(img + ((long)y0 * Ws + x1) * C));
        const float4 v10 = __ldg(reinterpret_cast<const float4*>(img + ((long)y1 * Ws + x0) * C));
        const float4 v11 = __ldg(reinterpret_cast<const float4*>(img + ((long)y1 * Ws + x1) * C));
        float4 o;
        { const float t = v00.x + wx * (v01.x - v00.x), u = v10.x + wx * (v11.x - v10.x); o.x = t + wy * (u - t); }
        { const float t = v00.y + wx * (v01.y - v00.y), u = v10.y + wx * (v11.y - v10.y); o.y = t + wy * (u - t); }
        { const float t = v00.z + wx * (v01.z - v00.z), u = v10.z + wx * (v11.z - v10.z); o.z = t + wy * (u - t); }
        { const float t = v00.w + wx * (v01.w - v00.w), u = v10.w + wx * (v11.w - v10.w); o.w = t + wy * (u - t); }
        *reinterpret_cast<float4*>(dst + i * 4) = o;
    }
}

__global__ void postprocess_kernel(const float* __restrict__ logits, const float* __restrict__ boxes, const long* __restrict__ idx,
                                   long* __restrict__ labels, float* __restrict__ out_boxes, float* __restrict__ scores,
                                   long* __restrict__ qidx, int B, int Q, int C, int K, float height, float width, int to_round) {
    pdl_entry();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * K) return;
    const int b = i / K;
    const long id = idx[i];
    const int q = (int)(id / C), c = (int)(id % C);
    labels[i] = c;
    if (qidx) qidx[i] = q;
    scores[i] = 1.f / (1.f + expf(-logits[((long)b * Q + q) * C + c]));
    const float* bx = boxes + ((long)b * Q + q) * 4;
    const float xc = bx[0] * width, yc = bx[1] * height, bw = bx[2] * width, bh = bx[3] * height;
    float x0 = xc - bw / 2, y0 = yc - bh / 2, x1 = xc + bw / 2, y1 = yc + bh / 2;
    if (to_round) {
        x0 = fmaxf(floorf(x0), 1.f); y0 = fmaxf(floorf(y0), 1.f);
        x1 = fminf(ceilf(x1), width - 1.f); y1 = fminf(ceilf(y1), height - 1.f);
    } else {
        x0 = fmaxf(x0, 0.f); y0 = fmaxf(y0, 0.f);
        x1 = fminf(x1, width); y1 = fminf(y1, height);
    }
    *reinterpret_cast<float4*>(out_boxes + (long)i * 4) = make_float4(x0, y0, x1, y1);
}

int resize_grid(long n) {
    long g = (n + 255) / 256;
    return (int)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g));
}

}  // namespace

// dst [B,H,W,C] float32 = resize(src [B,Hs,Ws,C] uint8) * mul (1/255 for [0,1] inputs), channels reversed when swap_rb
// (BGR -> RGB, C == 3).  Bilinear with half-pixel centres (align_corners = False, no antialias); Hs == H and Ws == W copies.
DFINE_API int dfine_preprocess_u8(const void* src, float* dst, int B, int Hs, int Ws, int H, int W, int C, float mul,
                                  int swap_rb, void* stream) {
    DFINE_REQUIRE(B >= 0 && Hs > 0 && Ws > 0 && H > 0 && W > 0 && C > 0 && C <= 4, "preprocess_u8: bad dims");
    if (B == 0) return 0;
    launch_k(resize_kernel<unsigned char>, resize_grid((long)B * H * W), 256, 0, (cudaStream_t)stream, 
        (const unsigned char*)src, dst, B, Hs, Ws, H, W, C, mul, swap_rb);
    DFINE_LAUNCH_CHECK("preprocess_u8");
    return 0;
}

DFINE_API int dfine_resize_bilinear_f32(const float* src, float* dst, int B, int Hs, int Ws, int H, int W, int C, void* stream) {
    DFINE_REQUIRE(B >= 0 && Hs > 0 && Ws > 0 && H > 0 && W > 0 && C > 0, "resize_bilinear_f32: bad dims");
    if (B == 0) return 0;
    if (C % 4 == 0 && ((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0)
        launch_k(resize_f32x4_kernel, resize_grid((long)B * H * W * (C / 4)), 256, 0, (cudaStream_t)stream, src, dst, B, Hs, Ws, H, W, C);
    else
        launch_k(resize_kernel<float>, resize_grid((long)B * H * W), 256, 0, (cudaStream_t)stream, src, dst, B, Hs, Ws, H, W, C, 1.f, 0);
    DFINE_LAUNCH_CHECK("resize_bilinear_f32");
    return 0;
}

// idx int64 [B,K]: indices into the flattened [Q*C] scores of every image (dfine_topk_rowmax on logits viewed [B, Q*C, 1]).
// Outputs: labels int64 [B,K], out_boxes [B,K,4] xyxy in input pixels, scores [B,K], qidx int64 [B,K] (optional: the
// query of every detection, for gathering masks).
DFINE_API int dfine_postprocess(const float* logits, const float* boxes, const long* idx, long* labels, float* out_boxes,
                                float* scores, long* qidx, int B, int Q, int C, int K, float height, float width,
                                int to_round, void* stream) {
    DFINE_REQUIRE(B >= 0 && Q > 0 && C > 0 && K > 0 && ((uintptr_t)out_boxes % 16) == 0, "postprocess: bad dims / alignment");
    if (B == 0) return 0;
    launch_k(postprocess_kernel, ceil_div((long)B * K, 128), 128, 0, (cudaStream_t)stream, logits, boxes, idx, labels, out_boxes,
                                                                                     scores, qidx, B, Q, C, K, height, width,
                                                                                     to_round);
    DFINE_LAUNCH_CHECK("postprocess");
    return 0;
}
