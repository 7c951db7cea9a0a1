// tcgen05 / TMEM / TMA implicit-GEMM kernels for sm_100a (kind::tf32, fp32 storage, fp32 accumulate).
//
// They carry the dense convolutions of HGNetv2 / HybridEncoder (1x1 and 3x3 stride-1, reference
// hgnetv2.py:35-80, hybrid_encoder.py:22-45) and every nn.Linear of AIFI / decoder / heads:
//
//   tc_fwd   : Y[pixel, n] = epi( sum_{tap, c} X[pixel + tap, c] * Wr[n, tap, c] )
//              A tile  = TH x TW output-pixel patch (<=128 pixels) x 32 channels, fetched by ONE 4-D TMA
//                        box per (tap, channel block) from the NHWC activation — the box origin is
//                        shifted by the tap and TMA's out-of-bounds zero fill implements the padding,
//                        so there is no im2col buffer and no halo logic;
//              B tile  = BN weight rows x 32 of the re-laid weight [Cout, taps*Cin] (2-D TMA);
//              both K-major with the 128-byte swizzle, consumed by tcgen05.mma.cta_group::1.kind::tf32
//              (UMMA M=128, N=BN, K=8), fp32 accumulator in TMEM (BN columns);
//              epilogue warps: tcgen05.ld -> registers -> shared transpose -> 128-byte row stores with
//              bias / activation, plus optional per-channel sum / sum-of-squares for train-mode
//              BatchNorm (removes the separate statistics pass over the conv output).
//              The data-gradient of these convs is the same kernel run on dY with flipped/transposed
//              weights.
//   tc_wgrad : dWr[co, tap, ci] += sum_{pixel} dY[pixel, co] * X[pixel + tap, ci]
//              both operands MN-major (the reduction runs over pixels, channels are contiguous):
//              32-pixel x 32-channel TMA boxes (SWIZZLE_128B_ATOM_32B), UMMA descriptors with a_major =
//              b_major = MN and the 128B_BASE32B layout (the only one defined for MN-major tf32);
//              split over pixel ranges, accumulated into dWr with red.global.add.f32.
//
// Warp roles: warp 0 = TMA producer, warp 1 = barrier init + TMEM alloc + MMA issuer, warps 2..5 = epilogue
// (TMEM lane quarter = warp % 4), warps 6..9 (3xTF32 mode only) = tf32 hi/lo converters.  Multi-stage smem ring;
// where two CTAs fit per SM one CTA's epilogue overlaps the other's main loop.  Every mbarrier wait is bounded and traps instead
// of hanging the device.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <mutex>
#include <cstdlib>

namespace {

// ------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 22)) __trap();  // a lost arrival must abort the launch, not hang the GPU
    }
}
// One lane of a converged warp.  The single-thread roles (TMA producer, MMA issuer) run their loops on the WHOLE warp and
// only the issuing instructions sit under this predicate: inside an `if (lane == 0)` region the compiler must treat every
// value as divergent, keeps descriptors / coordinates in vector registers and moves each one to the uniform datapath
// through an ELECT + R2UR.BROADCAST waterfall loop — ncu's source view showed the producer thread spending 80 % of its
// samples in that integer code (~1000 clk per k-block, the mainloop's actual bound), not waiting for a free stage.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, 128-byte swizzle (layout type 2), descriptor version 1 (sm_100).
// K-major : rows of 128 B (32 fp32 along K), 8-row atoms; SBO = 1024 B between 8-row groups.
// MN-major: rows of 128 B (32 fp32 along M/N), 8 k-rows per atom; SBO = 1024 B between k groups,
//           LBO = byte distance between consecutive 32-element M/N chunks.
// layout_type: 2 = SWIZZLE_128B (K-major operands), 1 = SWIZZLE_128B_BASE32B — the only layout the
// hardware accepts for MN-major 32-bit (tf32) operands: 128-byte rows, 32-byte swizzle granularity,
// 4 k-rows per atom (TMA side: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}
// instruction descriptor: D=f32, A=B=tf32, M=128, N=n; major bits 15/16 (0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t make_idesc(int n, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// the same for kind::f16 with bf16 operands (a_format = b_format = 1), K-major, D = f32
__host__ __device__ constexpr uint32_t make_idesc_bf16(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// kind::f16 with fp16 operands (a_format = b_format = 0), K-major, D = f32
__host__ __device__ constexpr uint32_t make_idesc_f16(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// {hi, lo} bf16 pair of an fp32 value: hi = RN bf16(x), lo = RN bf16(x - hi); x = hi + lo up to 2^-17 relative
__device__ __forceinline__ void bf16_split(float x, uint32_t& hi, uint32_t& lo) {
    const uint32_t h = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x));
    const float r = x - __uint_as_float(h << 16);
    hi = h;
    lo = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(r));
}

// two fp32 -> one packed bf16x2 word (element a in the low half: lower address)
__device__ __forceinline__ uint32_t bf16x2_rn(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

// two fp32 -> one packed f16x2 word (element a in the low half), and the two halves back as fp32
__device__ __forceinline__ uint32_t f16x2_rn(float a, float b) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ float f16_lo(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v & 0xffffu))); }
__device__ __forceinline__ float f16_hi(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v >> 16))); }

constexpr int BM = 128, BK = 32;
constexpr int EPI_LD = 33;
constexpr int MAX_TAPS = 9;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// Implicit-GEMM geometry of one launch.  The output "tile domain" is OH x OW pixels per image; output pixel
// (oh, ow) gathers, for tap t, the input pixel (oh*in_stride + dh[t], ow*in_stride + dw[t]) (zero outside the
// input: TMA out-of-bounds fill) against weight columns [wk[t], wk[t] + Cin), and is stored at pixel
// (oh*osy + ooy, ow*osx + oox) of a YH x YW output tensor.  Forward convs use os = 1; the data gradient of a
// stride-2 conv is four launches (one per output parity) with os = 2.
struct FwdParams {
    int n_taps;
    int dh[MAX_TAPS], dw[MAX_TAPS], wk[MAX_TAPS];
    int in_stride;
    int Cin;                  // channels per tap
    int TW, TH;               // output patch per CTA (TW*TH <= 128)
    int tiles_w, tiles_h;     // patches per image
    int OH, OW, N;            // tile domain, N = Cout
    int YH, YW, osy, osx, ooy, oox;
    long ldy;                 // output pixel stride (elements)
    int act;
    int dbg;                  // bring-up switch for the 3xTF32 converter (0 = normal)
    const float* x;           // input tensor geometry again, for the L2 prefetch warp
    int B, H, W;
    long ldx;
    int prefetch;             // k-blocks the prefetch warp may run ahead of the TMA producer (0 = off)
    int w_planes;             // 3xTF32: map_w is a 3-D {K, Cout, 2} map over adjacent hi / lo weight planes
    int wk2[MAX_TAPS];        // hybrid mode: first weight column of tap t in the tap-padded bf16 planes
    int half16;               // 16-bit plane modes (X3 = 2): 0 = bf16 planes (3xBF16), 1 = fp16 planes (3xFP16)
    float out_scale;          // accumulator scale applied first in the epilogue (3xFP16 weights are stored x 2^8)
    int w_row_off, w_k_off;   // per-image weights ("grouped" launches: the mask product): image i reads weight rows
                              // [n0 + i*w_row_off, ...) and weight columns [k + i*w_k_off, ...) of one stacked matrix
    int lab;                  // deploy-mode epilogue: y = lab_s * act(acc + bias) + lab_b (LearnableAffineBlock, hgnetv2.py:25-32)
    float lab_s, lab_b;
    const float* res;         // optional tensor added to the output in the epilogue (same pixel geometry as y,
    long ldres;               // pixel stride ldres): fuses the gradient-accumulation add of a multi-consumer tensor
    const float* ch_scale;    // optional per-channel scale applied before the bias (tc_fwd_ts only): a frozen / eval-mode
                              // BatchNorm folded into the epilogue, y = act(acc * scale[c] + bias[c])
    unsigned long long* trace;   // bring-up: per-CTA phase timestamps (globaltimer ns), 16 slots per CTA (dfine_tc_trace)
    // Train-mode BatchNorm finalize in the kernel's tail (tc_fwd_ts only; fin_counter != null): the CTA that retires last
    // (ticket on fin_counter, zeroed by the caller with `stats`) turns the complete per-channel sums into mean / invstd /
    // scale / shift and updates the running statistics — bn_finalize_kernel's arithmetic without its launch.
    unsigned int* fin_counter;
    const float* fin_w;       // BatchNorm weight / bias (null: 1 / 0)
    const float* fin_b;
    float* fin_rmean;         // running statistics (null: not tracked)
    float* fin_rvar;
    float* fin_mean;          // outputs, [N] each
    float* fin_invstd;
    float* fin_scale;
    float* fin_shift;
    long fin_M;               // pixels the statistics were taken over
    float fin_momentum, fin_eps;
    // ... and the BatchNorm apply pass (fin_y != null): every CTA waits for the finalize (a release / acquire flag next to
    // the ticket; the grid is at most one CTA per SM, so all CTAs are resident), then normalises the tiles IT wrote,
    // y = lab_s * act(conv * scale[c] + shift[c]) + lab_b (+ post_add), re-reading its own accumulators from L2:
    // bn_apply_kernel's arithmetic without its launch.  conv output `y` stays the saved pre-normalisation tensor.
    float* fin_y;             // normalised output, pixel stride fin_ldy
    long fin_ldy;
    const float* fin_post;    // optional residual added last (pixel stride fin_ldpost)
    long fin_ldpost;
    const float* fin_lab_s;   // optional LAB scalars (device)
    const float* fin_lab_b;
    int fin_act;
};

// X3 = 0: one kind::tf32 MMA per k-step (operands truncated to tf32 by the tensor core).
// X3 = 1: error-compensated "3xTF32": a = a_hi + a_lo, w = w_hi + w_lo with hi = round-to-nearest tf32;
//         D += a_hi*w_hi + a_lo*w_hi + a_hi*w_lo.  a_hi / a_lo are produced in shared memory by four converter
//         warps between the TMA landing and the MMA issue; w_hi / w_lo arrive pre-split from global memory.
//         Restores fp32-class accuracy (the dropped a_lo*w_lo term is 2^-22 relative) at 3x the MMA work,
//         which the HBM-bound layers of this network hide.
template <int BN, int X3, int STAGES>
struct FwdSmem {
    alignas(1024) float a[STAGES][BM * BK];
    alignas(1024) float b[STAGES][BN * BK];
    alignas(1024) float alo[X3 ? STAGES : 1][X3 ? BM * BK : 32];
    alignas(1024) float blo[X3 ? STAGES : 1][X3 ? BN * BK : 32];
    uint64_t full[STAGES], empty[STAGES], conv[STAGES], acc_full;   // (the epilogue's transpose buffers alias a[0])
    uint32_t tmem_base;
};

// warps: 0 TMA, 1 MMA, 2..5 epilogue, then the converter warps (4 for 3xTF32, 8 for the bf16-plane modes whose
// converter would otherwise bound the k-block time), then one L2-prefetch warp (persistent kernel)
constexpr int conv_warps(int x3) { return x3 == 0 ? 0 : (x3 == 1 ? 4 : 8); }
constexpr int fwd_threads(int x3) { return 32 * (6 + conv_warps(x3) + 1); }

template <int BN, int X3, int STAGES>
__global__ void __launch_bounds__(fwd_threads(X3)) tc_fwd_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                 const __grid_constant__ CUtensorMap map_w,
                                                                 const __grid_constant__ CUtensorMap map_wlo,
                                                                 float* __restrict__ y, const float* __restrict__ bias,
                                                                 double* __restrict__ stats, FwdParams p) {
    extern __shared__ uint8_t raw[];
    using Smem = FwdSmem<BN, X3, STAGES>;
    Smem& sm = *reinterpret_cast<Smem*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    // tile coordinates
    const int tiles_per_img = p.tiles_w * p.tiles_h;
    const int img = blockIdx.x / tiles_per_img, t = blockIdx.x % tiles_per_img;
    const int h0 = (t / p.tiles_w) * p.TH, w0 = (t % p.tiles_w) * p.TW;
    const int n0 = blockIdx.y * BN;
    const int cblocks = (p.Cin + BK - 1) / BK;
    const int num_k = p.n_taps * cblocks;

    if (warp == 0 && lane == 0) { prefetch_tmap(&map_x); prefetch_tmap(&map_w); if (X3) prefetch_tmap(&map_wlo); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); mbar_init(&sm.conv[s], 128); }
            mbar_init(&sm.acc_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(&sm.tmem_base, BN);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    pdl_entry();      // barriers / tensor memory / descriptor prefetch above overlap the previous kernel's tail

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(&sm.empty[s], ph ^ 1);
                const int tap = kb / cblocks, c0 = (kb % cblocks) * BK;
                // the x box is TW*TH (<=128) rows of 128 B; smem rows beyond that keep stale data that only
                // feeds accumulator rows the epilogue never stores.  expect_tx counts the box bytes:
                mbar_expect_tx(&sm.full[s], (uint32_t)((p.TW * p.TH + (X3 ? 2 : 1) * BN) * BK * sizeof(float)));
                tma_load_4d(sm.a[s], &map_x, &sm.full[s], c0, w0 * p.in_stride + p.dw[tap], h0 * p.in_stride + p.dh[tap],
                            img);
                tma_load_2d(sm.b[s], &map_w, &sm.full[s], p.wk[tap] + c0, n0);
                if (X3) tma_load_2d(sm.blo[s], &map_wlo, &sm.full[s], p.wk[tap] + c0, n0);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN, 0, 0);
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(X3 ? &sm.conv[s] : &sm.full[s], ph);
                tc_fence_after();
                const uint64_t da = make_desc(smem_u32(sm.a[s]), 16, 1024);
                const uint64_t db = make_desc(smem_u32(sm.b[s]), 16, 1024);
                if (X3) {
                    const uint64_t dal = make_desc(smem_u32(sm.alo[s]), 16, 1024);
                    const uint64_t dbl = make_desc(smem_u32(sm.blo[s]), 16, 1024);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        umma_tf32(tmem, dal + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        umma_tf32(tmem, da + 2 * k, dbl + 2 * k, idesc, 1);
                        umma_tf32(tmem, da + 2 * k, db + 2 * k, idesc, 1);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k)  // UMMA_K = 8 tf32 = 32 bytes -> +2 in 16-byte units
                        umma_tf32(tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                }
                umma_commit(&sm.empty[s]);
            }
            umma_commit(&sm.acc_full);
        }
        __syncwarp();
    } else if (warp < 6) {
        const int q = warp % 4;  // TMEM lane quarter
        // all MMAs (hence all TMA loads) have completed once acc_full fires: stage 0 is free to reuse
        float* buf = sm.a[0] + q * 32 * EPI_LD;
        // rows handled by this lane in the transposed store phase: r = 4*i + lane/8, i = 0..7
        long row_off[8];
        bool row_ok[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = 32 * q + 4 * i + lane / 8;
            const int th = r / p.TW, tw = r % p.TW;
            const int oh = h0 + th, ow = w0 + tw;
            row_ok[i] = th < p.TH && oh < p.OH && ow < p.OW;
            row_off[i] = (((long)img * p.YH + (long)oh * p.osy + p.ooy) * p.YW + (long)ow * p.osx + p.oox) * p.ldy;
        }
        mbar_wait(&sm.acc_full, 0);
        tc_fence_after();
        for (int c0 = 0; c0 < BN; c0 += 32) {
            if (n0 + c0 >= p.N) break;
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int c = 0; c < 32; ++c) buf[lane * EPI_LD + c] = __uint_as_float(v[c]);
            __syncwarp();
            const int col = n0 + c0 + 4 * (lane % 8);
            const bool col_ok = col < p.N;  // N % 4 == 0 is required by the host wrapper
            float4 bv = make_float4(0, 0, 0, 0);
            if (bias && col_ok) bv = __ldg(reinterpret_cast<const float4*>(bias + col));
            float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = 4 * i + lane / 8;
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = buf[r * EPI_LD + 4 * (lane % 8) + j];
                if (row_ok[i] && col_ok) {
                    if (stats) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) { s1[j] += o[j]; s2[j] += o[j] * o[j]; }
                    }
                    o[0] = act_fwd(o[0] + bv.x, p.act); o[1] = act_fwd(o[1] + bv.y, p.act);
                    o[2] = act_fwd(o[2] + bv.z, p.act); o[3] = act_fwd(o[3] + bv.w, p.act);
                    *reinterpret_cast<float4*>(y + row_off[i] + col) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
            if (stats) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 8); s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16);
                    s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 8); s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
                }
                if (lane < 8 && col_ok) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        atomicAdd(stats + col + j, (double)s1[j]);
                        atomicAdd(stats + p.N + col + j, (double)s2[j]);
                    }
                }
            }
            __syncwarp();
        }
    } else if (X3 && warp < 10) {
        // converter warps 6..9: split the landed activation tile into tf32 hi (in place) and lo parts.
        // Explicit ld.shared / st.shared (not generic accesses through the shared window) so that
        // fence.proxy.async.shared::cta orders exactly these writes before the tensor core's async-proxy reads;
        // every converter thread fences its own writes and arrives itself (barrier count 128).
        const int ct = threadIdx.x - 192;  // 0..127
        for (int kb = 0; kb < num_k; ++kb) {
            const int s = kb % STAGES, ph = (kb / STAGES) & 1;
            mbar_wait(&sm.full[s], ph);
            const uint32_t a_base = smem_u32(sm.a[s]), l_base = smem_u32(sm.alo[s]);
#pragma unroll
            for (int i = 0; i < BM * BK / 4 / 128; ++i) {
                const uint32_t off = (uint32_t)(ct + i * 128) * 16u;
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a_base + off) : "memory");
                float4 hi, lo;
                hi.x = tf32_rna(v.x); hi.y = tf32_rna(v.y); hi.z = tf32_rna(v.z); hi.w = tf32_rna(v.w);
                lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
                if (p.dbg == 1) { lo = v; hi = make_float4(0.f, 0.f, 0.f, 0.f); }       // everything through alo
                if (p.dbg == 3) { lo = make_float4(0.f, 0.f, 0.f, 0.f); hi = make_float4(0.f, 0.f, 0.f, 0.f); }
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_base + off), "f"(hi.x), "f"(hi.y),
                             "f"(hi.z), "f"(hi.w) : "memory");
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(l_base + off), "f"(lo.x), "f"(lo.y),
                             "f"(lo.z), "f"(lo.w) : "memory");
            }
            fence_proxy_async();
            mbar_arrive(&sm.conv[s]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, BN);
}

// ------------------------------------------------------------------------------------------- persistent forward
// Same math as tc_fwd_kernel, restructured the Blackwell way: ONE CTA per SM loops over output tiles
// (tile = blockIdx.x + i*gridDim.x, N tiles of the same pixel patch adjacent so the A patch is fetched from
// HBM once and re-served by L2), the smem ring runs continuously across tile boundaries, and the fp32
// accumulator is DOUBLE-BUFFERED in TMEM (2*BN columns): the MMA warp starts tile i+1 while the four
// epilogue warps drain tile i.  Removes the per-tile prologue (barrier init, TMEM alloc) that dominates the
// small-K layers (stem / stage-1 maps with 10^4 tiles) and exposes no epilogue latency.
constexpr int EPL = 36;   // epilogue transpose pitch (floats): 16-byte aligned rows, conflict-free float4 phases
template <int BN> struct StatT { typedef double type; };
template <> struct StatT<256> { typedef float type; };
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// X3 = 2: error-compensated "3xBF16": a = a_hi + a_lo, w = w_hi + w_lo with bf16 parts (16 mantissa bits per
//         operand, products exact in the fp32 accumulator; dropped a_lo*w_lo and representation terms 2^-17 relative).
//         Same three MMAs per k-step as 3xTF32 but kind::f16 runs at twice the tf32 rate and both bf16 planes of
//         an operand take the bytes of ONE fp32 plane: half the tensor time and 2/3 of the L2->SM bytes of X3 = 1.
//         The fp32 activation tile lands by TMA (128-byte swizzle) and four converter warps write its bf16 hi / lo
//         planes (64-byte rows, 64-byte swizzle); weights arrive pre-split as bf16 planes.
template <int BN, int X3, int STAGES>
struct PersistSmem {
    // X3 = 3 ("hybrid"): a_hi*w_hi as kind::tf32 (operands RN tf32), the two cross terms a_lo*w_hi + a_hi*w_lo as
    //         kind::f16 on bf16 copies (a 2^-9 relative error on a 2^-12 relative term): 3xTF32-class accuracy for
    //         8 instead of 12 tf32-MMA times per k-block.  b = [w_hi fp32 | bf16(w) | bf16(w_lo)],
    //         a = tf32 hi in place, a2 = [bf16(a_lo) | bf16(a)].
    static constexpr int B_PLANE = X3 == 2 ? BN * BK * 2 : BN * BK * 4;        // bytes of the first weight plane
    static constexpr int B16_PLANE = BN * BK * 2;
    static constexpr int B_BYTES = X3 == 3 ? B_PLANE + 2 * B16_PLANE : (X3 ? 2 : 1) * B_PLANE;   // [hi | lo]
    static constexpr int A2_PLANE = X3 >= 2 ? BM * BK * 2 : BM * BK * 4;
    static constexpr int A2_BYTES = X3 == 1 ? A2_PLANE : (X3 >= 2 ? 2 * A2_PLANE : 16);   // tf32 lo | two bf16 planes
    alignas(1024) float a[STAGES][BM * BK];                                      // TMA landing buffer (fp32)
    alignas(1024) uint8_t b[STAGES][B_BYTES];
    alignas(X3 ? 1024 : 16) uint8_t a2[X3 ? STAGES : 1][A2_BYTES];
    alignas(16) float epi[4][32 * EPL];
    // warp-private BatchNorm partial sums (sum | sum of squares per channel of the N tile), plain adds, merged
    // and flushed to global memory once per CTA.  (fp32 slots on the 256-wide tile, where smem is exhausted
    // and a CTA sees at most ~6 tiles; double elsewhere.)
    typename StatT<BN>::type statw[4][2 * BN];
    uint64_t full[STAGES], empty[STAGES], conv[STAGES], tfull[2], tempty[2];
    uint32_t tmem_base;
    volatile uint32_t produced;   // k-blocks whose TMA loads have been issued (progress hint for the prefetch warp)
};

struct TileSched {
    int n_tiles, total;   // N tiles per pixel patch, total tiles
};

template <int BN, int X3, int STAGES>
__global__ void __launch_bounds__(fwd_threads(X3), 1) tc_fwd_persist(const __grid_constant__ CUtensorMap map_x,
                                                                    const __grid_constant__ CUtensorMap map_w,
                                                                    const __grid_constant__ CUtensorMap map_wlo,
                                                                    float* __restrict__ y, const float* __restrict__ bias,
                                                                    double* __restrict__ stats, FwdParams p, TileSched ts) {
    extern __shared__ uint8_t raw[];
    using Smem = PersistSmem<BN, X3, STAGES>;
    Smem& sm = *reinterpret_cast<Smem*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int tiles_per_img = p.tiles_w * p.tiles_h;
    const int cblocks = (p.Cin + BK - 1) / BK;
    const int num_k = p.n_taps * cblocks;
    constexpr uint32_t TMEM_COLS = 2 * BN;

    if (stats)
        for (int i = threadIdx.x; i < 4 * 2 * BN; i += blockDim.x) (&sm.statw[0][0])[i] = 0;
    if (warp == 0 && lane == 0) { prefetch_tmap(&map_x); prefetch_tmap(&map_w); if (X3) prefetch_tmap(&map_wlo); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); mbar_init(&sm.conv[s], 32 * conv_warps(X3)); }
            for (int a = 0; a < 2; ++a) { mbar_init(&sm.tfull[a], 1); mbar_init(&sm.tempty[a], 4); }
            sm.produced = 0;
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(&sm.tmem_base, TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    pdl_entry();      // barriers / tensor memory / descriptor prefetch above overlap the previous kernel's tail

    if (warp == 0) {
        // whole warp, warp-uniform values, one elected lane issues (see elect_one)
        uint32_t g = 0, s = 0, ph = 0;
        const uint32_t a_bytes = (uint32_t)(p.TW * p.TH * BK * sizeof(float));
        for (int t = blockIdx.x; t < ts.total; t += gridDim.x) {
            const int mt = t / ts.n_tiles, n0 = (t % ts.n_tiles) * BN;
            const int img = mt / tiles_per_img, tt = mt % tiles_per_img;
            const int h0 = (tt / p.tiles_w) * p.TH * p.in_stride, w0 = (tt % p.tiles_w) * p.TW * p.in_stride;
            const int wn = n0 + img * p.w_row_off, wk_img = img * p.w_k_off;
            for (int tap = 0; tap < p.n_taps; ++tap) {
                const int xw = w0 + p.dw[tap], xh = h0 + p.dh[tap], wk = p.wk[tap] + wk_img, wk2 = p.wk2[tap] + wk_img;
                for (int c0 = 0; c0 < p.Cin; c0 += BK) {
                    mbar_wait(&sm.empty[s], ph ^ 1);
                    if (elect_one()) {
                        if (p.dbg == 11) {            // bring-up: no loads (the MMAs run on whatever smem holds)
                            mbar_arrive(&sm.full[s]);
                        } else if (p.dbg == 12) {     // bring-up: weights only
                            mbar_expect_tx(&sm.full[s], (uint32_t)(BN * BK * sizeof(float)));
                            tma_load_2d(sm.b[s], &map_w, &sm.full[s], wk + c0, n0);
                        } else if (p.dbg == 13) {     // bring-up: activations only
                            mbar_expect_tx(&sm.full[s], a_bytes);
                            tma_load_4d(sm.a[s], &map_x, &sm.full[s], c0, xw, xh, img);
                        } else {
                            mbar_expect_tx(&sm.full[s], a_bytes + (uint32_t)Smem::B_BYTES);
                            tma_load_4d(sm.a[s], &map_x, &sm.full[s], c0, xw, xh, img);
                            if (X3 == 3) {       // tf32 hi plane (fp32 map) + the two bf16 planes (one 3-D box, tap-padded K)
                                tma_load_2d(sm.b[s], &map_w, &sm.full[s], wk + c0, wn);
                                tma_load_3d(sm.b[s] + Smem::B_PLANE, &map_wlo, &sm.full[s], wk2 + c0, wn, 0);
                            } else if (X3 == 2 || (X3 && p.w_planes)) {   // hi and lo weight planes are adjacent in memory: one 3-D box
                                tma_load_3d(sm.b[s], &map_w, &sm.full[s], wk + c0, wn, 0);
                            } else {
                                tma_load_2d(sm.b[s], &map_w, &sm.full[s], wk + c0, wn);
                                if (X3) tma_load_2d(sm.b[s] + Smem::B_PLANE, &map_wlo, &sm.full[s], wk + c0, wn);
                            }
                        }
                        sm.produced = ++g;
                    } else {
                        ++g;
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
        if (lane == 0) sm.produced = 0x7fffffffu;
        __syncwarp();
    } else if (warp == 1) {
        // whole warp, one elected lane issues the MMAs and commits (see elect_one)
        constexpr uint32_t idesc = make_idesc(BN, 0, 0);
        uint32_t s = 0, ph = 0, i = 0;
        for (int t = blockIdx.x; t < ts.total; t += gridDim.x, ++i) {
            const uint32_t acc = i & 1, aph = (i >> 1) & 1;
            mbar_wait(&sm.tempty[acc], aph ^ 1);          // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d = tmem + acc * BN;
            for (int kb = 0; kb < num_k; ++kb) {
                mbar_wait(X3 ? &sm.conv[s] : &sm.full[s], ph);
                tc_fence_after();
                const uint64_t da = make_desc(smem_u32(sm.a[s]), 16, 1024);
                const uint64_t db = make_desc(smem_u32(sm.b[s]), 16, 1024);
                if (elect_one()) {
                    if constexpr (X3 == 2) {
                        // bf16 planes: 64-byte rows (32 channels), 64-byte swizzle (layout 4), 8-row atoms of 512 B;
                        // UMMA_K = 16 bf16 = 32 bytes -> +2 in 16-byte units per k-step
                        const uint32_t idesc16 = p.half16 ? make_idesc_f16(BN) : make_idesc_bf16(BN);
                        const uint64_t ah = make_desc(smem_u32(sm.a2[s]), 16, 512, 4);
                        const uint64_t al = make_desc(smem_u32(sm.a2[s] + Smem::A2_PLANE), 16, 512, 4);
                        const uint64_t bh = make_desc(smem_u32(sm.b[s]), 16, 512, 4);
                        const uint64_t bl = make_desc(smem_u32(sm.b[s] + Smem::B_PLANE), 16, 512, 4);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            umma_bf16(d, al + 2 * k, bh + 2 * k, idesc16, (kb | k) != 0);
                            umma_bf16(d, ah + 2 * k, bl + 2 * k, idesc16, 1);
                            umma_bf16(d, ah + 2 * k, bh + 2 * k, idesc16, 1);
                        }
                    } else if constexpr (X3 == 3) {
                        constexpr uint32_t idesc16 = make_idesc_bf16(BN);
                        const uint64_t al = make_desc(smem_u32(sm.a2[s]), 16, 512, 4);                        // bf16(a_lo)
                        const uint64_t ab = make_desc(smem_u32(sm.a2[s] + Smem::A2_PLANE), 16, 512, 4);     // bf16(a)
                        const uint64_t wb = make_desc(smem_u32(sm.b[s] + Smem::B_PLANE), 16, 512, 4);       // bf16(w)
                        const uint64_t wl = make_desc(smem_u32(sm.b[s] + Smem::B_PLANE + Smem::B16_PLANE), 16, 512, 4);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            umma_bf16(d, al + 2 * k, wb + 2 * k, idesc16, (kb | k) != 0);
                            umma_bf16(d, ab + 2 * k, wl + 2 * k, idesc16, 1);
                        }
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) umma_tf32(d, da + 2 * k, db + 2 * k, idesc, 1);
                    } else if constexpr (X3 == 1) {
                        const uint64_t dal = make_desc(smem_u32(sm.a2[s]), 16, 1024);
                        const uint64_t dbl = make_desc(smem_u32(sm.b[s] + Smem::B_PLANE), 16, 1024);
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) {
                            umma_tf32(d, dal + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                            umma_tf32(d, da + 2 * k, dbl + 2 * k, idesc, 1);
                            umma_tf32(d, da + 2 * k, db + 2 * k, idesc, 1);
                        }
                    } else if (p.dbg != 10) {     // dbg 10 = bring-up: loads only, no MMAs
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) umma_tf32(d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&sm.empty[s]);
                }
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            if (elect_one()) umma_commit(&sm.tfull[acc]);
        }
        __syncwarp();
    } else if (warp < 6) {
        // Epilogue: TMEM -> registers (one accumulator row per lane) -> shared transpose (float4, pitch 36) ->
        // 128-byte row stores.  Shared accesses are explicit st/ld.shared.v4 (8 + 8 per 32-column chunk).
        const int q = warp % 4;  // TMEM lane quarter
        const uint32_t wbase = smem_u32(sm.epi[q]);
        const bool plain = (bias == nullptr) && p.act == 0 && !p.lab;
        const bool local_stats = stats != nullptr && ts.n_tiles == 1;
        typename StatT<BN>::type* sw = sm.statw[q];
        uint32_t i = 0;
        for (int t = blockIdx.x; t < ts.total; t += gridDim.x, ++i) {
            const uint32_t acc = i & 1, aph = (i >> 1) & 1;
            const int mt = t / ts.n_tiles, n0 = (t % ts.n_tiles) * BN;
            const int img = mt / tiles_per_img, tt = mt % tiles_per_img;
            const int h0 = (tt / p.tiles_w) * p.TH, w0 = (tt % p.tiles_w) * p.TW;
            long row_off[8];       // output pixel index of the row (multiplied by the pixel stride at the access)
            bool row_ok[8];
#pragma unroll
            for (int r8 = 0; r8 < 8; ++r8) {
                const int r = 32 * q + 4 * r8 + lane / 8;
                const int th = r / p.TW, tw = r % p.TW;
                const int oh = h0 + th, ow = w0 + tw;
                row_ok[r8] = th < p.TH && oh < p.OH && ow < p.OW;
                row_off[r8] = ((long)img * p.YH + (long)oh * p.osy + p.ooy) * p.YW + (long)ow * p.osx + p.oox;
            }
            // Residual tile (data-gradient launches only, X3 = 0): software-pipelined one 32-column chunk ahead, the
            // first chunk requested BEFORE waiting for the accumulator so that its HBM latency hides behind the
            // tile's MMAs (issued inside the chunk loop the dependent loads made the epilogue the bottleneck).
            float4 rcur[8], rnxt[8];
            auto load_res = [&](int c0, float4 (&dst)[8]) {
                const int col = n0 + c0 + 4 * (lane % 8);
#pragma unroll
                for (int r8 = 0; r8 < 8; ++r8)
                    dst[r8] = (row_ok[r8] && col < p.N)
                                  ? __ldg(reinterpret_cast<const float4*>(p.res + row_off[r8] * p.ldres + col))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            const bool has_res = X3 == 0 && p.res != nullptr;
            if (X3 == 0 && has_res) load_res(0, rcur);
            mbar_wait(&sm.tfull[acc], aph);
            tc_fence_after();
            for (int c0 = 0; c0 < BN; c0 += 32) {
                if (n0 + c0 >= p.N) break;
                if (X3 == 0 && has_res && c0 + 32 < BN && n0 + c0 + 32 < p.N) load_res(c0 + 32, rnxt);
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + acc * BN + (uint32_t)c0, v);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4)
                    sts128(wbase + (uint32_t)(lane * EPL + 4 * c4) * 4u, __uint_as_float(v[4 * c4]),
                           __uint_as_float(v[4 * c4 + 1]), __uint_as_float(v[4 * c4 + 2]), __uint_as_float(v[4 * c4 + 3]));
                __syncwarp();
                const int col = n0 + c0 + 4 * (lane % 8);
                const bool col_ok = col < p.N;  // N % 4 == 0 is required by the host wrapper
                float4 bv = make_float4(0, 0, 0, 0);
                if (bias && col_ok) bv = __ldg(reinterpret_cast<const float4*>(bias + col));
                float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
#pragma unroll
                for (int r8 = 0; r8 < 8; ++r8) {
                    const int r = 4 * r8 + lane / 8;
                    float4 o = lds128(wbase + (uint32_t)(r * EPL + 4 * (lane % 8)) * 4u);
                    if (X3 == 2) { o.x *= p.out_scale; o.y *= p.out_scale; o.z *= p.out_scale; o.w *= p.out_scale; }
                    if (row_ok[r8] && col_ok) {
                        if (stats) {
                            s1[0] += o.x; s1[1] += o.y; s1[2] += o.z; s1[3] += o.w;
                            s2[0] += o.x * o.x; s2[1] += o.y * o.y; s2[2] += o.z * o.z; s2[3] += o.w * o.w;
                        }
                        if (!plain) {
                            o.x = act_fwd(o.x + bv.x, p.act); o.y = act_fwd(o.y + bv.y, p.act);
                            o.z = act_fwd(o.z + bv.z, p.act); o.w = act_fwd(o.w + bv.w, p.act);
                            if (p.lab) {
                                o.x = fmaf(o.x, p.lab_s, p.lab_b); o.y = fmaf(o.y, p.lab_s, p.lab_b);
                                o.z = fmaf(o.z, p.lab_s, p.lab_b); o.w = fmaf(o.w, p.lab_s, p.lab_b);
                            }
                        }
                        if (X3 == 0 && has_res) { o.x += rcur[r8].x; o.y += rcur[r8].y; o.z += rcur[r8].z; o.w += rcur[r8].w; }
                        *reinterpret_cast<float4*>(y + row_off[r8] * p.ldy + col) = o;
                    }
                }
                if (X3 == 0 && has_res) {
#pragma unroll
                    for (int r8 = 0; r8 < 8; ++r8) rcur[r8] = rnxt[r8];
                }
                if (stats) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 8); s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16);
                        s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 8); s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
                    }
                    if (lane < 8 && col_ok) {
                        if (local_stats) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {      // warp-private slots: plain read-modify-write
                                sw[c0 + 4 * lane + j] += s1[j];
                                sw[BN + c0 + 4 * lane + j] += s2[j];
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {      // wide-Cout layers: few tiles per channel
                                atomicAdd(stats + col + j, (double)s1[j]);
                                atomicAdd(stats + p.N + col + j, (double)s2[j]);
                            }
                        }
                    }
                }
                __syncwarp();
            }
            // every tcgen05.ld of this warp has completed (tmem_ld32 waits): hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.tempty[acc]);
        }
    } else if (warp == 6 + conv_warps(X3)) {
        // L2 prefetch warp.  TMA keeps only ~32 KB of 128-byte row requests in flight per SM (measured:
        // ~28 GB/s per SM when the activation tile streams from HBM, profiles/README.md), which cannot cover
        // HBM latency.  This warp walks the producer's (tile, k-block) sequence a bounded distance AHEAD of
        // it and issues fire-and-forget prefetch.global.L2 for every 128-byte row of the coming activation
        // boxes, so the TMA loads hit L2.  Purely a hint: no barrier, no effect on results.
        if (p.prefetch > 0) {
            const uint32_t my_tiles = (uint32_t)((ts.total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);
            const uint32_t total_k = my_tiles * (uint32_t)num_k;
            const uint32_t dmin = STAGES, dmax = STAGES + (uint32_t)p.prefetch;
            uint32_t gp = 0;
            while (gp < total_k) {
                const uint32_t done = sm.produced;
                if (done >= total_k) break;
                if (gp < done + dmin) gp = done + dmin;          // already being fetched by TMA: skip ahead
                if (gp >= total_k) break;
                if (gp > done + dmax) { __nanosleep(200); continue; }
                const int t = (int)blockIdx.x + (int)(gp / (uint32_t)num_k) * (int)gridDim.x;
                const int kb = (int)(gp % (uint32_t)num_k);
                const int mt = t / ts.n_tiles;
                const int img = mt / tiles_per_img, tt = mt % tiles_per_img;
                const int h0 = (tt / p.tiles_w) * p.TH, w0 = (tt % p.tiles_w) * p.TW;
                const int tap = kb / cblocks, c0 = (kb % cblocks) * BK;
                // N tiles of one pixel patch re-read the same A boxes: prefetch for the first of them only
                if (t % ts.n_tiles == 0 && img < p.B) {
#pragma unroll
                    for (int rr = 0; rr < BM / 32; ++rr) {
                        const int r = lane + 32 * rr;
                        const int th = r / p.TW, tw = r % p.TW;
                        const int ih = (h0 + th) * p.in_stride + p.dh[tap], iw = (w0 + tw) * p.in_stride + p.dw[tap];
                        if (th < p.TH && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
                            const float* a = p.x + (((long)img * p.H + ih) * p.W + iw) * p.ldx + c0;
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
                            if (((uintptr_t)a & 127) != 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 31));
                        }
                    }
                }
                ++gp;
            }
        }
    } else if (X3 && warp < 6 + conv_warps(X3)) {
        constexpr int CT = 32 * conv_warps(X3);   // converter threads
        const int ct = threadIdx.x - 192;  // 0..CT-1
        uint32_t g = 0;
        for (int t = blockIdx.x; t < ts.total; t += gridDim.x) {
            for (int kb = 0; kb < num_k; ++kb, ++g) {
                const int s = g % STAGES, ph = (g / STAGES) & 1;
                mbar_wait(&sm.full[s], ph);
                const uint32_t a_base = smem_u32(sm.a[s]), l_base = smem_u32(sm.a2[s]);
                if constexpr (X3 >= 2) {
                    // fp32 tile: 128-byte rows, 16-byte chunk j of row r at physical chunk j ^ (r & 7).  bf16 planes:
                    // 64-byte rows, 16-byte chunk c (channels 8c..8c+7) at physical chunk c ^ ((r >> 1) & 3).  The two
                    // fp32 chunks of one bf16 chunk are the physical PAIR k = c ^ ((r & 7) >> 1) — the very index
                    // of the destination chunk — in swapped order on odd rows.  So item idx = 4 r + k reads 32
                    // contiguous bytes at idx * 32 and writes 16 contiguous bytes at idx * 16 of each plane.
                    constexpr int ITEMS = BM * 4 / CT;
                    float4 v0[ITEMS], v1[ITEMS];
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) {      // all loads first: one shared-memory round trip per k-block
                        const uint32_t idx = (uint32_t)(ct + i * CT);
                        v0[i] = lds128(a_base + idx * 32u);
                        v1[i] = lds128(a_base + idx * 32u + 16u);
                    }
#pragma unroll
                    for (int i = 0; i < ITEMS; ++i) {
                        const uint32_t idx = (uint32_t)(ct + i * CT);
                        float4 x0 = v0[i], x1 = v1[i], r0, r1;       // plane 1 source (a) and plane 0 source (residual)
                        if constexpr (X3 == 3) {
                            float4 h0, h1;                           // tf32 hi, written back in place (same order)
                            h0.x = tf32_rna(x0.x); h0.y = tf32_rna(x0.y); h0.z = tf32_rna(x0.z); h0.w = tf32_rna(x0.w);
                            h1.x = tf32_rna(x1.x); h1.y = tf32_rna(x1.y); h1.z = tf32_rna(x1.z); h1.w = tf32_rna(x1.w);
                            sts128(a_base + idx * 32u, h0.x, h0.y, h0.z, h0.w);
                            sts128(a_base + idx * 32u + 16u, h1.x, h1.y, h1.z, h1.w);
                            r0 = make_float4(x0.x - h0.x, x0.y - h0.y, x0.z - h0.z, x0.w - h0.w);
                            r1 = make_float4(x1.x - h1.x, x1.y - h1.y, x1.z - h1.z, x1.w - h1.w);
                        }
                        if ((idx >> 2) & 1) {                        // odd row: the pair is stored high chunk first
                            float4 tmp = x0; x0 = x1; x1 = tmp;
                            if constexpr (X3 == 3) { tmp = r0; r0 = r1; r1 = tmp; }
                        }
                        uint32_t b0, b1, b2, b3, l0, l1, l2, l3;
                        if (X3 == 2 && p.half16) {                   // fp16 planes: hi = RN f16(x), lo = RN f16(x - hi)
                            b0 = f16x2_rn(x0.x, x0.y); b1 = f16x2_rn(x0.z, x0.w);
                            b2 = f16x2_rn(x1.x, x1.y); b3 = f16x2_rn(x1.z, x1.w);
                            l0 = f16x2_rn(x0.x - f16_lo(b0), x0.y - f16_hi(b0));
                            l1 = f16x2_rn(x0.z - f16_lo(b1), x0.w - f16_hi(b1));
                            l2 = f16x2_rn(x1.x - f16_lo(b2), x1.y - f16_hi(b2));
                            l3 = f16x2_rn(x1.z - f16_lo(b3), x1.w - f16_hi(b3));
                        } else {
                            b0 = bf16x2_rn(x0.x, x0.y); b1 = bf16x2_rn(x0.z, x0.w);
                            b2 = bf16x2_rn(x1.x, x1.y); b3 = bf16x2_rn(x1.z, x1.w);
                        }
                        if (X3 == 2 && p.half16) {
                        } else if constexpr (X3 == 2) {              // residual against the bf16 value itself
                            l0 = bf16x2_rn(x0.x - __uint_as_float(b0 << 16), x0.y - __uint_as_float(b0 & 0xffff0000u));
                            l1 = bf16x2_rn(x0.z - __uint_as_float(b1 << 16), x0.w - __uint_as_float(b1 & 0xffff0000u));
                            l2 = bf16x2_rn(x1.x - __uint_as_float(b2 << 16), x1.y - __uint_as_float(b2 & 0xffff0000u));
                            l3 = bf16x2_rn(x1.z - __uint_as_float(b3 << 16), x1.w - __uint_as_float(b3 & 0xffff0000u));
                        } else {                                     // residual against the tf32 hi part
                            l0 = bf16x2_rn(r0.x, r0.y); l1 = bf16x2_rn(r0.z, r0.w);
                            l2 = bf16x2_rn(r1.x, r1.y); l3 = bf16x2_rn(r1.z, r1.w);
                        }
                        // plane order: X3 = 2 -> [hi | lo]; X3 = 3 -> [bf16(a_lo) | bf16(a)]
                        const uint32_t pa = l_base + idx * 16u, pb = l_base + (uint32_t)Smem::A2_PLANE + idx * 16u;
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(X3 == 2 ? pa : pb), "r"(b0), "r"(b1),
                                     "r"(b2), "r"(b3) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(X3 == 2 ? pb : pa), "r"(l0), "r"(l1),
                                     "r"(l2), "r"(l3) : "memory");
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < BM * BK / 4 / 128; ++i) {
                        const uint32_t off = (uint32_t)(ct + i * 128) * 16u;
                        float4 v;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a_base + off) : "memory");
                        float4 hi, lo;
                        hi.x = tf32_rna(v.x); hi.y = tf32_rna(v.y); hi.z = tf32_rna(v.z); hi.w = tf32_rna(v.w);
                        lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
                        if (p.dbg == 1) { lo = v; hi = make_float4(0.f, 0.f, 0.f, 0.f); }       // everything through alo
                        if (p.dbg == 3) { lo = make_float4(0.f, 0.f, 0.f, 0.f); hi = make_float4(0.f, 0.f, 0.f, 0.f); }
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_base + off), "f"(hi.x), "f"(hi.y),
                                     "f"(hi.z), "f"(hi.w) : "memory");
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(l_base + off), "f"(lo.x), "f"(lo.y),
                                     "f"(lo.z), "f"(lo.w) : "memory");
                    }
                }
                fence_proxy_async();
                mbar_arrive(&sm.conv[s]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
    if (stats && ts.n_tiles == 1) {
        for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) {
            const int c = i % BN, which = i / BN;
            const double v = (double)sm.statw[0][i] + (double)sm.statw[1][i] + (double)sm.statw[2][i] + (double)sm.statw[3][i];
            if (c < p.N && v != 0.0) atomicAdd(stats + (long)which * p.N + c, v);
        }
    }
}

// ------------------------------------------------------------------------------------------- CTA-pair forward
// The plain-tf32 persistent kernel on a CTA PAIR (cluster of 2, tcgen05 cta_group::2): one UMMA covers 256 output
// pixels x BN channels — CTA r of the pair owns pixel tile 2*mp + r (its A tile, its 128 TMEM lanes, its epilogue) and
// loads only HALF of the weight tile (BN/2 rows); the tensor core reads A from both CTAs and the two weight halves from
// the CTA that holds them.  Per CTA and k-block that is 16 KB of A + BN/8 KB of weights instead of 16 + BN/4: half the
// weight bytes from L2 and from shared memory (the single-CTA kernel re-fetches the whole weight tile for every pixel
// tile and reads 118 B/clk of shared memory per 128x128 MMA — at the port's limit), and stages of 32 KB (BN = 256)
// allow a 6-deep ring.  Only the leader CTA (rank 0) issues MMAs; its tcgen05.commit multicasts the "stage free" /
// "accumulator full" arrivals to both CTAs; every TMA load of either CTA signals the LEADER's full barrier
// (.cta_group::2 form, peer bit of the mbarrier address cleared); the epilogue warps of both CTAs release the
// accumulator on the leader's barrier.  Used for the data gradients (and the plain-tf32 forward mode).
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;     // shared::cluster address with the CTA-pair rank bit cleared = the leader's

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive::one on the barrier at the same shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((unsigned short)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                             int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
        "%6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// instruction descriptor for the pair MMA: M = 256
__host__ __device__ constexpr uint32_t make_idesc_m256(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int BN, int STAGES>
struct PairSmem {
    alignas(1024) float a[STAGES][BM * BK];            // this CTA's 128 pixel rows
    alignas(1024) float b[STAGES][(BN / 2) * BK];      // this CTA's half of the weight tile
    alignas(16) float epi[4][32 * EPL];
    typename StatT<BN>::type statw[4][2 * BN];
    uint64_t full[STAGES], empty[STAGES], tfull[2], tempty[2];
    uint32_t tmem_base;
};
constexpr int PAIR_THREADS = 32 * 6;

template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1)
    tc_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, float* __restrict__ y,
                   const float* __restrict__ bias, double* __restrict__ stats, FwdParams p, TileSched ts) {
    extern __shared__ uint8_t raw[];
    using Smem = PairSmem<BN, STAGES>;
    Smem& sm = *reinterpret_cast<Smem*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x / 2, n_pairs = gridDim.x / 2;
    const int tiles_per_img = p.tiles_w * p.tiles_h;
    const int m_tiles = p.B * tiles_per_img;
    const int pair_total = ((m_tiles + 1) / 2) * ts.n_tiles;        // (pixel-tile pair, N tile), N tiles of a pair adjacent
    const int cblocks = (p.Cin + BK - 1) / BK;
    const int num_k = p.n_taps * cblocks;
    constexpr uint32_t TMEM_COLS = 2 * BN;
    constexpr uint32_t STAGE_BYTES_CTA_B = (BN / 2) * BK * sizeof(float);

    if (stats)
        for (int i = threadIdx.x; i < 4 * 2 * BN; i += blockDim.x) (&sm.statw[0][0])[i] = 0;
    if (warp == 0 && lane == 0) { prefetch_tmap(&map_x); prefetch_tmap(&map_w); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
            for (int a = 0; a < 2; ++a) { mbar_init(&sm.tfull[a], 1); mbar_init(&sm.tempty[a], 8); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc2(&sm.tmem_base, TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();          // both CTAs' barriers are initialised before any remote arrival / peer-signalling TMA load
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    pdl_entry();      // barriers / tensor memory / descriptor prefetch above overlap the previous kernel's tail

    if (warp == 0) {
        // whole warp, warp-uniform values, one elected lane issues (see elect_one)
        uint32_t s = 0, ph = 0;
        const uint32_t tx_bytes = 2u * ((uint32_t)(p.TW * p.TH * BK * sizeof(float)) + STAGE_BYTES_CTA_B);
        for (int t = pair; t < pair_total; t += n_pairs) {
            const int mp = t / ts.n_tiles, n0 = (t % ts.n_tiles) * BN;
            const int mt = 2 * mp + (int)rank;                 // may be == m_tiles (odd count): coordinates fall outside
            const int img = mt / tiles_per_img, tt = mt % tiles_per_img;      // the batch dimension -> TMA zero fill
            const int h0 = (tt / p.tiles_w) * p.TH * p.in_stride, w0 = (tt % p.tiles_w) * p.TW * p.in_stride;
            const int wn = n0 + (int)rank * (BN / 2) + img * p.w_row_off, wk_img = img * p.w_k_off;
            for (int tap = 0; tap < p.n_taps; ++tap) {
                const int xw = w0 + p.dw[tap], xh = h0 + p.dh[tap], wk = p.wk[tap] + wk_img;
                for (int c0 = 0; c0 < p.Cin; c0 += BK) {
                    mbar_wait(&sm.empty[s], ph ^ 1);
                    if (elect_one()) {
                        if (leader)      // the leader's barrier counts the bytes of BOTH CTAs' loads of this stage
                            mbar_expect_tx(&sm.full[s], tx_bytes);
                        tma2_load_4d(sm.a[s], &map_x, &sm.full[s], c0, xw, xh, img);
                        tma2_load_2d(sm.b[s], &map_w, &sm.full[s], wk + c0, wn);
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (leader) {
            constexpr uint32_t idesc = make_idesc_m256(BN);
            uint32_t s = 0, ph = 0, i = 0;
            for (int t = pair; t < pair_total; t += n_pairs, ++i) {
                const uint32_t acc = i & 1, aph = (i >> 1) & 1;
                mbar_wait(&sm.tempty[acc], aph ^ 1);          // the epilogue warps of both CTAs have drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem + acc * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&sm.full[s], ph);
                    tc_fence_after();
                    const uint64_t da = make_desc(smem_u32(sm.a[s]), 16, 1024);
                    const uint64_t db = make_desc(smem_u32(sm.b[s]), 16, 1024);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) umma_tf32_2cta(d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        umma_commit_pair(&sm.empty[s]);
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                if (elect_one()) umma_commit_pair(&sm.tfull[acc]);
            }
        }
        __syncwarp();
    } else {
        const int q = warp % 4;  // TMEM lane quarter
        const uint32_t wbase = smem_u32(sm.epi[q]);
        const bool plain = (bias == nullptr) && p.act == 0;
        const bool local_stats = stats != nullptr && ts.n_tiles == 1;
        typename StatT<BN>::type* sw = sm.statw[q];
        uint32_t i = 0;
        for (int t = pair; t < pair_total; t += n_pairs, ++i) {
            const uint32_t acc = i & 1, aph = (i >> 1) & 1;
            const int mp = t / ts.n_tiles, n0 = (t % ts.n_tiles) * BN;
            const int mt = 2 * mp + (int)rank;
            const bool tile_ok = mt < m_tiles;
            const int img = mt / tiles_per_img, tt = mt % tiles_per_img;
            const int h0 = (tt / p.tiles_w) * p.TH, w0 = (tt % p.tiles_w) * p.TW;
            long row_off[8];
            bool row_ok[8];
#pragma unroll
            for (int r8 = 0; r8 < 8; ++r8) {
                const int r = 32 * q + 4 * r8 + lane / 8;
                const int th = r / p.TW, tw = r % p.TW;
                const int oh = h0 + th, ow = w0 + tw;
                row_ok[r8] = tile_ok && th < p.TH && oh < p.OH && ow < p.OW;
                row_off[r8] = ((long)img * p.YH + (long)oh * p.osy + p.ooy) * p.YW + (long)ow * p.osx + p.oox;
            }
            mbar_wait(&sm.tfull[acc], aph);
            tc_fence_after();
            for (int c0 = 0; c0 < BN; c0 += 32) {
                if (n0 + c0 >= p.N) break;
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + acc * BN + (uint32_t)c0, v);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4)
                    sts128(wbase + (uint32_t)(lane * EPL + 4 * c4) * 4u, __uint_as_float(v[4 * c4]),
                           __uint_as_float(v[4 * c4 + 1]), __uint_as_float(v[4 * c4 + 2]), __uint_as_float(v[4 * c4 + 3]));
                __syncwarp();
                const int col = n0 + c0 + 4 * (lane % 8);
                const bool col_ok = col < p.N;
                float4 bv = make_float4(0, 0, 0, 0);
                if (bias && col_ok) bv = __ldg(reinterpret_cast<const float4*>(bias + col));
                float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
#pragma unroll
                for (int r8 = 0; r8 < 8; ++r8) {
                    const int r = 4 * r8 + lane / 8;
                    float4 o = lds128(wbase + (uint32_t)(r * EPL + 4 * (lane % 8)) * 4u);
                    if (row_ok[r8] && col_ok) {
                        if (stats) {
                            s1[0] += o.x; s1[1] += o.y; s1[2] += o.z; s1[3] += o.w;
                            s2[0] += o.x * o.x; s2[1] += o.y * o.y; s2[2] += o.z * o.z; s2[3] += o.w * o.w;
                        }
                        if (!plain) {
                            o.x = act_fwd(o.x + bv.x, p.act); o.y = act_fwd(o.y + bv.y, p.act);
                            o.z = act_fwd(o.z + bv.z, p.act); o.w = act_fwd(o.w + bv.w, p.act);
                        }
                        *reinterpret_cast<float4*>(y + row_off[r8] * p.ldy + col) = o;
                    }
                }
                if (stats) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 8); s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16);
                        s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 8); s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
                    }
                    if (lane < 8 && col_ok) {
                        if (local_stats) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) { sw[c0 + 4 * lane + j] += s1[j]; sw[BN + c0 + 4 * lane + j] += s2[j]; }
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                atomicAdd(stats + col + j, (double)s1[j]);
                                atomicAdd(stats + p.N + col + j, (double)s2[j]);
                            }
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&sm.tempty[acc]);      // (the leader waits for all 8 epilogue warps of the pair)
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();          // neither CTA may leave (or free TMEM) while the peer's MMAs / arrivals can still touch it
    if (warp == 1) tmem_dealloc2(tmem, TMEM_COLS);
    if (stats && ts.n_tiles == 1) {
        for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) {
            const int c = i % BN, which = i / BN;
            const double v = (double)sm.statw[0][i] + (double)sm.statw[1][i] + (double)sm.statw[2][i] + (double)sm.statw[3][i];
            if (c < p.N && v != 0.0) atomicAdd(stats + (long)which * p.N + c, v);
        }
    }
}

// ---- 16-bit split forward with the activation planes in TENSOR MEMORY (tcgen05.mma, A operand from TMEM) ---------------
// tc_fwd_persist<.., 2, ..> is bound by shared-memory bandwidth, not by the tensor pipe: per 32-channel k-block and
// 128 x 128 tile it moves 112 KB through shared memory (TMA writes the fp32 pixel tile and the weight planes, the converter
// warps read the tile and write two 16-bit planes, the six MMAs read 4 KB of A and 4 KB of B each) = 875 clk at 128 B/clk
// against 384 clk of kind::f16 tensor time.  Here the converter warps write the hi / lo planes straight into TMEM
// (tcgen05.st, one pixel row per thread = one TMEM lane, 8 columns per 16-channel k-step and plane) and the MMAs take A from
// there: 72 KB per k-block (562 clk), and a stage shrinks to the fp32 landing tile + the weight planes (32 KB at BN = 128:
// a 6-deep ring).  TMEM: accumulators in columns [0, 2 BN), activation planes of stage s in [2 BN + 32 s, 2 BN + 32 s + 32)
// as [plane][k-step][8 columns].  BN <= 128 (two accumulators of 256 columns would leave no room for the planes).
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void trace_mark(unsigned long long* trace, int slot) {
    if (trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        trace[(size_t)blockIdx.x * 16 + slot] = t;
    }
}
template <int BN, int STAGES>
struct TsSmem {
    static constexpr int B_PLANE = BN * BK * 2;
    alignas(1024) float a[STAGES][BM * BK];                 // fp32 landing buffer (128-byte swizzle)
    alignas(1024) uint8_t b[STAGES][2 * B_PLANE];           // [hi | lo] weight planes (64-byte swizzle)
    alignas(16) float epi[4][32 * EPL];
    typename StatT<BN>::type statw[4][2 * BN];
    uint64_t full[STAGES], empty[STAGES], conv[STAGES], tfull[2], tempty[2];
    uint32_t tmem_base;
    volatile uint32_t produced;
    uint32_t last_cta;
};
constexpr int TS_CONV_WARPS = 8;
constexpr int TS_THREADS = 32 * (6 + TS_CONV_WARPS + 1);

template <int BN, int STAGES>
__global__ void __launch_bounds__(TS_THREADS, 1) tc_fwd_ts(const __grid_constant__ CUtensorMap map_x,
                                                           const __grid_constant__ CUtensorMap map_w, float* __restrict__ y,
                                                           const float* __restrict__ bias, double* __restrict__ stats, FwdParams p,
                                                           TileSched ts) {
    static_assert(2 * BN + 32 * STAGES <= 512, "tc_fwd_ts: tensor memory columns");
    extern __shared__ uint8_t raw[];
    using Smem = TsSmem<BN, STAGES>;
    Smem& sm = *reinterpret_cast<Smem*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int tiles_per_img = p.tiles_w * p.tiles_h;
    const int cblocks = (p.Cin + BK - 1) / BK;
    const int num_k = p.n_taps * cblocks;
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t A_COL0 = 2 * BN;
    if (threadIdx.x == 0) trace_mark(p.trace, 0);

    if (stats)
        for (int i = threadIdx.x; i < 4 * 2 * BN; i += blockDim.x) (&sm.statw[0][0])[i] = 0;
    if (warp == 0 && lane == 0) { prefetch_tmap(&map_x); prefetch_tmap(&map_w); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); mbar_init(&sm.conv[s], TS_CONV_WARPS); }
            for (int a = 0; a < 2; ++a) { mbar_init(&sm.tfull[a], 1); mbar_init(&sm.tempty[a], 4); }
            sm.produced = 0;
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(&sm.tmem_base, TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    if (threadIdx.x == 0) trace_mark(p.trace, 1);
    pdl_entry();      // barriers / tensor memory / descriptor prefetch above overlap the previous kernel's tail
    if (threadIdx.x == 0) trace_mark(p.trace, 2);

    if (warp == 0) {
        // TMA producer: the whole warp walks (tile, tap, channel block) with warp-uniform values, one elected lane issues
        uint32_t g = 0, s = 0, ph = 0;
        const uint32_t tx_bytes = (uint32_t)(p.TW * p.TH * BK * sizeof(float)) + 2u * (uint32_t)Smem::B_PLANE;
        for (int t = blockIdx.x; t < ts.total; t += gridDim.x) {
            const int mt = t / ts.n_tiles, n0 = (t % ts.n_tiles) * BN;
            const int img = mt / tiles_per_img, tt = mt % tiles_per_img;
            const int h0 = (tt / p.tiles_w) * p.TH * p.in_stride, w0 = (tt % p.tiles_w) * p.TW * p.in_stride;
            const int wn = n0 + img * p.w_row_off, wk_img = img * p.w_k_off;
            for (int tap = 0; tap < p.n_taps; ++tap) {
                const int xw = w0 + p.dw[tap], xh = h0 + p.dh[tap], wk = p.wk[tap] + wk_img;
                for (int c0 = 0; c0 < p.Cin; c0 += BK) {
                    mbar_wait(&sm.empty[s], ph ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&sm.full[s], tx_bytes);
                        tma_load_4d(sm.a[s], &map_x, &sm.full[s], c0, xw, xh, img);
                        tma_load_3d(sm.b[s], &map_w, &sm.full[s], wk + c0, wn, 0);
                        if (g == 0) trace_mark(p.trace, 3);
                        sm.produced = ++g;
                    } else {
                        ++g;
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
        if (lane == 0) sm.produced = 0x7fffffffu;
        __syncwarp();
    } else if (warp == 1) {
        // MMA issuer: same scheme (tcgen05.mma / commit must come from ONE thread: elect.sync picks the same lane every time)
        const uint32_t idesc16 = p.half16 ? make_idesc_f16(BN) : make_idesc_bf16(BN);
        uint32_t s = 0, ph = 0, i = 0;
        for (int t = blockIdx.x; t < ts.total; t += gridDim.x, ++i) {
            const uint32_t acc = i & 1, aph = (i >> 1) & 1;
            mbar_wait(&sm.tempty[acc], aph ^ 1);
            tc_fence_after();
            const uint32_t d = tmem + acc * BN;
            for (int kb = 0; kb < num_k; ++kb) {
                mbar_wait(&sm.conv[s], ph);          // planes of this stage are in TMEM (and, before that, the weights landed)
                if (i == 0 && kb == 0 && lane == 0) trace_mark(p.trace, 5);
                tc_fence_after();
                const uint32_t ah = tmem + A_COL0 + s * 32u, al = ah + 16u;
                const uint64_t bh = make_desc(smem_u32(sm.b[s]), 16, 512, 4);
                const uint64_t bl = make_desc(smem_u32(sm.b[s] + Smem::B_PLANE), 16, 512, 4);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        umma_f16_ts(d, al + 8 * k, bh + 2 * k, idesc16, (kb | k) != 0);
                        umma_f16_ts(d, ah + 8 * k, bl + 2 * k, idesc16, 1);
                        umma_f16_ts(d, ah + 8 * k, bh + 2 * k, idesc16, 1);
                    }
                    umma_commit(&sm.empty[s]);
                }
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            if (elect_one()) umma_commit(&sm.tfull[acc]);
        }
        __syncwarp();
    } else if (warp < 6) {
        const int q = warp % 4;
        const uint32_t wbase = smem_u32(sm.epi[q]);
        const bool plain = (bias == nullptr) && p.act == 0 && !p.lab && p.ch_scale == nullptr;
        // warp-private shared slots whenever all tiles of this CTA cover the same channels (the host sizes the grid as a
        // multiple of the N-tile count): global fp64 atomics per tile made the 256-wide layers 2x slower
        const bool local_stats = stats != nullptr && (gridDim.x % ts.n_tiles) == 0;
        typename StatT<BN>::type* sw = sm.statw[q];
        uint32_t i = 0;
        for (int t = blockIdx.x; t < ts.total; t += gridDim.x, ++i) {
            const uint32_t acc = i & 1, aph = (i >> 1) & 1;
            const int mt = t / ts.n_tiles, n0 = (t % ts.n_tiles) * BN;
            const int img = mt / tiles_per_img, tt = mt % tiles_per_img;
            const int h0 = (tt / p.tiles_w) * p.TH, w0 = (tt % p.tiles_w) * p.TW;
            long row_off[8];
            bool row_ok[8];
#pragma unroll
            for (int r8 = 0; r8 < 8; ++r8) {
                const int r = 32 * q + 4 * r8 + lane / 8;
                const int th = r / p.TW, tw = r % p.TW;
                const int oh = h0 + th, ow = w0 + tw;
                row_ok[r8] = th < p.TH && oh < p.OH && ow < p.OW;
                row_off[r8] = ((long)img * p.YH + (long)oh * p.osy + p.ooy) * p.YW + (long)ow * p.osx + p.oox;
            }
            mbar_wait(&sm.tfull[acc], aph);
            if (i == 0 && warp == 2 && lane == 0) trace_mark(p.trace, 6);
            tc_fence_after();
            for (int c0 = 0; c0 < BN; c0 += 32) {
                if (n0 + c0 >= p.N) break;
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + acc * BN + (uint32_t)c0, v);
                const bool tr = p.trace && i == 0 && c0 == 0 && warp == 2 && lane == 0;
                if (tr) trace_mark(p.trace, 12);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4)
                    sts128(wbase + (uint32_t)(lane * EPL + 4 * c4) * 4u, __uint_as_float(v[4 * c4]),
                           __uint_as_float(v[4 * c4 + 1]), __uint_as_float(v[4 * c4 + 2]), __uint_as_float(v[4 * c4 + 3]));
                __syncwarp();
                if (tr) trace_mark(p.trace, 13);
                const int col = n0 + c0 + 4 * (lane % 8);
                const bool col_ok = col < p.N;
                float4 bv = make_float4(0, 0, 0, 0), sv = make_float4(1.f, 1.f, 1.f, 1.f);
                if (bias && col_ok) bv = __ldg(reinterpret_cast<const float4*>(bias + col));
                if (p.ch_scale && col_ok) sv = __ldg(reinterpret_cast<const float4*>(p.ch_scale + col));
                float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
#pragma unroll
                for (int r8 = 0; r8 < 8; ++r8) {
                    const int r = 4 * r8 + lane / 8;
                    float4 o = lds128(wbase + (uint32_t)(r * EPL + 4 * (lane % 8)) * 4u);
                    o.x *= p.out_scale; o.y *= p.out_scale; o.z *= p.out_scale; o.w *= p.out_scale;
                    if (row_ok[r8] && col_ok) {
                        if (stats) {
                            s1[0] += o.x; s1[1] += o.y; s1[2] += o.z; s1[3] += o.w;
                            s2[0] += o.x * o.x; s2[1] += o.y * o.y; s2[2] += o.z * o.z; s2[3] += o.w * o.w;
                        }
                        if (!plain) {
                            o.x = act_fwd(fmaf(o.x, sv.x, bv.x), p.act); o.y = act_fwd(fmaf(o.y, sv.y, bv.y), p.act);
                            o.z = act_fwd(fmaf(o.z, sv.z, bv.z), p.act); o.w = act_fwd(fmaf(o.w, sv.w, bv.w), p.act);
                            if (p.lab) {
                                o.x = fmaf(o.x, p.lab_s, p.lab_b); o.y = fmaf(o.y, p.lab_s, p.lab_b);
                                o.z = fmaf(o.z, p.lab_s, p.lab_b); o.w = fmaf(o.w, p.lab_s, p.lab_b);
                            }
                        }
                        *reinterpret_cast<float4*>(y + row_off[r8] * p.ldy + col) = o;
                    }
                }
                if (tr) trace_mark(p.trace, 14);
                if (stats) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 8); s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16);
                        s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 8); s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
                    }
                    if (lane < 8 && col_ok) {
                        if (local_stats) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) { sw[c0 + 4 * lane + j] += s1[j]; sw[BN + c0 + 4 * lane + j] += s2[j]; }
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                atomicAdd(stats + col + j, (double)s1[j]);
                                atomicAdd(stats + p.N + col + j, (double)s2[j]);
                            }
                        }
                    }
                }
                __syncwarp();
                if (tr) trace_mark(p.trace, 15);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.tempty[acc]);
            if (warp == 2 && lane == 0) trace_mark(p.trace, i == 0 ? 7 : 8);
        }
    } else if (warp == 6 + TS_CONV_WARPS) {
        if (p.prefetch > 0) {          // L2 prefetch warp (see tc_fwd_persist)
            const uint32_t my_tiles = (uint32_t)((ts.total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);
            const uint32_t total_k = my_tiles * (uint32_t)num_k;
            const uint32_t dmin = STAGES, dmax = STAGES + (uint32_t)p.prefetch;
            uint32_t gp = 0;
            while (gp < total_k) {
                const uint32_t done = sm.produced;
                if (done >= total_k) break;
                if (gp < done + dmin) gp = done + dmin;
                if (gp >= total_k) break;
                if (gp > done + dmax) { __nanosleep(200); continue; }
                const int t = (int)blockIdx.x + (int)(gp / (uint32_t)num_k) * (int)gridDim.x;
                const int kb = (int)(gp % (uint32_t)num_k);
                const int mt = t / ts.n_tiles;
                const int img = mt / tiles_per_img, tt = mt % tiles_per_img;
                const int h0 = (tt / p.tiles_w) * p.TH, w0 = (tt % p.tiles_w) * p.TW;
                const int tap = kb / cblocks, c0 = (kb % cblocks) * BK;
                if (t % ts.n_tiles == 0 && img < p.B) {
#pragma unroll
                    for (int rr = 0; rr < BM / 32; ++rr) {
                        const int r = lane + 32 * rr;
                        const int th = r / p.TW, tw = r % p.TW;
                        const int ih = (h0 + th) * p.in_stride + p.dh[tap], iw = (w0 + tw) * p.in_stride + p.dw[tap];
                        if (th < p.TH && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
                            const float* a = p.x + (((long)img * p.H + ih) * p.W + iw) * p.ldx + c0;
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
                            if (((uintptr_t)a & 127) != 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 31));
                        }
                    }
                }
                ++gp;
            }
        }
    } else {
        // Converter warps (CTA warps 6..13): a warp can touch the TMEM lanes of quarter (warp % 4) only, so warp w converts
        // pixel rows 32 (w % 4) .. +31 — one row per lane — and the two warps of a quarter split the k-block's two
        // 16-channel k-steps.  Row r of the fp32 tile: 128 bytes, 16-byte chunk j at physical chunk j ^ (r & 7): the
        // eight lanes of an LDS.128 phase read eight different chunks (no bank conflict).
        const int q = warp % 4, ks = (warp - 6) / 4;
        const int r = 32 * q + lane;
        const uint32_t row_off = (uint32_t)r * 128u, sw7 = (uint32_t)(r & 7);
        // The TMEM stores of k-block g are waited for (tcgen05.wait::st) and signalled one iteration LATER, after the
        // loads and the arithmetic of k-block g + 1: their latency is off the warp's per-k-block critical path.
        uint32_t s = 0, ph = 0;
        int pending = -1;           // stage whose planes have been stored but not yet signalled
        for (int t = blockIdx.x; t < ts.total; t += gridDim.x) {
            for (int kb = 0; kb < num_k; ++kb) {
                mbar_wait(&sm.full[s], ph);
                if (pending < 0 && warp == 6 && lane == 0) trace_mark(p.trace, 4);
                const uint32_t a_row = smem_u32(sm.a[s]) + row_off;
                float4 x[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) x[i] = lds128(a_row + ((((uint32_t)(4 * ks + i)) ^ sw7) << 4));
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (p.half16) {
                        hi[2 * i] = f16x2_rn(x[i].x, x[i].y); hi[2 * i + 1] = f16x2_rn(x[i].z, x[i].w);
                        lo[2 * i] = f16x2_rn(x[i].x - f16_lo(hi[2 * i]), x[i].y - f16_hi(hi[2 * i]));
                        lo[2 * i + 1] = f16x2_rn(x[i].z - f16_lo(hi[2 * i + 1]), x[i].w - f16_hi(hi[2 * i + 1]));
                    } else {
                        hi[2 * i] = bf16x2_rn(x[i].x, x[i].y); hi[2 * i + 1] = bf16x2_rn(x[i].z, x[i].w);
                        lo[2 * i] = bf16x2_rn(x[i].x - __uint_as_float(hi[2 * i] << 16), x[i].y - __uint_as_float(hi[2 * i] & 0xffff0000u));
                        lo[2 * i + 1] = bf16x2_rn(x[i].z - __uint_as_float(hi[2 * i + 1] << 16),
                                                  x[i].w - __uint_as_float(hi[2 * i + 1] & 0xffff0000u));
                    }
                }
                if (pending >= 0) {
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.conv[pending]);
                }
                const uint32_t ta = tmem + ((uint32_t)(32 * q) << 16) + A_COL0 + s * 32u + 8u * (uint32_t)ks;
                tmem_st8(ta, hi);
                tmem_st8(ta + 16u, lo);
                pending = (int)s;
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
        if (pending >= 0) {
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.conv[pending]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) trace_mark(p.trace, 9);
    if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
    if (stats && (gridDim.x % ts.n_tiles) == 0) {
        const int n0_cta = (int)(blockIdx.x % ts.n_tiles) * BN;       // the one N tile this CTA worked on
        for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) {
            const int c = n0_cta + i % BN, which = i / BN;
            const double v = (double)sm.statw[0][i] + (double)sm.statw[1][i] + (double)sm.statw[2][i] + (double)sm.statw[3][i];
            if (c < p.N && v != 0.0) atomicAdd(stats + (long)which * p.N + c, v);
        }
    }
    if (threadIdx.x == 0) trace_mark(p.trace, 10);
    if (stats && p.fin_counter) {
        // last-CTA finalize: every thread's statistics atomics are ordered before the CTA's ticket (fence + barrier);
        // the CTA drawing the last ticket sees all of them (fence after the ticket, L2 loads)
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) sm.last_cta = atomicAdd(p.fin_counter, 1u) == gridDim.x - 1 ? 1u : 0u;
        __syncthreads();
        if (sm.last_cta) {
            __threadfence();
            for (int c = threadIdx.x; c < p.N; c += blockDim.x)
                bn_finalize_channel(__ldcg(stats + c), __ldcg(stats + p.N + c), c, p.fin_w, p.fin_b, p.fin_rmean, p.fin_rvar,
                                    p.fin_mean, p.fin_invstd, p.fin_scale, p.fin_shift, p.fin_M, p.fin_momentum, p.fin_eps);
        }
        if (threadIdx.x == 0) trace_mark(p.trace, 11);
        if (p.fin_y) {
            unsigned int* flag = p.fin_counter + 1;
            if (sm.last_cta) {
                __threadfence();
                __syncthreads();
                if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(1u) : "memory");
            } else {
                if (threadIdx.x == 0) {
                    // bounded: the CTAs of this grid are co-resident by construction (<= one per SM); if that ever fails the
                    // launch ends in a trap (a loud error) instead of a hang
                    unsigned int v = 0, polls = 0;
                    while (true) {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                        if (v) break;
                        if (++polls > (1u << 23)) __trap();
                        __nanosleep(64);
                    }
                }
                __syncthreads();
            }
            // flat, grid-strided pass over the whole [M, N] tensor (every CTA's tiles are visible: each fenced before its
            // ticket), eight 16-byte loads in flight per thread
            const int VC = p.N / 4;
            const long n4 = p.fin_M * VC, stride = (long)gridDim.x * blockDim.x;
            const float ls = p.fin_lab_s ? __ldg(p.fin_lab_s) : 1.f, lb = p.fin_lab_s ? __ldg(p.fin_lab_b) : 0.f;
            constexpr int U = 8;
            for (long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += U * stride) {
                float4 v[U];
                long row[U];
                int col[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const long i = i0 + u * stride;
                    row[u] = i / VC;
                    col[u] = (int)(i - row[u] * VC) * 4;
                    if (i < n4) v[u] = __ldcg(reinterpret_cast<const float4*>(y + row[u] * p.ldy + col[u]));
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (i0 + u * stride >= n4) break;
                    const float4 sc = __ldcg(reinterpret_cast<const float4*>(p.fin_scale + col[u]));
                    const float4 sh = __ldcg(reinterpret_cast<const float4*>(p.fin_shift + col[u]));
                    float4 o = make_float4(act_fwd(fmaf(v[u].x, sc.x, sh.x), p.fin_act), act_fwd(fmaf(v[u].y, sc.y, sh.y), p.fin_act),
                                           act_fwd(fmaf(v[u].z, sc.z, sh.z), p.fin_act), act_fwd(fmaf(v[u].w, sc.w, sh.w), p.fin_act));
                    if (p.fin_lab_s) { o.x = ls * o.x + lb; o.y = ls * o.y + lb; o.z = ls * o.z + lb; o.w = ls * o.w + lb; }
                    if (p.fin_post) {
                        const float4 a = __ldg(reinterpret_cast<const float4*>(p.fin_post + row[u] * p.fin_ldpost + col[u]));
                        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
                    }
                    *reinterpret_cast<float4*>(p.fin_y + row[u] * p.fin_ldy + col[u]) = o;
                }
            }
        }
    }
}

// ---- CTA-pair forward on 16-bit split operands (the 3xFP16 / 3xBF16 modes) ------------------------------------------
// tc_pair_kernel's layout with the converter stage of tc_fwd_persist<.., 2, ..>: each CTA lands its own fp32 pixel tile and
// its HALF of the two weight planes on its OWN full barrier, its eight converter warps write the hi / lo 16-bit planes of
// the pixel tile and then signal the LEADER's conv barrier (one elected cluster-scope arrive per warp, 16 per stage); the
// leader issues three kind::f16 cta_group::2 MMAs (256 pixels x BN channels) per 16-channel k-step.  A stage is
// 16 KB (fp32 landing) + 16 KB (pixel planes) + BN/16 KB (weight planes): 48 KB at BN = 256 where the single-CTA kernel
// needs 64 KB, so the ring is 4 deep instead of 3 and the weight planes cost half the L2 -> SM bytes.
__device__ __forceinline__ void mbar_arrive_leader_release(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (++spins > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc16_m256(int n, int bf16) {
    return (1u << 4) | ((uint32_t)bf16 << 7) | ((uint32_t)bf16 << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int BN, int STAGES>
struct Pair16Smem {
    static constexpr int B_PLANE = (BN / 2) * BK * 2;       // bytes of one 16-bit plane of this CTA's half weight tile
    static constexpr int A2_PLANE = BM * BK * 2;
    alignas(1024) float a[STAGES][BM * BK];                 // fp32 landing buffer of this CTA's 128 pixel rows
    alignas(1024) uint8_t a2[STAGES][2 * A2_PLANE];         // [hi | lo] planes written by the converter warps
    alignas(1024) uint8_t b[STAGES][2 * B_PLANE];           // [hi | lo] planes of BN/2 weight rows
    alignas(16) float epi[4][32 * EPL];
    typename StatT<BN>::type statw[4][2 * BN];
    uint64_t full[STAGES], empty[STAGES], conv[STAGES], tfull[2], tempty[2];
    uint32_t tmem_base;
    volatile uint32_t produced;
};
constexpr int PAIR16_CONV_WARPS = 8;
constexpr int PAIR16_THREADS = 32 * (6 + PAIR16_CONV_WARPS + 1);

template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR16_THREADS, 1)
    tc_pair16_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, float* __restrict__ y,
                     const float* __restrict__ bias, double* __restrict__ stats, FwdParams p, TileSched ts) {
    extern __shared__ uint8_t raw[];
    using Smem = Pair16Smem<BN, STAGES>;
    Smem& sm = *reinterpret_cast<Smem*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x / 2, n_pairs = gridDim.x / 2;
    const int tiles_per_img = p.tiles_w * p.tiles_h;
    const int m_tiles = p.B * tiles_per_img;
    const int pair_total = ((m_tiles + 1) / 2) * ts.n_tiles;
    const int cblocks = (p.Cin + BK - 1) / BK;
    const int num_k = p.n_taps * cblocks;
    constexpr uint32_t TMEM_COLS = 2 * BN;

    if (stats)
        for (int i = threadIdx.x; i < 4 * 2 * BN; i += blockDim.x) (&sm.statw[0][0])[i] = 0;
    if (warp == 0 && lane == 0) { prefetch_tmap(&map_x); prefetch_tmap(&map_w); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); mbar_init(&sm.conv[s], 2 * PAIR16_CONV_WARPS);
            }
            for (int a = 0; a < 2; ++a) { mbar_init(&sm.tfull[a], 1); mbar_init(&sm.tempty[a], 8); }
            sm.produced = 0;
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc2(&sm.tmem_base, TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    pdl_entry();      // barriers / tensor memory / descriptor prefetch above overlap the previous kernel's tail

    if (warp == 0) {
        if (lane == 0) {
            uint32_t g = 0;
            for (int t = pair; t < pair_total; t += n_pairs) {
                const int mp = t / ts.n_tiles, n0 = (t % ts.n_tiles) * BN;
                const int mt = 2 * mp + (int)rank;
                const int img = mt / tiles_per_img, tt = mt % tiles_per_img;
                const int h0 = (tt / p.tiles_w) * p.TH, w0 = (tt % p.tiles_w) * p.TW;
                const int wn_img = img * p.w_row_off, wk_img = img * p.w_k_off;
                for (int kb = 0; kb < num_k; ++kb, ++g) {
                    const int s = g % STAGES, ph = (g / STAGES) & 1;
                    mbar_wait(&sm.empty[s], ph ^ 1);
                    const int tap = kb / cblocks, c0 = (kb % cblocks) * BK;
                    mbar_expect_tx(&sm.full[s], (uint32_t)(p.TW * p.TH * BK * sizeof(float)) + 2u * (uint32_t)Smem::B_PLANE);
                    tma_load_4d(sm.a[s], &map_x, &sm.full[s], c0, w0 * p.in_stride + p.dw[tap], h0 * p.in_stride + p.dh[tap], img);
                    tma_load_3d(sm.b[s], &map_w, &sm.full[s], p.wk[tap] + c0 + wk_img, n0 + (int)rank * (BN / 2) + wn_img, 0);
                    sm.produced = g + 1;
                }
            }
            sm.produced = 0x7fffffffu;
        }
        __syncwarp();
    } else if (warp == 1) {
        if (leader && lane == 0) {
            const uint32_t idesc16 = make_idesc16_m256(BN, p.half16 ? 0 : 1);
            uint32_t g = 0, i = 0;
            for (int t = pair; t < pair_total; t += n_pairs, ++i) {
                const uint32_t acc = i & 1, aph = (i >> 1) & 1;
                mbar_wait_cluster(&sm.tempty[acc], aph ^ 1);
                tc_fence_after();
                const uint32_t d = tmem + acc * BN;
                for (int kb = 0; kb < num_k; ++kb, ++g) {
                    const int s = g % STAGES, ph = (g / STAGES) & 1;
                    mbar_wait_cluster(&sm.conv[s], ph);          // both CTAs' planes written, both weight halves landed
                    tc_fence_after();
                    const uint64_t ah = make_desc(smem_u32(sm.a2[s]), 16, 512, 4);
                    const uint64_t al = make_desc(smem_u32(sm.a2[s] + Smem::A2_PLANE), 16, 512, 4);
                    const uint64_t bh = make_desc(smem_u32(sm.b[s]), 16, 512, 4);
                    const uint64_t bl = make_desc(smem_u32(sm.b[s] + Smem::B_PLANE), 16, 512, 4);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        umma_f16_2cta(d, al + 2 * k, bh + 2 * k, idesc16, (kb | k) != 0);
                        umma_f16_2cta(d, ah + 2 * k, bl + 2 * k, idesc16, 1);
                        umma_f16_2cta(d, ah + 2 * k, bh + 2 * k, idesc16, 1);
                    }
                    umma_commit_pair(&sm.empty[s]);
                }
                umma_commit_pair(&sm.tfull[acc]);
            }
        }
        __syncwarp();
    } else if (warp < 6) {
        const int q = warp % 4;
        const uint32_t wbase = smem_u32(sm.epi[q]);
        const bool plain = (bias == nullptr) && p.act == 0 && !p.lab;
        const bool local_stats = stats != nullptr && ts.n_tiles == 1;
        typename StatT<BN>::type* sw = sm.statw[q];
        uint32_t i = 0;
        for (int t = pair; t < pair_total; t += n_pairs, ++i) {
            const uint32_t acc = i & 1, aph = (i >> 1) & 1;
            const int mp = t / ts.n_tiles, n0 = (t % ts.n_tiles) * BN;
            const int mt = 2 * mp + (int)rank;
            const bool tile_ok = mt < m_tiles;
            const int img = mt / tiles_per_img, tt = mt % tiles_per_img;
            const int h0 = (tt / p.tiles_w) * p.TH, w0 = (tt % p.tiles_w) * p.TW;
            long row_off[8];
            bool row_ok[8];
#pragma unroll
            for (int r8 = 0; r8 < 8; ++r8) {
                const int r = 32 * q + 4 * r8 + lane / 8;
                const int th = r / p.TW, tw = r % p.TW;
                const int oh = h0 + th, ow = w0 + tw;
                row_ok[r8] = tile_ok && th < p.TH && oh < p.OH && ow < p.OW;
                row_off[r8] = ((long)img * p.YH + (long)oh * p.osy + p.ooy) * p.YW + (long)ow * p.osx + p.oox;
            }
            mbar_wait(&sm.tfull[acc], aph);
            tc_fence_after();
            for (int c0 = 0; c0 < BN; c0 += 32) {
                if (n0 + c0 >= p.N) break;
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + acc * BN + (uint32_t)c0, v);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4)
                    sts128(wbase + (uint32_t)(lane * EPL + 4 * c4) * 4u, __uint_as_float(v[4 * c4]),
                           __uint_as_float(v[4 * c4 + 1]), __uint_as_float(v[4 * c4 + 2]), __uint_as_float(v[4 * c4 + 3]));
                __syncwarp();
                const int col = n0 + c0 + 4 * (lane % 8);
                const bool col_ok = col < p.N;
                float4 bv = make_float4(0, 0, 0, 0);
                if (bias && col_ok) bv = __ldg(reinterpret_cast<const float4*>(bias + col));
                float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
#pragma unroll
                for (int r8 = 0; r8 < 8; ++r8) {
                    const int r = 4 * r8 + lane / 8;
                    float4 o = lds128(wbase + (uint32_t)(r * EPL + 4 * (lane % 8)) * 4u);
                    o.x *= p.out_scale; o.y *= p.out_scale; o.z *= p.out_scale; o.w *= p.out_scale;
                    if (row_ok[r8] && col_ok) {
                        if (stats) {
                            s1[0] += o.x; s1[1] += o.y; s1[2] += o.z; s1[3] += o.w;
                            s2[0] += o.x * o.x; s2[1] += o.y * o.y; s2[2] += o.z * o.z; s2[3] += o.w * o.w;
                        }
                        if (!plain) {
                            o.x = act_fwd(o.x + bv.x, p.act); o.y = act_fwd(o.y + bv.y, p.act);
                            o.z = act_fwd(o.z + bv.z, p.act); o.w = act_fwd(o.w + bv.w, p.act);
                            if (p.lab) {
                                o.x = fmaf(o.x, p.lab_s, p.lab_b); o.y = fmaf(o.y, p.lab_s, p.lab_b);
                                o.z = fmaf(o.z, p.lab_s, p.lab_b); o.w = fmaf(o.w, p.lab_s, p.lab_b);
                            }
                        }
                        *reinterpret_cast<float4*>(y + row_off[r8] * p.ldy + col) = o;
                    }
                }
                if (stats) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 8); s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16);
                        s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 8); s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
                    }
                    if (lane < 8 && col_ok) {
                        if (local_stats) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) { sw[c0 + 4 * lane + j] += s1[j]; sw[BN + c0 + 4 * lane + j] += s2[j]; }
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                atomicAdd(stats + col + j, (double)s1[j]);
                                atomicAdd(stats + p.N + col + j, (double)s2[j]);
                            }
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader_release(&sm.tempty[acc]);
        }
    } else if (warp == 6 + PAIR16_CONV_WARPS) {
        // L2 prefetch warp (see tc_fwd_persist): walks this CTA's (tile, k-block) sequence ahead of its producer
        if (p.prefetch > 0) {
            const uint32_t my_tiles = (uint32_t)((pair_total - pair + n_pairs - 1) / n_pairs);
            const uint32_t total_k = my_tiles * (uint32_t)num_k;
            const uint32_t dmin = STAGES, dmax = STAGES + (uint32_t)p.prefetch;
            uint32_t gp = 0;
            while (gp < total_k) {
                const uint32_t done = sm.produced;
                if (done >= total_k) break;
                if (gp < done + dmin) gp = done + dmin;
                if (gp >= total_k) break;
                if (gp > done + dmax) { __nanosleep(200); continue; }
                const int t = pair + (int)(gp / (uint32_t)num_k) * n_pairs;
                const int kb = (int)(gp % (uint32_t)num_k);
                const int mt = 2 * (t / ts.n_tiles) + (int)rank;
                const int img = mt / tiles_per_img, tt = mt % tiles_per_img;
                const int h0 = (tt / p.tiles_w) * p.TH, w0 = (tt % p.tiles_w) * p.TW;
                const int tap = kb / cblocks, c0 = (kb % cblocks) * BK;
                if (t % ts.n_tiles == 0 && img < p.B) {
#pragma unroll
                    for (int rr = 0; rr < BM / 32; ++rr) {
                        const int r = lane + 32 * rr;
                        const int th = r / p.TW, tw = r % p.TW;
                        const int ih = (h0 + th) * p.in_stride + p.dh[tap], iw = (w0 + tw) * p.in_stride + p.dw[tap];
                        if (th < p.TH && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
                            const float* a = p.x + (((long)img * p.H + ih) * p.W + iw) * p.ldx + c0;
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
                            if (((uintptr_t)a & 127) != 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 31));
                        }
                    }
                }
                ++gp;
            }
        }
    } else {
        // converter warps: this CTA's fp32 pixel tile -> hi / lo 16-bit planes (index algebra in tc_fwd_persist)
        constexpr int CT = 32 * PAIR16_CONV_WARPS;
        const int ct = threadIdx.x - 192;
        uint32_t g = 0;
        for (int t = pair; t < pair_total; t += n_pairs) {
            for (int kb = 0; kb < num_k; ++kb, ++g) {
                const int s = g % STAGES, ph = (g / STAGES) & 1;
                mbar_wait(&sm.full[s], ph);
                const uint32_t a_base = smem_u32(sm.a[s]), l_base = smem_u32(sm.a2[s]);
                constexpr int ITEMS = BM * 4 / CT;
                float4 v0[ITEMS], v1[ITEMS];
#pragma unroll
                for (int i = 0; i < ITEMS; ++i) {
                    const uint32_t idx = (uint32_t)(ct + i * CT);
                    v0[i] = lds128(a_base + idx * 32u);
                    v1[i] = lds128(a_base + idx * 32u + 16u);
                }
#pragma unroll
                for (int i = 0; i < ITEMS; ++i) {
                    const uint32_t idx = (uint32_t)(ct + i * CT);
                    float4 x0 = v0[i], x1 = v1[i];
                    if ((idx >> 2) & 1) { float4 tmp = x0; x0 = x1; x1 = tmp; }
                    uint32_t b0, b1, b2, b3, l0, l1, l2, l3;
                    if (p.half16) {
                        b0 = f16x2_rn(x0.x, x0.y); b1 = f16x2_rn(x0.z, x0.w);
                        b2 = f16x2_rn(x1.x, x1.y); b3 = f16x2_rn(x1.z, x1.w);
                        l0 = f16x2_rn(x0.x - f16_lo(b0), x0.y - f16_hi(b0));
                        l1 = f16x2_rn(x0.z - f16_lo(b1), x0.w - f16_hi(b1));
                        l2 = f16x2_rn(x1.x - f16_lo(b2), x1.y - f16_hi(b2));
                        l3 = f16x2_rn(x1.z - f16_lo(b3), x1.w - f16_hi(b3));
                    } else {
                        b0 = bf16x2_rn(x0.x, x0.y); b1 = bf16x2_rn(x0.z, x0.w);
                        b2 = bf16x2_rn(x1.x, x1.y); b3 = bf16x2_rn(x1.z, x1.w);
                        l0 = bf16x2_rn(x0.x - __uint_as_float(b0 << 16), x0.y - __uint_as_float(b0 & 0xffff0000u));
                        l1 = bf16x2_rn(x0.z - __uint_as_float(b1 << 16), x0.w - __uint_as_float(b1 & 0xffff0000u));
                        l2 = bf16x2_rn(x1.x - __uint_as_float(b2 << 16), x1.y - __uint_as_float(b2 & 0xffff0000u));
                        l3 = bf16x2_rn(x1.z - __uint_as_float(b3 << 16), x1.w - __uint_as_float(b3 & 0xffff0000u));
                    }
                    const uint32_t pa = l_base + idx * 16u, pb = l_base + (uint32_t)Smem::A2_PLANE + idx * 16u;
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pa), "r"(b0), "r"(b1), "r"(b2), "r"(b3) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pb), "r"(l0), "r"(l1), "r"(l2), "r"(l3) : "memory");
                }
                fence_proxy_async();
                __syncwarp();
                // (every thread's plane writes are fenced to the async proxy above and ordered before lane 0's arrive by
                // the warp barrier; a cluster-scope release on this arrive doubled the k-block time: the planes of the
                // PEER are only ever read by the peer's own tensor core, after the leader has seen this arrival and issued
                // the MMA — DFINE_TC_DBG=22 restores the cluster-scope release for A/B runs)
                if (lane == 0) {
                    if (p.dbg == 22) mbar_arrive_leader_release(&sm.conv[s]);
                    else mbar_arrive_leader(&sm.conv[s]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 1) tmem_dealloc2(tmem, TMEM_COLS);
    if (stats && ts.n_tiles == 1) {
        for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) {
            const int c = i % BN, which = i / BN;
            const double v = (double)sm.statw[0][i] + (double)sm.statw[1][i] + (double)sm.statw[2][i] + (double)sm.statw[3][i];
            if (c < p.N && v != 0.0) atomicAdd(stats + (long)which * p.N + c, v);
        }
    }
}

// ------------------------------------------------------------------------------------------- wgrad
constexpr int WG_THREADS = 192;
constexpr int WG_STAGES = 3;

struct WgradParams {
    int taps_h, taps_w, pad_t, pad_l, stride;
    int Cin, Cout;
    int OH, OW, B;
    int wchunks;          // ceil(OW / 32)
    long steps_total;     // B*OH*wchunks reduction steps of 32 pixels
    long steps_per_split;
    int cin_tiles;        // ceil(Cin / BN)
    int dy5, x5;          // operand fetched as ONE 5-D box per step (channel count % 32 == 0) instead of one per 32-ch chunk
};

template <int BN>
struct WgradSmem {
    alignas(1024) float a[WG_STAGES][BM * BK];  // 4 chunks of [32 pixels][32 cout]
    alignas(1024) float b[WG_STAGES][BN * BK];  // BN/32 chunks of [32 pixels][32 cin]
    uint64_t full[WG_STAGES], empty[WG_STAGES], acc_full;
    uint32_t tmem_base;
};

template <int BN>
__global__ void __launch_bounds__(WG_THREADS) tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy,
                                                              const __grid_constant__ CUtensorMap map_x,
                                                              float* __restrict__ dwr, WgradParams p) {
    constexpr int STAGES = WG_STAGES;
    extern __shared__ uint8_t raw[];
    WgradSmem<BN>& sm = *reinterpret_cast<WgradSmem<BN>*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int co0 = blockIdx.x * BM;
    const int tap = blockIdx.y / p.cin_tiles, ci0 = (blockIdx.y % p.cin_tiles) * BN;
    const int kh = tap / p.taps_w, kw = tap % p.taps_w;
    const long s_begin = (long)blockIdx.z * p.steps_per_split;
    const long s_end = s_begin + p.steps_per_split < p.steps_total ? s_begin + p.steps_per_split : p.steps_total;
    const int num_k = (int)(s_end - s_begin);
    // chunks of 32 output channels that exist (the rest of the 128-row accumulator is never stored)
    const int co_chunks = (p.Cout - co0 + 31) / 32 < BM / 32 ? (p.Cout - co0 + 31) / 32 : BM / 32;

    if (warp == 0 && lane == 0) { prefetch_tmap(&map_dy); prefetch_tmap(&map_x); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
            mbar_init(&sm.acc_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(&sm.tmem_base, BN);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    pdl_entry();      // barriers / tensor memory / descriptor prefetch above overlap the previous kernel's tail

    if (warp == 0) {
        // whole warp, warp-uniform values, one elected lane issues (see elect_one); the (image, row, 32-pixel chunk) of a
        // step advances incrementally (two 64-bit divisions per step sat on the producer's critical path)
        long step0 = s_begin;
        int wc = (int)(step0 % p.wchunks); step0 /= p.wchunks;
        int oh = (int)(step0 % p.OH);
        int b = (int)(step0 / p.OH);
        uint32_t s = 0, ph = 0;
        const uint32_t tx_bytes = (uint32_t)(((p.dy5 ? BM / 32 : co_chunks) * 32 + BN) * BK * sizeof(float));
        for (int kb = 0; kb < num_k; ++kb) {
            mbar_wait(&sm.empty[s], ph ^ 1);
            if (elect_one()) {
                // one 5-D box {32 ch, 32 px, chunks, row, image} per operand where the channel count allows it
                // (TMA instruction issue, not bytes, bounded the per-chunk version: 8 boxes of 4 KB per step)
                mbar_expect_tx(&sm.full[s], tx_bytes);
                if (p.dy5) {
                    tma_load_5d(sm.a[s], &map_dy, &sm.full[s], 0, wc * 32, co0 / 32, oh, b);
                } else {
                    for (int c = 0; c < co_chunks; ++c)
                        tma_load_4d(sm.a[s] + c * 32 * BK, &map_dy, &sm.full[s], co0 + 32 * c, wc * 32, oh, b);
                }
                if (p.x5) {
                    tma_load_5d(sm.b[s], &map_x, &sm.full[s], 0, wc * 32 * p.stride + kw - p.pad_l, ci0 / 32,
                                oh * p.stride + kh - p.pad_t, b);
                } else {
#pragma unroll
                    for (int c = 0; c < BN / 32; ++c)
                        tma_load_4d(sm.b[s] + c * 32 * BK, &map_x, &sm.full[s], ci0 + 32 * c,
                                    wc * 32 * p.stride + kw - p.pad_l, oh * p.stride + kh - p.pad_t, b);
                }
            }
            if (++s == STAGES) { s = 0; ph ^= 1; }
            if (++wc == p.wchunks) { wc = 0; if (++oh == p.OH) { oh = 0; ++b; } }
        }
        __syncwarp();
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(BN, 1, 1);
        uint32_t s = 0, ph = 0;
        for (int kb = 0; kb < num_k; ++kb) {
            mbar_wait(&sm.full[s], ph);
            tc_fence_after();
            // chunk (32 channels) stride = 32 pixels * 128 B = 4096 B (LBO); 4-pixel k-atom = 512 B (SBO)
            const uint64_t da = make_desc(smem_u32(sm.a[s]), 32 * BK * 4, 512, 1);
            const uint64_t db = make_desc(smem_u32(sm.b[s]), 32 * BK * 4, 512, 1);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 32 / 8; ++k)  // 8 pixels (two k-atoms) per UMMA -> +1024 B = +64 in 16-byte units
                    umma_tf32(tmem, da + 64 * k, db + 64 * k, idesc, (kb | k) != 0);
                umma_commit(&sm.empty[s]);
            }
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit(&sm.acc_full);
        __syncwarp();
    } else if (num_k > 0) {
        // Epilogue: TMEM -> registers (one Cout row per lane) -> shared transpose -> vector reductions.  A lane-per-row
        // scalar atomicAdd touches 32 different sectors per instruction (4.9 M sector atomics per 3x3 layer launch,
        // ~50 us of L2 atomic time); transposed, eight lanes cover one 128-byte run of a row with
        // red.global.add.v4.f32: 8x fewer sector operations.
        const int q = warp % 4;
        mbar_wait(&sm.acc_full, 0);
        tc_fence_after();
        // every MMA (hence every TMA load) has completed: the stage ring is free to host the transpose buffers
        const uint32_t wbase = smem_u32(sm.a[0]) + (uint32_t)q * 32u * EPL * 4u;
        const long ldw = (long)p.taps_h * p.taps_w * p.Cin;
        if (q < co_chunks) {
            for (int c0 = 0; c0 < BN; c0 += 32) {
                if (ci0 + c0 >= p.Cin) break;
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)c0, v);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4)
                    sts128(wbase + (uint32_t)(lane * EPL + 4 * c4) * 4u, __uint_as_float(v[4 * c4]),
                           __uint_as_float(v[4 * c4 + 1]), __uint_as_float(v[4 * c4 + 2]), __uint_as_float(v[4 * c4 + 3]));
                __syncwarp();
                const int ci = ci0 + c0 + 4 * (lane % 8);
#pragma unroll
                for (int r8 = 0; r8 < 8; ++r8) {
                    const int r = 4 * r8 + lane / 8;
                    const int co = co0 + 32 * q + r;
                    const float4 o = lds128(wbase + (uint32_t)(r * EPL + 4 * (lane % 8)) * 4u);
                    if (co < p.Cout && ci < p.Cin) {     // Cin % 4 == 0: a float4 group is entirely in or out
                        float* dst = dwr + (long)co * ldw + (long)tap * p.Cin + ci;
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(o.x), "f"(o.y),
                                     "f"(o.z), "f"(o.w) : "memory");
                    }
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, BN);
}

// ------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

// Element type of the tensor maps.  FLOAT32 is a plain copy (the tensor core then truncates the low 13
// mantissa bits); TFLOAT32 makes the TMA unit round to nearest tf32 while copying, which measured ~16 B/clk per
// SM (one 128-byte row per ~8.5 cycles) on B200 and throttled every GEMM of the network (profiles/README.md).
// DFINE_TMA_TF32=1 restores the rounding maps for A/B measurements.
CUtensorMapDataType map_dtype() {
    static const CUtensorMapDataType t = [] {
        const char* e = getenv("DFINE_TMA_TF32");
        return (e && e[0] == '1') ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    }();
    return t;
}

// 4-D fp32 tensor map (C, W, H, B) over an NHWC activation with pixel stride ld (elements), zero OOB fill.
// The box covers box_w x box_h pixels sampled every `estride` pixels (strided convolutions read their
// input through the map's element strides; there is no im2col buffer).
int make_map4(CUtensorMap* m, const float* base, long C, long W, long H, long B, long ld, int box_c, int box_w,
              int box_h, int estride, const char* who, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B,
              CUtensorMapDataType dtype = map_dtype()) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { dfine_set_error("%s: cuTensorMapEncodeTiled unavailable", who); return -2; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)ld * 4 * W, (cuuint64_t)ld * 4 * W * H};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)(box_w * estride), (cuuint32_t)(box_h * estride), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
    if (box[1] > 256 || box[2] > 256) { dfine_set_error("%s: TMA box %u x %u exceeds 256", who, box[1], box[2]); return -1; }
    CUresult r = enc(m, dtype, 4, (void*)base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        dfine_set_error("%s: cuTensorMapEncodeTiled(4d) failed (%d) C=%ld W=%ld H=%ld B=%ld ld=%ld box=%d,%d,%d es=%d",
                        who, (int)r, C, W, H, B, ld, box_c, box_w, box_h, estride);
        return -2;
    }
    return 0;
}
// 5-D view {32 channels of a chunk, W, chunk index, H, B} of an NHWC activation whose channel count is a multiple
// of 32: one box {32, box_w*estride, box_chunks, 1, 1} lands in shared memory as [chunk][pixel][32 ch] — the
// MN-major operand layout of the weight-gradient kernel — with a single TMA instruction.
int make_map5(CUtensorMap* m, const float* base, long C, long W, long H, long B, long ld, int box_w, int box_chunks,
              int estride, const char* who) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { dfine_set_error("%s: cuTensorMapEncodeTiled unavailable", who); return -2; }
    cuuint64_t dims[5] = {32, (cuuint64_t)W, (cuuint64_t)(C / 32), (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[4] = {(cuuint64_t)ld * 4, 128, (cuuint64_t)ld * 4 * W, (cuuint64_t)ld * 4 * W * H};
    cuuint32_t box[5] = {32, (cuuint32_t)(box_w * estride), (cuuint32_t)box_chunks, 1, 1};
    cuuint32_t es[5] = {1, (cuuint32_t)estride, 1, 1, 1};
    CUresult r = enc(m, map_dtype(), 5, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        dfine_set_error("%s: cuTensorMapEncodeTiled(5d) failed (%d) C=%ld W=%ld H=%ld B=%ld ld=%ld", who, (int)r, C, W, H, B, ld);
        return -2;
    }
    return 0;
}
int make_map2(CUtensorMap* m, const float* base, long inner, long rows, long ld, int box_inner, int box_rows,
              const char* who) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { dfine_set_error("%s: cuTensorMapEncodeTiled unavailable", who); return -2; }
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, map_dtype(), 2, (void*)base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        dfine_set_error("%s: cuTensorMapEncodeTiled(2d) failed (%d) inner=%ld rows=%ld ld=%ld", who, (int)r, inner, rows,
                        ld);
        return -2;
    }
    return 0;
}

void pick_patch(int OH, int OW, int max_span, int* TW, int* TH) {
    // widest patch row <= 128 that wastes the fewest rows; prefer full-width rows for narrow maps
    if (OH == 1) { *TW = 128 < max_span ? 128 : max_span; *TH = 1; return; }
    int best_tw = 1, best_th = 1;
    double best = -1.0;
    for (int tw = 1; tw <= 128 && tw <= max_span; ++tw) {
        if (tw > OW) break;
        int th = 128 / tw;
        if (th > max_span) th = max_span;
        if (th < 1) continue;
        const long tiles = (long)((OW + tw - 1) / tw) * ((OH + th - 1) / th);
        const double eff = (double)OH * OW / ((double)tiles * 128.0);
        if (eff > best + 1e-9 || (eff > best - 1e-9 && tw > best_tw)) { best = eff; best_tw = tw; best_th = th; }
    }
    *TW = best_tw;
    *TH = best_th;
}

template <int BN, int X3, int STAGES>
int launch_fwd(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& mwlo, float* y, const float* bias,
               double* stats, const FwdParams& p, int B, cudaStream_t st) {
    const int smem = (int)sizeof(FwdSmem<BN, X3, STAGES>) + 1024;
    DFINE_SET_SMEM_ONCE((tc_fwd_kernel<BN, X3, STAGES>), smem, "tc_fwd");
    dim3 grid(B * p.tiles_w * p.tiles_h, ceil_div(p.N, BN));
    launch_k(tc_fwd_kernel<BN, X3, STAGES>, grid, fwd_threads(X3), smem, st, mx, mw, mwlo, y, bias, stats, p);
    return 0;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <int BN, int X3, int STAGES>
int launch_persist(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& mwlo, float* y, const float* bias,
                   double* stats, const FwdParams& p, int B, cudaStream_t st) {
    const int smem = (int)sizeof(PersistSmem<BN, X3, STAGES>) + 1024;
    DFINE_SET_SMEM_ONCE((tc_fwd_persist<BN, X3, STAGES>), smem, "tc_fwd_persist");
    TileSched ts;
    ts.n_tiles = ceil_div(p.N, BN);
    ts.total = B * p.tiles_w * p.tiles_h * ts.n_tiles;
    const int grid = ts.total < sm_count() ? ts.total : sm_count();
    launch_k(tc_fwd_persist<BN, X3, STAGES>, grid, fwd_threads(X3), smem, st, mx, mw, mwlo, y, bias, stats, p, ts);
    return 0;
}

// Few pixel tiles (the decoder's 8000-row linears, the 20x20 maps): a narrower N tile puts more CTAs on the 148 SMs and
// shortens every CTA's epilogue; the pixel tile is then re-read by the N tiles of its patch from L2.  DFINE_TC_FILL=0: off.
int fill_bn(int bn, int m_tiles, int Cout) {
    static const bool on = [] { const char* e = getenv("DFINE_TC_FILL"); return !(e && e[0] == '0'); }();
    while (on && bn > 64 && (long)m_tiles * ceil_div(Cout, bn) < 120) bn /= 2;
    return bn;
}

template <int BN, int STAGES>
int launch_pair(const CUtensorMap& mx, const CUtensorMap& mw, float* y, const float* bias, double* stats, const FwdParams& p,
                int B, cudaStream_t st) {
    const int smem = (int)sizeof(PairSmem<BN, STAGES>) + 1024;
    DFINE_SET_SMEM_ONCE((tc_pair_kernel<BN, STAGES>), smem, "tc_pair");
    TileSched ts;
    ts.n_tiles = ceil_div(p.N, BN);
    const int m_tiles = B * p.tiles_w * p.tiles_h;
    ts.total = ((m_tiles + 1) / 2) * ts.n_tiles;
    // as many CTA pairs as the device can keep resident at once (GPCs with an odd SM count leave an SM without a partner)
    static const int max_pairs = [&] {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(sm_count(), 1, 1);
        cfg.blockDim = dim3(PAIR_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, tc_pair_kernel<BN, STAGES>, &cfg) != cudaSuccess || n < 1) {
            cudaGetLastError();
            n = sm_count() / 2;
        }
        return n < sm_count() / 2 ? n : sm_count() / 2;
    }();
    int pairs = ts.total < max_pairs ? ts.total : max_pairs;
    if (pairs < 1) pairs = 1;
    launch_k(tc_pair_kernel<BN, STAGES>, 2 * pairs, PAIR_THREADS, smem, st, mx, mw, y, bias, stats, p, ts);
    return 0;
}

template <int BN, int STAGES>
int launch_ts(const CUtensorMap& mx, const CUtensorMap& mw, float* y, const float* bias, double* stats, const FwdParams& p, int B,
              cudaStream_t st) {
    static_assert(sizeof(TsSmem<BN, STAGES>) + 1024 <= 232448, "tc_fwd_ts: shared memory");
    const int smem = (int)sizeof(TsSmem<BN, STAGES>) + 1024;
    DFINE_SET_SMEM_ONCE((tc_fwd_ts<BN, STAGES>), smem, "tc_fwd_ts");
    TileSched ts;
    ts.n_tiles = ceil_div(p.N, BN);
    ts.total = B * p.tiles_w * p.tiles_h * ts.n_tiles;
    int grid = ts.total < sm_count() ? ts.total : sm_count();
    if (grid >= ts.n_tiles) grid -= grid % ts.n_tiles;      // every CTA then stays on ONE N tile (CTA-local BN statistics)
    launch_k(tc_fwd_ts<BN, STAGES>, grid, TS_THREADS, smem, st, mx, mw, y, bias, stats, p, ts);
    return 0;
}

template <int BN, int STAGES>
int launch_pair16(const CUtensorMap& mx, const CUtensorMap& mw, float* y, const float* bias, double* stats, const FwdParams& p,
                  int B, cudaStream_t st) {
    static_assert(sizeof(Pair16Smem<BN, STAGES>) + 1024 <= 232448, "tc_pair16: shared memory");
    const int smem = (int)sizeof(Pair16Smem<BN, STAGES>) + 1024;
    DFINE_SET_SMEM_ONCE((tc_pair16_kernel<BN, STAGES>), smem, "tc_pair16");
    TileSched ts;
    ts.n_tiles = ceil_div(p.N, BN);
    const int m_tiles = B * p.tiles_w * p.tiles_h;
    ts.total = ((m_tiles + 1) / 2) * ts.n_tiles;
    static const int max_pairs = [&] {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(sm_count(), 1, 1);
        cfg.blockDim = dim3(PAIR16_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, tc_pair16_kernel<BN, STAGES>, &cfg) != cudaSuccess || n < 1) {
            cudaGetLastError();
            n = sm_count() / 2;
        }
        return n < sm_count() / 2 ? n : sm_count() / 2;
    }();
    int pairs = ts.total < max_pairs ? ts.total : max_pairs;
    if (pairs < 1) pairs = 1;
    launch_k(tc_pair16_kernel<BN, STAGES>, 2 * pairs, PAIR16_THREADS, smem, st, mx, mw, y, bias, stats, p, ts);
    return 0;
}

template <int BN>
int launch_wgrad(const CUtensorMap& mdy, const CUtensorMap& mx, float* dwr, WgradParams p, cudaStream_t st) {
    const int smem = (int)sizeof(WgradSmem<BN>) + 1024;
    DFINE_SET_SMEM_ONCE((tc_wgrad_kernel<BN>), smem, "tc_wgrad");
    p.cin_tiles = ceil_div(p.Cin, BN);
    const int gx = ceil_div(p.Cout, BM), gy = p.taps_h * p.taps_w * p.cin_tiles;
    long splits = (148L * 2 + (long)gx * gy - 1) / ((long)gx * gy);
    if (splits < 1) splits = 1;
    long sps = (p.steps_total + splits - 1) / splits;
    if (sps < 8) sps = 8;
    p.steps_per_split = sps;
    dim3 grid(gx, gy, ceil_div(p.steps_total, sps));
    launch_k(tc_wgrad_kernel<BN>, grid, WG_THREADS, smem, st, mdy, mx, dwr, p);
    return 0;
}

}  // namespace

// 1 if the tensor-core path accepts this conv/linear geometry: channel counts and pixel strides multiples of
// 4 floats (16-byte TMA granularity), stride 1 or 2, at most 9 taps.  (The 3-channel image conv and 1-wide
// heads stay on the CUDA-core kernels.)
DFINE_API int dfine_conv_tc_supported(int Cin, int Cout, int KH, int KW, int stride, int pad_t, int pad_l, int pad_b,
                                      int pad_r, long ldx, long ldy) {
    (void)pad_t; (void)pad_l; (void)pad_b; (void)pad_r;
    if (stride != 1 && stride != 2) return 0;
    if (KH * KW > MAX_TAPS || KH < 1 || KW < 1) return 0;
    if (Cin % 4 || Cout % 4 || ldx % 4 || ldy % 4) return 0;
    if (Cin < 4 || Cout < 4) return 0;
    return 1;
}

// General tensor-core implicit GEMM (see FwdParams):
//   y[b, oh*osy+ooy, ow*osx+oox, :Cout] = act( sum_t sum_c x[b, oh*in_stride+dh[t], ow*in_stride+dw[t], c] * w[n, wk[t]+c] + bias )
// for (oh, ow) in [0,OH) x [0,OW).  x: [B,H,W,Cin] pixel stride ldx; w: [Cout, ldw] row-major (tap-major K);
// y: [B,YH,YW,*] pixel stride ldy.  taps = n_taps x (dh, dw, wk) host ints.  stats (optional, double [2*Cout],
// zeroed by the caller) receives per-channel sum / sum of squares of the raw accumulator (train-mode BatchNorm).
// w_lo == null: plain kind::tf32.  w_lo != null: 3xTF32 — `w` must then hold tf32-rounded weights and w_lo the
// remainders (dfine_tf32_split); activations are split inside the kernel.  nn.Linear on [rows, K]: B=1, H=1, W=rows.
// Bring-up / profiling: when `buf` is non-null every tc_fwd_ts launch (the 3xFP16 forward kernel) writes per-CTA phase
// timestamps (globaltimer, ns) to buf[cta * 16 + slot]; slots: 0 entry, 1 prologue done, 2 dependency wait done,
// 3 first TMA issued, 4 first tile landed (converter), 5 first planes converted (MMA issuer), 6 first accumulator ready,
// 7 first tile stored, 8 last tile stored, 9 all roles done, 10 statistics flushed, 11 finalize done.  buf must hold
// 16 * grid unsigned 64-bit words; null switches the trace off.  (tools/trace_conv.py)
static unsigned long long* g_tc_trace = nullptr;
DFINE_API int dfine_tc_trace(unsigned long long* buf) {
    g_tc_trace = buf;
    return 0;
}

namespace {
struct BnFinArgs {      // the fused train-mode BatchNorm finalize (FwdParams::fin_*)
    unsigned int* counter;
    const float *w, *b;
    float *rmean, *rvar, *mean, *invstd, *scale, *shift;
    float momentum, eps;
    float* y_out;             // optional fused apply pass (FwdParams::fin_y ...)
    long ld_out;
    const float* post;
    long ld_post;
    const float *lab_s, *lab_b;
    int act;
};
int conv_tc_impl(const float* x, const float* w, const float* w_lo, const void* w_bf16, const float* bias, float* y,
                 double* stats, int B, int H, int W, int Cin, long ldx, int OH, int OW, int Cout, long ldy,
                 int YH, int YW, int osy, int osx, int ooy, int oox, int in_stride, int n_taps,
                 const int* taps, long ldw, int act, void* stream, long ldw16 = 0, const float* res = nullptr,
                 long ldres = 0, int half16 = 0, float out_scale = 1.f, long plane_stride16 = 0, int lab = 0,
                 float lab_s = 1.f, float lab_b = 0.f, int w_rows = 0, int w_row_off = 0, int w_k_off = 0,
                 const float* ch_scale = nullptr, const BnFinArgs* fin = nullptr) {
    DFINE_REQUIRE(n_taps >= 1 && n_taps <= MAX_TAPS, "conv_tc: %d taps unsupported", n_taps);
    DFINE_REQUIRE(in_stride == 1 || in_stride == 2, "conv_tc: input stride %d unsupported", in_stride);
    DFINE_REQUIRE(Cin % 4 == 0 && Cout % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && ldw % ((w_bf16 && !w) ? 8 : 4) == 0 &&
                      Cin >= 4 && Cout >= 4,
                  "conv_tc: unsupported geometry Cin=%d Cout=%d ldx=%ld ldy=%ld ldw=%ld", Cin, Cout, ldx, ldy, ldw);
    DFINE_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)w % 16) == 0 && ((uintptr_t)y % 16) == 0 &&
                      ((uintptr_t)w_lo % 16) == 0 && ((uintptr_t)w_bf16 % 16) == 0 &&
                      (!bias || ((uintptr_t)bias % 16) == 0),
                  "conv_tc: pointers must be 16-byte aligned");
    if ((long)B * OH * OW == 0) return 0;
    FwdParams p;
    p.n_taps = n_taps;
    for (int t = 0; t < MAX_TAPS; ++t) {
        p.dh[t] = t < n_taps ? taps[3 * t] : 0;
        p.dw[t] = t < n_taps ? taps[3 * t + 1] : 0;
        p.wk[t] = t < n_taps ? taps[3 * t + 2] : 0;
    }
    static const int dbg = [] { const char* e = getenv("DFINE_TC_DBG"); return e ? atoi(e) : 0; }();
    p.dbg = dbg;
    static const int pf = [] { const char* e = getenv("DFINE_TC_PREFETCH"); return e ? atoi(e) : 0; }();   // (measured: 0-5 % slower with the L2 prefetch warp on)
    p.prefetch = pf; p.x = x; p.B = B; p.H = H; p.W = W; p.ldx = ldx;
    DFINE_REQUIRE(!res || (ldres % 4 == 0 && ldres >= Cout && ((uintptr_t)res % 16) == 0), "conv_tc: residual stride %ld", ldres);
    p.res = res; p.ldres = ldres;
    p.half16 = half16; p.out_scale = out_scale;
    p.lab = lab; p.lab_s = lab_s; p.lab_b = lab_b;
    p.w_row_off = w_row_off; p.w_k_off = w_k_off;
    p.ch_scale = ch_scale;
    p.trace = g_tc_trace;
    p.fin_counter = nullptr;
    if (fin) {
        DFINE_REQUIRE(stats && fin->counter && fin->mean && fin->invstd && fin->scale && fin->shift && half16 && w_bf16 && !w,
                      "conv_tc: the fused BatchNorm finalize needs statistics, its outputs and the 3xFP16 tensor-memory kernel");
        p.fin_counter = fin->counter; p.fin_w = fin->w; p.fin_b = fin->b; p.fin_rmean = fin->rmean; p.fin_rvar = fin->rvar;
        p.fin_mean = fin->mean; p.fin_invstd = fin->invstd; p.fin_scale = fin->scale; p.fin_shift = fin->shift;
        p.fin_M = (long)B * OH * OW; p.fin_momentum = fin->momentum; p.fin_eps = fin->eps;
        p.fin_y = fin->y_out; p.fin_ldy = fin->ld_out; p.fin_post = fin->post; p.fin_ldpost = fin->ld_post;
        p.fin_lab_s = fin->lab_s; p.fin_lab_b = fin->lab_b; p.fin_act = fin->act;
        if (fin->y_out) {
            DFINE_REQUIRE(fin->ld_out % 4 == 0 && fin->ld_out >= Cout && ((uintptr_t)fin->y_out % 16) == 0 &&
                              (!fin->post || (fin->ld_post % 4 == 0 && fin->ld_post >= Cout && ((uintptr_t)fin->post % 16) == 0)) &&
                              (fin->lab_s == nullptr) == (fin->lab_b == nullptr) && YH == OH && YW == OW && osy == 1 && osx == 1 &&
                              ooy == 0 && oox == 0,
                          "conv_tc: fused BatchNorm apply geometry (ld_out=%ld ld_post=%ld)", fin->ld_out, fin->ld_post);
        }
    }
    const int WR = w_rows > 0 ? w_rows : Cout;        // rows of the weight matrix the maps cover (stacked per-image weights)
    p.in_stride = in_stride; p.Cin = Cin;
    p.OH = OH; p.OW = OW; p.N = Cout; p.ldy = ldy; p.act = act;
    p.YH = YH; p.YW = YW; p.osy = osy; p.osx = osx; p.ooy = ooy; p.oox = oox;
    pick_patch(OH, OW, 256 / in_stride, &p.TW, &p.TH);
    p.tiles_w = ceil_div(OW, p.TW);
    p.tiles_h = ceil_div(OH, p.TH);
    CUtensorMap mx, mw, mwlo;
    // TFLOAT32 maps make the TMA unit round fp32 -> tf32 (nearest) while it copies: right for the plain path,
    // but the 3xTF32 converter needs the untouched fp32 activations to form a_lo = a - tf32(a)
    int rc = make_map4(&mx, x, Cin, W, H, B, ldx, BK, p.TW, p.TH, in_stride, "conv_tc(x)", CU_TENSOR_MAP_SWIZZLE_128B,
                       (w_lo || w_bf16) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : map_dtype());
    if (rc) return rc;
    if (w_bf16) {
        // 3xBF16 / hybrid: two adjacent bf16 weight planes [2][Cout][ld16] (dfine_bf16_split); one 3-D box
        // {32 k, bn rows, 2 planes} with the 64-byte swizzle lands as [plane][row][64 B]
        const bool hybrid = w != nullptr;
        const long ld16 = hybrid ? ldw16 : ldw;
        // 256-wide tiles (the A patch is then read and converted once for Cout = 256) on the two-plane 16-bit modes
        static const bool wide16 = [] { const char* e = getenv("DFINE_TC_WIDE16"); return !(e && e[0] == '0'); }();
        int bn = Cout <= 32 ? 32 : (Cout <= 64 ? 64 : ((Cout <= 128 || hybrid || !wide16) ? 128 : 256));
        bn = fill_bn(bn, B * p.tiles_w * p.tiles_h, Cout);
        // activation planes in tensor memory (tc_fwd_ts): N tiles of at most 128.  DFINE_TC_TS=0: planes in shared memory.
        static const bool use_ts = [] { const char* e = getenv("DFINE_TC_TS"); return !(e && e[0] == '0'); }();
        const bool ts_path = use_ts && !hybrid;
        DFINE_REQUIRE(ts_path || !ch_scale, "conv_tc: the per-channel epilogue scale needs the tensor-memory kernel (DFINE_TC_TS)");
        DFINE_REQUIRE(ts_path || !fin, "conv_tc: the fused BatchNorm finalize needs the tensor-memory kernel (DFINE_TC_TS)");
        if (ts_path && bn > 128) bn = 128;
        EncodeTiledFn enc = get_encode();
        if (!enc) { dfine_set_error("conv_tc: cuTensorMapEncodeTiled unavailable"); return -2; }
        if (hybrid) {
            // tf32 hi plane through a plain fp32 map; the taps' columns in the tap-padded bf16 planes
            DFINE_REQUIRE(ld16 % 8 == 0 && ld16 % n_taps == 0, "conv_tc(hybrid): ldw16=%ld taps=%d", ld16, n_taps);
            const int cin_p = (int)(ld16 / n_taps);
            for (int t = 0; t < MAX_TAPS; ++t) {
                DFINE_REQUIRE(p.wk[t] % Cin == 0, "conv_tc(hybrid): tap column %d is not a multiple of Cin=%d", p.wk[t], Cin);
                p.wk2[t] = p.wk[t] / Cin * cin_p;
            }
            rc = make_map2(&mwlo, w, ldw, Cout, ldw, BK, bn, "conv_tc(w_hi)");     // placeholder var; swapped below
            if (rc) return rc;
        }
        cuuint64_t dims[3] = {(cuuint64_t)ld16, (cuuint64_t)WR, 2};
        // (the two planes of a weight are adjacent, or sit in two arenas `plane_stride16` elements apart)
        const cuuint64_t pstride = plane_stride16 > 0 ? (cuuint64_t)plane_stride16 * 2 : (cuuint64_t)ld16 * 2 * WR;
        DFINE_REQUIRE(pstride % 16 == 0, "conv_tc: 16-bit plane stride %ld must be a multiple of 8 elements", plane_stride16);
        cuuint64_t strides[2] = {(cuuint64_t)ld16 * 2, pstride};
        // CTA-pair kernel (cta_group::2): every CTA fetches HALF of the N tile's rows.  Opt-in (DFINE_TC_PAIR16=1): worth
        // 0.2 ms of the 40 ms D-FINE-m step and not yet cleared by the parity suite in every mode.
        static const bool use_pair16 = [] { const char* e = getenv("DFINE_TC_PAIR16"); return e && e[0] == '1'; }();
        // (per-image weights: both CTAs of a pair must sit in the same image — an even tile count per image)
        const bool pair_grouped_ok = (w_row_off == 0 && w_k_off == 0) || (p.tiles_w * p.tiles_h) % 2 == 0;
        // (measured, profiles/r2_pair16_microbench.txt: +6..13 % on the 256-wide tiles, -8..13 % on 128-wide ones, where
        // waiting for the converter warps of BOTH CTAs costs more than the halved weight traffic saves)
        const bool pair16 = use_pair16 && !hybrid && bn == 256 && (long)B * p.tiles_w * p.tiles_h >= 2 && pair_grouped_ok;
        cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)(pair16 ? bn / 2 : bn), 2};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&mw, half16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                         const_cast<void*>(w_bf16), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            dfine_set_error("conv_tc: cuTensorMapEncodeTiled(bf16 planes) failed (%d) ldw=%ld Cout=%d", (int)r, ldw, Cout);
            return -2;
        }
        p.w_planes = 1;
        cudaStream_t st = (cudaStream_t)stream;
        if (ts_path) {
            rc = bn == 32  ? launch_ts<32, 8>(mx, mw, y, bias, stats, p, B, st)
               : bn == 64  ? launch_ts<64, 8>(mx, mw, y, bias, stats, p, B, st)
                           : launch_ts<128, 6>(mx, mw, y, bias, stats, p, B, st);
            if (rc) return rc;
            DFINE_LAUNCH_CHECK("conv_tc(16-bit planes, A in TMEM)");
            return 0;
        }
        if (pair16) {
            rc = launch_pair16<256, 4>(mx, mw, y, bias, stats, p, B, st);
            if (rc) return rc;
            DFINE_LAUNCH_CHECK("conv_tc(16-bit planes, pair)");
            return 0;
        }
        if (hybrid) {     // map_w = fp32 hi plane (encoded into mwlo above), map_wlo = the bf16 planes (in mw)
            rc = bn == 32 ? launch_persist<32, 3, 4>(mx, mwlo, mw, y, bias, stats, p, B, st)
               : bn == 64 ? launch_persist<64, 3, 4>(mx, mwlo, mw, y, bias, stats, p, B, st)
                          : launch_persist<128, 3, 3>(mx, mwlo, mw, y, bias, stats, p, B, st);
        } else {
            rc = bn == 32  ? launch_persist<32, 2, 4>(mx, mw, mw, y, bias, stats, p, B, st)
               : bn == 64  ? launch_persist<64, 2, 4>(mx, mw, mw, y, bias, stats, p, B, st)
               : bn == 128 ? launch_persist<128, 2, 4>(mx, mw, mw, y, bias, stats, p, B, st)
                           : launch_persist<256, 2, 3>(mx, mw, mw, y, bias, stats, p, B, st);
        }
        if (rc) return rc;
        DFINE_LAUNCH_CHECK("conv_tc(bf16 planes)");
        return 0;
    }
    DFINE_REQUIRE(!ch_scale, "conv_tc: the per-channel epilogue scale exists on the 16-bit-plane forward only");
    static const bool persist_bn = [] { const char* e = getenv("DFINE_TC_PERSIST"); return !(e && e[0] == '0'); }();
    // N tile: one tile covers Cout when it can (the A patch is then read exactly once); 256-wide tiles only on
    // the persistent plain-tf32 kernel (its smem ring has room for 48 KB stages)
    int bn = Cout <= 32 ? 32 : (Cout <= 64 ? 64 : ((Cout <= 128 || w_lo || !persist_bn) ? 128 : 256));
    bn = fill_bn(bn, B * p.tiles_w * p.tiles_h, Cout);
    rc = make_map2(&mw, w, ldw, WR, ldw, BK, bn, "conv_tc(w)");
    if (rc) return rc;
    mwlo = mw;
    p.w_planes = 0;
    if (w_lo) {
        rc = make_map2(&mwlo, w_lo, ldw, Cout, ldw, BK, bn, "conv_tc(w_lo)");
        if (rc) return rc;
        const long plane_dist = (long)(w_lo - w);              // floats between the hi and the lo plane
        if (persist_bn && plane_dist > 0 && plane_dist % 4 == 0 && plane_dist * 4 < (1L << 40)) {
            // the planes sit at a fixed distance (adjacent per weight, or the optimizer's hi / lo arenas): both are
            // fetched by ONE 3-D box whose outer dimension steps from the hi to the lo plane
            EncodeTiledFn enc = get_encode();
            cuuint64_t dims[3] = {(cuuint64_t)ldw, (cuuint64_t)Cout, 2};
            cuuint64_t strides[2] = {(cuuint64_t)ldw * 4, (cuuint64_t)plane_dist * 4};
            cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)bn, 2};
            cuuint32_t es[3] = {1, 1, 1};
            CUtensorMap m3;
            if (enc && enc(&m3, map_dtype(), 3, (void*)w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
                mw = m3;
                p.w_planes = 1;
            }
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
    static const bool persist = [] { const char* e = getenv("DFINE_TC_PERSIST"); return !(e && e[0] == '0'); }();
    DFINE_REQUIRE((persist && !w_lo && !w_bf16) || !res, "conv_tc: the fused residual add needs the persistent plain-tf32 kernel");
    // CTA-pair kernel (cta_group::2) for the plain-tf32 launches with at least a few pixel-tile pairs per SM pair and no
    // fused residual: the weight map then delivers HALF an N tile per CTA.  DFINE_TC_PAIR=0 keeps the single-CTA kernel.
    static const bool use_pair = [] { const char* e = getenv("DFINE_TC_PAIR"); return !(e && e[0] == '0'); }();
    const bool pair_grouped_ok = (w_row_off == 0 && w_k_off == 0) || (p.tiles_w * p.tiles_h) % 2 == 0;
    if (persist && use_pair && !w_lo && !res && bn >= 64 && (long)B * p.tiles_w * p.tiles_h >= 2 && pair_grouped_ok) {
        CUtensorMap mwh;
        rc = make_map2(&mwh, w, ldw, WR, ldw, BK, bn / 2, "conv_tc(w half)");
        if (rc) return rc;
        rc = bn == 64  ? launch_pair<64, 8>(mx, mwh, y, bias, stats, p, B, st)
           : bn == 128 ? launch_pair<128, 8>(mx, mwh, y, bias, stats, p, B, st)
                       : launch_pair<256, 6>(mx, mwh, y, bias, stats, p, B, st);
        if (rc) return rc;
        DFINE_LAUNCH_CHECK("conv_tc(pair)");
        return 0;
    }
    if (persist) {
        if (w_lo) {
            rc = bn == 32 ? launch_persist<32, 1, 4>(mx, mw, mwlo, y, bias, stats, p, B, st)
               : bn == 64 ? launch_persist<64, 1, 4>(mx, mw, mwlo, y, bias, stats, p, B, st)
                          : launch_persist<128, 1, 3>(mx, mw, mwlo, y, bias, stats, p, B, st);
        } else {
            rc = bn == 32  ? launch_persist<32, 0, 8>(mx, mw, mwlo, y, bias, stats, p, B, st)
               : bn == 64  ? launch_persist<64, 0, 8>(mx, mw, mwlo, y, bias, stats, p, B, st)
               : bn == 128 ? launch_persist<128, 0, 6>(mx, mw, mwlo, y, bias, stats, p, B, st)
                           : launch_persist<256, 0, 4>(mx, mw, mwlo, y, bias, stats, p, B, st);
        }
    } else if (w_lo) {
        rc = bn == 32 ? launch_fwd<32, 1, 2>(mx, mw, mwlo, y, bias, stats, p, B, st)
           : bn == 64 ? launch_fwd<64, 1, 2>(mx, mw, mwlo, y, bias, stats, p, B, st)
                      : launch_fwd<128, 1, 3>(mx, mw, mwlo, y, bias, stats, p, B, st);
    } else {
        rc = bn == 32 ? launch_fwd<32, 0, 4>(mx, mw, mwlo, y, bias, stats, p, B, st)
           : bn == 64 ? launch_fwd<64, 0, 4>(mx, mw, mwlo, y, bias, stats, p, B, st)
                      : launch_fwd<128, 0, 3>(mx, mw, mwlo, y, bias, stats, p, B, st);
    }
    if (rc) return rc;
    DFINE_LAUNCH_CHECK("conv_tc");
    return 0;
}
}  // namespace

// Per-image weights ("grouped" launch, the segmentation head's mask product): `w` stacks the images' matrices — w_rows
// total rows; image i uses rows [i*w_row_off, i*w_row_off + Cout) and columns [i*w_k_off, ...) — 0 / 0 / 0 = one shared weight.
// `res` (optional, pixel stride ldres, same pixel geometry as y): y = act(conv + bias) + res — the data-gradient
// launches use it to fold the gradient-accumulation add of a tensor with two consumers into the epilogue.
DFINE_API int dfine_conv_tc(const float* x, const float* w, const float* w_lo, const float* bias, float* y,
                            double* stats, int B, int H, int W, int Cin, long ldx, int OH, int OW, int Cout, long ldy,
                            int YH, int YW, int osy, int osx, int ooy, int oox, int in_stride, int n_taps,
                            const int* taps, long ldw, int act, const float* res, long ldres, int w_rows,
                            int w_row_off, int w_k_off, void* stream) {
    DFINE_REQUIRE(w_lo == nullptr || (w_row_off == 0 && w_k_off == 0), "conv_tc: per-image weights are not wired for 3xTF32");
    return conv_tc_impl(x, w, w_lo, nullptr, bias, y, stats, B, H, W, Cin, ldx, OH, OW, Cout, ldy, YH, YW, osy, osx, ooy,
                        oox, in_stride, n_taps, taps, ldw, act, stream, 0, res, ldres, 0, 1.f, 0, 0, 1.f, 0.f, w_rows,
                        w_row_off, w_k_off);
}

// The same contract with error-compensated 3xBF16 operands (see PersistSmem): `w_planes` = the two bf16 planes
// [2][Cout][ldw] written by dfine_bf16_split (each tap's channels padded to a multiple of 8, so ldw % 8 == 0 and
// the taps' weight offsets wk[] are multiples of 8); activations are split in the kernel.
DFINE_API int dfine_conv_tc_bf16x3(const float* x, const void* w_planes, const float* bias, float* y, double* stats,
                                   int B, int H, int W, int Cin, long ldx, int OH, int OW, int Cout, long ldy, int YH,
                                   int YW, int osy, int osx, int ooy, int oox, int in_stride, int n_taps,
                                   const int* taps, long ldw, int act, void* stream) {
    DFINE_REQUIRE(w_planes != nullptr, "conv_tc_bf16x3: null weight planes");
    return conv_tc_impl(x, nullptr, nullptr, w_planes, bias, y, stats, B, H, W, Cin, ldx, OH, OW, Cout, ldy, YH, YW, osy,
                        osx, ooy, oox, in_stride, n_taps, taps, ldw, act, stream);
}

// Error-compensated 3xFP16: the 3xBF16 contract with fp16 planes — 11 significand bits per part (the very precision of
// tf32), so a = a_hi + a_lo carries 22 bits like the 3xTF32 split while the three MMAs run on kind::f16 at twice the
// tf32 rate and both planes of an operand take the bytes of one fp32 plane.  fp16's narrow exponent is handled by
// scale: `w_planes` = dfine_f16_split(w * w_scale) with w_scale a power of two that lifts the small weights' lo parts
// out of the subnormal range; `out_scale` = 1 / w_scale is applied to the accumulator first thing in the epilogue
// (exact).  lab != 0: the deploy-mode epilogue y = lab_scale * act(acc + bias) + lab_bias.  `plane_stride` = elements between the hi and the lo plane (0: adjacent, [2][Cout][ldw]).  Activations are split unscaled inside the kernel (post-normalisation values are O(1); |a| must stay below
// 65504, lo parts below 2^-14 lose relative — not absolute (2^-25) — precision).  ch_scale (optional, [Cout]): per-channel
// scale applied before the bias — a frozen / eval-mode BatchNorm folded into the epilogue, y = act(acc * scale + bias).
DFINE_API int dfine_conv_tc_f16x3(const float* x, const void* w_planes, const float* bias, float* y, double* stats,
                                  int B, int H, int W, int Cin, long ldx, int OH, int OW, int Cout, long ldy, int YH,
                                  int YW, int osy, int osx, int ooy, int oox, int in_stride, int n_taps,
                                  const int* taps, long ldw, int act, float out_scale, long plane_stride,
                                  int lab, float lab_scale, float lab_bias, int w_rows, int w_row_off,
                                  const float* ch_scale, void* stream) {
    DFINE_REQUIRE(w_planes != nullptr, "conv_tc_f16x3: null weight planes");
    return conv_tc_impl(x, nullptr, nullptr, w_planes, bias, y, stats, B, H, W, Cin, ldx, OH, OW, Cout, ldy, YH, YW, osy,
                        osx, ooy, oox, in_stride, n_taps, taps, ldw, act, stream, 0, nullptr, 0, 1, out_scale, plane_stride,
                        lab, lab_scale, lab_bias, w_rows, w_row_off, 0, ch_scale);
}

// dfine_conv_tc_f16x3 for a conv followed by a TRAIN-mode BatchNorm: besides the per-channel statistics (`stats`,
// double [2*Cout], zeroed by the caller) the kernel's last CTA runs dfine_bn_finalize's arithmetic in its tail — mean /
// invstd / scale / shift [Cout] and the running-statistics update (hgnetv2.py:65, torch BatchNorm2d momentum rule) —
// so the separate finalize launch disappears.  `counter`: one zeroed unsigned int (the retirement ticket).
// y_out != null: the kernel also runs dfine_bn_apply's pass on the tiles each CTA wrote (after a grid-wide wait for the
// finalize; `counter` then points at TWO zeroed unsigned ints, ticket and flag):
// y_out = lab_scale * act(y * scale[c] + shift[c]) + lab_bias (+ post_add), pixel strides ld_out / ld_post; `y` keeps the
// raw conv output the backward pass needs.  lab_scale / lab_bias: device scalars or null.
DFINE_API int dfine_conv_tc_f16x3_bn(const float* x, const void* w_planes, float* y, double* stats, unsigned int* counter,
                                     const float* bn_weight, const float* bn_bias, float* running_mean,
                                     float* running_var, float* mean, float* invstd, float* scale, float* shift,
                                     float momentum, float eps, int B, int H, int W, int Cin, long ldx, int OH, int OW,
                                     int Cout, long ldy, int in_stride, int n_taps, const int* taps, long ldw,
                                     float out_scale, long plane_stride, float* y_out, long ld_out,
                                     const float* post_add, long ld_post, const float* lab_scale, const float* lab_bias,
                                     int act, void* stream) {
    DFINE_REQUIRE(w_planes != nullptr, "conv_tc_f16x3_bn: null weight planes");
    BnFinArgs fin{counter, bn_weight, bn_bias, running_mean, running_var, mean, invstd, scale, shift, momentum, eps,
                  y_out, ld_out, post_add, ld_post, lab_scale, lab_bias, act};
    return conv_tc_impl(x, nullptr, nullptr, w_planes, nullptr, y, stats, B, H, W, Cin, ldx, OH, OW, Cout, ldy, OH, OW, 1, 1,
                        0, 0, in_stride, n_taps, taps, ldw, 0, stream, 0, nullptr, 0, 1, out_scale, plane_stride, 0, 1.f,
                        0.f, 0, 0, 0, nullptr, &fin);
}

// Hybrid operands (see PersistSmem, X3 = 3): a_hi*w_hi on kind::tf32, the cross terms on bf16 copies.  `w_hi` = the
// tf32-rounded fp32 weight matrix [Cout][ldw] (dfine_tf32_split's hi plane), `w_planes16` = bf16 planes
// [2][Cout][ldw16] = [bf16(w) | bf16(w - tf32(w))] written by dfine_bf16_split(mode 1), every tap's channel run
// padded to a multiple of 8 (ldw16 = n_taps * Cin_p).  Taps must address whole channel runs (wk[t] % Cin == 0).
DFINE_API int dfine_conv_tc_hybrid(const float* x, const float* w_hi, const void* w_planes16, const float* bias, float* y,
                                   double* stats, int B, int H, int W, int Cin, long ldx, int OH, int OW, int Cout,
                                   long ldy, int YH, int YW, int osy, int osx, int ooy, int oox, int in_stride,
                                   int n_taps, const int* taps, long ldw, long ldw16, int act, void* stream) {
    DFINE_REQUIRE(w_hi != nullptr && w_planes16 != nullptr, "conv_tc_hybrid: null weights");
    return conv_tc_impl(x, w_hi, nullptr, w_planes16, bias, y, stats, B, H, W, Cin, ldx, OH, OW, Cout, ldy, YH, YW, osy,
                        osx, ooy, oox, in_stride, n_taps, taps, ldw, act, stream, ldw16);
}

// The weight half of the 3xBF16 split.  w: [rows][taps * Cin] fp32 (row stride ldw); planes: [2][rows][taps * Cin_p]
// bf16 with every tap's channel run padded to Cin_p = Cin rounded up to 8 (a TMA box must start on a 16-byte
// boundary: tap * Cin_p * 2 bytes); planes[0] = RN bf16(w), planes[1] = RN bf16(w - planes[0]), pads zero.
namespace {
__global__ void bf16_split_kernel(const float* __restrict__ w, long ldw, unsigned short* __restrict__ planes,
                                  long rows, int taps, int Cin, int Cin_p, int mode) {
    pdl_entry();
    const long ldp = (long)taps * Cin_p, n = rows * ldp;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const long r = i / ldp;
        const int k = (int)(i % ldp), tap = k / Cin_p, c = k % Cin_p;
        uint32_t hi = 0, lo = 0;
        if (c < Cin) {
            const float v = w[r * ldw + (long)tap * Cin + c];
            bf16_split(v, hi, lo);
            if (mode == 1) lo = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v - tf32_rna(v)));
        }
        planes[i] = (unsigned short)hi;
        planes[n + i] = (unsigned short)lo;
    }
}
}  // namespace
namespace {
__global__ void f16_split_kernel(const float* __restrict__ w, long ldw, unsigned short* __restrict__ planes, long rows,
                                 int taps, int Cin, int Cin_p, float scale) {
    pdl_entry();
    const long ldp = (long)taps * Cin_p, n = rows * ldp;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const long r = i / ldp;
        const int k = (int)(i % ldp), tap = k / Cin_p, c = k % Cin_p;
        unsigned short hi = 0, lo = 0;
        if (c < Cin) {
            const float v = w[r * ldw + (long)tap * Cin + c] * scale;
            const __half h = __float2half_rn(v);
            hi = __half_as_ushort(h);
            lo = __half_as_ushort(__float2half_rn(v - __half2float(h)));
        }
        planes[i] = hi;
        planes[n + i] = lo;
    }
}
}  // namespace
// The weight half of the 3xFP16 split: planes [2][rows][taps * Cin_p] fp16 of w * scale (scale a power of two),
// planes[0] = RN f16(w * scale), planes[1] = RN f16(w * scale - planes[0]); tap runs padded to Cin_p like dfine_bf16_split.
DFINE_API int dfine_f16_split(const float* w, long ldw, void* planes, long rows, int taps, int Cin, int Cin_p,
                              float scale, void* stream) {
    DFINE_REQUIRE(Cin_p >= Cin && Cin_p % 8 == 0 && taps >= 1 && ((uintptr_t)planes % 16) == 0 && scale > 0.f,
                  "f16_split: Cin=%d Cin_p=%d taps=%d scale=%g", Cin, Cin_p, taps, (double)scale);
    const long n = rows * taps * Cin_p;
    if (n == 0) return 0;
    long g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    launch_k(f16_split_kernel, (int)g, 256, 0, (cudaStream_t)stream, w, ldw, (unsigned short*)planes, rows, taps, Cin, Cin_p, scale);
    DFINE_LAUNCH_CHECK("f16_split");
    return 0;
}

// mode 0: planes[1] = RN bf16(w - planes[0]) (3xBF16); mode 1: planes[1] = RN bf16(w - RN tf32(w)) (hybrid).
DFINE_API int dfine_bf16_split(const float* w, long ldw, void* planes, long rows, int taps, int Cin, int Cin_p,
                               int mode, void* stream) {
    DFINE_REQUIRE(Cin_p >= Cin && Cin_p % 8 == 0 && taps >= 1 && ((uintptr_t)planes % 16) == 0,
                  "bf16_split: Cin=%d Cin_p=%d taps=%d", Cin, Cin_p, taps);
    const long n = rows * taps * Cin_p;
    if (n == 0) return 0;
    long g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    launch_k(bf16_split_kernel, (int)g, 256, 0, (cudaStream_t)stream, w, ldw, (unsigned short*)planes, rows, taps, Cin, Cin_p,
                                                               mode);
    DFINE_LAUNCH_CHECK("bf16_split");
    return 0;
}

// hi = round-to-nearest tf32(w), lo = w - hi (exact): the weight half of the 3xTF32 split, run once per
// weight version over a flat arena.
namespace {
__global__ void tf32_split_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, long n) {
    pdl_entry();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float v = w[i], h = tf32_rna(v);
        hi[i] = h;
        lo[i] = v - h;
    }
}
}  // namespace
DFINE_API int dfine_tf32_split(const float* w, float* hi, float* lo, long n, void* stream) {
    if (n == 0) return 0;
    long g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    launch_k(tf32_split_kernel, (int)g, 256, 0, (cudaStream_t)stream, w, hi, lo, n);
    DFINE_LAUNCH_CHECK("tf32_split");
    return 0;
}

// dwr[Cout, KH*KW*Cin] += sum over output pixels dy[b,oh,ow,co] * x[b,oh*stride+kh-pad_t,ow*stride+kw-pad_l,ci];
// dwr zeroed (or holding a running gradient) on entry.  x: [B,H,W,Cin] stride ldx, dy: [B,OH,OW,Cout] stride ldy.
DFINE_API int dfine_conv_wgrad_tc(const float* dy, const float* x, float* dwr, int B, int H, int W, int Cin, int OH,
                                  int OW, int Cout, int KH, int KW, int stride, int pad_t, int pad_l, long ldx,
                                  long ldy, void* stream) {
    DFINE_REQUIRE(dfine_conv_tc_supported(Cin, Cout, KH, KW, stride, pad_t, pad_l, 0, 0, ldx, ldy),
                  "conv_wgrad_tc: unsupported geometry Cin=%d Cout=%d k=%dx%d stride=%d", Cin, Cout, KH, KW, stride);
    DFINE_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)dy % 16) == 0, "conv_wgrad_tc: alignment");
    if ((long)B * OH * OW == 0) return 0;
    WgradParams p;
    p.taps_h = KH; p.taps_w = KW; p.pad_t = pad_t; p.pad_l = pad_l; p.stride = stride;
    p.Cin = Cin; p.Cout = Cout; p.OH = OH; p.OW = OW; p.B = B;
    p.wchunks = ceil_div(OW, 32);
    p.steps_total = (long)B * OH * p.wchunks;
    CUtensorMap mdy, mx;
    static const bool box5 = [] { const char* e = getenv("DFINE_WGRAD_BOX5"); return !(e && e[0] == '0'); }();
    const int bn = Cin <= 32 ? 32 : (Cin <= 64 ? 64 : 128);
    p.dy5 = box5 && Cout % 32 == 0;
    p.x5 = box5 && Cin % 32 == 0;
    static bool box5_ok = true;    // cleared if the driver rejects the 5-D encoding (then one box per 32-channel chunk)
    int rc = 0;
    if (p.dy5 && box5_ok && make_map5(&mdy, dy, Cout, OW, OH, B, ldy, 32, BM / 32, 1, "conv_wgrad_tc(dy5)") != 0) box5_ok = false;
    if (!box5_ok) p.dy5 = 0;
    if (!p.dy5) {
        rc = make_map4(&mdy, dy, Cout, OW, OH, B, ldy, 32, 32, 1, 1, "conv_wgrad_tc(dy)", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc) return rc;
    }
    if (p.x5 && box5_ok && make_map5(&mx, x, Cin, W, H, B, ldx, 32, bn / 32, stride, "conv_wgrad_tc(x5)") != 0) box5_ok = false;
    if (!box5_ok) p.x5 = 0;
    if (!p.x5) {
        rc = make_map4(&mx, x, Cin, W, H, B, ldx, 32, 32, 1, stride, "conv_wgrad_tc(x)", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc) return rc;
    }
    if (!box5_ok && p.dy5) {       // the x encoding failed after dy was encoded 5-D: re-encode dy per chunk
        p.dy5 = 0;
        rc = make_map4(&mdy, dy, Cout, OW, OH, B, ldy, 32, 32, 1, 1, "conv_wgrad_tc(dy)", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    rc = Cin <= 32 ? launch_wgrad<32>(mdy, mx, dwr, p, st)
       : Cin <= 64 ? launch_wgrad<64>(mdy, mx, dwr, p, st)
                   : launch_wgrad<128>(mdy, mx, dwr, p, st);
    if (rc) return rc;
    DFINE_LAUNCH_CHECK("conv_wgrad_tc");
    return 0;
}
