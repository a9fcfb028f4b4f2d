// sf_common.cuh -- shared device helpers for the per-pixel head-to-metric kernels (sm_100a).
//
// Numeric contract (pinned by oracle/segpath_oracle.c against the reference):
//   bilinear taps   : scale = (float)in/out; r = fmaf(d+0.5, scale, -0.5); r = max(r,0);
//                     i0 = min(floor(r), in-1); l1 = r - i0; l0 = 1 - l1; i1 = i0 + (i0 < in-1)
//                     in == out: i0 = i1 = d, l0 = 1, l1 = 0;  bf16/fp16: l1, l0 rounded to the dtype
//   interpolation   : t_k = fmaf(v[k][j0], lw0, rn(v[k][j1]*lw1)); z = fmaf(t_i0, lh0, rn(t_i1*lh1)),
//                     rounded to the logits' dtype.  Every step uses explicit *_rn intrinsics so nvcc's
//                     -fmad contraction cannot change the association (argmax must be bit-exact).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/segfactory.h"

namespace sf {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2   = 0.6931471805599453f;
constexpr int   kMaxCtas = 1024;          // upper bound on persistent grid sizes (partials slots)
constexpr int   kPartialStride = 8;       // doubles per CTA partial record
constexpr int   kRadixBins = 2048;

// ---------------------------------------------------------------------------------------------
// Workspace (caller-allocated device memory, see sf_workspace_bytes)
// ---------------------------------------------------------------------------------------------
struct Workspace {
    sf_loss_state state;                               // must stay first (Python views it)
    unsigned int  n_partials;                          // grid size of the last sf_ce_forward
    unsigned int  pad0[15];
    double        partials[kMaxCtas][kPartialStride];  // {sum_loss, sum_w, sum_hard, n_valid, n_hard}
    double        sel_partials[kMaxCtas];              // pass-D partial sums of the radix select
    unsigned int  radix_hist[3][kRadixBins];
    // Order-free totals of the one-pass kernel (sf_onepass.cu): its grid is one CTA per tile (thousands), so instead of
    // per-CTA slots every CTA adds its partial sums as 128-bit fixed point (2^-40 units, two's complement, {lo, hi}) with
    // integer atomics -- integer addition is associative, so the loss stays bitwise reproducible.  n_partials == kFxMarker
    // tells sf_loss_finalize to read these instead of `partials`.
    unsigned long long fx_loss[2], fx_w[2], fx_hard[2];
    unsigned long long fx_n_valid, fx_n_hard;
    unsigned int  fx_flags;                            // non-finite partials: 3 bits (nan, +inf, -inf) per sum
    unsigned int  pad1[3];
};
constexpr unsigned int kFxMarker = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------
// Logit element traits: load-as-float, round-to-dtype, store
// ---------------------------------------------------------------------------------------------
template <typename T> struct Logit;
template <> struct Logit<float> {
    static constexpr int code = SF_F32;
    static __device__ __forceinline__ float ld(const float* p) { return *p; }
    static __device__ __forceinline__ float rnd(float v) { return v; }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct Logit<__nv_bfloat16> {
    static constexpr int code = SF_BF16;
    static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ float rnd(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};
template <> struct Logit<__half> {
    static constexpr int code = SF_F16;
    static __device__ __forceinline__ float ld(const __half* p) { return __half2float(*p); }
    static __device__ __forceinline__ float rnd(float v) { return __half2float(__float2half_rn(v)); }
    static __device__ __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
};

// ---------------------------------------------------------------------------------------------
// Bilinear taps (host + device: the host uses them to size TMA boxes and tiles)
// ---------------------------------------------------------------------------------------------
struct Axis {          // one spatial dimension of the resize
    int   in, out;
    float scale;       // (float)in / out
    int   ident;       // in == out
};

__host__ __device__ __forceinline__ float sf_fma(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}

// unclamped floor of the source coordinate: -1 for the rows/cols that clamp to r = 0
__host__ __device__ __forceinline__ int src_floor(const Axis& ax, int d) {
    if (ax.ident) return d;
    const float r = sf_fma((float)d + 0.5f, ax.scale, -0.5f);
#ifdef __CUDA_ARCH__
    int u = __float2int_rd(r);
#else
    int u = (int)floorf(r);
#endif
    return u < -1 ? -1 : (u > ax.in - 1 ? ax.in - 1 : u);
}
__host__ __device__ __forceinline__ int src_i0(const Axis& ax, int d) {
    const int u = src_floor(ax, d);
    return u < 0 ? 0 : u;
}
// smallest d in [0, out] with src_floor(d) >= u   (src_floor is monotone non-decreasing in d)
__host__ __device__ inline int first_dst_of(const Axis& ax, int u) {
    if (u <= -1) return 0;
    if (u >= ax.in) return ax.out;
    if (ax.ident) return u;
    float g = ((float)u + 0.5f) / ax.scale - 0.5f;
    int d = (int)g;
    if (d < 0) d = 0;
    if (d > ax.out) d = ax.out;
    while (d > 0 && src_floor(ax, d - 1) >= u) --d;
    while (d < ax.out && src_floor(ax, d) < u) ++d;
    return d;
}

struct Tap { int i0, i1; float l0, l1; };

template <typename LT>
__device__ __forceinline__ Tap make_tap(const Axis& ax, int d) {
    Tap t;
    if (ax.ident) { t.i0 = t.i1 = d; t.l0 = 1.0f; t.l1 = 0.0f; return t; }
    float r = __fmaf_rn((float)d + 0.5f, ax.scale, -0.5f);
    r = fmaxf(r, 0.0f);
    int i0 = __float2int_rd(r);
    i0 = min(i0, ax.in - 1);
    float l1 = __fsub_rn(r, (float)i0);
    l1 = fminf(fmaxf(l1, 0.0f), 1.0f);
    l1 = Logit<LT>::rnd(l1);
    t.l0 = Logit<LT>::rnd(__fsub_rn(1.0f, l1));
    t.l1 = l1;
    t.i0 = i0;
    t.i1 = i0 + (i0 < ax.in - 1 ? 1 : 0);
    return t;
}

// one horizontal lerp / one vertical lerp with the pinned association
__device__ __forceinline__ float lerp_pinned(float a, float la, float b, float lb) {
    return __fmaf_rn(a, la, __fmul_rn(b, lb));
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA (cp.async.bulk.tensor), fast exp2
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a TMA that never lands (bad descriptor) traps instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// log2 of a NORMAL positive float (softmax sums are >= 1): the bare MUFU.LG2, without __log2f's denormal pre/post-scaling
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// shared-memory `*p += v` when `on`, as ONE predicated instruction (no branch / reconvergence point in the row loop)
__device__ __forceinline__ void red_shared_add_if(uint32_t saddr, int v, bool on) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q red.shared.add.u32 [%0], %1;\n\t}"
                 :: "r"(saddr), "r"(v), "r"((int)on) : "memory");
}

// NaN-propagating maximum of three (FMNMX3.NAN): unlike fmaxf, a NaN input makes the result NaN
__device__ __forceinline__ float max3_nan(float a, float b, float c) {
    float d;
    asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// FocalLoss (util/losses.py:17-25): f = a (1-pt)^g ce with pt = exp(-ce).  Returns f; *dfdce = a[(1-pt)^g + g ce pt (1-pt)^(g-1)].
__device__ __forceinline__ float focal_value(float ce, float alpha, float gamma, float* dfdce) {
    const float pt = expf(-ce);
    const float om = -expm1f(-ce);                                  // 1 - pt without the cancellation at small ce
    if (gamma == 0.0f) { *dfdce = alpha; return alpha * ce; }
    if (!(om > 0.0f)) { *dfdce = 0.0f; return 0.0f; }               // ce == 0 (also every ignored pixel)
    const float w = powf(om, gamma);
    *dfdce = alpha * (w + gamma * ce * pt * (w / om));
    return alpha * w * ce;
}

__device__ __forceinline__ long long ld_label(const void* labels, int is_u8, long long idx) {
    return is_u8 ? (long long)__ldg(reinterpret_cast<const unsigned char*>(labels) + idx)
                 : __ldg(reinterpret_cast<const long long*>(labels) + idx);
}

// Low 32 bits of a label (two's complement).  The integer-scale kernels keep their prefetched labels in one register each:
// every value in [-2^31, 2^31) -- classes, 255, -100, -1 -- keeps its meaning; labels beyond int32 are outside the contract
// (the reference would index its bincount out of range with them).
__device__ __forceinline__ int ld_label_lo32(const void* labels, int is_u8, long long idx) {
    return is_u8 ? (int)__ldg(reinterpret_cast<const unsigned char*>(labels) + idx)
                 : __ldg(reinterpret_cast<const int*>(labels) + 2 * idx);
}

// 128-bit fixed-point accumulation of a double (see Workspace::fx_*).  |v| < 2^87; quantum 2^-40 per call.
__device__ __forceinline__ void fx_add(unsigned long long* acc, unsigned int* flags, int which, double v) {
    if (!isfinite(v)) { atomicOr(flags, (isnan(v) ? 1u : v > 0 ? 2u : 4u) << (3 * which)); return; }
    const double s = v * 0x1p40;
    const double h = floor(s * 0x1p-64);
    const double l = s - h * 0x1p64;                                // exact, in [0, 2^64)
    const unsigned long long lo = __double2ull_rz(l);
    long long hi = __double2ll_rz(h);
    const unsigned long long old = atomicAdd(&acc[0], lo);
    if (old + lo < old) ++hi;                                       // carry out of the low word
    if (hi) atomicAdd(&acc[1], (unsigned long long)hi);
}
__device__ __forceinline__ double fx_get(const unsigned long long* acc, unsigned int flags, int which) {
    const unsigned int f = (flags >> (3 * which)) & 7u;
    if (f) return (f & 1u) || ((f & 2u) && (f & 4u)) ? __longlong_as_double(0x7ff8000000000000LL)
                                                     : (f & 2u) ? __longlong_as_double(0x7ff0000000000000LL)
                                                                : __longlong_as_double(0xfff0000000000000LL);
    return (double)(long long)__ldcg(&acc[1]) * 0x1p24 + (double)__ldcg(&acc[0]) * 0x1p-40;
}

// deterministic block-wide sum of doubles (fixed shuffle tree + fixed warp order); result valid on thread 0
template <int kThreads>
__device__ __forceinline__ double block_sum(double v, double* scratch /* >= kThreads/32 doubles */) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < kThreads / 32; ++i) r += scratch[i];
    return r;
}

}  // namespace sf
