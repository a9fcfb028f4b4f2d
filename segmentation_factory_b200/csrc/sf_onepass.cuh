// sf_onepass.cuh -- ONE walk over the label pixels for forward AND backward (integer-scale heads, C <= 24).
//
// Replaces, in a single launch, models/build_models.py:65 (upsample), util/losses.py:32-36 / 52-57 (per-pixel CE, the
// OHEM `loss > thresh` test), util/metrics.py:25-27 (argmax + bincount) and autograd's nll_loss2d_backward +
// _log_softmax_backward_data + upsample_bilinear2d_backward.
//
// Why one pass is legal: `keep = loss > thresh` (OHEM threshold branch) and `keep = label != ignore` (plain CE) are
// pixel-local; only the 1/K (or 1/sum w) scale is global.  So the walk that computes max -> sum exp -> loss / argmax
// also emits (softmax - onehot) * keep * w_y into the low-res gradient, reusing the C exponentials, and the global
// scale is applied afterwards (sf_ce_onepass_backward: grad = acc * grad_out / K).  Versus the fwd + bwd kernel pair
// every pixel is interpolated once instead of twice, exponentiated once, its label read once, and the lse_px
// round trip (4 B/px written + read) is gone; loss_px is still written under OHEM because the top-k fallback
// (util/losses.py:58-59) needs the loss vector.  If sf_loss_finalize picks that fallback (branch 2) the speculative
// gradient is wrong and sf_ce_onepass_backward re-walks in OP_BWD mode: keep comes from loss_px against the k-th value
// and the softmax is rebuilt only for rows in which some lane of the warp keeps its pixel.
//
// Schedule = the gather-form tile of sf_backward_fast.cu: tile = TH x (128/S - 1) low-res cells x all C channels,
// 128 threads = the 128 label columns that interpolate from it, walking the S*TH + S dependant rows; vertical part of
// the transposed bilinear in registers, horizontal part through shared memory; every gradient element is written by
// exactly one thread (no atomics).  Tiles overlap by one row group and one cell column (factor 1.16 at S = 4): the
// halo pixels are recomputed (their gradient is needed) but only the OWNER tile of a pixel -- the one that owns the
// cell (i0, j0) of its top-left tap, or the rim -- adds it to the loss sums, the matrix and loss_px.
//
// Argmax index (first maximum, torch rule) without a compare/select chain: d_k = z_k - m is +0 exactly when z_k == m
// and has its sign bit set otherwise, so one funnel shift per channel collects the "less than max" bits and the index
// is the lowest clear bit.  d_k also feeds the exponentials (2^(d_k log2 e)), so the subtraction is not extra work on
// the ALU pipe, which the compare/select chain (2 ALU ops per channel) loaded most.  +-0 maxima and NaNs take an
// exact slow path (first NaN wins, -0 == +0).
#pragma once

#include "sf_fast.cuh"

#include <cstdlib>

#include "sf_launch.h"

namespace sf {

constexpr int kOpThreads = 128;                // label columns per tile

enum { OP_FUSED = 0, OP_BWD = 1 };

struct OpParams {
    const void* labels;
    const float* class_w;
    float* loss_px;                             // FUSED + OHEM: written (owner pixels); BWD: read
    const float* grad_out;                      // BWD
    Workspace* ws;
    unsigned long long* hist;                   // FUSED + HIST
    void* grad;                                 // FUSED: float accumulators (unscaled); BWD: LT, final
    int B, h, w, H, W, ignore, label_u8;
    float thresh;
    int tiles_x, tiles_y;
    long long lab_stride;                       // bytes between label rows (W << 3 for int64, W for uint8)
    int only_branch2;                           // BWD: unless sf_loss_state.branch == 2 only scale grad_acc into grad
    const float* grad_acc;                      // BWD + only_branch2: the fused walk's unscaled sums
    long long n_grad;                           //   B*C*h*w
    int C, th, n_chunks, ohem;                  // sf_onepass_chunk.cuh: runtime class count, tile height, 16-class chunks, OHEM on
    int xchg;                                   // FUSED: boundary rows are exchanged (two float atomics) instead of recomputed
};

// TH = owned low-res rows per tile (8 or 16); the box stages rows ia-1 .. ia+TH
template <typename LT, int C, int S, int TH>
struct OpGeom {
    static constexpr int BH = TH + 2;
    static constexpr int TBW = kOpThreads / S - 1;                  // owned low-res columns per tile
    static constexpr int JL = TBW < 32 ? 32 : 64;                   // gather lanes along j (>= TBW)
    static constexpr int KQ = kOpThreads / JL;                      // channel phases of the gather
    static constexpr int XAL = 16 / (int)sizeof(LT);
    static constexpr int BW = round_up(XAL - 1 + TBW + 2, XAL);     // box cols: floor_al(ja-1) .. ja+TBW
    static constexpr int PLANE = BH * BW;
    static constexpr int SRC_BYTES = round_up(C * PLANE * (int)sizeof(LT), 128);
    static constexpr int NP = C / 2;
    static constexpr bool ODD = (C & 1) != 0;
    static constexpr int HIST_BYTES = round_up((C + 2) * C * 4, 16);   // C x C matrix + the spare row of not-counted pixels (+ slack)
    static constexpr int ROWS = S * TH + S;                         // label rows a tile walks (at most)
    static constexpr size_t smem(bool hist) {
        return (size_t)SRC_BYTES + (size_t)2 * C * kOpThreads * 4 + (hist ? HIST_BYTES : 0) + 16;
    }
};

template <int C, int NPA>
__device__ __forceinline__ void op_onehot_sub2(int y, float2 (&a0)[NPA], float& a0l, float2 (&a1)[NPA], float& a1l,
                                               float v0, float v1) {
#define SF_CASE(K)                                                                      \
    case K:                                                                             \
        if ((K) < (C / 2) * 2) {                                                        \
            if ((K) & 1) { a0[(K) / 2 < C / 2 ? (K) / 2 : 0].y -= v0; a1[(K) / 2 < C / 2 ? (K) / 2 : 0].y -= v1; } \
            else         { a0[(K) / 2 < C / 2 ? (K) / 2 : 0].x -= v0; a1[(K) / 2 < C / 2 ? (K) / 2 : 0].x -= v1; } \
        } else if ((K) == C - 1) { a0l -= v0; a1l -= v1; }                              \
        break;
    switch (y) {
        SF_CASE(0) SF_CASE(1) SF_CASE(2) SF_CASE(3) SF_CASE(4) SF_CASE(5) SF_CASE(6) SF_CASE(7)
        SF_CASE(8) SF_CASE(9) SF_CASE(10) SF_CASE(11) SF_CASE(12) SF_CASE(13) SF_CASE(14) SF_CASE(15)
        SF_CASE(16) SF_CASE(17) SF_CASE(18) SF_CASE(19) SF_CASE(20) SF_CASE(21) SF_CASE(22) SF_CASE(23)
        default: break;
    }
#undef SF_CASE
}

__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// out[i] = acc[i] * (grad_out * inv_denom) over a grid-strided slice (this CTA's share), unless the loss state says branch 2
// (top-k): then nothing is written and false is returned.  A thread's share is small (~18 float4 at the bench shape with one
// slice per tile CTA), so the time of a CTA is a latency chain, not bandwidth: all loads of a batch of U are issued BEFORE
// the state is read (they are wasted on branch 2, which re-walks anyway), and the tail is one predicated batch instead of
// one load at a time.
template <typename LT>
__device__ __forceinline__ void scale_put4(LT* o, float4 v, float s) {
    if (sizeof(LT) == 4) {
        *reinterpret_cast<float4*>(o) = make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
    } else {
        unsigned lo, hi;                                             // 4 x 16 bit in one 8-byte store
        if (Logit<LT>::code == SF_BF16) {
            const __nv_bfloat162 a = __floats2bfloat162_rn(v.x * s, v.y * s), b = __floats2bfloat162_rn(v.z * s, v.w * s);
            lo = *reinterpret_cast<const unsigned*>(&a); hi = *reinterpret_cast<const unsigned*>(&b);
        } else {
            const __half2 a = __floats2half2_rn(v.x * s, v.y * s), b = __floats2half2_rn(v.z * s, v.w * s);
            lo = *reinterpret_cast<const unsigned*>(&a); hi = *reinterpret_cast<const unsigned*>(&b);
        }
        *reinterpret_cast<uint2*>(o) = make_uint2(lo, hi);
    }
}

template <typename LT, int U>
__device__ __forceinline__ bool scale_slice(const float* __restrict__ acc, LT* __restrict__ out, long long n,
                                            const float* __restrict__ grad_out, const sf_loss_state* __restrict__ stp,
                                            long long first, long long stride) {
    const bool vec = (((uintptr_t)acc | (uintptr_t)out) & 15) == 0;
    const long long n4 = vec ? n >> 2 : 0;
    const float4* a4 = reinterpret_cast<const float4*>(acc);
    float4 v[U];
    #pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long i = first + u * stride;
        if (i < n4) v[u] = __ldcs(a4 + i);
    }
    if (stp->branch == 2) return false;
    const float s = (grad_out ? __ldg(grad_out) : 1.0f) * stp->inv_denom;
    long long i0 = first;
    while (true) {
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i < n4) scale_put4<LT>(out + 4 * i, v[u], s);
        }
        i0 += U * stride;
        if (i0 >= n4) break;
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i < n4) v[u] = __ldcs(a4 + i);
        }
    }
    for (long long j = 4 * n4 + first; j < n; j += stride) Logit<LT>::st(out + j, acc[j] * s);
    return true;
}

template <typename LT, int C, int S, int TH, int MINB, int UNR, int MODE, bool OHEM, bool HIST>
__global__ void __launch_bounds__(kOpThreads, MINB)
onepass_kernel(const __grid_constant__ CUtensorMap tmap, const OpParams p) {
    static_assert(C >= 2 && C <= 24, "op_onehot_sub2 covers 24 classes");
    using G = OpGeom<LT, C, S, TH>;
    using TP = ScaleTap<S>;
    constexpr int NP = G::NP;
    constexpr int NPA = NP > 0 ? NP : 1;
    constexpr bool FUSED = MODE == OP_FUSED;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem[];
    const LT* Sg = reinterpret_cast<const LT*>(smem);
    float* Rbuf = reinterpret_cast<float*>(smem + G::SRC_BYTES);                // [2][C][128]
    int* hist_s = reinterpret_cast<int*>(smem + G::SRC_BYTES + 2 * C * kOpThreads * 4);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + G::SRC_BYTES + 2 * C * kOpThreads * 4 + (HIST ? G::HIST_BYTES : 0));
    __shared__ float cw_s[32];
    __shared__ double red_s[kOpThreads / 32][5];

    const sf_loss_state* stp = &p.ws->state;
    if (!FUSED && p.only_branch2) {
        // unless the loss took the top-k branch the speculative gradient sums of the fused walk stand:
        // grad = acc * grad_out / denom, this CTA's slice
        const long long cta = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        const long long n_cta = (long long)gridDim.x * gridDim.y * gridDim.z;
        if (scale_slice<LT, 20>(p.grad_acc, reinterpret_cast<LT*>(p.grad), p.n_grad, p.grad_out, stp,
                                cta * kOpThreads + threadIdx.x, n_cta * kOpThreads)) return;
    }

    // grid = (tiles_x, tiles_y, B): tile coordinates come straight from the block index, so everything derived from them
    // (row range, base pointers of the label / loss streams) stays in uniform registers
    const int tid = threadIdx.x;
    uint32_t hist_sa = smem_u32(hist_s);                            // shared-window address, converted once:
    asm volatile("mov.b32 %0, %0;" : "+r"(hist_sa));               // opaque, or ptxas rebuilds it (4 uniform instructions) per row
    const int txi = blockIdx.x, tyi = blockIdx.y, b = blockIdx.z;
    const int ia = tyi * TH, ib = min(ia + TH, p.h);
    const int ja = txi * G::TBW, jb = min(ja + G::TBW, p.w);
    const int jO = (ja - 1) & ~(G::XAL - 1);                        // box x-origin: 16-byte aligned, -XAL for the first tile
    if (tid < 32) cw_s[tid] = (p.class_w && tid < C) ? __ldg(p.class_w + tid) : 1.0f;
    if (HIST)
        for (int k = tid; k < (C + 2) * C; k += kOpThreads) hist_s[k] = 0;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_expect_tx(bar, (uint32_t)(C * G::PLANE * (int)sizeof(LT)));
        tma_load_3d(smem, &tmap, bar, jO, ia - 1, b * C);
    }

    // ---- this thread's label column ----
    const int X = S * ja - S / 2 + tid;
    const bool colok = X >= 0 && X < p.W;
    const bool own_col = colok && (tid >= S || ja == 0);            // the cell of the left tap belongs to this tile (or the rim)
    int j0, j1; float lw0, lw1;
    TP::get(min(max(X, 0), p.W - 1), p.w, j0, j1, lw0, lw1);      // idle threads keep an in-box column
    const int cj0 = j0 - jO, dj = j1 - j0;
    const LT* c0p = Sg + cj0;
    const LT* c1p = c0p + dj;
    const float2 lw0_2 = splat(lw0), lw1_2 = splat(lw1);

    // BWD: the loss state is final
    const float kth = FUSED ? 0.f : stp->kth, tie = FUSED ? 0.f : stp->tie_share;
    const float gscale = FUSED ? 1.0f : (p.grad_out ? __ldg(p.grad_out) : 1.0f) * stp->inv_denom;
    const bool bwd_has_loss = !FUSED && stp->branch != 0;           // CE: every valid pixel is kept, loss_px may be absent
    const float2 l2e2 = splat(kLog2e);

    // ---- per-pixel streams (label, loss_px): one 32-bit pixel offset `off` against per-tile base pointers ----
    // First group walked.  Halo mode: ia - 1, recomputed only for its lower vertical part into row ia.  Exchange mode (FUSED,
    // p.xchg): ia -- the tile above adds its lower part of row ia itself, and this tile adds its own into row ib.
    const bool xchg = FUSED && p.xchg != 0;
    const int u0 = (xchg && ia > 0) ? ia : ia - 1;
    const int Ya = TP::first(u0 < -1 ? -1 : u0);
    const int Yb = ib >= p.h ? p.H : TP::first(ib);                // exclusive
    const int lab_shift = p.label_u8 ? 0 : 3;
    const long long base_px = ((long long)b * p.H + Ya) * p.W;
    const char* lab_base = reinterpret_cast<const char*>(p.labels) + (base_px << lab_shift);
    float* loss_base = p.loss_px + base_px;                         // never dereferenced when p.loss_px is null
    {
        // Pull the tile's label (and loss) lines into L2 now: one row of work does not cover a DRAM miss, and doing it
        // here keeps prefetch address arithmetic out of the row loop.
        const int x0 = max(S * ja - S / 2, 0), x1 = min(S * ja - S / 2 + kOpThreads, p.W);     // columns [x0, x1)
        constexpr int LPR = kOpThreads * 8 / 128 + 1;                // 128-byte lines per row of int64 labels (uint8: clamped repeats)
        for (int i = tid; i < (Yb - Ya) * LPR; i += kOpThreads) {
            const int r = i / LPR, l = i - r * LPR;
            const long long px = (long long)r * p.W + min(x0 + ((l * 128) >> lab_shift), x1 - 1);
            prefetch_l2(lab_base + (px << lab_shift));
        }
        if (bwd_has_loss) {
            constexpr int LPR4 = kOpThreads * 4 / 128 + 1;
            for (int i = tid; i < (Yb - Ya) * LPR4; i += kOpThreads) {
                const int r = i / LPR4, l = i - r * LPR4;
                prefetch_l2(loss_base + (long long)r * p.W + min(x0 + l * 32, x1 - 1));
            }
        }
    }
    // Running 64-bit pointers (not base + offset: under register pressure ptxas rematerialises the 64-bit base products
    // inside the row loop, ~35 instructions per row).
    int nlo = -1; float nloss = 0.f;
    const char* labp = lab_base + ((size_t)(colok ? X : 0) << lab_shift);          // label of the next row to fetch
    float* lossp = loss_base + (colok ? X : 0);                     // FUSED: row being consumed (store); BWD: next row to fetch
    int rows_left = Yb - Ya;                                        // rows still to fetch; the row being consumed is one behind
    // No guard around the loads (a guard is one more branch per row): the pointers stop at the tile's last row, which is
    // then simply read again, and idle columns (outside the image) read column 0 and are masked by `colok`.
    auto fetch = [&]() {
        nlo = p.label_u8 ? (int)__ldg(reinterpret_cast<const unsigned char*>(labp)) : __ldg(reinterpret_cast<const int*>(labp));
        if (bwd_has_loss) nloss = __ldg(lossp);
        --rows_left;
        if (rows_left > 0) {                                         // CTA-uniform
            labp += p.lab_stride;
            if (!FUSED) lossp += p.W;
        }
    };
    fetch();
    const int ign = p.ignore;
    const float inv_s = 1.0f / S;

    float2 ta[NPA], tb[NPA], accA[NPA], accB[NPA];
    float tal = 0.f, tbl = 0.f, accAl = 0.f, accBl = 0.f;
    #pragma unroll
    for (int q = 0; q < NPA; ++q) { accA[q] = make_float2(0.f, 0.f); accB[q] = make_float2(0.f, 0.f); }
    auto hrow = [&](int row, float2 (&t)[NPA], float& tl) {
        const LT* q0 = c0p + row * G::BW;
        const LT* q1 = c1p + row * G::BW;
        #pragma unroll
        for (int q = 0; q < NP; ++q) {
            const float2 a = make_float2(Logit<LT>::ld(q0 + (2 * q) * G::PLANE), Logit<LT>::ld(q0 + (2 * q + 1) * G::PLANE));
            const float2 c = make_float2(Logit<LT>::ld(q1 + (2 * q) * G::PLANE), Logit<LT>::ld(q1 + (2 * q + 1) * G::PLANE));
            t[q] = lerp2_pinned(a, lw0_2, c, lw1_2);
        }
        if (G::ODD) tl = lerp_pinned(Logit<LT>::ld(q0 + (C - 1) * G::PLANE), lw0, Logit<LT>::ld(q1 + (C - 1) * G::PLANE), lw1);
    };

    int ecount = 0;
    // emit the finished source row i held in `acc`, then clear it.  Called uniformly by all threads.
    auto emit_row = [&](int i, float2 (&acc)[NPA], float& accl, bool shared_row) {
        if (i >= ia) {
            float* buf = Rbuf + (ecount & 1) * C * kOpThreads;
            ++ecount;
            #pragma unroll
            for (int q = 0; q < NP; ++q) {
                buf[(2 * q) * kOpThreads + tid] = acc[q].x;
                buf[(2 * q + 1) * kOpThreads + tid] = acc[q].y;
            }
            if (G::ODD) buf[(C - 1) * kOpThreads + tid] = accl;
            __syncthreads();
            const int jl = tid % G::JL, kq = tid / G::JL;
            const bool lrim = ja + jl == 0, rrim = ja + jl == p.w - 1;
            if (jl < jb - ja) {
                float wgt[2 * S];
                #pragma unroll
                for (int d = 0; d < 2 * S; ++d) {
                    float v = d < S ? ((float)d + 0.5f) * (1.0f / S) : 1.0f - ((float)(d - S) + 0.5f) * (1.0f / S);
                    if (d < S) v = lrim ? (d >= S / 2 ? 1.0f : 0.0f) : v;                 // clamped left rim: l0 = 1
                    else v = rrim ? (d < S + S / 2 ? 1.0f : 0.0f) : v;                    // clamped right rim: l0 + l1 = 1
                    wgt[d] = v;
                }
                const long long go = (((long long)b * C + kq) * p.h + i) * p.w + ja + jl;
                const float* rk = buf + kq * kOpThreads + S * jl;
                #pragma unroll
                for (int k = 0; k < (C + G::KQ - 1) / G::KQ; ++k) {
                    if (kq + k * G::KQ < C) {
                        float sum = 0.f;
                        #pragma unroll
                        for (int d4 = 0; d4 < 2 * S; d4 += 4) {
                            const float4 v = *reinterpret_cast<const float4*>(rk + k * G::KQ * kOpThreads + d4);
                            sum = fmaf(v.x, wgt[d4], sum); sum = fmaf(v.y, wgt[d4 + 1], sum);
                            sum = fmaf(v.z, wgt[d4 + 2], sum); sum = fmaf(v.w, wgt[d4 + 3], sum);
                        }
                        const long long o = go + (long long)k * G::KQ * p.h * p.w;
                        if (FUSED) {
                            // a row shared with the neighbouring tile has exactly two addends (this tile's part and the
                            // neighbour's) on a zeroed element: float atomics in either order give the same bits
                            if (shared_row) atomicAdd(reinterpret_cast<float*>(p.grad) + o, sum);
                            else reinterpret_cast<float*>(p.grad)[o] = sum;
                        } else {
                            Logit<LT>::st(reinterpret_cast<LT*>(p.grad) + o, sum);
                        }
                    }
                }
            }
        }
        #pragma unroll
        for (int q = 0; q < NPA; ++q) acc[q] = make_float2(0.f, 0.f);
        accl = 0.f;
    };

    // per-thread loss statistics of the owned pixels (FUSED).  OHEM needs {n_valid, n_hard, sum_hard}; CE {n_valid, sum, sum_w}.
    float stat_a = 0.f, stat_b = 0.f;                               // OHEM: sum_hard, -        CE: sum_loss, sum_w
    int cnts = 0;                                                   // n_valid | n_hard << 16 of this tile's owned pixels (<= 68 rows)
    int hrun_key = C * C, hrun_cnt = 0;                             // confusion-matrix run of this thread's column (starts in the spare row)

    // one row group u: top/bot = lerped logits of rows i0/i1, at/ab = gradient accumulators of rows i0/i1
    auto do_group = [&](int u, const float2 (&top)[NPA], float topl, const float2 (&bot)[NPA], float botl,
                        float2 (&at)[NPA], float& atl, float2 (&ab)[NPA], float& abl) {
        const bool own = own_col && (u >= ia || u < 0);              // FUSED: this tile accounts for the group's pixels
        const int i0 = max(u, 0), i1 = min(i0 + 1, p.h - 1);
        const LT* g0 = c0p + (i0 - (ia - 1)) * G::BW;                // label-logit taps of this group, channel 0
        const int di = (i1 - i0) * G::BW;
        // State that follows the label of this column.  Consecutive rows mostly share it, so everything that depends on
        // the label alone -- the horizontal lerps of the label channel, the class weight and the "- onehot" sums (applied to
        // the label's accumulator lane through a 24-way switch) -- is refreshed only when the label changes.
        int cur_y = -1; float run0 = 0.f, run1 = 0.f, ty0 = 0.f, ty1 = 0.f, wy = 1.f;
        float lh1 = u < 0 ? 0.0f : 0.5f * inv_s;                     // (r + 0.5) / S, stepped by 1/S per row
        const float lh1_step = u < 0 ? 0.0f : inv_s;
        #pragma unroll UNR
        for (int r = TP::count(u, p.h); r > 0; --r, lh1 += lh1_step) {
            const int lo = nlo; const float loss_in = nloss;
            fetch();
            const bool valid = colok && (unsigned)lo < (unsigned)C && lo != ign;
            float l = 0.f;
            if (__any_sync(FULL, valid)) {
                if (valid && lo != cur_y) {
                    if (cur_y >= 0) op_onehot_sub2<C, NPA>(cur_y, at, atl, ab, abl, run0, run1);
                    run0 = 0.f; run1 = 0.f; cur_y = lo;
                    wy = cw_s[lo];
                    if (FUSED) {
                        const LT* qk = g0 + lo * G::PLANE;
                        ty0 = lerp_pinned(Logit<LT>::ld(qk), lw0, Logit<LT>::ld(qk + dj), lw1);
                        ty1 = lerp_pinned(Logit<LT>::ld(qk + di), lw0, Logit<LT>::ld(qk + di + dj), lw1);
                    }
                }
                float coef = 0.f;                                    // weight of (softmax - onehot) for this pixel
                bool work = valid;
                if (!FUSED) {
                    float keep = 1.f;
                    if (bwd_has_loss) keep = loss_in > kth ? 1.f : (loss_in == kth ? tie : 0.f);
                    coef = (valid && keep != 0.f) ? keep * wy * gscale : 0.f;          // never 0 * inf
                    work = coef != 0.f;
                }
                if (FUSED || __any_sync(FULL, work)) {
                    const float lh0 = 1.0f - lh1;
                    const float2 lh0_2 = splat(lh0), lh1_2 = splat(lh1);
                    float2 z[NPA];
                    float zl = 0.f;
                    #pragma unroll
                    for (int q = 0; q < NP; ++q) z[q] = rnd2<LT>(lerp2_pinned(top[q], lh0_2, bot[q], lh1_2));
                    if (G::ODD) zl = Logit<LT>::rnd(lerp_pinned(topl, lh0, botl, lh1));
                    // maximum: two independent chains (a single chain is ten dependent FMNMX3)
                    float m = G::ODD ? zl : z[NP - 1].y, mb = z[0].x;
                    #pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        if (q & 1) mb = fmaxf(mb, fmaxf(z[q].x, z[q].y)); else m = fmaxf(m, fmaxf(z[q].x, z[q].y));
                    }
                    m = fmaxf(m, mb);
                    // d_k = z_k - m (<= 0, +0 exactly at the maximum), e_k = 2^(d_k log2 e)
                    const float2 nm = splat(-m);
                    unsigned bits_lo = 0u, bits_hi = 0u;             // "z_k < m" bits: channels [0, C/2) and [C/2, C)
                    constexpr int CH = C / 2;                        // split point of the two funnel-shift chains
                    if (G::ODD) {
                        zl = zl - m;
                        if (FUSED && HIST) bits_hi = __funnelshift_l(__float_as_uint(zl), bits_hi, 1);
                    }
                    #pragma unroll
                    for (int q = NP - 1; q >= 0; --q) {
                        z[q] = __fadd2_rn(z[q], nm);
                        if (FUSED && HIST) {
                            if (2 * q + 1 >= CH) bits_hi = __funnelshift_l(__float_as_uint(z[q].y), bits_hi, 1);
                            else bits_lo = __funnelshift_l(__float_as_uint(z[q].y), bits_lo, 1);
                            if (2 * q >= CH) bits_hi = __funnelshift_l(__float_as_uint(z[q].x), bits_hi, 1);
                            else bits_lo = __funnelshift_l(__float_as_uint(z[q].x), bits_lo, 1);
                        }
                    }
                    float2 a2 = make_float2(0.f, 0.f), b2 = make_float2(0.f, 0.f);
                    #pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        const float2 e = __fmul2_rn(z[q], l2e2);
                        z[q] = make_float2(ex2_approx(e.x), ex2_approx(e.y));
                        if (q & 1) b2 = __fadd2_rn(b2, z[q]); else a2 = __fadd2_rn(a2, z[q]);
                    }
                    a2 = __fadd2_rn(a2, b2);
                    float ssum = a2.x + a2.y;
                    if (G::ODD) { zl = ex2_approx(zl * kLog2e); ssum += zl; }

                    if (FUSED) {
                        // Straight-line, predicated bookkeeping: in the profile of the first version a third of all stall
                        // samples sat on branches (BRA / BSYNC / VOTE), one taken-or-not decision per feature and row.
                        // label logit: same ops, same bits as the z entry of that channel before the subtraction
                        const float zy = Logit<LT>::rnd(lerp_pinned(ty0, lh0, ty1, lh1));
                        const float lv = wy * (lg2_approx(ssum) * kLn2 - (zy - m));      // ssum >= 1: the max term contributes 2^0
                        l = valid ? lv : 0.f;
                        const bool keep = OHEM ? (valid && l > p.thresh) : valid;
                        coef = keep ? wy : 0.f;
                        work = keep;
                        const bool ov = own && valid, ok = own && keep;
                        cnts += (ov ? 1 : 0) + ((OHEM && ok) ? 0x10000 : 0);
                        if (OHEM) stat_a += ok ? l : 0.f;
                        else { stat_a += ov ? l : 0.f; stat_b += ov ? wy : 0.f; }
                        if (HIST) {
                            int idx = __ffs(~((bits_hi << CH) | bits_lo)) - 1;
                            if (!(fabsf(m * ssum) > 0.f)) {          // ssum is NaN, or the maximum is +-0 (or denormal: harmless)
                                // exact slow path: a NaN beats everything and the first NaN wins (torch); -0 == +0
                                idx = -1;
                                int first_eq = C - 1;
                                #pragma unroll 1
                                for (int k = C - 1; k >= 0; --k) {
                                    const LT* qk = g0 + k * G::PLANE;
                                    const float t0 = lerp_pinned(Logit<LT>::ld(qk), lw0, Logit<LT>::ld(qk + dj), lw1);
                                    const float t1 = lerp_pinned(Logit<LT>::ld(qk + di), lw0, Logit<LT>::ld(qk + di + dj), lw1);
                                    const float zk = Logit<LT>::rnd(lerp_pinned(t0, lh0, t1, lh1));
                                    if (zk != zk) idx = k;
                                    if (zk == m) first_eq = k;
                                }
                                if (idx < 0) idx = first_eq;
                            }
                            // run of (GT, prediction) cells down this column; pixels that are not counted (ignored or out-of-range
                            // label, halo lanes) run into the spare row C of the shared matrix: no validity select, no run guard
                            const int key = ((own && valid) ? lo : C) * C + idx;
                            const bool chg = key != hrun_key;
                            red_shared_add_if(hist_sa + 4u * (uint32_t)hrun_key, hrun_cnt, chg);
                            hrun_cnt = chg ? 1 : hrun_cnt + 1;
                            hrun_key = key;
                        }
                    }
                    // FUSED: no second vote -- on the threshold branch nearly every valid pixel is kept (steps that settle on
                    // the top-k branch are routed to the forward / backward kernels by the caller's branch hint)
                    if (FUSED || __any_sync(FULL, work)) {
                        const float cf = work ? coef * rcp_approx(ssum) : 0.f;
                        const float v0 = cf * lh0, v1 = cf * lh1;
                        const float2 v0_2 = splat(v0), v1_2 = splat(v1);
                        #pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            at[q] = __ffma2_rn(z[q], v0_2, at[q]);
                            ab[q] = __ffma2_rn(z[q], v1_2, ab[q]);
                        }
                        if (G::ODD) { atl = fmaf(zl, v0, atl); abl = fmaf(zl, v1, abl); }
                        run0 = fmaf(coef, lh0, run0); run1 = fmaf(coef, lh1, run1);     // coef == 0 where the pixel is dropped
                    }
                }
            } else if (FUSED && HIST) {
                // a row in which the whole warp is ignored still ends this column's matrix run
                const bool chg = hrun_key != C * C;
                red_shared_add_if(hist_sa + 4u * (uint32_t)hrun_key, hrun_cnt, chg);
                hrun_cnt = chg ? 1 : hrun_cnt + 1;
                hrun_key = C * C;
            }
            if (FUSED && OHEM) {
                if (own) *lossp = l;
                lossp += p.W;
            }
        }
        if (cur_y >= 0) op_onehot_sub2<C, NPA>(cur_y, at, atl, ab, abl, run0, run1);
        if (u == p.h - 1) {                                          // i1 == i0: the clamped last row takes both parts (l0 + l1 = 1)
            #pragma unroll
            for (int q = 0; q < NP; ++q) at[q] = __fadd2_rn(at[q], ab[q]);
            if (G::ODD) atl += abl;
        }
    };

    __syncthreads();
    mbar_wait(bar, 0);

    // ---- walk groups u = u0 .. ib-1, two per iteration with swapped roles (no register copies) ----
    int u = u0;
    {
        const int i0 = max(u, 0), i1 = min(i0 + 1, p.h - 1);
        hrow(i0 - (ia - 1), ta, tal);
        hrow(i1 - (ia - 1), tb, tbl);
        if (u < 0) {                                                 // rim group u = -1 shares rows (0, 1) with u = 0
            do_group(u, ta, tal, tb, tbl, accA, accAl, accB, accBl);
            ++u;
        }
    }
    const bool tail_row = xchg && ib < p.h;                          // exchange mode: row ib receives this tile's lower part
    while (u < ib) {
        do_group(u, ta, tal, tb, tbl, accA, accAl, accB, accBl);
        emit_row(u, accA, accAl, xchg && u == ia && ia > 0);         // row u complete; accB now collects row u+1
        ++u;
        if (u >= ib) { if (tail_row) emit_row(u, accB, accBl, true); break; }
        hrow(min(u + 1, p.h - 1) - (ia - 1), ta, tal);               // ta: new bottom row
        do_group(u, tb, tbl, ta, tal, accB, accBl, accA, accAl);
        emit_row(u, accB, accBl, false);                             // u > ia here
        ++u;
        if (u >= ib) { if (tail_row) emit_row(u, accA, accAl, true); break; }
        hrow(min(u + 1, p.h - 1) - (ia - 1), tb, tbl);
    }

    if (FUSED) {
        if (HIST) {
            red_shared_add_if(hist_sa + 4u * (uint32_t)hrun_key, hrun_cnt, true);
            __syncthreads();
            for (int k = tid; k < C * C; k += kOpThreads) {
                const int v = hist_s[k];
                if (v) atomicAdd(p.hist + k, (unsigned long long)v);
            }
        }
        // per-CTA sums in a fixed order, then order-free integer atomics (Workspace::fx_*)
        double v[4] = {(double)stat_a, (double)stat_b, (double)(cnts & 0xffff), (double)(cnts >> 16)};
        #pragma unroll
        for (int f = 0; f < 4; ++f) {
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[f] += __shfl_down_sync(FULL, v[f], o);
        }
        if ((tid & 31) == 0) {
            #pragma unroll
            for (int f = 0; f < 4; ++f) red_s[tid >> 5][f] = v[f];
        }
        __syncthreads();
        if (tid < 4) {
            double r = 0.0;
            for (int wq = 0; wq < kOpThreads / 32; ++wq) r += red_s[wq][tid];
            Workspace* ws = p.ws;
            if (tid == 0) { if (OHEM) fx_add(ws->fx_hard, &ws->fx_flags, 2, r); else fx_add(ws->fx_loss, &ws->fx_flags, 0, r); }
            else if (tid == 1) { if (!OHEM) fx_add(ws->fx_w, &ws->fx_flags, 1, r); }
            else if (tid == 2) atomicAdd(&ws->fx_n_valid, (unsigned long long)r);
            else if (OHEM) atomicAdd(&ws->fx_n_hard, (unsigned long long)r);
        }
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0) p.ws->n_partials = kFxMarker;
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
template <typename LT, int C, int S, int TH, int MINB, int UNR, int MODE, bool OHEM, bool HIST>
inline int op_launch(const OnepassArgs& a) {
    using G = OpGeom<LT, C, S, TH>;
    if (a.w < G::BW || a.h < G::BH) return SF_ERR_UNSUPPORTED;      // TMA box must fit inside the tensor
    OpParams p{};
    p.labels = a.labels; p.class_w = a.class_w; p.loss_px = a.loss_px; p.grad_out = a.grad_out;
    p.ws = reinterpret_cast<Workspace*>(a.ws); p.hist = a.hist;
    p.grad = MODE == OP_FUSED ? (void*)a.grad_acc : a.grad_logits;
    p.B = a.B; p.h = a.h; p.w = a.w; p.H = a.H; p.W = a.W; p.ignore = a.ignore; p.label_u8 = a.ydtype == SF_U8;
    p.thresh = a.thresh; p.only_branch2 = a.only_branch2;
    p.lab_stride = (long long)a.W << (a.ydtype == SF_U8 ? 0 : 3);
    p.grad_acc = a.grad_acc; p.n_grad = (long long)a.B * a.C * a.h * a.w;
    p.tiles_x = (a.w + G::TBW - 1) / G::TBW;
    p.tiles_y = (a.h + TH - 1) / TH;
    if (p.tiles_y > 65535 || a.B > 65535) return SF_ERR_UNSUPPORTED;
    const dim3 grid(p.tiles_x, p.tiles_y, a.B);
    CUtensorMap tmap{};
    const int rc = make_logits_tmap(&tmap, a.logits, a.ldtype, a.B * a.C, a.h, a.w, C, G::BH, G::BW);
    if (rc != 0) return rc;
    size_t smem = G::smem(HIST);
#ifdef SF_OP_EXPERIMENT
    if (const char* pad = std::getenv("SF_DEBUG_SMEM_PAD_KB")) smem += (size_t)atoi(pad) * 1024;   // occupancy sensitivity knob
#endif
    auto k = onepass_kernel<LT, C, S, TH, MINB, UNR, MODE, OHEM, HIST>;
    static SmemLimitOnce smem_once;                                  // per instantiation
    const cudaError_t ea = set_smem_limit_once(k, smem, smem_once);
    if (ea != cudaSuccess) return (int)ea;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a.stream);
    if (MODE == OP_FUSED) {
        Workspace* w = reinterpret_cast<Workspace*>(a.ws);
        cudaError_t e = cudaMemsetAsync(w->fx_loss, 0, offsetof(Workspace, pad1) - offsetof(Workspace, fx_loss), st);
        if (e != cudaSuccess) return (int)e;
        // Exchange mode: the first row of every tile is shared with the tile above (two atomic addends on a zeroed element).
        // With h a multiple of TH those rows are every TH-th row of the whole [B*C*h][w] array: one strided memset.
        p.xchg = p.tiles_y > 1 && a.h % TH == 0 && !env_no_xchg();
        if (p.xchg) {
            e = cudaMemset2DAsync(a.grad_acc, (size_t)TH * a.w * sizeof(float), 0, (size_t)a.w * sizeof(float),
                                  (size_t)a.B * a.C * (a.h / TH), st);
            if (e != cudaSuccess) return (int)e;
        }
    }
    k<<<grid, kOpThreads, smem, st>>>(tmap, p);
    return (int)cudaGetLastError();
}

// resident CTAs per SM of the shipped configuration (SF_DEBUG_OP_VARIANT selects other tile heights / occupancies in experiment builds)
constexpr int kOpMinB = 3;

inline int op_variant() {
    static int v = []() { const char* e = std::getenv("SF_DEBUG_OP_VARIANT"); return e ? atoi(e) : 0; }();
    return v;
}

template <typename LT, int C, int S, int TH, int MINB, int UNR>
inline int op_mode2(const OnepassArgs& a) {
    if (a.mode == OP_BWD) return op_launch<LT, C, S, TH, MINB, 1, OP_BWD, false, false>(a);
    const bool ohem = a.thresh >= 0.f;
    if (a.hist) return ohem ? op_launch<LT, C, S, TH, MINB, UNR, OP_FUSED, true, true>(a) : op_launch<LT, C, S, TH, MINB, UNR, OP_FUSED, false, true>(a);
    return ohem ? op_launch<LT, C, S, TH, MINB, UNR, OP_FUSED, true, false>(a) : op_launch<LT, C, S, TH, MINB, UNR, OP_FUSED, false, false>(a);
}
template <typename LT, int C, int S>
inline int op_mode(const OnepassArgs& a) {
#ifdef SF_OP_EXPERIMENT
    switch (op_variant()) {
        case 1: return op_mode2<LT, C, S, 8, 4, 1>(a);
        case 2: return op_mode2<LT, C, S, 16, 3, 1>(a);
        case 3: return op_mode2<LT, C, S, 8, 3, 1>(a);
        case 4: return op_mode2<LT, C, S, 16, 2, 2>(a);
        case 5: return op_mode2<LT, C, S, 8, 2, 2>(a);
        default: break;
    }
#endif
    // 16-row tiles recompute fewer halo rows (17/16 vs 9/8 of the pixels) but need enough tiles to fill the GPU and must
    // leave room for kOpMinB CTAs per SM in shared memory
    const long long tiles16 = (long long)a.B * ((a.h + 15) / 16) * ((a.w + OpGeom<LT, C, S, 16>::TBW - 1) / OpGeom<LT, C, S, 16>::TBW);
    constexpr bool fits16 = OpGeom<LT, C, S, 16>::smem(true) * kOpMinB <= 227 * 1024;   // e.g. C = 21 fp32: 78 KB -> only two CTAs per SM
    if (fits16 && a.h >= 18 && tiles16 >= 4LL * kOpMinB * sm_count()) return op_mode2<LT, C, S, 16, kOpMinB, 1>(a);
    return op_mode2<LT, C, S, 8, kOpMinB, 1>(a);
}
template <typename LT, int C>
inline int op_scale(const OnepassArgs& a, int s) {
    switch (s) {
        case 4: return op_mode<LT, C, 4>(a);
        case 8: return op_mode<LT, C, 8>(a);
        default: return SF_ERR_UNSUPPORTED;
    }
}
template <typename LT>
inline int op_classes(const OnepassArgs& a, int s) {
    switch (a.C) {
        case 19: return op_scale<LT, 19>(a, s);
#ifndef SF_OP_EXPERIMENT
        case 21: return op_scale<LT, 21>(a, s);
        case 9: return op_scale<LT, 9>(a, s);                       // Synapse (train_gpu.py:48-49)
        case 2: return op_scale<LT, 2>(a, s);                       // Kvasir (binary)
#endif
        default: return SF_ERR_UNSUPPORTED;
    }
}

}  // namespace sf
