// sf_onepass.cu -- ONE walk over the label pixels for forward AND backward (integer-scale heads, C <= 24).
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
#include "sf_fast.cuh"

#include "sf_launch.h"

namespace sf {

constexpr int kOpThreads = 128;                // label columns per tile
constexpr int kOpTH = 8;                       // owned low-res rows per tile
constexpr int kOpBH = kOpTH + 2;               // staged source rows: ia-1 .. ia+TH
constexpr int kOpPrefetchRows = 3;             // L2 prefetch distance of the per-pixel streams, in label rows

enum { OP_FUSED = 0, OP_BWD = 1 };

struct OpParams {
    const void* labels;
    const float* class_w;
    float* loss_px;                             // FUSED: written (owner pixels); BWD: read
    const float* grad_out;                      // BWD
    Workspace* ws;
    unsigned long long* hist;
    void* grad;                                 // FUSED: float accumulators (unscaled); BWD: LT, final
    int B, h, w, H, W, ignore, label_u8;
    float thresh;
    int tiles_x, tiles_y;
    int only_branch2;                           // BWD: exit at once unless sf_loss_state.branch == 2
};

template <typename LT, int C, int S>
struct OpGeom {
    static constexpr int TBW = kOpThreads / S - 1;                  // owned low-res columns per tile
    static constexpr int JL = TBW < 32 ? 32 : 64;                   // gather lanes along j (>= TBW)
    static constexpr int KQ = kOpThreads / JL;                      // channel phases of the gather
    static constexpr int XAL = 16 / (int)sizeof(LT);
    static constexpr int BW = round_up(XAL - 1 + TBW + 2, XAL);     // box cols: floor_al(ja-1) .. ja+TBW
    static constexpr int PLANE = kOpBH * BW;
    static constexpr int SRC_BYTES = round_up(C * PLANE * (int)sizeof(LT), 128);
    static constexpr int NP = C / 2;
    static constexpr bool ODD = (C & 1) != 0;
    static constexpr int HIST_BYTES = round_up(C * C * 4, 16);
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

#ifndef SF_OP_CTAS_PER_SM
#define SF_OP_CTAS_PER_SM 3
#endif

template <typename LT, int C, int S, int MODE, bool OHEM, bool HIST>
__global__ void __launch_bounds__(kOpThreads, SF_OP_CTAS_PER_SM)
onepass_kernel(const __grid_constant__ CUtensorMap tmap, const OpParams p) {
    static_assert(C >= 2 && C <= 24, "op_onehot_sub2 covers 24 classes");
    using G = OpGeom<LT, C, S>;
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
    if (!FUSED && p.only_branch2 && stp->branch != 2) return;       // the speculative gradient of the fused walk stands

    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int txi = bid % p.tiles_x; bid /= p.tiles_x;
    const int tyi = bid % p.tiles_y;
    const int b = bid / p.tiles_y;
    const int ia = tyi * kOpTH, ib = min(ia + kOpTH, p.h);
    const int ja = txi * G::TBW, jb = min(ja + G::TBW, p.w);
    const int jO = (ja - 1) & ~(G::XAL - 1);                        // box x-origin: 16-byte aligned, -XAL for the first tile
    if (tid < 32) cw_s[tid] = (p.class_w && tid < C) ? __ldg(p.class_w + tid) : 1.0f;
    const bool want_hist = HIST && p.hist != nullptr;
    if (HIST)
        for (int k = tid; k < C * C; k += kOpThreads) hist_s[k] = 0;
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
    const int cj0 = j0 - jO, cj1 = j1 - jO;
    const LT* c0p = Sg + cj0;
    const LT* c1p = Sg + cj1;
    const float2 lw0_2 = splat(lw0), lw1_2 = splat(lw1);

    // BWD: the loss state is final
    const float kth = FUSED ? 0.f : stp->kth, tie = FUSED ? 0.f : stp->tie_share;
    const float gscale = FUSED ? 1.0f : (p.grad_out ? __ldg(p.grad_out) : 1.0f) * stp->inv_denom;
    const bool bwd_has_loss = !FUSED && stp->branch != 0;           // CE: every valid pixel is kept, loss_px may be absent
    const float2 l2e2 = splat(kLog2e);

    // ---- row walk state ----
    const int Ya = TP::first(ia - 1 < -1 ? -1 : ia - 1);
    const int Yb = ib >= p.h ? p.H : TP::first(ib);                // exclusive
    int nlo = -1; float nloss = 0.f;
    long long pxo = ((long long)b * p.H + Ya) * p.W + (colok ? X : 0);           // pixel offset of the next row to fetch
    int to_fetch = colok ? Yb - Ya : 0;                             // rows still to fetch; the row being consumed is one behind
    const int lab_shift = p.label_u8 ? 0 : 3;
    const long long pf_px = (long long)kOpPrefetchRows * p.W;
    auto prefetch = [&]() {
        if (to_fetch > 0) {
            const char* labp = reinterpret_cast<const char*>(p.labels) + (pxo << lab_shift);
            nlo = p.label_u8 ? (int)__ldg(reinterpret_cast<const unsigned char*>(labp)) : __ldg(reinterpret_cast<const int*>(labp));
            if (bwd_has_loss) nloss = __ldg(p.loss_px + pxo);
            if (to_fetch > kOpPrefetchRows) {
                prefetch_l2(labp + (pf_px << lab_shift));
                if (bwd_has_loss) prefetch_l2(p.loss_px + pxo + pf_px);
            }
        }
        pxo += p.W; --to_fetch;
    };
    prefetch();
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
    auto emit_row = [&](int i, float2 (&acc)[NPA], float& accl) {
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
                        if (FUSED) reinterpret_cast<float*>(p.grad)[o] = sum;
                        else Logit<LT>::st(reinterpret_cast<LT*>(p.grad) + o, sum);
                    }
                }
            }
        }
        #pragma unroll
        for (int q = 0; q < NPA; ++q) acc[q] = make_float2(0.f, 0.f);
        accl = 0.f;
    };

    // per-thread loss statistics of the owned pixels (FUSED)
    float tile_loss = 0.f, tile_w = 0.f, tile_hard = 0.f;
    int cnt_valid = 0, cnt_hard = 0;
    int hrun_key = -1, hrun_cnt = 0;                                // confusion-matrix run of this thread's column

    // one row group u: top/bot = lerped logits of rows i0/i1, at/ab = gradient accumulators of rows i0/i1
    auto do_group = [&](int u, const float2 (&top)[NPA], float topl, const float2 (&bot)[NPA], float botl,
                        float2 (&at)[NPA], float& atl, float2 (&ab)[NPA], float& abl) {
        const int cnt = TP::count(u, p.h);
        const bool merge = u == p.h - 1;                             // i1 == i0: both parts land in row h-1
        const bool own = own_col && (u >= ia || u < 0);              // FUSED: this tile accounts for the group's pixels
        const int i0 = max(u, 0), i1 = min(i0 + 1, p.h - 1);
        const int o0 = (i0 - (ia - 1)) * G::BW, o1 = (i1 - (ia - 1)) * G::BW;
        // "- onehot" part: consecutive rows of a column mostly share their label, so its two sums are kept as a
        // run and applied to the label's accumulator lane (a 24-way switch) only when the label changes
        int run_y = -1; float run0 = 0.f, run1 = 0.f;
        float lh1 = u < 0 ? 0.0f : 0.5f * inv_s;                     // (r + 0.5) / S, stepped by 1/S per row
        const float lh1_step = u < 0 ? 0.0f : inv_s;
        #pragma unroll 1
        for (int r = 0; r < cnt; ++r, lh1 += lh1_step) {
            const int lo = nlo; const float loss_in = nloss;
            const bool alive = to_fetch >= 0;                        // prefetch() runs one row ahead
            prefetch();                                              // pxo now points two rows past the current one
            const bool valid = alive && (unsigned)lo < (unsigned)C && lo != ign;
            const float wy = cw_s[valid ? lo : 0];
            float coef = 0.f;                                        // weight of (softmax - onehot) for this pixel
            bool work = valid;
            if (!FUSED) {
                float keep = 1.f;
                if (bwd_has_loss) keep = loss_in > kth ? 1.f : (loss_in == kth ? tie : 0.f);
                coef = (valid && keep != 0.f) ? keep * wy * gscale : 0.f;          // never 0 * inf
                work = coef != 0.f;
            }
            float l = 0.f;
            if (__any_sync(FULL, work)) {
                const float lh0 = 1.0f - lh1;
                const float2 lh0_2 = splat(lh0), lh1_2 = splat(lh1);
                float2 z[NPA];
                float zl = 0.f;
                #pragma unroll
                for (int q = 0; q < NP; ++q) z[q] = rnd2<LT>(lerp2_pinned(top[q], lh0_2, bot[q], lh1_2));
                if (G::ODD) zl = Logit<LT>::rnd(lerp_pinned(topl, lh0, botl, lh1));
                float m = G::ODD ? zl : z[NP - 1].y;
                #pragma unroll
                for (int q = 0; q < NP; ++q) m = fmaxf(m, fmaxf(z[q].x, z[q].y));
                // d_k = z_k - m (<= 0), e_k = 2^(d_k log2 e)
                const float2 nm = splat(-m);
                unsigned bits_lo = 0u, bits_hi = 0u;                 // "z_k < m" bits: channels [0, C/2) and [C/2, C)
                constexpr int CH = C / 2;                            // split point of the two funnel-shift chains
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
                    // label logit: same ops, same bits as the z entry of that channel before the subtraction
                    float zy = 0.f;
                    if (valid) {
                        const LT* qk = Sg + lo * G::PLANE;
                        const float t0 = lerp_pinned(Logit<LT>::ld(qk + o0 + cj0), lw0, Logit<LT>::ld(qk + o0 + cj1), lw1);
                        const float t1 = lerp_pinned(Logit<LT>::ld(qk + o1 + cj0), lw0, Logit<LT>::ld(qk + o1 + cj1), lw1);
                        zy = Logit<LT>::rnd(lerp_pinned(t0, lh0, t1, lh1));
                        const float lg = lg2_approx(ssum) * kLn2;                  // ssum >= 1: the max term contributes 2^0
                        l = wy * (lg - (zy - m));
                    }
                    const bool keep = OHEM ? (valid && l > p.thresh) : valid;
                    coef = keep ? wy : 0.f;
                    if (own) {
                        if (valid) {
                            tile_loss += l; tile_w += wy; ++cnt_valid;
                            if (OHEM && keep) { tile_hard += l; ++cnt_hard; }
                        }
                        if (HIST && want_hist) {
                            int idx = __ffs(~((bits_hi << CH) | bits_lo)) - 1;
                            if (ssum != ssum || m == 0.f) {
                                // exact slow path: a NaN beats everything and the first NaN wins (torch); -0 == +0
                                idx = -1;
                                int first_eq = C - 1;
                                #pragma unroll 1
                                for (int k = C - 1; k >= 0; --k) {
                                    const LT* qk = Sg + k * G::PLANE;
                                    const float t0 = lerp_pinned(Logit<LT>::ld(qk + o0 + cj0), lw0, Logit<LT>::ld(qk + o0 + cj1), lw1);
                                    const float t1 = lerp_pinned(Logit<LT>::ld(qk + o1 + cj0), lw0, Logit<LT>::ld(qk + o1 + cj1), lw1);
                                    const float zk = Logit<LT>::rnd(lerp_pinned(t0, lh0, t1, lh1));
                                    if (zk != zk) idx = k;
                                    if (zk == m) first_eq = k;
                                }
                                if (idx < 0) idx = first_eq;
                            }
                            const int key = valid ? lo * C + idx : -1;
                            if (key != hrun_key) {
                                if (hrun_key >= 0) atomicAdd(&hist_s[hrun_key], hrun_cnt);
                                hrun_key = key; hrun_cnt = 0;
                            }
                            ++hrun_cnt;
                        }
                    }
                    work = keep;
                }
                if (__any_sync(FULL, work)) {
                    const float cf = work ? __fdividef(coef, ssum) : 0.f;
                    const float aw0 = merge ? 1.0f : lh0, aw1 = merge ? 0.0f : lh1;   // clamped last row: l0 + l1 = 1
                    const float v0 = cf * aw0, v1 = cf * aw1;
                    const float2 v0_2 = splat(v0), v1_2 = splat(v1);
                    #pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        at[q] = __ffma2_rn(z[q], v0_2, at[q]);
                        ab[q] = __ffma2_rn(z[q], v1_2, ab[q]);
                    }
                    if (G::ODD) { atl = fmaf(zl, v0, atl); abl = fmaf(zl, v1, abl); }
                    if (work) {
                        if (lo != run_y) {
                            if (run_y >= 0) op_onehot_sub2<C, NPA>(run_y, at, atl, ab, abl, run0, run1);
                            run_y = lo; run0 = 0.f; run1 = 0.f;
                        }
                        run0 += coef * aw0; run1 += coef * aw1;
                    }
                }
            } else if (FUSED && HIST && want_hist && own) {
                // a row in which the whole warp is ignored still ends this column's matrix run
                if (hrun_key >= 0) { atomicAdd(&hist_s[hrun_key], hrun_cnt); hrun_key = -1; hrun_cnt = 0; }
                ++hrun_cnt;
            }
            if (FUSED && own && p.loss_px) p.loss_px[pxo - 2 * (long long)p.W] = l;
        }
        if (run_y >= 0) op_onehot_sub2<C, NPA>(run_y, at, atl, ab, abl, run0, run1);
    };

    __syncthreads();
    mbar_wait(bar, 0);

    // ---- walk groups u = ia-1 .. ib-1, two per iteration with swapped roles (no register copies) ----
    int u = ia - 1;
    {
        const int i0 = max(u, 0), i1 = min(i0 + 1, p.h - 1);
        hrow(i0 - (ia - 1), ta, tal);
        hrow(i1 - (ia - 1), tb, tbl);
        if (u < 0) {                                                 // rim group u = -1 shares rows (0, 1) with u = 0
            do_group(u, ta, tal, tb, tbl, accA, accAl, accB, accBl);
            ++u;
        }
    }
    while (u < ib) {
        do_group(u, ta, tal, tb, tbl, accA, accAl, accB, accBl);
        emit_row(u, accA, accAl);                                    // row u complete; accB now collects row u+1
        ++u;
        if (u >= ib) break;
        hrow(min(u + 1, p.h - 1) - (ia - 1), ta, tal);               // ta: new bottom row
        do_group(u, tb, tbl, ta, tal, accB, accBl, accA, accAl);
        emit_row(u, accB, accBl);
        ++u;
        if (u >= ib) break;
        hrow(min(u + 1, p.h - 1) - (ia - 1), tb, tbl);
    }

    if (FUSED) {
        if (HIST && want_hist) {
            if (hrun_key >= 0) atomicAdd(&hist_s[hrun_key], hrun_cnt);
            __syncthreads();
            for (int k = tid; k < C * C; k += kOpThreads) {
                const int v = hist_s[k];
                if (v) atomicAdd(p.hist + k, (unsigned long long)v);
            }
        }
        // per-CTA sums in a fixed order, then order-free integer atomics (Workspace::fx_*)
        double v[5] = {(double)tile_loss, (double)tile_w, (double)tile_hard, (double)cnt_valid, (double)cnt_hard};
        #pragma unroll
        for (int f = 0; f < 5; ++f) {
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[f] += __shfl_down_sync(FULL, v[f], o);
        }
        if ((tid & 31) == 0) {
            #pragma unroll
            for (int f = 0; f < 5; ++f) red_s[tid >> 5][f] = v[f];
        }
        __syncthreads();
        if (tid < 5) {
            double r = 0.0;
            for (int wq = 0; wq < kOpThreads / 32; ++wq) r += red_s[wq][tid];
            Workspace* ws = p.ws;
            if (tid == 0) fx_add(ws->fx_loss, &ws->fx_flags, 0, r);
            else if (tid == 1) fx_add(ws->fx_w, &ws->fx_flags, 1, r);
            else if (tid == 2) { if (OHEM) fx_add(ws->fx_hard, &ws->fx_flags, 2, r); }
            else if (tid == 3) atomicAdd(&ws->fx_n_valid, (unsigned long long)r);
            else if (OHEM) atomicAdd(&ws->fx_n_hard, (unsigned long long)r);
        }
        if (blockIdx.x == 0 && tid == 0) p.ws->n_partials = kFxMarker;
    }
}

// grad = acc * (grad_out * inv_denom), unless the fused walk's speculation failed (branch 2: OP_BWD wrote grad)
template <typename LT>
__global__ void __launch_bounds__(256)
onepass_scale_kernel(const float4* __restrict__ acc, LT* __restrict__ out, long long n4, long long n,
                     const float* __restrict__ grad_out, const Workspace* __restrict__ ws) {
    if (ws->state.branch == 2) return;
    const float s = (grad_out ? __ldg(grad_out) : 1.0f) * ws->state.inv_denom;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldcs(acc + i);
        LT* o = out + 4 * i;
        if (sizeof(LT) == 4) {
            *reinterpret_cast<float4*>(o) = make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
        } else {
            Logit<LT>::st(o, v.x * s); Logit<LT>::st(o + 1, v.y * s); Logit<LT>::st(o + 2, v.z * s); Logit<LT>::st(o + 3, v.w * s);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long i = 4 * n4; i < n; ++i) Logit<LT>::st(out + i, reinterpret_cast<const float*>(acc)[i] * s);
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
template <typename LT, int C, int S, int MODE, bool OHEM, bool HIST>
static int op_launch(const OnepassArgs& a) {
    using G = OpGeom<LT, C, S>;
    if (a.w < G::BW || a.h < kOpBH) return SF_ERR_UNSUPPORTED;      // TMA box must fit inside the tensor
    OpParams p{};
    p.labels = a.labels; p.class_w = a.class_w; p.loss_px = a.loss_px; p.grad_out = a.grad_out;
    p.ws = reinterpret_cast<Workspace*>(a.ws); p.hist = a.hist;
    p.grad = MODE == OP_FUSED ? (void*)a.grad_acc : a.grad_logits;
    p.B = a.B; p.h = a.h; p.w = a.w; p.H = a.H; p.W = a.W; p.ignore = a.ignore; p.label_u8 = a.ydtype == SF_U8;
    p.thresh = a.thresh; p.only_branch2 = a.only_branch2;
    p.tiles_x = (a.w + G::TBW - 1) / G::TBW;
    p.tiles_y = (a.h + kOpTH - 1) / kOpTH;
    const long long grid = (long long)a.B * p.tiles_x * p.tiles_y;
    if (grid >= (1LL << 31)) return SF_ERR_SHAPE;
    CUtensorMap tmap{};
    const int rc = make_logits_tmap(&tmap, a.logits, a.ldtype, a.B * a.C, a.h, a.w, C, kOpBH, G::BW);
    if (rc != 0) return rc;
    const size_t smem = G::smem(HIST);
    auto k = onepass_kernel<LT, C, S, MODE, OHEM, HIST>;
    static bool attr_set_dev[64] = {};                               // per instantiation and device; idempotent, so a race is harmless
    int dev = 0; cudaGetDevice(&dev);
    bool& attr_set = attr_set_dev[dev & 63];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a.stream);
    if (MODE == OP_FUSED) {
        Workspace* w = reinterpret_cast<Workspace*>(a.ws);
        cudaError_t e = cudaMemsetAsync(w->fx_loss, 0, offsetof(Workspace, pad1) - offsetof(Workspace, fx_loss), st);
        if (e != cudaSuccess) return (int)e;
    }
    k<<<(int)grid, kOpThreads, smem, st>>>(tmap, p);
    return (int)cudaGetLastError();
}

template <typename LT, int C, int S>
static int op_mode(const OnepassArgs& a) {
    if (a.mode == OP_BWD) return op_launch<LT, C, S, OP_BWD, false, false>(a);
    const bool ohem = a.thresh >= 0.f;
    if (a.hist) return ohem ? op_launch<LT, C, S, OP_FUSED, true, true>(a) : op_launch<LT, C, S, OP_FUSED, false, true>(a);
    return ohem ? op_launch<LT, C, S, OP_FUSED, true, false>(a) : op_launch<LT, C, S, OP_FUSED, false, false>(a);
}
template <typename LT, int C>
static int op_scale(const OnepassArgs& a, int s) {
    switch (s) {
        case 4: return op_mode<LT, C, 4>(a);
        case 8: return op_mode<LT, C, 8>(a);
        default: return SF_ERR_UNSUPPORTED;
    }
}
template <typename LT>
static int op_classes(const OnepassArgs& a, int s) {
    switch (a.C) {
        case 19: return op_scale<LT, 19>(a, s);
        case 21: return op_scale<LT, 21>(a, s);
        default: return SF_ERR_UNSUPPORTED;
    }
}

bool onepass_supported(int ldtype, const void* logits, int C, int h, int w, int H, int W) {
    if (env_no_tma() || env_no_fast() || h <= 0 || w <= 0 || H % h || W % w) return false;
    const int s = H / h;
    if (s != W / w || (s != 4 && s != 8)) return false;
    if (C != 19 && C != 21) return false;
    const int esz = ldtype == SF_F32 ? 4 : 2;
    if (((size_t)w * esz) % 16 || ((uintptr_t)logits % 16)) return false;
    const int xal = 16 / esz;
    const int bw = round_up(xal - 1 + (kOpThreads / s - 1) + 2, xal);
    return w >= bw && h >= kOpBH;
}

// SF_ERR_UNSUPPORTED when the shape is not covered (the caller then takes the forward / finalize / backward route).
int launch_onepass(const OnepassArgs& a) {
    if (!onepass_supported(a.ldtype, a.logits, a.C, a.h, a.w, a.H, a.W)) return SF_ERR_UNSUPPORTED;
    const int s = a.H / a.h;
    if (a.ldtype == SF_F32) return op_classes<float>(a, s);
    if (a.ldtype == SF_BF16) return op_classes<__nv_bfloat16>(a, s);
    return op_classes<__half>(a, s);
}

int launch_onepass_scale(const float* acc, void* out, int ldtype, long long n, const float* grad_out, const void* ws,
                         void* stream) {
    const long long n4 = (((uintptr_t)acc | (uintptr_t)out) % 16) == 0 ? n / 4 : 0;
    const int grid = (int)std::min<long long>(std::max<long long>((n4 + 255) / 256, 1), 8LL * sm_count());
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const float4* a4 = reinterpret_cast<const float4*>(acc);
    const Workspace* w = reinterpret_cast<const Workspace*>(ws);
    if (ldtype == SF_F32) onepass_scale_kernel<float><<<grid, 256, 0, st>>>(a4, reinterpret_cast<float*>(out), n4, n, grad_out, w);
    else if (ldtype == SF_BF16) onepass_scale_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(a4, reinterpret_cast<__nv_bfloat16*>(out), n4, n, grad_out, w);
    else onepass_scale_kernel<__half><<<grid, 256, 0, st>>>(a4, reinterpret_cast<__half*>(out), n4, n, grad_out, w);
    return (int)cudaGetLastError();
}

}  // namespace sf
