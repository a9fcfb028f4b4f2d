// sf_onepass_chunk.cuh -- the one-pass forward+backward walk for MANY classes (24 < C <= 255: ADE20K 150, COCO-Stuff 171).
//
// Same contract as sf_onepass.cuh (models/build_models.py:65 upsample, util/losses.py:32-36 / 52-57 per-pixel CE and the
// OHEM `loss > thresh` test, util/metrics.py:25-27 argmax + bincount, and autograd's backward of all of it), for class
// counts whose logits no longer fit one thread's registers.
//
// Schedule.  A CTA owns TH x (NT/S - 1) low-res cells; its NT threads are the NT label columns that interpolate from them.
// The walk goes down the tile one ROW GROUP at a time (the S label rows between two low-res rows) and makes two passes
// over the classes per group, in chunks of 16 staged by a TMA/mbarrier ring ([16 classes][2 rows][BW] per stage, re-read
// from L2 for the second pass):
//   pass A  online softmax per pixel: running max m, running sum s of 2^(z log2e - rn(m log2e)), the label logit;
//           then, per pixel and ONCE per group instead of once per chunk: loss, the OHEM keep test, the loss sums, loss_px;
//   pass B  the chunk's logits are rebuilt, e_k = 2^(z_k log2e - rn(m log2e)) and the two vertical parts
//           e_k * keep * w_y / s * {lh0, lh1} summed over the group's rows; these 2 x 16 values per column go through shared
//           memory, where thread (cell, class phase) sums the 2S columns that depend on its cell (the horizontal part of
//           the transposed bilinear).  The lower part is carried -- per cell, in shared memory -- to the next group, the
//           upper part plus the carry is the finished gradient row.  Every element is written once, no atomics.
//           The argmax index also falls out of pass B: d_k = z_k - m is +0 exactly at the maximum, one funnel shift per
//           class collects the sign bits, the first chunk with a clear bit holds the first maximum (torch's rule).
// Versus sf_forward_chunk.cu + sf_backward_fast.cu (chunked): the per-pixel scalar work (label, keep, weights, "- onehot")
// runs once per group instead of once per chunk and kernel, the lse_px round trip is gone, labels are read once.
//
// The matrix is counted with warp-aggregated 64-bit reds straight into the C x C table (it would take 90-117 KB of shared
// memory, i.e. the occupancy of the walk).
#pragma once

#include "sf_onepass.cuh"

namespace sf {

constexpr int kOcKC = 16;                      // classes per chunk (8 fp32x2 pairs)
constexpr int kOcStages = 4;

template <typename LT, int S, int NT>
struct OcGeom {
    static constexpr int TBW = NT / S - 1;                           // owned low-res columns per tile
    static constexpr int JL = NT / S;                                // gather lanes along j (TBW + 1)
    static constexpr int KQ = S;                                     // class phases of the gather (NT / JL)
    static constexpr int NK = kOcKC / KQ;                            // classes per gather thread and chunk
    static constexpr int XAL = 16 / (int)sizeof(LT);
    static constexpr int BW = round_up(XAL - 1 + TBW + 2, XAL);      // box cols: floor_al(ja-1) .. ja+TBW
    static constexpr int PLANE = 2 * BW;                             // two source rows per class
    static constexpr int STAGE_BYTES = round_up(kOcKC * PLANE * (int)sizeof(LT), 128);
    static constexpr int RBUF_BYTES = 2 * 2 * kOcKC * NT * 4;        // [2 buffers][top, bottom][16][NT] floats
    static size_t smem(int n_chunks) {
        return (size_t)kOcStages * STAGE_BYTES + RBUF_BYTES + (size_t)n_chunks * kOcKC * JL * 4 + 2 * kOcStages * 8 + 16;
    }
};

template <typename LT, int S, int NT, int MINB, int MODE, bool HIST>
__global__ void __launch_bounds__(NT, MINB)
onepass_chunk_kernel(const __grid_constant__ CUtensorMap tmap, const OpParams p) {
    using G = OcGeom<LT, S, NT>;
    using TP = ScaleTap<S>;
    constexpr int KC = kOcKC, NP = KC / 2, NS = kOcStages;
    constexpr bool FUSED = MODE == OP_FUSED;
    constexpr bool WANT_IDX = FUSED && HIST;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem[];
    float* Rbuf = reinterpret_cast<float*>(smem + NS * G::STAGE_BYTES);
    float* carry = reinterpret_cast<float*>(smem + NS * G::STAGE_BYTES + G::RBUF_BYTES);      // [n_chunks * 16][JL]
    uint64_t* full = reinterpret_cast<uint64_t*>(carry + p.n_chunks * KC * G::JL);
    uint64_t* empty = full + NS;
    __shared__ float cw_s[256];
    __shared__ double red_s[NT / 32][4];

    const sf_loss_state* stp = &p.ws->state;
    if (!FUSED && p.only_branch2) {
        // unless the loss took the top-k branch the speculative gradient sums of the fused walk stand
        const long long cta = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        const long long n_cta = (long long)gridDim.x * gridDim.y * gridDim.z;
        if (scale_slice<LT, 16>(p.grad_acc, reinterpret_cast<LT*>(p.grad), p.n_grad, p.grad_out, stp,
                                cta * NT + threadIdx.x, n_cta * NT)) return;
    }

    const int tid = threadIdx.x, lane = tid & 31;
    const int C = p.C, NCH = p.n_chunks;
    const int txi = blockIdx.x, tyi = blockIdx.y, b = blockIdx.z;
    const int ia = tyi * p.th, ib = min(ia + p.th, p.h);
    const int ja = txi * G::TBW, jb = min(ja + G::TBW, p.w);
    const int jO = (ja - 1) & ~(G::XAL - 1);                        // box x-origin: 16-byte aligned, -XAL for the first tile
    for (int k = tid; k < 256; k += NT) cw_s[k] = (p.class_w && k < C) ? __ldg(p.class_w + k) : 1.0f;
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NT / 32); }
        mbar_fence_init();
    }
    __syncthreads();

    // ---- the item stream: (group, pass, chunk), in the order it is consumed; thread 0 is the producer ----
    const int total_items = (ib - ia + 1) * 2 * NCH;
    int is_n = 0, is_slot = 0, is_c = 0, is_pass = 0, is_u = ia - 1;
    auto issue_next = [&]() {
        mbar_expect_tx(&full[is_slot], (uint32_t)(KC * G::PLANE * (int)sizeof(LT)));
        // rows max(u, 0) and the one below (zero fill past the last row: then both vertical taps are the first row);
        // a tail chunk reads past its classes (the next image's, or zero fill): masked below
        tma_load_3d(smem + is_slot * G::STAGE_BYTES, &tmap, &full[is_slot], jO, max(is_u, 0), b * C + is_c * KC);
        ++is_n;
        if (++is_slot == NS) is_slot = 0;
        if (++is_c == NCH) { is_c = 0; if (++is_pass == 2) { is_pass = 0; ++is_u; } }
    };
    if (tid == 0)
        for (int n = 0; n < NS && n < total_items; ++n) issue_next();
    int n_item = 0, slot = 0, phase = 0, prev_slot = 0, prev_phase = 0;
    auto acquire = [&]() -> const LT* {
        mbar_wait(&full[slot], (uint32_t)phase);
        return reinterpret_cast<const LT*>(smem + slot * G::STAGE_BYTES);
    };
    // Done with the stage: every warp arrives on its `empty` barrier; the producer refills the stage of the PREVIOUS item (one
    // item of slack, so it rarely waits for the slowest warp).
    auto release = [&]() {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
        if (tid == 0 && n_item >= 1 && is_n < total_items) {
            mbar_wait(&empty[prev_slot], (uint32_t)prev_phase);
            issue_next();
        }
        prev_slot = slot; prev_phase = phase;
        ++n_item;
        if (++slot == NS) { slot = 0; phase ^= 1; }
    };

    // ---- this thread's label column ----
    const int X = S * ja - S / 2 + tid;
    const bool colok = X >= 0 && X < p.W;
    const bool own_col = colok && (tid >= S || ja == 0);            // the cell of the left tap belongs to this tile (or the rim)
    int j0, j1; float lw0, lw1;
    TP::get(min(max(X, 0), p.W - 1), p.w, j0, j1, lw0, lw1);      // idle threads keep an in-box column
    const int cj0 = j0 - jO, dj = j1 - j0;
    const float2 lw0_2 = splat(lw0), lw1_2 = splat(lw1);
    const float2 l2e2 = splat(kLog2e);

    // BWD: the loss state is final
    const float kth = FUSED ? 0.f : stp->kth, tie = FUSED ? 0.f : stp->tie_share;
    const float gscale = FUSED ? 1.0f : (p.grad_out ? __ldg(p.grad_out) : 1.0f) * stp->inv_denom;
    const bool bwd_has_loss = !FUSED && stp->branch != 0;           // CE: every valid pixel is kept, loss_px may be absent
    const bool ohem = p.ohem != 0;

    // ---- per-pixel streams (label, loss_px) ----
    const int Ya = TP::first(ia - 1 < -1 ? -1 : ia - 1);
    const int Yb = ib >= p.h ? p.H : TP::first(ib);                // exclusive
    const int lab_shift = p.label_u8 ? 0 : 3;
    const long long base_px = ((long long)b * p.H + Ya) * p.W;
    const char* lab_base = reinterpret_cast<const char*>(p.labels) + (base_px << lab_shift);
    float* loss_base = p.loss_px + base_px;                         // never dereferenced when p.loss_px is null
    {
        // pull the tile's label (and loss) lines into L2 now
        const int x0 = max(S * ja - S / 2, 0), x1 = min(S * ja - S / 2 + NT, p.W);
        constexpr int LPR = NT * 8 / 128 + 1;
        for (int i = tid; i < (Yb - Ya) * LPR; i += NT) {
            const int r = i / LPR, l = i - r * LPR;
            const long long px = (long long)r * p.W + min(x0 + ((l * 128) >> lab_shift), x1 - 1);
            prefetch_l2(lab_base + (px << lab_shift));
        }
        if (bwd_has_loss) {
            constexpr int LPR4 = NT * 4 / 128 + 1;
            for (int i = tid; i < (Yb - Ya) * LPR4; i += NT) {
                const int r = i / LPR4, l = i - r * LPR4;
                prefetch_l2(loss_base + (long long)r * p.W + min(x0 + l * 32, x1 - 1));
            }
        }
    }
    const char* labp = lab_base + ((size_t)(colok ? X : 0) << lab_shift);          // first row of the next group to fetch
    float* lossp = loss_base + (colok ? X : 0);                                     // first row of the group being consumed
    const float* lossq = lossp;                                                     // BWD: next group to fetch
    int nlab[S]; float nloss[S];
    auto fetch_group = [&](int u) {                                  // CTA-uniform trip counts
        const int cn = TP::count(u, p.h);
        #pragma unroll
        for (int r = 0; r < S; ++r) {
            nlab[r] = -1; nloss[r] = 0.f;
            if (r < cn) {
                const char* q = labp + (long long)r * p.lab_stride;
                nlab[r] = p.label_u8 ? (int)__ldg(reinterpret_cast<const unsigned char*>(q)) : __ldg(reinterpret_cast<const int*>(q));
                if (bwd_has_loss) nloss[r] = __ldg(lossq + (long long)r * p.W);
            }
        }
        labp += (long long)cn * p.lab_stride;
        lossq += (long long)cn * p.W;
    };
    fetch_group(ia - 1);
    const int ign = p.ignore;

    // gather role: owned cell jl, classes kq + i * KQ of every chunk
    const int jl = tid % G::JL, kq = tid / G::JL;
    const bool gather_on = jl < jb - ja;
    const bool lrim = ja + jl == 0, rrim = ja + jl == p.w - 1;

    float stat_a = 0.f, stat_b = 0.f;                               // OHEM: sum_hard, -        CE: sum_loss, sum_w
    int cnts = 0;                                                   // n_valid | n_hard << 16 of this tile's owned pixels
    int ecount = 0;
    const unsigned lane_lt = (1u << lane) - 1u;

    for (int u = ia - 1; u < ib; ++u) {
        const int cnt = TP::count(u, p.h);
        const int i0 = max(u, 0), i1 = min(i0 + 1, p.h - 1);
        const int di = (i1 - i0) * G::BW;
        const bool merge = u == p.h - 1;                             // i1 == i0: both vertical parts land in row h-1
        const bool own = own_col && (u >= ia || u < 0);              // FUSED: this tile accounts for the group's pixels
        int lab[S]; float loss_in[S];
        unsigned vmask = 0u;
        #pragma unroll
        for (int r = 0; r < S; ++r) {
            lab[r] = nlab[r]; loss_in[r] = nloss[r];
            if (r < cnt && colok && (unsigned)lab[r] < (unsigned)C && lab[r] != ign) vmask |= 1u << r;
        }
        if (u + 1 < ib) fetch_group(u + 1);
        const bool warp_valid = __any_sync(FULL, vmask != 0u);

        // ---------------- pass A: online softmax ----------------
        float m[S], s[S], zy[S];
        #pragma unroll
        for (int r = 0; r < S; ++r) { m[r] = -INFINITY; s[r] = 0.f; zy[r] = 0.f; }
        for (int c = 0; c < NCH; ++c) {
            const LT* Sg = acquire();
            if (warp_valid) {
                const int c0 = c * KC;
                const int nvalid = min(KC, C - c0);
                const LT* q00 = Sg + cj0;
                float2 top[NP], bot[NP];
                #pragma unroll
                for (int q = 0; q < NP; ++q) {
                    const LT* qa = q00 + (2 * q) * G::PLANE;
                    const LT* qb = qa + G::PLANE;
                    const float2 a = make_float2(Logit<LT>::ld(qa), Logit<LT>::ld(qb));
                    const float2 cc = make_float2(Logit<LT>::ld(qa + dj), Logit<LT>::ld(qb + dj));
                    top[q] = lerp2_pinned(a, lw0_2, cc, lw1_2);
                    const float2 d = make_float2(Logit<LT>::ld(qa + di), Logit<LT>::ld(qb + di));
                    const float2 e = make_float2(Logit<LT>::ld(qa + di + dj), Logit<LT>::ld(qb + di + dj));
                    bot[q] = lerp2_pinned(d, lw0_2, e, lw1_2);
                }
                #pragma unroll
                for (int r = 0; r < S; ++r) {
                    if (r < cnt && __any_sync(FULL, (vmask >> r) & 1u)) {
                        const float lh1 = u < 0 ? 0.0f : ((float)r + 0.5f) * (1.0f / S);
                        const float lh0 = 1.0f - lh1;
                        const float2 lh0_2 = splat(lh0), lh1_2 = splat(lh1);
                        float2 z[NP];
                        #pragma unroll
                        for (int q = 0; q < NP; ++q) z[q] = rnd2<LT>(lerp2_pinned(top[q], lh0_2, bot[q], lh1_2));
                        if (nvalid < KC) {                           // tail chunk: classes past C never win nor add
                            #pragma unroll
                            for (int q = 0; q < NP; ++q) {
                                if (2 * q >= nvalid) z[q].x = -INFINITY;
                                if (2 * q + 1 >= nvalid) z[q].y = -INFINITY;
                            }
                        }
                        float cm = fmaxf(z[0].x, z[0].y), cmb = fmaxf(z[1].x, z[1].y);
                        #pragma unroll
                        for (int q = 2; q < NP; ++q) {
                            if (q & 1) cmb = fmaxf(cmb, fmaxf(z[q].x, z[q].y)); else cm = fmaxf(cm, fmaxf(z[q].x, z[q].y));
                        }
                        cm = fmaxf(cm, cmb);
                        // every term stays relative to ONE rounded exponent rn(m log2e); when m grows the sum is rescaled by the
                        // difference of the rounded exponents (as sf_forward_chunk.cu)
                        const float mn = fmaxf(m[r], cm);
                        const float m2n = __fmul_rn(mn, kLog2e);
                        const float sc = (m[r] == mn) ? 1.0f : ex2_approx(__fsub_rn(__fmul_rn(m[r], kLog2e), m2n));
                        const float2 nm2 = splat(mn == -INFINITY ? 0.0f : -m2n);
                        float2 a2 = make_float2(0.f, 0.f), b2 = make_float2(0.f, 0.f);
                        #pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            const float2 e = __ffma2_rn(z[q], l2e2, nm2);
                            const float2 ex = make_float2(ex2_approx(e.x), ex2_approx(e.y));
                            if (q & 1) b2 = __fadd2_rn(b2, ex); else a2 = __fadd2_rn(a2, ex);
                        }
                        a2 = __fadd2_rn(a2, b2);
                        s[r] = fmaf(s[r], sc, a2.x + a2.y);
                        m[r] = mn;
                        if (FUSED) {
                            const int yk = lab[r] - c0;
                            if (((vmask >> r) & 1u) && (unsigned)yk < (unsigned)KC) {   // label logit: same ops, same bits as its z[] entry
                                const LT* qk = q00 + yk * G::PLANE;
                                const float t0 = lerp_pinned(Logit<LT>::ld(qk), lw0, Logit<LT>::ld(qk + dj), lw1);
                                const float t1 = lerp_pinned(Logit<LT>::ld(qk + di), lw0, Logit<LT>::ld(qk + di + dj), lw1);
                                zy[r] = Logit<LT>::rnd(lerp_pinned(t0, lh0, t1, lh1));
                            }
                        }
                    }
                }
            }
            release();
        }

        // ---------------- per pixel, once per group: loss, keep, weights of pass B ----------------
        float nm2r[S], v0[S], v1[S], coefr[S];
        int idx[S];
        unsigned wmask = 0u, smask = 0u;                             // rows with a kept pixel / rows that take the exact argmax
        #pragma unroll
        for (int r = 0; r < S; ++r) {
            nm2r[r] = 0.f; v0[r] = 0.f; v1[r] = 0.f; coefr[r] = 0.f; idx[r] = -1;
            if (r < cnt) {
                const bool valid = (vmask >> r) & 1u;
                const float lh1 = u < 0 ? 0.0f : ((float)r + 0.5f) * (1.0f / S);
                const float lh0 = 1.0f - lh1;
                const float m2 = __fmul_rn(m[r], kLog2e);
                const float wy = cw_s[valid ? lab[r] : 0];
                float coef = 0.f;
                if (FUSED) {
                    const float corr = __fmaf_rn(m[r], kLog2e, -m2);                  // exact residual of rn(m log2e)
                    const float lg = (lg2_approx(s[r]) - corr) * kLn2;
                    const float lv = wy * (lg - (zy[r] - m[r]));
                    const float l = valid ? lv : 0.f;
                    const bool keep = ohem ? (valid && l > p.thresh) : valid;
                    coef = keep ? wy : 0.f;
                    const bool ov = own && valid, ok = own && keep;
                    cnts += (ov ? 1 : 0) + ((ohem && ok) ? 0x10000 : 0);
                    if (ohem) stat_a += ok ? l : 0.f;
                    else { stat_a += ov ? l : 0.f; stat_b += ov ? wy : 0.f; }
                    if (ohem && own) lossp[(long long)r * p.W] = l;
                    if (WANT_IDX && valid && !(fabsf(m[r] * s[r]) > 0.f)) smask |= 1u << r;   // NaN / Inf somewhere, or a +-0 maximum
                } else {
                    float keep = 1.f;
                    if (bwd_has_loss) keep = loss_in[r] > kth ? 1.f : (loss_in[r] == kth ? tie : 0.f);
                    coef = (valid && keep != 0.f) ? keep * wy * gscale : 0.f;         // never 0 * inf
                }
                if (coef != 0.f) wmask |= 1u << r;
                const float cf = coef != 0.f ? coef * rcp_approx(s[r]) : 0.f;
                const float aw0 = merge ? 1.0f : lh0, aw1 = merge ? 0.0f : lh1;       // clamped last row: l0 + l1 = 1
                v0[r] = cf * aw0; v1[r] = cf * aw1;
                coefr[r] = coef;
                nm2r[r] = m[r] == -INFINITY ? 0.0f : -m2;
            }
        }
        lossp += (long long)cnt * p.W;
        const unsigned rowmask = WANT_IDX ? vmask : wmask;          // rows whose logits pass B rebuilds
        const bool warp_work = __any_sync(FULL, rowmask != 0u);
        unsigned found_nan = 0u;                                     // exact argmax: rows whose index is a NaN's (final)

        // ---------------- pass B: gradient parts, argmax ----------------
        for (int c = 0; c < NCH; ++c) {
            const LT* Sg = acquire();
            const int c0 = c * KC;
            const int nvalid = min(KC, C - c0);
            float2 at[NP], ab[NP];
            #pragma unroll
            for (int q = 0; q < NP; ++q) { at[q] = make_float2(0.f, 0.f); ab[q] = make_float2(0.f, 0.f); }
            if (warp_work) {
                const LT* q00 = Sg + cj0;
                float2 top[NP], bot[NP];
                #pragma unroll
                for (int q = 0; q < NP; ++q) {
                    const LT* qa = q00 + (2 * q) * G::PLANE;
                    const LT* qb = qa + G::PLANE;
                    const float2 a = make_float2(Logit<LT>::ld(qa), Logit<LT>::ld(qb));
                    const float2 cc = make_float2(Logit<LT>::ld(qa + dj), Logit<LT>::ld(qb + dj));
                    top[q] = lerp2_pinned(a, lw0_2, cc, lw1_2);
                    const float2 d = make_float2(Logit<LT>::ld(qa + di), Logit<LT>::ld(qb + di));
                    const float2 e = make_float2(Logit<LT>::ld(qa + di + dj), Logit<LT>::ld(qb + di + dj));
                    bot[q] = lerp2_pinned(d, lw0_2, e, lw1_2);
                }
                #pragma unroll
                for (int r = 0; r < S; ++r) {
                    if (r < cnt && __any_sync(FULL, (rowmask >> r) & 1u)) {
                        const float lh1 = u < 0 ? 0.0f : ((float)r + 0.5f) * (1.0f / S);
                        const float lh0 = 1.0f - lh1;
                        const float2 lh0_2 = splat(lh0), lh1_2 = splat(lh1);
                        float2 z[NP];
                        #pragma unroll
                        for (int q = 0; q < NP; ++q) z[q] = rnd2<LT>(lerp2_pinned(top[q], lh0_2, bot[q], lh1_2));
                        if (nvalid < KC) {
                            #pragma unroll
                            for (int q = 0; q < NP; ++q) {
                                if (2 * q >= nvalid) z[q].x = -INFINITY;
                                if (2 * q + 1 >= nvalid) z[q].y = -INFINITY;
                            }
                        }
                        if (WANT_IDX) {
                            // d_k = z_k - m: +0 exactly at the maximum, sign bit set below it
                            const float2 nm = splat(-m[r]);
                            unsigned bits = 0u;
                            #pragma unroll
                            for (int q = NP - 1; q >= 0; --q) {
                                const float2 d = __fadd2_rn(z[q], nm);
                                bits = __funnelshift_l(__float_as_uint(d.y), bits, 1);
                                bits = __funnelshift_l(__float_as_uint(d.x), bits, 1);
                            }
                            const unsigned at_max = ~bits & 0xffffu;
                            const bool special = (smask >> r) & 1u;
                            if (!special && idx[r] < 0 && at_max) idx[r] = c0 + __ffs(at_max) - 1;
                            if (__any_sync(FULL, special)) {
                                // exact path: a NaN beats everything and the first NaN wins (torch); -0 == +0; +-Inf compare equal to m
                                if (special && !((found_nan >> r) & 1u)) {
                                    int kn = -1, ke = -1;
                                    #pragma unroll
                                    for (int q = NP - 1; q >= 0; --q) {
                                        if (z[q].y != z[q].y) kn = 2 * q + 1;
                                        if (z[q].y == m[r]) ke = 2 * q + 1;
                                        if (z[q].x != z[q].x) kn = 2 * q;
                                        if (z[q].x == m[r]) ke = 2 * q;
                                    }
                                    if (kn >= 0) { idx[r] = c0 + kn; found_nan |= 1u << r; }
                                    else if (ke >= 0 && idx[r] < 0) idx[r] = c0 + ke;
                                }
                            }
                        }
                        const float2 nm2 = splat(nm2r[r]), v0_2 = splat(v0[r]), v1_2 = splat(v1[r]);
                        #pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            const float2 e = __ffma2_rn(z[q], l2e2, nm2);
                            const float2 ex = make_float2(ex2_approx(e.x), ex2_approx(e.y));
                            at[q] = __ffma2_rn(ex, v0_2, at[q]);
                            ab[q] = __ffma2_rn(ex, v1_2, ab[q]);
                        }
                    }
                }
            }
            release();                                               // the stage is free; what follows lives in Rbuf / carry

            // ---- emission: R[part][k][column] -> shared memory, "- onehot" on the label's class, then the horizontal gather ----
            float* buf = Rbuf + (ecount & 1) * (2 * KC * NT);
            ++ecount;
            #pragma unroll
            for (int q = 0; q < NP; ++q) {
                buf[(2 * q) * NT + tid] = at[q].x;
                buf[(2 * q + 1) * NT + tid] = at[q].y;
                buf[(KC + 2 * q) * NT + tid] = ab[q].x;
                buf[(KC + 2 * q + 1) * NT + tid] = ab[q].y;
            }
            #pragma unroll
            for (int r = 0; r < S; ++r) {
                const int yk = lab[r] - c0;
                if (((wmask >> r) & 1u) && (unsigned)yk < (unsigned)KC) {
                    const float lh1 = u < 0 ? 0.0f : ((float)r + 0.5f) * (1.0f / S);
                    const float lh0 = 1.0f - lh1;
                    const float aw0 = merge ? 1.0f : lh0, aw1 = merge ? 0.0f : lh1;
                    buf[yk * NT + tid] -= coefr[r] * aw0;            // own column: plain read-modify-write
                    buf[(KC + yk) * NT + tid] -= coefr[r] * aw1;
                }
            }
            __syncthreads();
            if (gather_on) {
                float wgt[2 * S];
                #pragma unroll
                for (int d = 0; d < 2 * S; ++d) {
                    float v = d < S ? ((float)d + 0.5f) * (1.0f / S) : 1.0f - ((float)(d - S) + 0.5f) * (1.0f / S);
                    if (d < S) v = lrim ? (d >= S / 2 ? 1.0f : 0.0f) : v;                 // clamped left rim: l0 = 1
                    else v = rrim ? (d < S + S / 2 ? 1.0f : 0.0f) : v;                    // clamped right rim: l0 + l1 = 1
                    wgt[d] = v;
                }
                const float* rk = buf + kq * NT + S * jl;
                float* cp = carry + (c0 + kq) * G::JL + jl;
                const long long go = (((long long)b * C + c0 + kq) * p.h + i0) * p.w + ja + jl;
                #pragma unroll
                for (int k = 0; k < G::NK; ++k) {
                    if (kq + k * G::KQ < nvalid) {
                        float st = 0.f, sb = 0.f;
                        #pragma unroll
                        for (int d4 = 0; d4 < 2 * S; d4 += 4) {
                            const float4 vt = *reinterpret_cast<const float4*>(rk + k * G::KQ * NT + d4);
                            const float4 vb = *reinterpret_cast<const float4*>(rk + (KC + k * G::KQ) * NT + d4);
                            st = fmaf(vt.x, wgt[d4], st); st = fmaf(vt.y, wgt[d4 + 1], st);
                            st = fmaf(vt.z, wgt[d4 + 2], st); st = fmaf(vt.w, wgt[d4 + 3], st);
                            sb = fmaf(vb.x, wgt[d4], sb); sb = fmaf(vb.y, wgt[d4 + 1], sb);
                            sb = fmaf(vb.z, wgt[d4 + 2], sb); sb = fmaf(vb.w, wgt[d4 + 3], sb);
                        }
                        float* cpk = cp + k * G::KQ * G::JL;
                        if (u < 0) {
                            *cpk = st;                               // rim group: its part of row 0 waits for group 0 (lower part is 0)
                        } else {
                            const float prev = u == ia - 1 ? 0.f : *cpk;     // the tile's first group has no carry
                            if (u >= ia) {
                                const long long o = go + (long long)k * G::KQ * p.h * p.w;
                                if (FUSED) reinterpret_cast<float*>(p.grad)[o] = prev + st;
                                else Logit<LT>::st(reinterpret_cast<LT*>(p.grad) + o, prev + st);
                            }
                            *cpk = sb;
                        }
                    }
                }
            }
        }

        if (WANT_IDX && warp_valid) {
            // confusion counts of the group's owned pixels: one red per distinct (GT, prediction) cell and warp row
            #pragma unroll
            for (int r = 0; r < S; ++r) {
                if (r < cnt) {
                    const int key = (own && ((vmask >> r) & 1u)) ? lab[r] * C + max(idx[r], 0) : -1;
                    const unsigned peers = __match_any_sync(FULL, key);
                    if (key >= 0 && (peers & lane_lt) == 0u) atomicAdd(p.hist + key, (unsigned long long)__popc(peers));
                }
            }
        }
    }

    if (FUSED) {
        // per-CTA sums in a fixed order, then order-free integer atomics (Workspace::fx_*)
        double v[4] = {(double)stat_a, (double)stat_b, (double)(cnts & 0xffff), (double)(cnts >> 16)};
        #pragma unroll
        for (int f = 0; f < 4; ++f) {
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[f] += __shfl_down_sync(FULL, v[f], o);
        }
        if (lane == 0) {
            #pragma unroll
            for (int f = 0; f < 4; ++f) red_s[tid >> 5][f] = v[f];
        }
        __syncthreads();
        if (tid < 4) {
            double r = 0.0;
            for (int wq = 0; wq < NT / 32; ++wq) r += red_s[wq][tid];
            Workspace* ws = p.ws;
            if (tid == 0) { if (ohem) fx_add(ws->fx_hard, &ws->fx_flags, 2, r); else fx_add(ws->fx_loss, &ws->fx_flags, 0, r); }
            else if (tid == 1) { if (!ohem) fx_add(ws->fx_w, &ws->fx_flags, 1, r); }
            else if (tid == 2) atomicAdd(&ws->fx_n_valid, (unsigned long long)r);
            else if (ohem) atomicAdd(&ws->fx_n_hard, (unsigned long long)r);
        }
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0) p.ws->n_partials = kFxMarker;
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
struct OcPlan { int th; long long cost; int occ; };

// Tile height for one (NT, occupancy): the walk is a few hundred long CTAs, so wave quantisation decides -- take the height
// that minimises waves x (TH + 1) groups (every tile also walks the group above its first row).
inline OcPlan oc_plan(int B, int h, int tiles_x, int occ) {
    OcPlan best{0, 0, occ};
    const long long slots = (long long)occ * sm_count();
    for (int th = 2; th <= 60; ++th) {
        if (th > h) break;
        const long long tiles = (long long)B * tiles_x * ((h + th - 1) / th);
        const long long waves = (tiles + slots - 1) / slots;
        const long long cost = waves * (th + 1);
        if (best.th == 0 || cost < best.cost) { best.th = th; best.cost = cost; }
    }
    return best;
}

template <typename LT, int S, int NT, int MINB, int MODE, bool HIST>
inline int oc_launch(const OnepassArgs& a, int th) {
    using G = OcGeom<LT, S, NT>;
    OpParams p{};
    p.labels = a.labels; p.class_w = a.class_w; p.loss_px = a.loss_px; p.grad_out = a.grad_out;
    p.ws = reinterpret_cast<Workspace*>(a.ws); p.hist = a.hist;
    p.grad = MODE == OP_FUSED ? (void*)a.grad_acc : a.grad_logits;
    p.B = a.B; p.h = a.h; p.w = a.w; p.H = a.H; p.W = a.W; p.ignore = a.ignore; p.label_u8 = a.ydtype == SF_U8;
    p.thresh = a.thresh; p.only_branch2 = a.only_branch2;
    p.lab_stride = (long long)a.W << (a.ydtype == SF_U8 ? 0 : 3);
    p.grad_acc = a.grad_acc; p.n_grad = (long long)a.B * a.C * a.h * a.w;
    p.C = a.C; p.th = th; p.n_chunks = (a.C + kOcKC - 1) / kOcKC; p.ohem = a.thresh >= 0.f;
    p.tiles_x = (a.w + G::TBW - 1) / G::TBW;
    p.tiles_y = (a.h + th - 1) / th;
    if (p.tiles_y > 65535 || a.B > 65535) return SF_ERR_UNSUPPORTED;
    const dim3 grid(p.tiles_x, p.tiles_y, a.B);
    CUtensorMap tmap{};
    const int rc = make_logits_tmap(&tmap, a.logits, a.ldtype, a.B * a.C, a.h, a.w, kOcKC, 2, G::BW);
    if (rc != 0) return rc;
    const size_t smem = G::smem(p.n_chunks);
    auto k = onepass_chunk_kernel<LT, S, NT, MINB, MODE, HIST>;
    static SmemLimitOnce smem_once;                                  // per instantiation
    const cudaError_t ea = set_smem_limit_once(k, smem, smem_once);
    if (ea != cudaSuccess) return (int)ea;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a.stream);
    if (MODE == OP_FUSED) {
        Workspace* w = reinterpret_cast<Workspace*>(a.ws);
        cudaError_t e = cudaMemsetAsync(w->fx_loss, 0, offsetof(Workspace, pad1) - offsetof(Workspace, fx_loss), st);
        if (e != cudaSuccess) return (int)e;
    }
    k<<<grid, NT, smem, st>>>(tmap, p);
    return (int)cudaGetLastError();
}

// resident CTAs per SM: shared memory (ring + Rbuf + carry, 1 KB reserved per CTA) against the register bound MINB
template <typename LT, int S, int NT, int MINB>
inline int oc_occupancy(int n_chunks) {
    const size_t per_cta = OcGeom<LT, S, NT>::smem(n_chunks) + 1024 + 1280;      // + static shared memory
    const int by_smem = (int)((size_t)228 * 1024 / per_cta);
    return std::max(1, std::min(MINB, by_smem));
}

template <typename LT, int S, int NT, int MINB>
inline int oc_mode(const OnepassArgs& a, int th) {
    if (a.mode == OP_BWD) return oc_launch<LT, S, NT, MINB, OP_BWD, false>(a, th);
    return a.hist ? oc_launch<LT, S, NT, MINB, OP_FUSED, true>(a, th) : oc_launch<LT, S, NT, MINB, OP_FUSED, false>(a, th);
}

// debug / test knobs, read per launch: force the tile width (128 / 192 columns) and height
inline int oc_forced_nt() { const char* e = std::getenv("SF_DEBUG_OC_NT"); return e ? atoi(e) : 0; }
inline int oc_forced_th() { const char* e = std::getenv("SF_DEBUG_OC_TH"); return e ? std::min(std::max(atoi(e), 0), 60) : 0; }

template <typename LT, int S>
inline int oc_scale(const OnepassArgs& a) {
    const int n_chunks = (a.C + kOcKC - 1) / kOcKC;
    using G1 = OcGeom<LT, S, 128>;
    using G2 = OcGeom<LT, S, 192>;
    const bool fit1 = a.w >= G1::BW, fit2 = a.w >= G2::BW;         // the TMA box must fit inside the tensor
    if (!fit1 || a.h < 2) return SF_ERR_UNSUPPORTED;
    OcPlan p1 = oc_plan(a.B, a.h, (a.w + G1::TBW - 1) / G1::TBW, oc_occupancy<LT, S, 128, 3>(n_chunks));
    OcPlan p2 = oc_plan(a.B, a.h, (a.w + G2::TBW - 1) / G2::TBW, oc_occupancy<LT, S, 192, 2>(n_chunks));
    // both shapes keep 12 warps per SM at full occupancy, where a group takes the same time: compare waves x groups
    bool wide = fit2 && p2.cost < p1.cost;
    if (oc_forced_nt() == 128) wide = false;
    if (oc_forced_nt() == 192 && fit2) wide = true;
    int th = wide ? p2.th : p1.th;
    if (oc_forced_th() > 0) th = oc_forced_th();
    return wide ? oc_mode<LT, S, 192, 2>(a, th) : oc_mode<LT, S, 128, 3>(a, th);
}

template <typename LT>
inline int oc_dispatch(const OnepassArgs& a, int s) {
    switch (s) {
        case 4: return oc_scale<LT, 4>(a);
        case 8: return oc_scale<LT, 8>(a);
        default: return SF_ERR_UNSUPPORTED;
    }
}

}  // namespace sf
