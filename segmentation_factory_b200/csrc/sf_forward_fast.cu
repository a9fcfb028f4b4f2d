// sf_forward_fast.cu -- integer-scale, all-channels-in-registers specialisation of the fused forward.
//
// Same contract and same arithmetic as sf_forward.cu (bit-identical argmax / matrix, same loss formula); used
// when H == S*h, W == S*w, S in {2,4,8} and C is one of the instantiated class counts.  What changes is the
// schedule (measured on B200, profiles/r01_v1_*: the generic kernel issued 34 instructions per pixel-channel,
// two thirds of them integer/addressing/setup):
//   * a thread owns one label column and walks a strip of GPS row groups (GPS*S rows): the horizontal lerps of a
//     source row are computed once and the bottom row of one group is reused as the top row of the next;
//   * geometry is constexpr (sf_fast.cuh), shared-memory offsets fold into LDS immediates;
//   * vertical lerp, exp argument and exp sum run as packed fp32x2 over channel pairs (FFMA2/FMUL2/FADD2);
//   * labels of the next group are prefetched while the current one is computed;
//   * a NaN in the logits is detected for free (the exp sum turns NaN) and resolved on a rare slow path.
// Tile = 64 columns x 4 strips x GPS groups; one TMA box [C][4*GPS+1][BW] per tile, 2-stage mbarrier pipeline,
// persistent CTAs (2 per SM).
#include "sf_fast.cuh"
#include <cstdlib>

#include "sf_dice_bins.cuh"

#include "sf_launch.h"

namespace sf {

// Experiment knob.  Measured: 3 CTAs per SM (80 registers, 100-470 B of spills) is SLOWER for the eval-only mode -- 0.128 vs
// 0.120 ms per 8 Cityscapes images -- so 2 (113 registers, no spills) ships.
#ifndef SF_FF_ARGMAX_CTAS
#define SF_FF_ARGMAX_CTAS 2
#endif
constexpr int kFfArgmaxCtas = SF_FF_ARGMAX_CTAS;                   // resident CTAs per SM of the eval-only (argmax + matrix) mode
constexpr int kFfTW = 64, kFfStrips = 4, kFfThreads = kFfTW * kFfStrips, kFfStages = 2;

struct FfParams {
    const void* labels;
    const float* class_w;
    float* loss_px;
    float* lse_px;
    unsigned long long* hist;
    Workspace* ws;
    int B, h, w, H, W, ignore, label_u8;
    float thresh;
    int tiles_x, tiles_y, total_tiles;
    DiceArgs dice;                                                  // MODE >= 2 (sf_dice.cu)
    float* fold_f32; long long* fold_i64; long long* fold_mat;      // FF_ARGMAX + fold: the per-call bookkeeping of sf_eval_update
    int fold;
};

enum { FF_ARGMAX = 0, FF_SOFTMAX = 1, FF_DICE_STATS = 2, FF_DICE_COEF = 3, FF_SOFTMAX_STATS = 4, FF_SOFTMAX_NOHIST = 5 };

// S labels of one byte each: 32 bits are enough for S <= 4 (64-bit shifts / selects cost two instructions each)
template <int S> struct FfLabPack { using type = unsigned long long; };
template <> struct FfLabPack<4> { using type = unsigned int; };
template <> struct FfLabPack<2> { using type = unsigned int; };

template <typename LT, int C, int S>
struct FfGeom {
    static constexpr int GPS = 16 / S >= 2 ? 16 / S : 2;            // 16 rows per strip (S=8: 2 groups)
    static constexpr int NGT = kFfStrips * GPS;                     // groups per tile
    static constexpr int BH = NGT + 1;
    static constexpr int XAL = 16 / (int)sizeof(LT);
    static constexpr int BW = round_up(XAL + kFfTW / S + 1, XAL);
    static constexpr int PLANE = BH * BW;
    static constexpr int STAGE_BYTES = round_up(C * PLANE * (int)sizeof(LT), 128);
    static constexpr int NP = C / 2;                                // channel pairs; channel C-1 is the odd tail
    static constexpr bool ODD = (C & 1) != 0;
};

// MODE: FF_ARGMAX / FF_SOFTMAX are the forward proper; FF_DICE_STATS / FF_DICE_COEF are the two dice walks of
// sf_dice.cu (same tiles, staging and strip walk, different per-pixel body: softmax from the stored lse, no argmax).
template <typename LT, int C, int S, int MODE>
__global__ void __launch_bounds__(kFfThreads, (MODE == FF_ARGMAX && C <= 21) ? kFfArgmaxCtas : 2)
fwd_fast_kernel(const __grid_constant__ CUtensorMap tmap, const FfParams p) {
    using G = FfGeom<LT, C, S>;
    using TP = ScaleTap<S>;
    constexpr int NP = G::NP;
    constexpr bool CSTAT = MODE == FF_SOFTMAX_STATS;                // CE forward that also gathers the dice sums (criterion)
    constexpr bool NOHIST = MODE == FF_SOFTMAX_NOHIST;              // loss only (training without metrics): no argmax index
    constexpr bool NEED_IDX = MODE == FF_ARGMAX || MODE == FF_SOFTMAX;   // the modes that can feed the confusion matrix
    constexpr bool SOFTMAX = MODE == FF_SOFTMAX || CSTAT || NOHIST;
    constexpr bool DSTAT = MODE == FF_DICE_STATS, DCOEF = MODE == FF_DICE_COEF, DICE = DSTAT || DCOEF;
    constexpr bool ACCP = DSTAT || CSTAT;                           // per-thread sum p_k accumulators + shared bins
    constexpr int CP = C + (C & 1);                                 // even pitch of the per-image dice tables (LDS.64 pairs)
    extern __shared__ __align__(128) unsigned char smem[];
    int* hist_s = reinterpret_cast<int*>(smem + kFfStages * G::STAGE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kFfStages * G::STAGE_BYTES + round_up((C + 2) * C * 4, 16));
    __shared__ double red_scratch[kFfThreads / 32];
    __shared__ uint64_t empty_bar[kFfStages];                       // consumer -> producer: "every warp is done with this stage"
    __shared__ float cw_s[32];                                      // class weights as LDS: an LDG in the row loop would share
                                                                    // a scoreboard slot with the label prefetch and serialise it
    const int tid = threadIdx.x;
    const int tx = tid % kFfTW, sy = tid / kFfTW;
    if (tid < 32) cw_s[tid] = (p.class_w && tid < C) ? __ldg(p.class_w + tid) : 1.0f;
    // FF_ARGMAX exists only to fill the matrix (sf_eval_hist refuses a null one): a compile-time `true` there keeps the
    // per-row test of a kernel parameter (LDCU + 2 UISETP + BRA.U) out of the eval-only row loop
    const bool want_hist = MODE == FF_ARGMAX ? true : (!DICE && !CSTAT && !NOHIST && p.hist != nullptr);
    uint32_t hist_sa = smem_u32(hist_s);
    asm volatile("mov.b32 %0, %0;" : "+r"(hist_sa));               // opaque: keeps the shared-window address in a register instead of
                                                                    // rebuilding it (S2UR + UMOV + UIADD3 + ULEA) in every row
    const DiceBins bins{reinterpret_cast<unsigned*>(hist_s), C};                        // DSTAT: in the histogram area
    float* tab_s = reinterpret_cast<float*>(full + kFfStages);                          // DCOEF: ta[B][CP], tb[B][CP]
    const int lane = tid & 31;

    if (want_hist)
        for (int k = tid; k < (C + 2) * C; k += kFfThreads) hist_s[k] = 0;   // row C collects the pixels that are not counted (row C + 1: slack)
    if (ACCP) bins.zero(tid, kFfThreads);
    if (DCOEF)
        for (int k = tid; k < 2 * p.B * CP; k += kFfThreads) {
            const int which = k / (p.B * CP), r = k % (p.B * CP), bb = r / CP, cc = r % CP;
            tab_s[k] = cc < C ? __ldg((which ? p.dice.tb : p.dice.ta) + bb * C + cc) : 0.f;
        }
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < kFfStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty_bar[s], kFfThreads / 32); }
        mbar_fence_init();
    }
    __syncthreads();

    const int per_img = p.tiles_x * p.tiles_y;
    int n_tiles = 0;
    if ((int)blockIdx.x < p.total_tiles) n_tiles = (p.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;

    auto issue = [&](int n) {
        const int tile = blockIdx.x + n * gridDim.x;
        const int b = tile / per_img, r = tile % per_img;
        const int ua = -1 + (r / p.tiles_x) * G::NGT;
        const int jO = (r % p.tiles_x) * (kFfTW / S) - G::XAL;
        uint64_t* bar = &full[n % kFfStages];
        mbar_expect_tx(bar, (uint32_t)(C * G::PLANE * (int)sizeof(LT)));
        tma_load_3d(smem + (n % kFfStages) * G::STAGE_BYTES, &tmap, bar, jO, ua, b * C);
    };
    if (tid == 0)
        for (int n = 0; n < kFfStages && n < n_tiles; ++n) issue(n);

    double acc_loss = 0.0, acc_w = 0.0, acc_hard = 0.0;
    unsigned int cnt_valid = 0, cnt_hard = 0;
    const float2 l2e2 = splat(kLog2e);
    int run_key = C * C, run_cnt = 0;                               // confusion-matrix run of this thread's column (starts in the spare row)

    // ---- dice state: sum p_k of this thread's roi pixels (DSTAT), flushed as integers so the result is order-free ----
    float2 accp[(ACCP && NP > 0) ? NP : 1];
    float accpl = 0.f;
    #pragma unroll
    for (int q = 0; q < ((ACCP && NP > 0) ? NP : 1); ++q) accp[q] = make_float2(0.f, 0.f);
    int cur_b = -1, since_flush = 0;
    const float inv_denom = (DCOEF && p.ws) ? p.ws->state.inv_denom : 0.f;
    auto redux3 = [&](float v) -> unsigned long long {               // warp sum of v (<= 2^7) in 2^40 fixed point
        const unsigned long long f = __float2ull_rn(v * 1099511627776.f);
        const unsigned a = __reduce_add_sync(0xffffffffu, (unsigned)(f & 0xFFFFFull));
        const unsigned bq = __reduce_add_sync(0xffffffffu, (unsigned)((f >> 20) & 0xFFFFFull));
        const unsigned cq = __reduce_add_sync(0xffffffffu, (unsigned)(f >> 40));
        return ((unsigned long long)cq << 40) + ((unsigned long long)bq << 20) + a;
    };
    auto flush_acc = [&]() {                                         // registers -> shared bins (all lanes of every warp)
        #pragma unroll
        for (int q = 0; q < NP; ++q) {
            const unsigned long long sx = redux3(accp[q].x), sy2 = redux3(accp[q].y);
            if (lane == 0) { bins.add_sum(2 * q, sx); bins.add_sum(2 * q + 1, sy2); }
            accp[q] = make_float2(0.f, 0.f);
        }
        if (G::ODD) {
            const unsigned long long sl = redux3(accpl);
            if (lane == 0) bins.add_sum(C - 1, sl);
            accpl = 0.f;
        }
        since_flush = 0;
    };
    auto flush_bins = [&](int b) { bins.flush(p.dice.stats, b, tid, kFfThreads); };   // CTA-uniform
    int tiles_in_bins = 0;

    for (int n = 0; n < n_tiles; ++n) {
        const int tile = blockIdx.x + n * gridDim.x;
        const int b = tile / per_img, rr = tile % per_img;
        const int ua = -1 + (rr / p.tiles_x) * G::NGT;
        const int Xa = (rr % p.tiles_x) * kFfTW;
        if (ACCP) {
            if (b != cur_b || since_flush >= 8) {                    // keep the fp32 partials short (<= 128 pixels)
                flush_acc();
                // a limb word takes <= 8 warps x 16 rows adds per tile
                if (cur_b >= 0 && (b != cur_b || tiles_in_bins >= kDiceFlushAdds / 128)) { flush_bins(cur_b); tiles_in_bins = 0; }
                cur_b = b;
            }
            ++since_flush; ++tiles_in_bins;
        }
        const int X = Xa + tx;
        const bool colok = X < p.W;
        int j0, j1; float lw0, lw1;
        TP::get(min(X, p.W - 1), p.w, j0, j1, lw0, lw1);            // idle threads keep an in-box column
        const int jO = Xa / S - G::XAL;
        const LT* Sg = reinterpret_cast<const LT*>(smem + (n % kFfStages) * G::STAGE_BYTES);
        const LT* c0p = Sg + (j0 - jO);                            // column taps inside the box
        const LT* c1p = Sg + (j1 - jO);
        const float2 lw0_2 = splat(lw0), lw1_2 = splat(lw1);

        const int u_first = ua + sy * G::GPS;
        // Labels travel as one byte each (valid classes < C <= 24 here, 255 = not counted), packed 8 per register
        // pair.  Raw 64-bit loads for the NEXT group are issued a whole group ahead and only decoded when needed.
        int raw[S];
        float lsn[DICE ? S : 1], lsc[DICE ? S : 1];                  // DICE: the stored lse travels with the labels
        // A thread's groups are consecutive, so its label rows are too: one running pointer per tile instead of rebuilding the
        // 64-bit pixel index per group and row (fetch_labels is always called with the group after the previous call's)
        const int lab_shift = p.label_u8 ? 0 : 3;
        const long long lab_stride = (long long)p.W << lab_shift;
        long long pxn = ((long long)b * p.H + TP::first(u_first)) * p.W + X;           // pixel index of the next row to fetch
        const char* labp = reinterpret_cast<const char*>(p.labels) + (pxn << lab_shift);
        auto fetch_labels = [&](int u) {
            #pragma unroll
            for (int r = 0; r < S; ++r) raw[r] = -1;
            if (u <= p.h - 1 && colok) {
                const int cnt = TP::count(u, p.h);
                #pragma unroll
                for (int r = 0; r < S; ++r)
                    if (r < cnt) {
                        const char* q = labp + r * lab_stride;
                        raw[r] = p.label_u8 ? (int)__ldg(reinterpret_cast<const unsigned char*>(q)) : __ldg(reinterpret_cast<const int*>(q));
                        if (DICE) lsn[r] = __ldg(p.dice.lse_in + pxn + (long long)r * p.W);
                    }
                labp += cnt * lab_stride;
                if (DICE) pxn += (long long)cnt * p.W;
            }
        };
        using LabPack = typename FfLabPack<S>::type;
        auto pack_labels = [&]() -> LabPack {
            LabPack pk = 0;
            #pragma unroll
            for (int r = 0; r < S; ++r) {
                if (DICE) lsc[r] = lsn[r];
                const int y = raw[r];
                const unsigned v = (y != p.ignore && y >= 0 && y < C) ? (unsigned)y : 255u;
                pk |= (LabPack)v << (8 * r);
            }
            return pk;
        };
        fetch_labels(u_first);

        mbar_wait(&full[n % kFfStages], (uint32_t)((n / kFfStages) & 1));

        float2 ta[NP > 0 ? NP : 1], tb[NP > 0 ? NP : 1];
        float tal = 0.f, tbl = 0.f;                                  // odd tail channel C-1
        auto hrow = [&](int row, float2 (&t)[NP > 0 ? NP : 1], float& tl) {   // horizontal lerps of one box row
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

        float tile_loss = 0.f, tile_w = 0.f, tile_hard = 0.f;

        // one row group: `top`/`bot` hold the horizontal lerps of source rows i0 / i1
        auto do_group = [&](int u, const float2 (&top)[NP > 0 ? NP : 1], float topl,
                            const float2 (&bot)[NP > 0 ? NP : 1], float botl, typename FfLabPack<S>::type labs) {
            const int i0 = max(u, 0), i1 = min(i0 + 1, p.h - 1);
            const int cnt = TP::count(u, p.h);
            const int o0 = (i0 - ua) * G::BW, o1 = (i1 - ua) * G::BW;
            const int cj0 = j0 - jO, cj1 = j1 - jO;
            long long px = ((long long)b * p.H + TP::first(u)) * p.W + X;
            float lh1 = u < 0 ? 0.0f : 0.5f / S;                     // (r + 0.5) / S, stepped by 1/S per row (exact in fp32)
            const float lh1_step = u < 0 ? 0.0f : 1.0f / S;
            #pragma unroll 1
            for (int r = 0; r < cnt; ++r, px += p.W, labs >>= 8, lh1 += lh1_step) {
                const float lh0 = 1.0f - lh1;
                const float2 lh0_2 = splat(lh0), lh1_2 = splat(lh1);
                float2 z[NP > 0 ? NP : 1];
                float zl = 0.f;
                #pragma unroll
                for (int q = 0; q < NP; ++q) z[q] = rnd2<LT>(lerp2_pinned(top[q], lh0_2, bot[q], lh1_2));
                if (G::ODD) zl = Logit<LT>::rnd(lerp_pinned(topl, lh0, botl, lh1));
                // max as a 3-input tree (FMNMX3), then the lowest index that attains it (first-max rule of torch.argmax)
                // Without the softmax (whose sum turns NaN by itself) the maximum is taken NaN-propagating (FMNMX3.NAN): a NaN
                // among the channels then shows up in m and sends the pixel to the exact path below, as do +-0 and +-Inf maxima
                float m = G::ODD ? zl : z[NP - 1].y;
                #pragma unroll
                for (int q = 0; q < NP; ++q) m = SOFTMAX ? fmaxf(m, fmaxf(z[q].x, z[q].y)) : max3_nan(m, z[q].x, z[q].y);
                // First index attaining the maximum without a compare/select chain (2 ALU-pipe ops per channel, the pipe this
                // kernel saturates): d_k = z_k - m is +0 exactly at the maximum and has its sign bit set elsewhere, so one funnel
                // shift per channel collects the "below the maximum" bits and the index is the lowest clear bit.  +-0 maxima
                // (-0 == +0 for torch) and NaN / Inf take the exact slow path below.
                int idx = 0;
                if (NEED_IDX) {
                    const float2 nm = splat(-m);
                    unsigned bits_lo = 0u, bits_hi = 0u;
                    constexpr int CH = C / 2;
                    if (G::ODD) bits_hi = __funnelshift_l(__float_as_uint(zl - m), bits_hi, 1);
                    #pragma unroll
                    for (int q = NP - 1; q >= 0; --q) {
                        const float2 d = __fadd2_rn(z[q], nm);
                        if (2 * q + 1 >= CH) bits_hi = __funnelshift_l(__float_as_uint(d.y), bits_hi, 1);
                        else bits_lo = __funnelshift_l(__float_as_uint(d.y), bits_lo, 1);
                        if (2 * q >= CH) bits_hi = __funnelshift_l(__float_as_uint(d.x), bits_hi, 1);
                        else bits_lo = __funnelshift_l(__float_as_uint(d.x), bits_lo, 1);
                    }
                    idx = __ffs(~((bits_hi << CH) | bits_lo)) - 1;
                }
                float ssum = 1.f;
                const float m2 = __fmul_rn(m, kLog2e);
                if (SOFTMAX) {
                    const float2 nm2 = splat(-m2);
                    float2 a2 = make_float2(0.f, 0.f), b2 = make_float2(0.f, 0.f);
                    #pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        const float2 e = __ffma2_rn(z[q], l2e2, nm2);
                        const float2 ex = make_float2(ex2_approx(e.x), ex2_approx(e.y));
                        if (q & 1) b2 = __fadd2_rn(b2, ex); else a2 = __fadd2_rn(a2, ex);
                        if (CSTAT) z[q] = ex;                        // the exponentials are needed again once the sum is known
                    }
                    a2 = __fadd2_rn(a2, b2);
                    ssum = a2.x + a2.y;
                    if (G::ODD) {
                        const float exl = ex2_approx(__fmaf_rn(zl, kLog2e, -m2));
                        ssum += exl;
                        if (CSTAT) zl = exl;
                    }
                }
                const bool weird = SOFTMAX ? ((ssum != ssum) || m == 0.f) : !(fabsf(m) > 0.f && fabsf(m) < INFINITY);
                if (weird && NEED_IDX) {                              // torch: NaN beats everything, first NaN wins; -0 == +0
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
                const int lab = (int)(labs & 0xffu);
                const bool valid = lab != 255;
                if (SOFTMAX && colok) {
                    const float corr = __fmaf_rn(m, kLog2e, -m2);          // exact residual of rn(m*log2e)
                    const float lg = (lg2_approx(ssum) - corr) * kLn2;     // ssum >= 1: the max term contributes 2^0
                    float l = 0.f;
                    if (valid) {
                        // label logit: same ops, same bits as the z[] entry of that channel
                        const LT* qk = Sg + lab * G::PLANE;
                        const float t0 = lerp_pinned(Logit<LT>::ld(qk + o0 + cj0), lw0, Logit<LT>::ld(qk + o0 + cj1), lw1);
                        const float t1 = lerp_pinned(Logit<LT>::ld(qk + o1 + cj0), lw0, Logit<LT>::ld(qk + o1 + cj1), lw1);
                        const float zy = Logit<LT>::rnd(lerp_pinned(t0, lh0, t1, lh1));
                        const float wy = cw_s[lab];
                        l = wy * (lg - (zy - m));
                        tile_loss += l; tile_w += wy; ++cnt_valid;
                        if (l > p.thresh) { tile_hard += l; ++cnt_hard; }
                    }
                    if (p.loss_px) p.loss_px[px] = l;
                    if (p.lse_px) p.lse_px[px] = m + lg;
                }
                if (CSTAT) {                                         // dice sums: p_k = e_k / sum over the roi pixels
                    const bool roi = colok && valid;
                    const float inv = roi ? __fdividef(1.0f, ssum) : 0.f;
                    const float2 inv2 = splat(inv);
                    #pragma unroll
                    for (int q = 0; q < NP; ++q) accp[q] = __ffma2_rn(z[q], inv2, accp[q]);
                    if (G::ODD) accpl = fmaf(zl, inv, accpl);
                    float py = 0.f;
                    if (roi) {
                        const LT* qk = Sg + lab * G::PLANE;
                        const float t0 = lerp_pinned(Logit<LT>::ld(qk + o0 + cj0), lw0, Logit<LT>::ld(qk + o0 + cj1), lw1);
                        const float t1 = lerp_pinned(Logit<LT>::ld(qk + o1 + cj0), lw0, Logit<LT>::ld(qk + o1 + cj1), lw1);
                        const float zy = Logit<LT>::rnd(lerp_pinned(t0, lh0, t1, lh1));
                        py = ex2_approx(__fmaf_rn(zy, kLog2e, -m2)) * inv;
                    }
                    const int key = roi ? lab : -1;
                    const unsigned peers = __match_any_sync(0xffffffffu, key);
                    const unsigned sum = __reduce_add_sync(peers, roi ? __float2uint_rn(py * 67108864.f) : 0u);
                    if (roi && lane == __ffs(peers) - 1) bins.add_inter(lab, sum, __popc(peers));
                }
                if (want_hist) {
                    // (GT, prediction) cell of this pixel.  A column's consecutive pixels mostly fall in the same cell, so the
                    // count is kept as a per-thread run and added to the shared matrix only when the cell changes
                    // (cheaper than a per-pixel __match_any_sync vote: no MATCH / FLO / POPC, and the adds are rarer)
                    // not-counted pixels (label byte 255: ignored, out of range, outside the image) go to the spare row C of
                    // the shared matrix, so the run needs no validity select and no "is there a run" guard
                    const int key = min(lab, C) * C + idx;
                    const bool chg = key != run_key;                 // straight-line: one predicated shared-memory red, no branch
                    red_shared_add_if(hist_sa + 4u * (uint32_t)run_key, run_cnt, chg);
                    run_cnt = chg ? 1 : run_cnt + 1;
                    run_key = key;
                }
            }
        };

        // dice walks: p_k = 2^(z_k log2e - rn(lse log2e)) * pcf from the stored lse; no max / argmax / histogram
        auto do_group_dice = [&](int u, const float2 (&top)[NP > 0 ? NP : 1], float topl,
                                 const float2 (&bot)[NP > 0 ? NP : 1], float botl, typename FfLabPack<S>::type labs) {
            const int i0 = max(u, 0), i1 = min(i0 + 1, p.h - 1);
            const int cnt = TP::count(u, p.h);
            const int o0 = (i0 - ua) * G::BW, o1 = (i1 - ua) * G::BW;
            const int cj0 = j0 - jO, cj1 = j1 - jO;
            long long px = ((long long)b * p.H + TP::first(u)) * p.W + X;
            #pragma unroll
            for (int r = 0; r < S; ++r) {
                if (r < cnt) {                                       // warp-uniform
                    const float lh1 = u < 0 ? 0.0f : ((float)r + 0.5f) * (1.0f / S);
                    const float lh0 = 1.0f - lh1;
                    const int lab = (int)((labs >> (8 * r)) & 0xffu);
                    const bool valid = lab != 255;                   // also false outside the image (raw = -1)
                    const float lse = valid ? lsc[r] : 0.f;
                    const float lse2 = __fmul_rn(lse, kLog2e);
                    const float pcf = valid ? __fmaf_rn(-__fmaf_rn(lse, kLog2e, -lse2), kLn2, 1.0f) : 0.f;
                    const float2 lh0_2 = splat(lh0), lh1_2 = splat(lh1), nl2 = splat(-lse2), pcf2 = splat(pcf);
                    float2 g2 = make_float2(0.f, 0.f);
                    if (__any_sync(0xffffffffu, valid)) {
                        #pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            const float2 z = rnd2<LT>(lerp2_pinned(top[q], lh0_2, bot[q], lh1_2));
                            const float2 e = __ffma2_rn(z, l2e2, nl2);
                            const float2 ex = make_float2(ex2_approx(e.x), ex2_approx(e.y));
                            if (DSTAT) accp[q] = __ffma2_rn(ex, pcf2, accp[q]);
                            else g2 = __ffma2_rn(ex, *reinterpret_cast<const float2*>(tab_s + (p.B + b) * CP + 2 * q), g2);
                        }
                        if (G::ODD) {
                            const float zl = Logit<LT>::rnd(lerp_pinned(topl, lh0, botl, lh1));
                            const float exl = ex2_approx(__fmaf_rn(zl, kLog2e, -lse2));
                            if (DSTAT) accpl = fmaf(exl, pcf, accpl);
                            else g2.x = fmaf(exl, tab_s[(p.B + b) * CP + C - 1], g2.x);
                        }
                    }
                    float py = 0.f;
                    if (valid) {                                     // label logit: same ops as the z entry of that channel
                        const LT* qk = Sg + lab * G::PLANE;
                        const float t0 = lerp_pinned(Logit<LT>::ld(qk + o0 + cj0), lw0, Logit<LT>::ld(qk + o0 + cj1), lw1);
                        const float t1 = lerp_pinned(Logit<LT>::ld(qk + o1 + cj0), lw0, Logit<LT>::ld(qk + o1 + cj1), lw1);
                        const float zy = Logit<LT>::rnd(lerp_pinned(t0, lh0, t1, lh1));
                        py = ex2_approx(__fmaf_rn(zy, kLog2e, -lse2)) * pcf;
                    }
                    if (DSTAT) {
                        const int key = valid ? lab : -1;
                        const unsigned peers = __match_any_sync(0xffffffffu, key);
                        const unsigned sum = __reduce_add_sync(peers, valid ? __float2uint_rn(py * 67108864.f) : 0u);
                        if (valid && lane == __ffs(peers) - 1) bins.add_inter(lab, sum, __popc(peers));
                    } else if (colok) {
                        float ca = 0.f, cb = 0.f;
                        if (valid) {
                            cb = fmaf(tab_s[b * CP + lab], py, inv_denom * cw_s[lab]);
                            ca = cb - (g2.x + g2.y) * pcf;
                        }
                        p.dice.coef_a[px + (long long)r * p.W] = ca;
                        p.dice.coef_b[px + (long long)r * p.W] = cb;
                    }
                }
            }
        };
        auto group = [&](int u, const float2 (&top)[NP > 0 ? NP : 1], float topl,
                         const float2 (&bot)[NP > 0 ? NP : 1], float botl, typename FfLabPack<S>::type labs) {
            if (DICE) do_group_dice(u, top, topl, bot, botl, labs); else do_group(u, top, topl, bot, botl, labs);
        };

        // strip walk, two groups per iteration so the bottom row of one group becomes the top row of the next
        // by swapping the roles of ta/tb (no register copies)
        int u = u_first;
        const int u_end = min(u_first + G::GPS, p.h);                // exclusive; uniform per warp (a warp shares sy)
        if (u < u_end) {
            const int i0 = max(u, 0), i1 = min(i0 + 1, p.h - 1);
            hrow(i0 - ua, ta, tal);
            hrow(i1 - ua, tb, tbl);
            if (u < 0) {                                             // rim group u = -1 shares (row 0, row 1) with u = 0
                const LabPack labs = pack_labels();
                fetch_labels(u + 1);
                group(u, ta, tal, tb, tbl, labs);
                ++u;
            }
        }
        while (u < u_end) {
            {
                const LabPack labs = pack_labels();
                fetch_labels(u + 1 < u_end ? u + 1 : p.h);
                group(u, ta, tal, tb, tbl, labs);
                ++u;
            }
            if (u >= u_end) break;
            {
                const int i1 = min(u + 1, p.h - 1);
                hrow(i1 - ua, ta, tal);                              // ta: new bottom row, tb: top row
                const LabPack labs = pack_labels();
                fetch_labels(u + 1 < u_end ? u + 1 : p.h);
                group(u, tb, tbl, ta, tal, labs);
                ++u;
            }
            if (u >= u_end) break;
            {
                const int i1 = min(u + 1, p.h - 1);
                hrow(i1 - ua, tb, tbl);                              // roles back: ta top, tb bottom
            }
        }
        acc_loss += (double)tile_loss; acc_w += (double)tile_w; acc_hard += (double)tile_hard;

        // Release the stage warp by warp instead of a CTA-wide barrier per tile (measured: 4.5 % of the stall samples): only
        // the producer thread waits until all 8 warps have left the tile before its TMA overwrites the stage; the other
        // warps run ahead into the next tile, whose stage landed a tile ago.
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[n % kFfStages]);
        if (tid == 0 && n + kFfStages < n_tiles) {
            mbar_wait(&empty_bar[n % kFfStages], (uint32_t)((n / kFfStages) & 1));
            issue(n + kFfStages);
        }
    }

    if (ACCP) {
        flush_acc();
        if (cur_b >= 0) flush_bins(cur_b);
    }
    if (want_hist) {
        atomicAdd(&hist_s[run_key], run_cnt);
        __syncthreads();
        for (int k = tid; k < C * C; k += kFfThreads) {
            const int v = hist_s[k];
            if (v) atomicAdd(p.hist + k, (unsigned long long)v);
        }
        if (MODE == FF_ARGMAX && p.fold) {
            // Per-call bookkeeping of Metrics / ConfusionMatrix in the LAST CTA instead of a second launch.  The ticket lives in
            // bits 48..63 of hist[0] (a call counts < 2^40 pixels, the grid has <= 1024 CTAs), so no extra caller memory is
            // needed and the scratch matrix is left zeroed as sf_accumulate_counts leaves it.
            __shared__ bool last_cta;
            __threadfence();                                         // this CTA's counts before its ticket
            __syncthreads();
            if (tid == 0) last_cta = (atomicAdd(p.hist, 1ull << 48) >> 48) == (unsigned long long)(gridDim.x - 1);
            __syncthreads();
            if (last_cta) {
                __threadfence();
                for (int k = tid; k < C * C; k += kFfThreads) {
                    unsigned long long c = __ldcg(p.hist + k);
                    if (k == 0) c &= (1ull << 48) - 1ull;
                    if (p.fold_f32) p.fold_f32[k] = __fadd_rn(p.fold_f32[k], __ull2float_rn(c));   // torch: float32 += int64
                    if (p.fold_i64) p.fold_i64[k] += (long long)c;
                    if (p.fold_mat) p.fold_mat[k] += (long long)c;
                    p.hist[k] = 0ull;
                }
            }
        }
    }
    if (SOFTMAX) {
        const double r0d = block_sum<kFfThreads>(acc_loss, red_scratch);
        const double r1d = block_sum<kFfThreads>(acc_w, red_scratch);
        const double r2d = block_sum<kFfThreads>(acc_hard, red_scratch);
        const double r3d = block_sum<kFfThreads>((double)cnt_valid, red_scratch);
        const double r4d = block_sum<kFfThreads>((double)cnt_hard, red_scratch);
        if (tid == 0) {
            double* o = p.ws->partials[blockIdx.x];
            o[0] = r0d; o[1] = r1d; o[2] = r2d; o[3] = r3d; o[4] = r4d;
            if (blockIdx.x == 0) p.ws->n_partials = gridDim.x;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
template <typename LT, int C, int S, int MODE>
static int ff_launch(const FwdArgs& a) {
    using G = FfGeom<LT, C, S>;
    FfParams p{};
    p.labels = a.labels; p.class_w = a.class_w; p.loss_px = a.loss_px; p.lse_px = a.lse_px; p.hist = a.hist;
    p.ws = reinterpret_cast<Workspace*>(a.ws);
    p.B = a.B; p.h = a.h; p.w = a.w; p.H = a.H; p.W = a.W; p.ignore = a.ignore; p.label_u8 = a.ydtype == SF_U8;
    p.thresh = a.thresh; p.dice = a.dice;
    p.fold_f32 = a.fold_f32; p.fold_i64 = a.fold_i64; p.fold_mat = a.fold_mat;
    p.fold = MODE == FF_ARGMAX && a.folded_out != nullptr && a.hist != nullptr;
    p.tiles_x = (a.W + kFfTW - 1) / kFfTW;
    p.tiles_y = (a.h + 1 + G::NGT - 1) / G::NGT;
    const long long total = (long long)a.B * p.tiles_x * p.tiles_y;
    if (total >= (1LL << 31)) return SF_ERR_SHAPE;
    p.total_tiles = (int)total;
    CUtensorMap tmap{};
    const int rc = make_logits_tmap(&tmap, a.logits, a.ldtype, a.B * C, a.h, a.w, C, G::BH, G::BW);
    if (rc != 0) return rc;
    size_t smem = (size_t)kFfStages * G::STAGE_BYTES + round_up((C + 2) * C * 4, 16) + kFfStages * 8 + 16;
    if (MODE == FF_DICE_COEF) {
        if ((long long)a.B * (C + 1) > 4096) return -1000;           // per-image tables live in shared memory
        smem += (size_t)2 * a.B * (C + (C & 1)) * 4;
    }
    auto k = fwd_fast_kernel<LT, C, S, MODE>;
    static SmemLimitOnce smem_once;                                  // per instantiation
    cudaError_t e = set_smem_limit_once(k, smem, smem_once);
    if (e != cudaSuccess) return (int)e;
    const char* gm = std::getenv("SF_DEBUG_CTAS_PER_SM");             // experiment knob (occupancy sensitivity)
    const int per_sm = gm ? std::max(1, atoi(gm)) : ((MODE == FF_ARGMAX && C <= 21) ? kFfArgmaxCtas : 2);
    const int grid = (int)std::min<long long>(total, std::min(kMaxCtas, per_sm * sm_count()));
    k<<<grid, kFfThreads, smem, reinterpret_cast<cudaStream_t>(a.stream)>>>(tmap, p);
    const int lrc = (int)cudaGetLastError();
    if (lrc == 0 && p.fold) *a.folded_out = 1;
    return lrc;
}

template <typename LT, int C, int S>
static int ff_mode(const FwdArgs& a) {
    if (a.dice_mode == DICE_STATS && a.softmax) return ff_launch<LT, C, S, FF_SOFTMAX_STATS>(a);   // CE forward + dice sums
    if (a.dice_mode == DICE_STATS) return ff_launch<LT, C, S, FF_DICE_STATS>(a);
    if (a.dice_mode == DICE_COEF) return ff_launch<LT, C, S, FF_DICE_COEF>(a);
    if (a.softmax && !a.hist) return ff_launch<LT, C, S, FF_SOFTMAX_NOHIST>(a);
    return a.softmax ? ff_launch<LT, C, S, FF_SOFTMAX>(a) : ff_launch<LT, C, S, FF_ARGMAX>(a);
}
template <typename LT, int C>
static int ff_scale(const FwdArgs& a, int s) {
    switch (s) {
        case 4: return ff_mode<LT, C, 4>(a);
        case 8: return ff_mode<LT, C, 8>(a);
        default: return -1000;
    }
}
template <typename LT>
static int ff_classes(const FwdArgs& a, int s) {
    switch (a.C) {
        case 19: return ff_scale<LT, 19>(a, s);
        case 21: return ff_scale<LT, 21>(a, s);
        case 9: return ff_scale<LT, 9>(a, s);                        // Synapse (train_gpu.py:48-49)
        case 2: return ff_scale<LT, 2>(a, s);                        // Kvasir
        default: return -1000;
    }
}

// Returns -1000 when the shape is not covered (the caller then takes the generic kernel).
int launch_forward_fast(const FwdArgs& a) {
    if (a.force_no_tma || a.h <= 0 || a.w <= 0 || a.H % a.h || a.W % a.w) return -1000;
    const int s = a.H / a.h;
    if (s != a.W / a.w) return -1000;
    const int esz = a.ldtype == SF_F32 ? 4 : 2;
    if (((size_t)a.w * esz) % 16 || ((uintptr_t)a.logits % 16)) return -1000;
    // the box must not be wider/taller than the tensor (see sf_forward.cu)
    const int xal = 16 / esz;
    if (a.w < round_up(xal + kFfTW / s + 1, xal) || a.h < 17) return -1000;
    if (a.ldtype == SF_F32) return ff_classes<float>(a, s);
    if (a.ldtype == SF_BF16) return ff_classes<__nv_bfloat16>(a, s);
    return ff_classes<__half>(a, s);
}

}  // namespace sf
