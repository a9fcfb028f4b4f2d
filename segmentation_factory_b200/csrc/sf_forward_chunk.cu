// sf_forward_chunk.cu -- integer-scale fused forward for MANY classes (ADE20K 150/151, COCO-Stuff 171/172).
//
// Same contract and arithmetic as sf_forward.cu / sf_forward_fast.cu.  With C > 24 the classes no longer fit one
// thread's registers, so they stream through in chunks of 16 (8 fp32x2 pairs) with an online softmax per pixel:
// running max m, running sum s of 2^(z*log2e - rn(m*log2e)) (rescaled by the difference of the ROUNDED exponents
// when m grows, so every term stays relative to one rn(m*log2e)), running first-max index, and the label logit.
// Work item = (tile, chunk); tile = 64 label columns x 8 row groups (512 threads, one row group of S rows per
// thread, so the per-pixel state is 4*S registers); box [16][9][BW] per item through a 3-stage TMA/mbarrier ring;
// persistent CTAs.  Geometry is constexpr (sf_fast.cuh), arithmetic is packed fp32x2, the max is an FMNMX3 tree and
// the in-chunk index is only recovered for pixels whose running max actually moved.
// The C x C int32 histogram (90-118 KB for 150-172 classes) shares the SM's shared memory with the ring, hence one
// 512-thread CTA per SM when the matrix is requested.
#include "sf_fast.cuh"
#include "sf_dice_bins.cuh"
#include "sf_launch.h"

namespace sf {

constexpr int kFcTW = 64, kFcStrips = 8, kFcThreads = kFcTW * kFcStrips, kFcStages = 3, kFcNCH = 16, kFcNP = kFcNCH / 2;

struct FcParams {
    const void* labels;
    const float* class_w;
    float* loss_px;
    float* lse_px;
    unsigned long long* hist;
    Workspace* ws;
    int B, C, h, w, H, W, ignore, label_u8;
    float thresh;
    int tiles_x, tiles_y, total_tiles, n_chunks, hist_bytes;
    DiceArgs dice;                                                  // MODE >= 2 (sf_dice.cu)
};

enum { FC_ARGMAX = 0, FC_SOFTMAX = 1, FC_DICE_STATS = 2, FC_DICE_COEF = 3, FC_SOFTMAX_NOHIST = 4 };

template <typename LT, int S>
struct FcGeom {
    static constexpr int NGT = kFcStrips;                           // row groups per tile (one per thread row)
    static constexpr int BH = NGT + 1;
    static constexpr int XAL = 16 / (int)sizeof(LT);
    static constexpr int BW = round_up(XAL + kFcTW / S + 1, XAL);
    static constexpr int PLANE = BH * BW;
    static constexpr int STAGE_BYTES = round_up(kFcNCH * PLANE * (int)sizeof(LT), 128);
};

// MODE: FC_ARGMAX / FC_SOFTMAX are the forward proper; FC_DICE_STATS / FC_DICE_COEF are the dice walks of sf_dice.cu
// (softmax from the stored lse: no running max, no argmax, no histogram).
template <typename LT, int S, int MODE>
__global__ void __launch_bounds__(kFcThreads, 1)
fwd_chunk_kernel(const __grid_constant__ CUtensorMap tmap, const FcParams p) {
    using G = FcGeom<LT, S>;
    using TP = ScaleTap<S>;
    constexpr int NP = kFcNP;
    constexpr bool NOHIST = MODE == FC_SOFTMAX_NOHIST;              // loss only (training without metrics): no argmax index
    constexpr bool SOFTMAX = MODE == FC_SOFTMAX || NOHIST;
    constexpr bool DSTAT = MODE == FC_DICE_STATS, DCOEF = MODE == FC_DICE_COEF, DICE = DSTAT || DCOEF;
    extern __shared__ __align__(128) unsigned char smem[];
    int* hist_s = reinterpret_cast<int*>(smem + kFcStages * G::STAGE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kFcStages * G::STAGE_BYTES + p.hist_bytes);
    __shared__ double red_scratch[kFcThreads / 32];
    __shared__ float cw_s[256];
    __shared__ uint64_t empty_bar[kFcStages];                       // consumer -> producer: "every warp is done with this stage"

    const int tid = threadIdx.x;
    const int tx = tid % kFcTW, sy = tid / kFcTW;
    const bool want_hist = !DICE && !NOHIST && p.hist != nullptr;
    const unsigned lane_lt = (1u << (tid & 31)) - 1u;
    const int C = p.C;
    const int CP = C + (C & 1);                                     // even pitch of the per-image dice tables
    const DiceBins bins{reinterpret_cast<unsigned*>(hist_s), C};                        // DSTAT
    float* tab_s = reinterpret_cast<float*>(full + kFcStages);                          // DCOEF: ta[B][CP], tb[B][CP + 16]

    if (want_hist)
        for (int k = tid; k < C * C; k += kFcThreads) hist_s[k] = 0;
    if (DSTAT) bins.zero(tid, kFcThreads);
    const int TBP = CP + kFcNCH;                                    // tb rows are padded so a tail chunk reads zeros
    if (DCOEF) {
        for (int k = tid; k < p.B * CP; k += kFcThreads) tab_s[k] = (k % CP) < C ? __ldg(p.dice.ta + (k / CP) * C + k % CP) : 0.f;
        for (int k = tid; k < p.B * TBP; k += kFcThreads) tab_s[p.B * CP + k] = (k % TBP) < C ? __ldg(p.dice.tb + (k / TBP) * C + k % TBP) : 0.f;
    }
    const float* tb_s = tab_s + p.B * CP;
    const float inv_denom = (DCOEF && p.ws) ? p.ws->state.inv_denom : 0.f;
    int cur_b = -1;
    auto flush_bins = [&](int bb) { bins.flush(p.dice.stats, bb, tid, kFcThreads); };   // CTA-uniform
    int tiles_in_bins = 0;
    if (tid < 256) cw_s[tid] = (p.class_w && tid < C) ? __ldg(p.class_w + tid) : 1.0f;
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < kFcStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty_bar[s], kFcThreads / 32); }
        mbar_fence_init();
    }
    __syncthreads();

    const int per_img = p.tiles_x * p.tiles_y;
    int n_tiles = 0;
    if ((int)blockIdx.x < p.total_tiles) n_tiles = (p.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;
    const int n_items = n_tiles * p.n_chunks;

    auto issue = [&](int n) {
        const int tile = blockIdx.x + (n / p.n_chunks) * gridDim.x;
        const int chunk = n % p.n_chunks;
        const int b = tile / per_img, r = tile % per_img;
        const int ua = -1 + (r / p.tiles_x) * G::NGT;
        const int jO = (r % p.tiles_x) * (kFcTW / S) - G::XAL;
        uint64_t* bar = &full[n % kFcStages];
        mbar_expect_tx(bar, (uint32_t)(kFcNCH * G::PLANE * (int)sizeof(LT)));
        tma_load_3d(smem + (n % kFcStages) * G::STAGE_BYTES, &tmap, bar, jO, ua, b * C + chunk * kFcNCH);
    };
    if (tid == 0)
        for (int n = 0; n < kFcStages && n < n_items; ++n) issue(n);

    double acc_loss = 0.0, acc_w = 0.0, acc_hard = 0.0;
    unsigned int cnt_valid = 0, cnt_hard = 0;
    const float2 l2e2 = splat(kLog2e);
    int run_key = -1, run_cnt = 0;

    // per-(tile) thread state
    int b = 0, ua = 0, X = 0, j0 = 0, j1 = 0, jO = 0, u = 0, cnt = 0, i0 = 0, i1 = 0;
    float lw0 = 0.f, lw1 = 0.f;
    bool colok = false;
    unsigned long long labs = 0;
    unsigned nanlock = 0;
    float m[S], s[S], zy[S];
    int idx[S];
    float lse2[DICE ? S : 1], pcf[DICE ? S : 1], Gs[DICE ? S : 1], py[DICE ? S : 1];

    // (tile, chunk) and the ring slot advance as counters: a division by the runtime chunk count per item and thread was
    // ~20 instructions (IABS / MUFU.RCP / ...) in a loop body of ~200
    int chunk = -1, tile_i = -1, slot = -1, phase = 1;
    for (int n = 0; n < n_items; ++n) {
        if (++chunk == p.n_chunks || n == 0) { chunk = 0; ++tile_i; }
        if (++slot == kFcStages || n == 0) { slot = 0; phase ^= 1; }
        const LT* Sg = reinterpret_cast<const LT*>(smem + slot * G::STAGE_BYTES);

        if (chunk == 0) {
            const int tile = blockIdx.x + tile_i * gridDim.x;
            b = tile / per_img;
            if (DSTAT) {                                             // a limb word takes <= 16 warps x S rows adds per tile
                if (cur_b >= 0 && (b != cur_b || tiles_in_bins >= kDiceFlushAdds / (16 * S))) { flush_bins(cur_b); tiles_in_bins = 0; }
                cur_b = b; ++tiles_in_bins;
            }
            const int rr = tile % per_img;
            ua = -1 + (rr / p.tiles_x) * G::NGT;
            const int Xa = (rr % p.tiles_x) * kFcTW;
            X = Xa + tx;
            colok = X < p.W;
            TP::get(min(X, p.W - 1), p.w, j0, j1, lw0, lw1);
            jO = Xa / S - G::XAL;
            u = ua + sy;
            cnt = (u <= p.h - 1 && colok) ? TP::count(u, p.h) : 0;
            i0 = max(u, 0); i1 = min(i0 + 1, p.h - 1);
            labs = 0; nanlock = 0;
            #pragma unroll
            for (int r = 0; r < S; ++r) { m[r] = -INFINITY; s[r] = 0.f; zy[r] = 0.f; idx[r] = 0; }
            if (cnt > 0) {
                const long long base = ((long long)b * p.H + TP::first(u)) * p.W + X;
                #pragma unroll
                for (int r = 0; r < S; ++r) {
                    unsigned v = 255u;
                    if (r < cnt) {
                        const long long y = ld_label(p.labels, p.label_u8, base + (long long)r * p.W);
                        if (y != p.ignore && y >= 0 && y < C) v = (unsigned)y;
                    }
                    labs |= (unsigned long long)v << (8 * r);
                    if (DICE) {
                        lse2[r] = 0.f; pcf[r] = 0.f; Gs[r] = 0.f; py[r] = 0.f;
                        if (v != 255u) {
                            const float lse = __ldg(p.dice.lse_in + base + (long long)r * p.W);
                            lse2[r] = __fmul_rn(lse, kLog2e);
                            pcf[r] = __fmaf_rn(-__fmaf_rn(lse, kLog2e, -lse2[r]), kLn2, 1.0f);
                        }
                    }
                }
            }
        }

        mbar_wait(&full[slot], (uint32_t)phase);

        if (DICE) {
            // warp-uniform guard (a warp shares sy): the reductions below need every lane, idle columns carry no valid pixel
            const int rows = (u <= p.h - 1) ? TP::count(u, p.h) : 0;
            if (rows > 0) {
                const int c0 = chunk * kFcNCH;
                const int nvalid = min(kFcNCH, C - c0);
                const LT* q00 = Sg + (i0 - ua) * G::BW + (j0 - jO);
                const LT* q01 = Sg + (i0 - ua) * G::BW + (j1 - jO);
                const LT* q10 = Sg + (i1 - ua) * G::BW + (j0 - jO);
                const LT* q11 = Sg + (i1 - ua) * G::BW + (j1 - jO);
                const float2 lw0_2 = splat(lw0), lw1_2 = splat(lw1);
                float2 top[NP], bot[NP], accp[NP];
                #pragma unroll
                for (int q = 0; q < NP; ++q) {
                    const float2 a = make_float2(Logit<LT>::ld(q00 + (2 * q) * G::PLANE), Logit<LT>::ld(q00 + (2 * q + 1) * G::PLANE));
                    const float2 c = make_float2(Logit<LT>::ld(q01 + (2 * q) * G::PLANE), Logit<LT>::ld(q01 + (2 * q + 1) * G::PLANE));
                    top[q] = lerp2_pinned(a, lw0_2, c, lw1_2);
                    const float2 d = make_float2(Logit<LT>::ld(q10 + (2 * q) * G::PLANE), Logit<LT>::ld(q10 + (2 * q + 1) * G::PLANE));
                    const float2 e = make_float2(Logit<LT>::ld(q11 + (2 * q) * G::PLANE), Logit<LT>::ld(q11 + (2 * q + 1) * G::PLANE));
                    bot[q] = lerp2_pinned(d, lw0_2, e, lw1_2);
                    accp[q] = make_float2(0.f, 0.f);
                }
                #pragma unroll
                for (int r = 0; r < S; ++r) {
                    if (r < rows) {
                        const int lab = (int)((labs >> (8 * r)) & 0xffu);
                        const bool valid = cnt > 0 && lab != 255;
                        const float lh1 = u < 0 ? 0.0f : ((float)r + 0.5f) * (1.0f / S);
                        const float lh0 = 1.0f - lh1;
                        const float pv = valid ? pcf[r] : 0.f;
                        const float l2 = valid ? lse2[r] : 0.f;
                        if (__any_sync(0xffffffffu, valid)) {
                            const float2 lh0_2 = splat(lh0), lh1_2 = splat(lh1), nl2 = splat(-l2), pv2 = splat(pv);
                            float2 g2 = make_float2(0.f, 0.f);
                            #pragma unroll
                            for (int q = 0; q < NP; ++q) {
                                float2 z = rnd2<LT>(lerp2_pinned(top[q], lh0_2, bot[q], lh1_2));
                                if (2 * q >= nvalid) z.x = -INFINITY;            // classes past C add nothing
                                if (2 * q + 1 >= nvalid) z.y = -INFINITY;
                                const float2 e = __ffma2_rn(z, l2e2, nl2);
                                const float2 ex = make_float2(ex2_approx(e.x), ex2_approx(e.y));
                                if (DSTAT) accp[q] = __ffma2_rn(ex, pv2, accp[q]);
                                else g2 = __ffma2_rn(ex, *reinterpret_cast<const float2*>(tb_s + b * TBP + c0 + 2 * q), g2);
                            }
                            if (DCOEF) Gs[r] = fmaf(g2.x + g2.y, pv, Gs[r]);
                        }
                        const int yk = lab - c0;
                        const bool mine = valid && yk >= 0 && yk < nvalid;     // this chunk holds the label's class
                        float pyv = 0.f;
                        if (mine) {                                              // label logit: same ops as its z[] entry
                            const float t0 = lerp_pinned(Logit<LT>::ld(q00 + yk * G::PLANE), lw0, Logit<LT>::ld(q01 + yk * G::PLANE), lw1);
                            const float t1 = lerp_pinned(Logit<LT>::ld(q10 + yk * G::PLANE), lw0, Logit<LT>::ld(q11 + yk * G::PLANE), lw1);
                            const float zyv = Logit<LT>::rnd(lerp_pinned(t0, lh0, t1, lh1));
                            pyv = ex2_approx(__fmaf_rn(zyv, kLog2e, -l2)) * pv;
                            if (DCOEF) py[r] = pyv;
                        }
                        if (DSTAT && __any_sync(0xffffffffu, mine)) {      // only the chunk that holds a lane's label class
                            const int key = mine ? lab : -1;
                            const unsigned peers = __match_any_sync(0xffffffffu, key);
                            const unsigned sum = __reduce_add_sync(peers, mine ? __float2uint_rn(pyv * 67108864.f) : 0u);
                            if (mine && (peers & lane_lt) == 0) bins.add_inter(lab, sum, __popc(peers));
                        }
                    }
                }
                if (DSTAT) {
                    // this item's partial sums (<= S pixels per lane) -> shared bins.  Transposing butterfly: each step
                    // halves the values a lane carries while summing over its partner, so 16 classes x 32 lanes take
                    // 8+4+2+1+1 shuffles (fixed tree => deterministic); lane l ends with class bits (l>>1)&15 reversed below
                    float v[kFcNCH];
                    #pragma unroll
                    for (int q = 0; q < NP; ++q) { v[2 * q] = accp[q].x; v[2 * q + 1] = accp[q].y; }
                    const int lane = tid & 31;
                    #pragma unroll
                    for (int half = kFcNCH / 2, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
                        const bool up = (lane & bit) != 0;
                        #pragma unroll
                        for (int i = 0; i < half; ++i) {
                            const float keep = up ? v[i + half] : v[i];
                            const float send = up ? v[i] : v[i + half];
                            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
                        }
                    }
                    const float tot = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
                    const int cls = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                    if ((lane & 1) == 0 && cls < nvalid) bins.add_sum(c0 + cls, __float2ull_rn(tot * 1099511627776.f));
                }
                if (DCOEF && chunk == p.n_chunks - 1 && cnt > 0) {
                    const long long px0 = ((long long)b * p.H + TP::first(u)) * p.W + X;
                    #pragma unroll
                    for (int r = 0; r < S; ++r) {
                        if (r < cnt) {
                            const int lab = (int)((labs >> (8 * r)) & 0xffu);
                            float ca = 0.f, cb = 0.f;
                            if (lab != 255) {
                                cb = fmaf(tab_s[b * CP + lab], py[r], inv_denom * cw_s[lab]);
                                ca = cb - Gs[r];
                            }
                            p.dice.coef_a[px0 + (long long)r * p.W] = ca;
                            p.dice.coef_b[px0 + (long long)r * p.W] = cb;
                        }
                    }
                }
            }
        } else
        if (cnt > 0) {
            const int c0 = chunk * kFcNCH;
            const int nvalid = min(kFcNCH, C - c0);
            const LT* q00 = Sg + (i0 - ua) * G::BW + (j0 - jO);
            const LT* q01 = Sg + (i0 - ua) * G::BW + (j1 - jO);
            const LT* q10 = Sg + (i1 - ua) * G::BW + (j0 - jO);
            const LT* q11 = Sg + (i1 - ua) * G::BW + (j1 - jO);
            const float2 lw0_2 = splat(lw0), lw1_2 = splat(lw1);
            float2 top[NP], bot[NP];
            #pragma unroll
            for (int q = 0; q < NP; ++q) {
                const float2 a = make_float2(Logit<LT>::ld(q00 + (2 * q) * G::PLANE), Logit<LT>::ld(q00 + (2 * q + 1) * G::PLANE));
                const float2 c = make_float2(Logit<LT>::ld(q01 + (2 * q) * G::PLANE), Logit<LT>::ld(q01 + (2 * q + 1) * G::PLANE));
                top[q] = lerp2_pinned(a, lw0_2, c, lw1_2);
                const float2 d = make_float2(Logit<LT>::ld(q10 + (2 * q) * G::PLANE), Logit<LT>::ld(q10 + (2 * q + 1) * G::PLANE));
                const float2 e = make_float2(Logit<LT>::ld(q11 + (2 * q) * G::PLANE), Logit<LT>::ld(q11 + (2 * q + 1) * G::PLANE));
                bot[q] = lerp2_pinned(d, lw0_2, e, lw1_2);
            }
            // NaN / Inf probe over the chunk's taps: x * 0 is NaN exactly for NaN and Inf.  Packed, no per-class predicate:
            // the classes past C of a tail chunk hold finite neighbours' data or TMA zero fill.
            float2 nan2 = make_float2(0.f, 0.f);
            const float2 zero2 = make_float2(0.f, 0.f);
            #pragma unroll
            for (int q = 0; q < NP; ++q) nan2 = __ffma2_rn(top[q], zero2, __ffma2_rn(bot[q], zero2, nan2));
            const float nanflag = nan2.x + nan2.y;
            const bool odd = nanflag != nanflag;                     // a NaN or Inf among this chunk's taps
            #pragma unroll
            for (int r = 0; r < S; ++r) {
                if (r < cnt) {
                    const float lh1 = u < 0 ? 0.0f : ((float)r + 0.5f) * (1.0f / S);
                    const float lh0 = 1.0f - lh1;
                    const float2 lh0_2 = splat(lh0), lh1_2 = splat(lh1);
                    float2 z[NP];
                    #pragma unroll
                    for (int q = 0; q < NP; ++q) z[q] = rnd2<LT>(lerp2_pinned(top[q], lh0_2, bot[q], lh1_2));
                    if (nvalid < kFcNCH) {                           // tail chunk: classes past C never win nor add
                        #pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            if (2 * q >= nvalid) z[q].x = -INFINITY;
                            if (2 * q + 1 >= nvalid) z[q].y = -INFINITY;
                        }
                    }
                    float cm = z[0].x;
                    #pragma unroll
                    for (int q = 0; q < NP; ++q) cm = fmaxf(cm, fmaxf(z[q].x, z[q].y));
                    if (cm > m[r]) {                                 // the running max moved: find its first index here
                        int ci = kFcNCH - 1;
                        #pragma unroll
                        for (int q = NP - 1; q >= 0; --q) {
                            ci = (z[q].y == cm) ? 2 * q + 1 : ci;
                            ci = (z[q].x == cm) ? 2 * q : ci;
                        }
                        if (SOFTMAX) s[r] *= ex2_approx(__fsub_rn(__fmul_rn(m[r], kLog2e), __fmul_rn(cm, kLog2e)));
                        m[r] = cm; idx[r] = c0 + ci;
                    }
                    if (odd && !((nanlock >> r) & 1u)) {             // torch: NaN beats everything, first NaN wins
                        int first = -1;
                        #pragma unroll
                        for (int q = NP - 1; q >= 0; --q) {
                            if (z[q].y != z[q].y) first = 2 * q + 1;
                            if (z[q].x != z[q].x) first = 2 * q;
                        }
                        if (first >= 0) { idx[r] = c0 + first; m[r] = INFINITY; s[r] = NAN; nanlock |= 1u << r; }
                    }
                    if (SOFTMAX) {
                        if (m[r] > -INFINITY) {
                            const float2 nm2 = splat(-__fmul_rn(m[r], kLog2e));
                            float2 a2 = make_float2(0.f, 0.f), b2 = make_float2(0.f, 0.f);
                            #pragma unroll
                            for (int q = 0; q < NP; ++q) {
                                const float2 e = __ffma2_rn(z[q], l2e2, nm2);
                                const float2 ex = make_float2(ex2_approx(e.x), ex2_approx(e.y));
                                if (q & 1) b2 = __fadd2_rn(b2, ex); else a2 = __fadd2_rn(a2, ex);
                            }
                            a2 = __fadd2_rn(a2, b2);
                            s[r] += a2.x + a2.y;
                        }
                        const int yk = (int)((labs >> (8 * r)) & 0xffu) - c0;
                        if (yk >= 0 && yk < nvalid) {                // label logit: same ops, same bits as its z[] entry
                            const float t0 = lerp_pinned(Logit<LT>::ld(q00 + yk * G::PLANE), lw0, Logit<LT>::ld(q01 + yk * G::PLANE), lw1);
                            const float t1 = lerp_pinned(Logit<LT>::ld(q10 + yk * G::PLANE), lw0, Logit<LT>::ld(q11 + yk * G::PLANE), lw1);
                            zy[r] = Logit<LT>::rnd(lerp_pinned(t0, lh0, t1, lh1));
                        }
                    }
                }
            }
        }

        if (!DICE && chunk == p.n_chunks - 1) {
            // ---- tile epilogue: per-pixel outputs, histogram, loss sums ----
            const long long px0 = ((long long)b * p.H + TP::first(min(u, p.h - 1))) * p.W + X;
            float tile_loss = 0.f, tile_w = 0.f, tile_hard = 0.f;
            const int rows = (u <= p.h - 1) ? TP::count(u, p.h) : 0;   // warp-uniform (a warp shares sy); cnt is 0 for idle columns
            #pragma unroll
            for (int r = 0; r < S; ++r) {
                if (r < rows) {
                    const int lab = (int)((labs >> (8 * r)) & 0xffu);
                    const bool valid = cnt > 0 && lab != 255;
                    if (SOFTMAX && cnt > 0) {
                        const float m2 = __fmul_rn(m[r], kLog2e);
                        const float corr = __fmaf_rn(m[r], kLog2e, -m2);
                        const float lg = (log2f(s[r]) - corr) * kLn2;
                        float l = 0.f;
                        if (valid) {
                            const float wy = cw_s[lab];
                            l = wy * (lg - (zy[r] - m[r]));
                            tile_loss += l; tile_w += wy; ++cnt_valid;
                            if (l > p.thresh) { tile_hard += l; ++cnt_hard; }
                        }
                        const long long px = px0 + (long long)r * p.W;
                        if (p.loss_px) p.loss_px[px] = l;
                        if (p.lse_px) p.lse_px[px] = m[r] + lg;
                    }
                    if (want_hist) {
                        const int key = valid ? lab * C + idx[r] : -1;   // per-thread runs (see sf_forward_fast.cu)
                        if (key != run_key) {
                            if (run_key >= 0) atomicAdd(&hist_s[run_key], run_cnt);
                            run_key = key; run_cnt = 0;
                        }
                        ++run_cnt;
                    }
                }
            }
            acc_loss += (double)tile_loss; acc_w += (double)tile_w; acc_hard += (double)tile_hard;
        }

        // Release the stage warp by warp (as sf_forward_fast.cu does) instead of a CTA-wide barrier per item -- 0.8 stall cycles
        // per issued instruction in the cfg-3 profile: only the producer thread waits until all 16 warps have left the item
        // before its TMA overwrites the stage; the other warps run ahead into the next item, whose stage landed earlier.
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty_bar[slot]);
        if (tid == 0 && n + kFcStages < n_items) {
            mbar_wait(&empty_bar[slot], (uint32_t)phase);
            issue(n + kFcStages);
        }
    }

    if (DSTAT && cur_b >= 0) flush_bins(cur_b);
    if (want_hist) {
        if (run_key >= 0) atomicAdd(&hist_s[run_key], run_cnt);
        __syncthreads();
        for (int k = tid; k < C * C; k += kFcThreads) {
            const int v = hist_s[k];
            if (v) atomicAdd(p.hist + k, (unsigned long long)v);
        }
    }
    if (SOFTMAX) {
        const double r0d = block_sum<kFcThreads>(acc_loss, red_scratch);
        const double r1d = block_sum<kFcThreads>(acc_w, red_scratch);
        const double r2d = block_sum<kFcThreads>(acc_hard, red_scratch);
        const double r3d = block_sum<kFcThreads>((double)cnt_valid, red_scratch);
        const double r4d = block_sum<kFcThreads>((double)cnt_hard, red_scratch);
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
template <typename LT, int S, int MODE>
static int fc_launch(const FwdArgs& a) {
    using G = FcGeom<LT, S>;
    if (a.w < G::BW || a.h < G::BH) return -1000;                   // the TMA box must fit inside the tensor
    FcParams p{};
    p.labels = a.labels; p.class_w = a.class_w; p.loss_px = a.loss_px; p.lse_px = a.lse_px; p.hist = a.hist;
    p.ws = reinterpret_cast<Workspace*>(a.ws);
    p.B = a.B; p.C = a.C; p.h = a.h; p.w = a.w; p.H = a.H; p.W = a.W; p.ignore = a.ignore; p.label_u8 = a.ydtype == SF_U8;
    p.thresh = a.thresh; p.dice = a.dice;
    p.tiles_x = (a.W + kFcTW - 1) / kFcTW;
    p.tiles_y = (a.h + 1 + G::NGT - 1) / G::NGT;
    p.n_chunks = (a.C + kFcNCH - 1) / kFcNCH;
    const long long total = (long long)a.B * p.tiles_x * p.tiles_y;
    if (total * p.n_chunks >= (1LL << 31)) return SF_ERR_SHAPE;
    p.total_tiles = (int)total;
    p.hist_bytes = MODE == FC_DICE_STATS ? round_up(3 * a.C * 8, 16) : (a.hist && MODE < 2) ? round_up(a.C * a.C * 4, 16) : 16;
    size_t smem = (size_t)kFcStages * G::STAGE_BYTES + p.hist_bytes + kFcStages * 8 + 16;
    if (MODE == FC_DICE_COEF) {
        const size_t tab = (size_t)a.B * (2 * (a.C + (a.C & 1)) + kFcNCH) * 4;
        if (tab > 64 * 1024) return -1000;                          // per-image tables live in shared memory
        smem += tab;
    }
    if (smem > 227 * 1024) return SF_ERR_TOO_MANY_CLASSES;
    CUtensorMap tmap{};
    const int rc = make_logits_tmap(&tmap, a.logits, a.ldtype, a.B * a.C, a.h, a.w, kFcNCH, G::BH, G::BW);
    if (rc != 0) return rc;
    auto k = fwd_chunk_kernel<LT, S, MODE>;
    static SmemLimitOnce smem_once;                                  // per instantiation
    cudaError_t e = set_smem_limit_once(k, smem, smem_once);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 1;                                                  // 512 threads x ~120 registers: one CTA per SM
    const int grid = (int)std::min<long long>(total, std::min(kMaxCtas, per_sm * sm_count()));
    k<<<grid, kFcThreads, smem, reinterpret_cast<cudaStream_t>(a.stream)>>>(tmap, p);
    return (int)cudaGetLastError();
}

template <typename LT, int S>
static int fc_mode(const FwdArgs& a) {
    if (a.dice_mode == DICE_STATS) return fc_launch<LT, S, FC_DICE_STATS>(a);
    if (a.dice_mode == DICE_COEF) return fc_launch<LT, S, FC_DICE_COEF>(a);
    if (a.softmax && !a.hist) return fc_launch<LT, S, FC_SOFTMAX_NOHIST>(a);
    return a.softmax ? fc_launch<LT, S, FC_SOFTMAX>(a) : fc_launch<LT, S, FC_ARGMAX>(a);
}
template <typename LT>
static int fc_scale(const FwdArgs& a, int s) {
    switch (s) {
        case 4: return fc_mode<LT, 4>(a);
        case 8: return fc_mode<LT, 8>(a);
        default: return -1000;
    }
}

// Returns -1000 when the shape is not covered (the caller then takes the generic kernel).
int launch_forward_chunk(const FwdArgs& a) {
    if (a.force_no_tma || a.C <= 24 || a.C > 255 || a.h <= 0 || a.w <= 0 || a.H % a.h || a.W % a.w) return -1000;
    const int s = a.H / a.h;
    if (s != a.W / a.w) return -1000;
    const int esz = a.ldtype == SF_F32 ? 4 : 2;
    if (((size_t)a.w * esz) % 16 || ((uintptr_t)a.logits % 16)) return -1000;
    if (a.ldtype == SF_F32) return fc_scale<float>(a, s);
    if (a.ldtype == SF_BF16) return fc_scale<__nv_bfloat16>(a, s);
    return fc_scale<__half>(a, s);
}

}  // namespace sf
