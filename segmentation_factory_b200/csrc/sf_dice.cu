// sf_dice.cu -- the dice term of engine.criterion (engine.py:10-15 -> util/losses.py:126-177, multiclass=True) on
// logits that stay at low resolution.  The reference materialises a [B,C,H,W] one-hot and a [B,C,H,W] softmax and
// loops B x C times in Python with a host sync each (`if sets_sum == 0`); here it is three launches:
//
//   dice_kernel<STATS>  walks every label pixel once (same tile / pass / chunk items and TMA staging as the generic
//                       forward, sf_fwd_generic.cuh), rebuilds p_c = 2^(z_c log2e - lse log2e) from the lse the CE
//                       forward stored and accumulates, per image and class over the non-ignored pixels,
//                       sum p_c, sum p_c [y == c] and count [y == c].
//   dice_finalize       D_bc = (2 inter + eps) / (sets + eps) (sets == 0 -> 2 inter), loss = 1 - mean D, plus the two
//                       backward tables ta = 2 / (BC (sets + eps)), tb = D / (BC (sets + eps)).
//   dice_kernel<COEF>   second walk: G = sum_j tb_j p_j per pixel, stores the two per-pixel scalars of the fused
//                       CE + dice gradient  d/dz_k = p_k (A + tb_k) - [k == y] Bc,
//                       Bc = w_y / sum(w) + ta_y p_y,  A = Bc - G            (sf_backward.cu consumes them).
//
// The sums are accumulated as integers (fixed point) so that they do not depend on the order in which warps and CTAs
// arrive: bitwise reproducible without a fixed-order reduction tree.  sum p_c: per-thread fp32 partial over <= 4
// pixels -> 2^40 fixed point -> two 32-bit warp reductions (REDUX) -> one 64-bit shared atomic per warp and class;
// inter: 2^26 fixed point per pixel, warp-aggregated per label with __match_any_sync; counts are exact.
#include "sf_dice_bins.cuh"
#include "sf_fwd_generic.cuh"

namespace sf {

constexpr float kFixS = 1099511627776.f;          // 2^40: sum p_c
constexpr float kFixI = 67108864.f;               // 2^26: p_y of one pixel
constexpr int   kDiceHeaderBytes = 64;

using DiceParams = DiceArgs;

template <typename LT, int NCH, int MODE>
__global__ void __launch_bounds__(kFwdThreads, 2)
dice_kernel(const __grid_constant__ CUtensorMap tmap, const FwdParams p, const DiceParams d) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* stages = smem;
    const DiceBins bins{reinterpret_cast<unsigned*>(smem + kFwdStages * p.stage_bytes), p.C};
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kFwdStages * p.stage_bytes + p.hist_bytes);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int tx = tid % kFwdTW, gy = tid / kFwdTW;

    if (MODE == DICE_STATS) bins.zero(tid, kFwdThreads);
    if (tid == 0 && p.use_tma) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < kFwdStages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    int n_my_tiles = 0;
    if ((int)blockIdx.x < p.total_tiles) n_my_tiles = (p.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;
    const int ipt = p.n_passes * p.n_chunks;
    const int n_items = n_my_tiles * ipt;

    if (p.use_tma && tid == 0)
        for (int n = 0; n < kFwdStages && n < n_items; ++n)
            issue_tma<LT, NCH>(p, &tmap, decode_item<NCH>(p, n), stages + n * p.stage_bytes, &full[n]);

    const float inv_denom = (MODE == DICE_COEF && p.ws) ? p.ws->state.inv_denom : 0.f;

    auto flush = [&](int b) { bins.flush(d.stats, b, tid, kFwdThreads); };   // CTA-uniform
    int since_flush = 0;
    int cur_b = -1;

    int X = 0, j0r = 0, j1r = 0, r0 = 0, r1 = 0, nrows = 0, Yb = 0;
    float lw0 = 0.f, lw1 = 0.f;
    float lh0[kFwdG], lh1[kFwdG], lse2[kFwdG], pcf[kFwdG], G[kFwdG], py[kFwdG];
    int lab[kFwdG];
    bool colok = false;

    for (int n = 0; n < n_items; ++n) {
        const int stage = n % kFwdStages;
        const ItemCoord c = decode_item<NCH>(p, n);
        LT* S = reinterpret_cast<LT*>(stages + stage * p.stage_bytes);

        if (c.chunk == 0) {
            if (MODE == DICE_STATS && c.pass == 0) {
                // a limb word takes <= 8 warps x 4 rows x n_passes adds per tile
                if (cur_b >= 0 && (c.b != cur_b || (since_flush += 32 * p.n_passes) >= kDiceFlushAdds)) { flush(cur_b); since_flush = 0; }
                cur_b = c.b;
            }
            X = c.Xa + tx;
            colok = X < p.ax.out;
            const Tap tc = make_tap<LT>(p.ax, colok ? X : p.ax.out - 1);
            const int jA = floor_al(src_i0(p.ax, c.Xa), p.xalign);
            j0r = tc.i0 - jA; j1r = tc.i1 - jA; lw0 = tc.l0; lw1 = tc.l1;
            const int u = c.ua + gy;
            int y0 = p.ay.out, y1 = p.ay.out;
            if (u < p.ay.in) { y0 = first_dst_of(p.ay, u); y1 = first_dst_of(p.ay, u + 1); }
            Yb = y0 + c.pass * kFwdG;
            nrows = min(kFwdG, max(0, y1 - Yb));
            const int i0 = u < 0 ? 0 : u;
            r0 = i0 - c.ua;
            r1 = r0 + ((i0 < p.ay.in - 1 && !p.ay.ident) ? 1 : 0);
            #pragma unroll
            for (int r = 0; r < kFwdG; ++r) {
                lab[r] = -1; lh0[r] = 1.f; lh1[r] = 0.f; lse2[r] = 0.f; pcf[r] = 0.f; G[r] = 0.f; py[r] = 0.f;
                if (r < nrows) {
                    const Tap tr = make_tap<LT>(p.ay, Yb + r);
                    lh0[r] = tr.l0; lh1[r] = tr.l1;
                    if (colok) {
                        const long long px = ((long long)c.b * p.ay.out + (Yb + r)) * p.ax.out + X;
                        const long long y = ld_label(p.labels, p.label_u8, px);
                        if (y != p.ignore && y >= 0 && y < p.C) {
                            lab[r] = (int)y;
                            const float lse = __ldg(d.lse_in + px);
                            lse2[r] = __fmul_rn(lse, kLog2e);
                            // p = 2^(z log2e - rn(lse log2e)) * 2^-(lse log2e - rn(lse log2e)), second factor to 1st order
                            pcf[r] = __fmaf_rn(-__fmaf_rn(lse, kLog2e, -lse2[r]), kLn2, 1.0f);
                        }
                    }
                }
            }
        }

        if (p.use_tma) {
            mbar_wait(&full[stage], (uint32_t)((n / kFwdStages) & 1));
        } else {
            load_plain<LT, NCH>(p, c, S);
            __syncthreads();
        }

        if (nrows > 0) {                                           // warp-uniform (a warp shares gy)
            const int c0 = c.chunk * NCH;
            const int nvalid = min(NCH, p.C - c0);
            const int plane = kFwdBH * p.box_w;
            float tt[NCH], tb[NCH], aux[NCH];                      // aux: STATS sum p_k of my pixels; COEF tb table
            {
                const LT* q0 = S + r0 * p.box_w;
                const LT* q1 = S + r1 * p.box_w;
                #pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    tt[k] = lerp_pinned(Logit<LT>::ld(q0 + k * plane + j0r), lw0, Logit<LT>::ld(q0 + k * plane + j1r), lw1);
                    tb[k] = lerp_pinned(Logit<LT>::ld(q1 + k * plane + j0r), lw0, Logit<LT>::ld(q1 + k * plane + j1r), lw1);
                    aux[k] = 0.f;
                    if (MODE == DICE_COEF && k < nvalid) aux[k] = __ldg(d.tb + (long long)c.b * p.C + c0 + k);
                }
            }
            #pragma unroll
            for (int r = 0; r < kFwdG; ++r) {
                if (r < nrows) {
                    const int yk = lab[r] - c0;                    // lab = -1 (ignored / outside the image): never matches
                    float pyv = 0.f;
                    if (lab[r] >= 0) {
                        float g = 0.f;
                        #pragma unroll
                        for (int k = 0; k < NCH; ++k) {
                            const float z = Logit<LT>::rnd(lerp_pinned(tt[k], lh0[r], tb[k], lh1[r]));
                            float pk = ex2_approx(__fmaf_rn(z, kLog2e, -lse2[r])) * pcf[r];
                            if (k >= nvalid) pk = 0.f;
                            if (k == yk) pyv = pk;
                            if (MODE == DICE_STATS) aux[k] += pk; else g = fmaf(aux[k], pk, g);
                        }
                        if (MODE == DICE_COEF) {
                            G[r] += g;
                            if (yk >= 0 && yk < nvalid) py[r] = pyv;
                        }
                    }
                    if (MODE == DICE_STATS && __any_sync(0xffffffffu, lab[r] >= 0 && yk >= 0 && yk < nvalid)) {
                        const int key = (lab[r] >= 0 && yk >= 0 && yk < nvalid) ? lab[r] : -1;
                        const unsigned peers = __match_any_sync(0xffffffffu, key);
                        const unsigned v = key >= 0 ? __float2uint_rn(pyv * kFixI) : 0u;
                        const unsigned sum = __reduce_add_sync(peers, v);
                        if (key >= 0 && lane == __ffs(peers) - 1) bins.add_inter(key, sum, __popc(peers));
                    }
                }
            }
            if (MODE == DICE_STATS) {
                #pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    const unsigned long long v = __float2ull_rn(aux[k] * kFixS);      // <= 4 * 2^40
                    const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)(v & 0xFFFFFull));
                    const unsigned hi = __reduce_add_sync(0xffffffffu, (unsigned)(v >> 20));
                    if (lane == 0 && k < nvalid) bins.add_sum(c0 + k, ((unsigned long long)hi << 20) + lo);
                }
            }
        }

        if (MODE == DICE_COEF && c.chunk == p.n_chunks - 1) {
            #pragma unroll
            for (int r = 0; r < kFwdG; ++r) {
                if (r < nrows && colok) {
                    const long long px = ((long long)c.b * p.ay.out + (Yb + r)) * p.ax.out + X;
                    float ca = 0.f, cb = 0.f;
                    if (lab[r] >= 0) {
                        const float wy = p.class_w ? __ldg(p.class_w + lab[r]) : 1.f;
                        cb = fmaf(__ldg(d.ta + (long long)c.b * p.C + lab[r]), py[r], inv_denom * wy);
                        ca = cb - G[r];
                    }
                    d.coef_a[px] = ca;
                    d.coef_b[px] = cb;
                }
            }
        }

        __syncthreads();
        if (p.use_tma && tid == 0 && n + kFwdStages < n_items)
            issue_tma<LT, NCH>(p, &tmap, decode_item<NCH>(p, n + kFwdStages), stages + stage * p.stage_bytes, &full[stage]);
    }

    if (MODE == DICE_STATS && cur_b >= 0) flush(cur_b);
}

// One CTA: per (image, class) dice coefficient, the loss and the backward tables.
__global__ void __launch_bounds__(256)
dice_finalize_kernel(unsigned char* dws, int B, int C, float eps, const Workspace* ws, float* loss_out) {
    __shared__ double scratch[8];
    sf_dice_state* st = reinterpret_cast<sf_dice_state*>(dws);
    const unsigned long long* stats = reinterpret_cast<const unsigned long long*>(dws + kDiceHeaderBytes);
    float* ta = reinterpret_cast<float*>(dws + kDiceHeaderBytes + (size_t)B * C * 24);
    float* tb = ta + (size_t)B * C;
    const double inv_bc = 1.0 / ((double)B * (double)C);
    double acc = 0.0;
    for (int e = threadIdx.x; e < B * C; e += 256) {
        const double sp = (double)stats[e * 3 + 0] * (1.0 / 1099511627776.0);
        const double si = (double)stats[e * 3 + 1] * (1.0 / 67108864.0);
        const double sn = (double)stats[e * 3 + 2];
        const double sets = sp + sn;
        const bool dead = sets == 0.0;                             // losses.py:152-153: sets_sum = 2 * inter, D constant
        const double den = (dead ? 2.0 * si : sets) + (double)eps;
        const double D = (2.0 * si + (double)eps) / den;
        acc += D;
        ta[e] = dead ? 0.f : (float)(2.0 / den * inv_bc);
        tb[e] = dead ? 0.f : (float)(D / den * inv_bc);
    }
    const double tot = block_sum<256>(acc, scratch);
    if (threadIdx.x == 0) {
        const float dl = (float)(1.0 - tot * inv_bc);
        const float ce = ws ? ws->state.loss : 0.f;
        st->dice_loss = dl; st->total_loss = ce + dl; st->B = B; st->C = C; st->epsilon = eps;
        if (loss_out) *loss_out = ce + dl;
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
size_t dice_ws_bytes(int B, int C) { return (size_t)kDiceHeaderBytes + (size_t)B * C * (24 + 8); }

template <typename LT, int NCH, int MODE>
static int launch_dice_one(const CUtensorMap& tmap, const FwdParams& p, const DiceParams& d, int grid, size_t smem, cudaStream_t st) {
    auto k = dice_kernel<LT, NCH, MODE>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    k<<<grid, kFwdThreads, smem, st>>>(tmap, p, d);
    return (int)cudaGetLastError();
}

template <typename LT, int MODE>
static int launch_dice_nch(int nch, const CUtensorMap& tmap, const FwdParams& p, const DiceParams& d, int grid, size_t smem, cudaStream_t st) {
    switch (nch) {
        case 19: return launch_dice_one<LT, 19, MODE>(tmap, p, d, grid, smem, st);
        case 16: return launch_dice_one<LT, 16, MODE>(tmap, p, d, grid, smem, st);
        default: return launch_dice_one<LT, 8, MODE>(tmap, p, d, grid, smem, st);
    }
}

template <int MODE>
static int launch_dice_walk(const FwdArgs& a, const DiceParams& d) {
    if (!a.logits || !a.labels || a.B <= 0 || a.C <= 0 || a.h <= 0 || a.w <= 0 || a.H <= 0 || a.W <= 0) return SF_ERR_ARG;
    if (a.ldtype < SF_F32 || a.ldtype > SF_F16 || a.ydtype < SF_I64 || a.ydtype > SF_U8) return SF_ERR_ARG;
    const int esz = a.ldtype == SF_F32 ? 4 : 2;
    if (((uintptr_t)a.logits % esz) || (a.ydtype == SF_I64 && ((uintptr_t)a.labels % 8))) return SF_ERR_ALIGN;
    if ((long long)a.H * a.W > (1LL << 23)) return SF_ERR_SHAPE;    // 2^40 fixed point of sum p_c must fit 64 bits
    if (!env_no_fast()) {
        FwdArgs f = a;
        f.dice_mode = MODE; f.dice = d;
        int rc = launch_dice_direct(f);                            // logits already at label resolution (sf_direct.cu)
        if (rc != -1000) return rc;
        rc = launch_forward_fast(f);                               // integer-scale specialisation (sf_forward_fast.cu)
        if (rc != -1000) return rc;
        rc = launch_forward_chunk(f);                              // integer scale, > 24 classes (sf_forward_chunk.cu)
        if (rc != -1000) return rc;
    }
    const int nch = pick_nch(a.C);
    const size_t aux = MODE == DICE_STATS ? (((size_t)3 * a.C * 8 + 15) & ~(size_t)15) : 16;
    FwdParams p{};
    CUtensorMap tmap{};
    int grid = 0;
    size_t smem = 0;
    const int rc = setup_generic_forward(a, nch, aux, &p, &tmap, &grid, &smem);
    if (rc != 0) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a.stream);
    if (a.ldtype == SF_F32) return launch_dice_nch<float, MODE>(nch, tmap, p, d, grid, smem, st);
    if (a.ldtype == SF_BF16) return launch_dice_nch<__nv_bfloat16, MODE>(nch, tmap, p, d, grid, smem, st);
    return launch_dice_nch<__half, MODE>(nch, tmap, p, d, grid, smem, st);
}

int launch_dice_stats(const FwdArgs& a, const float* lse_px, void* dice_ws) {
    unsigned char* dws = reinterpret_cast<unsigned char*>(dice_ws);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a.stream);
    cudaError_t e = cudaMemsetAsync(dws + kDiceHeaderBytes, 0, (size_t)a.B * a.C * 24, st);
    if (e != cudaSuccess) return (int)e;
    DiceParams d{};
    d.lse_in = lse_px;
    d.stats = reinterpret_cast<unsigned long long*>(dws + kDiceHeaderBytes);
    return launch_dice_walk<DICE_STATS>(a, d);
}

// CE forward that also gathers the dice sums: one fused launch where the integer-scale kernel covers the shape, else
// the CE forward followed by the stats walk.
int launch_ce_dice_forward(const FwdArgs& a, void* dice_ws) {
    unsigned char* dws = reinterpret_cast<unsigned char*>(dice_ws);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a.stream);
    if (!env_no_fast() && (long long)a.H * a.W <= (1LL << 23)) {
        cudaError_t e = cudaMemsetAsync(dws + kDiceHeaderBytes, 0, (size_t)a.B * a.C * 24, st);
        if (e != cudaSuccess) return (int)e;
        FwdArgs f = a;
        f.softmax = true; f.hist = nullptr; f.loss_px = nullptr;
        f.dice_mode = DICE_STATS;
        f.dice = DiceArgs{};
        f.dice.stats = reinterpret_cast<unsigned long long*>(dws + kDiceHeaderBytes);
        const int rc = launch_forward_fast(f);                     // FF_SOFTMAX_STATS
        if (rc != -1000) return rc;
    }
    FwdArgs f = a;
    f.softmax = true; f.dice_mode = DICE_OFF;
    const int rc = launch_forward(f);
    if (rc != 0) return rc;
    FwdArgs g = a;
    g.softmax = false; g.hist = nullptr; g.loss_px = nullptr; g.lse_px = nullptr;
    return launch_dice_stats(g, a.lse_px, dice_ws);
}

int launch_dice_finalize(int B, int C, float eps, const void* ws, void* dice_ws, float* loss_out, void* stream) {
    dice_finalize_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<unsigned char*>(dice_ws), B, C, eps, reinterpret_cast<const Workspace*>(ws), loss_out);
    return (int)cudaGetLastError();
}

int launch_dice_coef(const FwdArgs& a, const float* lse_px, const void* dice_ws, float* coef_a, float* coef_b) {
    const unsigned char* dws = reinterpret_cast<const unsigned char*>(dice_ws);
    DiceParams d{};
    d.lse_in = lse_px;
    d.ta = reinterpret_cast<const float*>(dws + kDiceHeaderBytes + (size_t)a.B * a.C * 24);
    d.tb = d.ta + (size_t)a.B * a.C;
    d.coef_a = coef_a; d.coef_b = coef_b;
    return launch_dice_walk<DICE_COEF>(a, d);
}

const float* dice_tb_table(const void* dice_ws, int B, int C) {
    const unsigned char* dws = reinterpret_cast<const unsigned char*>(dice_ws);
    return reinterpret_cast<const float*>(dws + kDiceHeaderBytes + (size_t)B * C * 24) + (size_t)B * C;
}

}  // namespace sf
