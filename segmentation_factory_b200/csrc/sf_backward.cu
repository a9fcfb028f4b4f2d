// sf_backward.cu -- fused softmax-minus-onehot + transposed-bilinear backward, gather form.
//
// Replaces autograd of util/losses.py:28-66 through models/build_models.py:65:
//   index/topk backward -> nll_loss2d_backward -> _log_softmax_backward_data -> upsample_bilinear2d_backward.
// d L / d z_low[b,c,i,j] = sum over label pixels px that interpolate from (i,j) of
//        W(px -> i,j) * coef_px * (softmax_c(px) - [c == y_px]),   coef_px = g * keep_px * w_y / denom.
//
// Every low-res gradient element is OWNED by one CTA (tile = kBwdTH rows x tbw columns x one channel chunk),
// which walks all label pixels that depend on its tile (a halo of one low-res cell), so there are no global
// atomics and the result is bitwise reproducible.  Thread t owns label column Xa+t for the whole walk:
//   * vertical part in registers: acc0/acc1[k] collect g*lh0 / g*lh1 for source rows i and i+1; when the walk
//     moves to the next source row, acc0 is complete ("row emission") and acc1 becomes the new acc0;
//   * horizontal part through shared memory: the emitted row R[k][x] is written once per thread, then each
//     owned (j,k) sums its <= dmax dependants with a precomputed dense weight table (lanes run along j, so
//     the gradient stores are coalesced; R is padded x + x/32 so the stride-s gather is bank-conflict free).
// The softmax is rebuilt from the staged low-res tile (TMA) and the per-pixel lse the forward stored:
// p = 2^(z*log2e - rn(lse*log2e)), the rounding residual of lse*log2e folded into the per-pixel coefficient.
// Rows where no lane of a warp has a kept pixel (ignored labels, OHEM-dropped pixels) skip the channel loop.
#include "sf_common.cuh"
#include "sf_launch.h"

namespace sf {

constexpr int kBwdThreads = 256;
constexpr int kBwdTH = 8;                    // owned low-res rows per tile
constexpr int kBwdBH = kBwdTH + 2;           // staged source rows
constexpr int kBwdTBWP = 64;                 // pitch of the weight table / max owned columns
constexpr int kBwdDMax = 64;                 // max dependants of one low-res column
constexpr int kBwdRP = kBwdThreads + kBwdThreads / 32;   // padded row pitch of R

struct BwdParams {
    const void* logits;
    const void* labels;
    const float* class_w;
    const float* loss_px;
    const float* lse_px;
    const float* grad_out;
    const Workspace* ws;
    const float* coef_a;                     // DICE: per-pixel A and Bc of sf_dice.cu (d/dz_k = p_k (A + tb_k) - [k == y] Bc)
    const float* coef_b;
    const float* dice_tb;                    // DICE: [B][C]
    void* grad;
    Axis ay, ax;
    int B, C, ignore, label_u8;
    int tbw, tiles_x, tiles_y, n_chunks;
    int box_w, dmax, use_tma, src_bytes;
    int xalign;                              // TMA inner coordinate granularity (16 bytes) in elements
};

__host__ __device__ __forceinline__ int first_dst_clamped(const Axis& ax, int i) {   // first d with src_i0(d) >= i
    return i <= 0 ? 0 : first_dst_of(ax, i);
}
__device__ __forceinline__ int swz(int x) { return x + (x >> 5); }
__host__ __device__ __forceinline__ int floor_al_b(int x, int a) { return x & ~(a - 1); }   // also right for x = -1

template <int NCH>
__device__ __forceinline__ void onehot_sub(int yk, float (&a0)[NCH], float (&a1)[NCH], float v0, float v1) {
#define SF_CASE(K) case K: a0[(K) < NCH ? (K) : 0] -= v0; a1[(K) < NCH ? (K) : 0] -= v1; break;
    switch (yk) {
        SF_CASE(0) SF_CASE(1) SF_CASE(2) SF_CASE(3) SF_CASE(4) SF_CASE(5) SF_CASE(6) SF_CASE(7)
        SF_CASE(8) SF_CASE(9) SF_CASE(10) SF_CASE(11) SF_CASE(12) SF_CASE(13) SF_CASE(14) SF_CASE(15)
        SF_CASE(16) SF_CASE(17) SF_CASE(18) SF_CASE(19) SF_CASE(20) SF_CASE(21) SF_CASE(22) SF_CASE(23)
        default: break;
    }
#undef SF_CASE
}

template <typename LT, int NCH, bool DICE>
__global__ void __launch_bounds__(kBwdThreads, 2)
bwd_kernel(const __grid_constant__ CUtensorMap tmap, const BwdParams p) {
    static_assert(NCH <= 24, "onehot_sub covers 24 channels per chunk");
    extern __shared__ __align__(128) unsigned char smem[];
    LT* S = reinterpret_cast<LT*>(smem);
    float* Rbuf = reinterpret_cast<float*>(smem + p.src_bytes);                 // [2][NCH][kBwdRP]
    float* Wd = Rbuf + 2 * NCH * kBwdRP;                                        // [dmax][kBwdTBWP]
    int* xs_tab = reinterpret_cast<int*>(Wd + p.dmax * kBwdTBWP);              // [kBwdTBWP]
    uint64_t* bar = reinterpret_cast<uint64_t*>(xs_tab + kBwdTBWP);

    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int chunk = bid % p.n_chunks; bid /= p.n_chunks;
    const int txi = bid % p.tiles_x; bid /= p.tiles_x;
    const int tyi = bid % p.tiles_y;
    const int b = bid / p.tiles_y;
    const int h = p.ay.in, w = p.ax.in;
    const int ia = tyi * kBwdTH, ib = min(ia + kBwdTH, h);
    const int ja = txi * p.tbw, jb = min(ja + p.tbw, w);
    const int nown = jb - ja;
    const int c0 = chunk * NCH;
    const int nvalid = min(NCH, p.C - c0);
    const int Xa = first_dst_clamped(p.ax, ja - 1), Xb = first_dst_clamped(p.ax, jb);
    const int Ya = first_dst_clamped(p.ay, ia - 1), Yb = first_dst_clamped(p.ay, ib);
    const int jO = floor_al_b(ja - 1, p.xalign);                    // box x-origin (16-byte aligned, may be negative)

    // ---- stage the low-res tile (rows ia-1 .. ia+TH, cols jO .. jO+box_w) ----
    if (p.use_tma) {
        if (tid == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
            mbar_expect_tx(bar, (uint32_t)(NCH * kBwdBH * p.box_w * (int)sizeof(LT)));
            tma_load_3d(S, &tmap, bar, jO, ia - 1, b * p.C + c0);
        }
    } else {
        const LT* src = reinterpret_cast<const LT*>(p.logits);
        const int n = NCH * kBwdBH * p.box_w;
        for (int e = tid; e < n; e += kBwdThreads) {
            const int col = e % p.box_w, row = (e / p.box_w) % kBwdBH, k = e / (p.box_w * kBwdBH);
            const int ch = c0 + k, i = ia - 1 + row, j = jO + col;
            LT v = LT(0.0f);
            if (ch < p.C && i >= 0 && i < h && j >= 0 && j < w) v = src[(((long long)b * p.C + ch) * h + i) * w + j];
            S[e] = v;
        }
    }

    // ---- per-tile tables: first dependant column and dense weights of every owned low-res column ----
    for (int jl = tid; jl < kBwdTBWP; jl += kBwdThreads)
        xs_tab[jl] = jl < nown ? first_dst_clamped(p.ax, ja + jl - 1) - Xa : 0;
    for (int e = tid; e < p.dmax * kBwdTBWP; e += kBwdThreads) {
        const int jl = e % kBwdTBWP, d = e / kBwdTBWP;
        float wgt = 0.f;
        if (jl < nown) {
            const int j = ja + jl;
            const int x = first_dst_clamped(p.ax, j - 1) + d;
            if (x < first_dst_clamped(p.ax, j + 1)) {
                const Tap t = make_tap<LT>(p.ax, x);
                wgt = (t.i0 == j ? t.l0 : 0.f) + ((t.i1 == j && t.i1 != t.i0) ? t.l1 : 0.f);
                if (t.i1 == t.i0 && t.i0 == j) wgt = t.l0 + t.l1;
            }
        }
        Wd[e] = wgt;
    }

    const int X = Xa + tid;
    const bool colok = X < Xb;
    const Tap tc = make_tap<LT>(p.ax, colok ? X : max(Xb - 1, 0));
    const int j0l = tc.i0 - jO, j1l = tc.i1 - jO;
    const float lw0 = tc.l0, lw1 = tc.l1;

    int branch = 0;
    float kth = 0.f, tie = 0.f;
    float gscale = p.grad_out ? __ldg(p.grad_out) : 1.0f;
    if (!DICE) {
        const sf_loss_state* stp = &p.ws->state;
        branch = stp->branch; kth = stp->kth; tie = stp->tie_share;
        gscale *= stp->inv_denom;
    }
    __shared__ float kb_s[24];                                      // DICE: tb[b][c0 .. c0+NCH) (LDS broadcast in the row loop)
    if (DICE && tid < NCH) kb_s[tid] = tid < nvalid ? __ldg(p.dice_tb + (long long)b * p.C + c0 + tid) : 0.f;

    __syncthreads();
    if (p.use_tma) mbar_wait(bar, 0);

    float acc0[NCH], acc1[NCH], tt[NCH], tb[NCH];
    #pragma unroll
    for (int k = 0; k < NCH; ++k) { acc0[k] = 0.f; acc1[k] = 0.f; tt[k] = 0.f; tb[k] = 0.f; }
    const int plane = kBwdBH * p.box_w;
    int ecount = 0;

    auto load_rows = [&](int i) {
        const int r0 = i - (ia - 1);
        const int r1 = r0 + ((i < h - 1 && !p.ay.ident) ? 1 : 0);
        const LT* q0 = S + r0 * p.box_w;
        const LT* q1 = S + r1 * p.box_w;
        #pragma unroll
        for (int k = 0; k < NCH; ++k) {
            tt[k] = lerp_pinned(Logit<LT>::ld(q0 + k * plane + j0l), lw0, Logit<LT>::ld(q0 + k * plane + j1l), lw1);
            tb[k] = lerp_pinned(Logit<LT>::ld(q1 + k * plane + j0l), lw0, Logit<LT>::ld(q1 + k * plane + j1l), lw1);
        }
    };

    // emit the finished source row i (acc0), then shift acc1 -> acc0.  Called uniformly by all threads.
    auto emit_row = [&](int i) {
        if (i >= ia) {
            float* buf = Rbuf + (ecount & 1) * NCH * kBwdRP;
            ++ecount;
            const int xp = swz(tid);
            #pragma unroll
            for (int k = 0; k < NCH; ++k) buf[k * kBwdRP + xp] = acc0[k];
            __syncthreads();
            LT* g = reinterpret_cast<LT*>(p.grad);
            const int n_out = nown * nvalid;
            for (int o = tid; o < n_out; o += kBwdThreads) {
                const int jl = o % nown, k = o / nown;
                const int xs = xs_tab[jl];
                const float* rk = buf + k * kBwdRP;
                float sum = 0.f;
                for (int d = 0; d < p.dmax; ++d) {
                    const int x = min(xs + d, kBwdThreads - 1);
                    sum = fmaf(rk[swz(x)], Wd[d * kBwdTBWP + jl], sum);
                }
                Logit<LT>::st(g + (((long long)b * p.C + c0 + k) * h + i) * w + ja + jl, sum);
            }
        }
        #pragma unroll
        for (int k = 0; k < NCH; ++k) { acc0[k] = acc1[k]; acc1[k] = 0.f; }
    };

    // i_cur: the source row acc0 currently collects for.  Starting one above the tile makes the first
    // transition(s) emit zero rows for owned rows that have no dependants above them (downscaling).
    int i_cur = ia - 1;
    bool loaded = false;

    for (int Y = Ya; Y < Yb; ++Y) {
        const Tap ty = make_tap<LT>(p.ay, Y);
        if (ty.i0 != i_cur || !loaded) {
            while (i_cur < ty.i0) { emit_row(i_cur); ++i_cur; }
            load_rows(i_cur);
            loaded = true;
        }
        // ---- per-pixel coefficient ----
        float coef = 0.f, lse = 0.f, cb = 0.f;
        int yk = -1;
        bool live = false;
        if (colok) {
            const long long px = ((long long)b * p.ay.out + Y) * p.ax.out + X;
            const long long y = ld_label(p.labels, p.label_u8, px);
            if (y != p.ignore && y >= 0 && y < p.C) {
                if (DICE) {
                    coef = __ldg(p.coef_a + px) * gscale;
                    cb = __ldg(p.coef_b + px) * gscale;
                    lse = __ldg(p.lse_px + px); yk = (int)y - c0; live = true;
                } else {
                    float keep = 1.f;
                    if (branch != 0) {
                        const float l = __ldg(p.loss_px + px);
                        keep = branch == 3 ? l : (l > kth ? 1.f : (l == kth ? tie : 0.f));   // 3: loss_px holds d focal / d ce
                    }
                    const float wy = p.class_w ? __ldg(p.class_w + y) : 1.f;
                    coef = keep != 0.f ? keep * wy * gscale : 0.f;       // never 0 * inf
                    if (coef != 0.f) { lse = __ldg(p.lse_px + px); yk = (int)y - c0; live = true; }
                }
            }
        }
        if (__any_sync(0xffffffffu, live)) {
            if (live) {
                const float lse2 = __fmul_rn(lse, kLog2e);
                const float corr = __fmaf_rn(lse, kLog2e, -lse2);          // exact residual of rn(lse*log2e)
                const float pf = __fmaf_rn(-corr, kLn2, 1.0f);             // 2^-corr to first order
                const float cf = coef * pf;
                if (!DICE) cb = coef;
                float lh0 = ty.l0, lh1 = ty.l1;
                float v0 = cf * lh0, v1 = cf * lh1, o0 = cb * lh0, o1 = cb * lh1;
                float u0 = DICE ? gscale * pf * lh0 : 0.f, u1 = DICE ? gscale * pf * lh1 : 0.f;
                if (ty.i1 == ty.i0) { v0 += v1; v1 = 0.f; o0 += o1; o1 = 0.f; u0 += u1; u1 = 0.f; }   // clamped bottom row / identity
                #pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    const float z = Logit<LT>::rnd(lerp_pinned(tt[k], lh0, tb[k], lh1));
                    const float pc = ex2_approx(__fmaf_rn(z, kLog2e, -lse2));
                    if (DICE) {
                        const float kbk = kb_s[k];
                        acc0[k] = fmaf(pc, fmaf(kbk, u0, v0), acc0[k]);
                        acc1[k] = fmaf(pc, fmaf(kbk, u1, v1), acc1[k]);
                    } else {
                        acc0[k] = fmaf(pc, v0, acc0[k]);
                        acc1[k] = fmaf(pc, v1, acc1[k]);
                    }
                }
                if (yk >= 0 && yk < nvalid) onehot_sub<NCH>(yk, acc0, acc1, o0, o1);
            }
        }
    }
    while (i_cur < ib) { emit_row(i_cur); ++i_cur; }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
static int pick_nch_bwd(int C) {
    if (C == 19) return 19;
    const int pad16 = (C + 15) / 16 * 16, pad8 = (C + 7) / 8 * 8;
    return pad8 < pad16 ? 8 : 16;
}

template <typename LT, int NCH, bool DICE>
static int launch_bwd_one(const CUtensorMap& tmap, const BwdParams& p, int grid, size_t smem, cudaStream_t st) {
    auto k = bwd_kernel<LT, NCH, DICE>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    k<<<grid, kBwdThreads, smem, st>>>(tmap, p);
    return (int)cudaGetLastError();
}

template <typename LT>
static int launch_bwd_nch(int nch, const CUtensorMap& tmap, const BwdParams& p, int grid, size_t smem, cudaStream_t st) {
    if (p.coef_a) {
        switch (nch) {
            case 19: return launch_bwd_one<LT, 19, true>(tmap, p, grid, smem, st);
            case 16: return launch_bwd_one<LT, 16, true>(tmap, p, grid, smem, st);
            default: return launch_bwd_one<LT, 8, true>(tmap, p, grid, smem, st);
        }
    }
    switch (nch) {
        case 19: return launch_bwd_one<LT, 19, false>(tmap, p, grid, smem, st);
        case 16: return launch_bwd_one<LT, 16, false>(tmap, p, grid, smem, st);
        default: return launch_bwd_one<LT, 8, false>(tmap, p, grid, smem, st);
    }
}

int launch_backward(const BwdArgs& a) {
    if (!a.logits || !a.labels || a.B <= 0 || a.C <= 0 || a.h <= 0 || a.w <= 0 || a.H <= 0 || a.W <= 0) return SF_ERR_ARG;
    if (a.ldtype < SF_F32 || a.ldtype > SF_F16 || a.ydtype < SF_I64 || a.ydtype > SF_U8) return SF_ERR_ARG;
    const int esz = a.ldtype == SF_F32 ? 4 : 2;
    if (((uintptr_t)a.logits % esz) || ((uintptr_t)a.grad_logits % esz) ||
        (a.ydtype == SF_I64 && ((uintptr_t)a.labels % 8))) return SF_ERR_ALIGN;

    if (a.coef_a && (!a.coef_b || !a.dice_tb)) return SF_ERR_ARG;
    if (!env_no_fast()) {
        int rc = launch_backward_direct(a);                        // logits already at label resolution (sf_direct.cu)
        if (rc != -1000) return rc;
        rc = launch_backward_fast(a);                              // integer-scale specialisation (sf_backward_fast.cu)
        if (rc != -1000) return rc;
    }
    BwdParams p{};
    p.logits = a.logits; p.labels = a.labels; p.class_w = a.class_w; p.loss_px = a.loss_px; p.lse_px = a.lse_px;
    p.grad_out = a.grad_out; p.ws = reinterpret_cast<const Workspace*>(a.ws); p.grad = a.grad_logits;
    p.coef_a = a.coef_a; p.coef_b = a.coef_b; p.dice_tb = a.dice_tb;
    p.ay = Axis{a.h, a.H, (float)a.h / (float)a.H, a.h == a.H};
    p.ax = Axis{a.w, a.W, (float)a.w / (float)a.W, a.w == a.W};
    p.B = a.B; p.C = a.C; p.ignore = a.ignore; p.label_u8 = a.ydtype == SF_U8;

    // dependants per low-res column, and the widest tile whose dependants fit one thread each
    int dmax = 1;
    for (int j = 0; j < a.w; ++j)
        dmax = std::max(dmax, first_dst_clamped(p.ax, j + 1) - first_dst_clamped(p.ax, j - 1));
    if (dmax > kBwdDMax) return SF_ERR_SHAPE;
    int tbw = std::min(kBwdTBWP, a.w);
    for (; tbw >= 1; --tbw) {
        bool ok = true;
        for (int ja = 0; ja < a.w && ok; ja += tbw)
            ok = first_dst_clamped(p.ax, std::min(ja + tbw, a.w)) - first_dst_clamped(p.ax, ja - 1) <= kBwdThreads;
        if (ok) break;
    }
    if (tbw < 1) return SF_ERR_SHAPE;
    p.tbw = tbw; p.dmax = dmax;
    p.tiles_x = (a.w + tbw - 1) / tbw;
    p.tiles_y = (a.h + kBwdTH - 1) / kBwdTH;
    const int nch = pick_nch_bwd(a.C);
    p.n_chunks = (a.C + nch - 1) / nch;
    const long long grid = (long long)a.B * p.tiles_x * p.tiles_y * p.n_chunks;
    if (grid >= (1LL << 31)) return SF_ERR_SHAPE;

    const int align = 16 / esz;
    p.xalign = align;
    int span = 1;
    for (int ja = 0; ja < a.w; ja += tbw) span = std::max(span, std::min(ja + tbw, a.w - 1) - floor_al_b(ja - 1, align) + 1);
    p.box_w = (span + align - 1) / align * align;
    if (p.box_w > 256) return SF_ERR_SHAPE;
    p.src_bytes = ((nch * kBwdBH * p.box_w * esz) + 127) & ~127;
    const size_t smem = (size_t)p.src_bytes + (size_t)2 * nch * kBwdRP * 4 + (size_t)dmax * kBwdTBWP * 4 + kBwdTBWP * 4 + 16;
    if (smem > 227 * 1024) return SF_ERR_SHAPE;

    CUtensorMap tmap{};
    p.use_tma = 0;
    if (!a.force_no_tma && ((size_t)a.w * esz) % 16 == 0 && ((uintptr_t)a.logits % 16) == 0 && p.box_w <= a.w && kBwdBH <= a.h) {
        const int rc = make_logits_tmap(&tmap, a.logits, a.ldtype, a.B * a.C, a.h, a.w, nch, kBwdBH, p.box_w);
        if (rc != 0) return rc;
        p.use_tma = 1;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a.stream);
    if (a.ldtype == SF_F32) return launch_bwd_nch<float>(nch, tmap, p, (int)grid, smem, st);
    if (a.ldtype == SF_BF16) return launch_bwd_nch<__nv_bfloat16>(nch, tmap, p, (int)grid, smem, st);
    return launch_bwd_nch<__half>(nch, tmap, p, (int)grid, smem, st);
}

}  // namespace sf
