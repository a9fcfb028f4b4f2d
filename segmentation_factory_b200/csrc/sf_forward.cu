// sf_forward.cu -- fused bilinear-upsample + log-softmax/CE + argmax + confusion-matrix forward.
//
// Replaces (reference file:line): models/build_models.py:65 (F.interpolate), util/losses.py:32,50,55
// (CrossEntropyLoss with ignore_index, reduction none/mean), util/metrics.py:25-27 (argmax + bincount),
// engine.py:90 (second argmax).  Full-resolution logits are never materialised.
//
// Decomposition
//   * work item = (tile, pass, channel chunk).  A tile is TW output columns x NGY "row groups"; a row
//     group is the set of output rows that share one (unclamped) source row u, so the two horizontal
//     lerps per channel are computed once per group and reused by all of its rows.
//   * the low-res logits needed by an item -- box [NCH][NGY+1][BW] -- are staged in shared memory by TMA
//     (cp.async.bulk.tensor.3d + mbarrier), kStages deep, one persistent CTA looping over its items.
//     Out-of-image box parts are zero-filled by TMA and never read (indices are clamped like ATen does).
//   * each thread owns one output column and one row group: per channel 2 FFMA-class ops per pixel for the
//     vertical lerp, then online softmax state (m, s, argmax) per pixel across channel chunks.
//   * GT x pred counts go to a shared-memory C x C int32 histogram as per-thread runs (a column's consecutive
//     pixels mostly share their cell), flushed once per CTA with 64-bit global atomics.
//   * loss sums are kept per thread in double, reduced per CTA in a fixed order and written to the
//     workspace (no atomics -> bitwise reproducible loss).
#include "sf_fwd_generic.cuh"

namespace sf {

template <typename LT, int NCH, bool SOFTMAX>
__global__ void __launch_bounds__(kFwdThreads, 2)
fwd_kernel(const __grid_constant__ CUtensorMap tmap, const FwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* stages = smem;
    int* hist_s = reinterpret_cast<int*>(smem + kFwdStages * p.stage_bytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kFwdStages * p.stage_bytes + p.hist_bytes);
    __shared__ double red_scratch[kFwdThreads / 32];

    const int tid = threadIdx.x;
    const int tx = tid % kFwdTW, gy = tid / kFwdTW;
    const bool want_hist = p.hist != nullptr;

    if (want_hist)
        for (int k = tid; k < p.C * p.C; k += kFwdThreads) hist_s[k] = 0;
    if (tid == 0) {
        if (p.use_tma) {
            tma_prefetch_desc(&tmap);
            for (int s = 0; s < kFwdStages; ++s) mbar_init(&full[s], 1);
            mbar_fence_init();
        }
    }
    __syncthreads();

    int n_my_tiles = 0;
    if ((int)blockIdx.x < p.total_tiles) n_my_tiles = (p.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;
    const int ipt = p.n_passes * p.n_chunks;
    const int n_items = n_my_tiles * ipt;

    if (p.use_tma && tid == 0)
        for (int n = 0; n < kFwdStages && n < n_items; ++n)
            issue_tma<LT, NCH>(p, &tmap, decode_item<NCH>(p, n), stages + n * p.stage_bytes, &full[n]);

    // per-thread loss accumulators
    double acc_loss = 0.0, acc_w = 0.0, acc_hard = 0.0;
    unsigned int cnt_valid = 0, cnt_hard = 0;
    int run_key = -1, run_cnt = 0;

    // per-(tile, pass) state
    int X = 0, j0r = 0, j1r = 0, r0 = 0, r1 = 0, nrows = 0, Yb = 0;
    float lw0 = 0.f, lw1 = 0.f;
    float lh0[kFwdG], lh1[kFwdG];
    int lab[kFwdG];
    float m[kFwdG], s[kFwdG], zy[kFwdG];
    int idx[kFwdG];
    unsigned nanlock = 0;                  // bit r: pixel r already holds its first-NaN argmax
    bool colok = false;

    for (int n = 0; n < n_items; ++n) {
        const int stage = n % kFwdStages;
        const ItemCoord c = decode_item<NCH>(p, n);
        LT* S = reinterpret_cast<LT*>(stages + stage * p.stage_bytes);

        if (c.chunk == 0) {
            // ---- (tile, pass) setup: column tap, the rows of my group handled in this pass, labels ----
            X = c.Xa + tx;
            colok = X < p.ax.out;
            const Tap tc = make_tap<LT>(p.ax, colok ? X : p.ax.out - 1);
            const int jA = floor_al(src_i0(p.ax, c.Xa), p.xalign);
            j0r = tc.i0 - jA; j1r = tc.i1 - jA; lw0 = tc.l0; lw1 = tc.l1;
            const int u = c.ua + gy;                               // my group's (unclamped) source row
            int y0 = p.ay.out, y1 = p.ay.out;
            if (u < p.ay.in) { y0 = first_dst_of(p.ay, u); y1 = first_dst_of(p.ay, u + 1); }
            Yb = y0 + c.pass * kFwdG;
            nrows = min(kFwdG, max(0, y1 - Yb));
            const int i0 = u < 0 ? 0 : u;
            r0 = i0 - c.ua;
            r1 = r0 + ((i0 < p.ay.in - 1 && !p.ay.ident) ? 1 : 0);
            nanlock = 0;
            #pragma unroll
            for (int r = 0; r < kFwdG; ++r) {
                lab[r] = -1; lh0[r] = 1.f; lh1[r] = 0.f;
                m[r] = -INFINITY; s[r] = 0.f; zy[r] = 0.f; idx[r] = 0;
                if (r < nrows) {
                    const Tap tr = make_tap<LT>(p.ay, Yb + r);
                    lh0[r] = tr.l0; lh1[r] = tr.l1;
                    if (colok) {
                        const long long y = ld_label(p.labels, p.label_u8, ((long long)c.b * p.ay.out + (Yb + r)) * p.ax.out + X);
                        lab[r] = (y != p.ignore && y >= 0 && y < p.C) ? (int)y : -1;
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

        if (nrows > 0 && colok) {
            const int c0 = c.chunk * NCH;
            const int nvalid = min(NCH, p.C - c0);
            const int plane = kFwdBH * p.box_w;
            float tt[NCH], tb[NCH];
            float nanflag = 0.f;
            {
                const LT* q0 = S + r0 * p.box_w;
                const LT* q1 = S + r1 * p.box_w;
                #pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    tt[k] = lerp_pinned(Logit<LT>::ld(q0 + k * plane + j0r), lw0, Logit<LT>::ld(q0 + k * plane + j1r), lw1);
                    tb[k] = lerp_pinned(Logit<LT>::ld(q1 + k * plane + j0r), lw0, Logit<LT>::ld(q1 + k * plane + j1r), lw1);
                    if (k < nvalid) nanflag = __fmaf_rn(tt[k], 0.f, __fmaf_rn(tb[k], 0.f, nanflag));
                }
            }
            const bool odd = nanflag != nanflag;                   // a NaN or Inf somewhere in this chunk's taps
            #pragma unroll
            for (int r = 0; r < kFwdG; ++r) {
                if (r < nrows) {
                    float z[NCH];
                    #pragma unroll
                    for (int k = 0; k < NCH; ++k) {
                        z[k] = Logit<LT>::rnd(lerp_pinned(tt[k], lh0[r], tb[k], lh1[r]));
                        if (k >= nvalid) z[k] = -INFINITY;
                    }
                    float cm = z[0]; int ci = 0;
                    #pragma unroll
                    for (int k = 1; k < NCH; ++k)
                        if (z[k] > cm) { cm = z[k]; ci = k; }
                    if (cm > m[r]) {
                        // rescale by the difference of the ROUNDED m*log2e values so every term of s stays
                        // relative to the same rn(m*log2e); the residual is undone once in the epilogue
                        if (SOFTMAX) s[r] *= ex2_approx(__fsub_rn(__fmul_rn(m[r], kLog2e), __fmul_rn(cm, kLog2e)));
                        m[r] = cm; idx[r] = c0 + ci;
                    }
                    if (odd && !(nanlock >> r & 1u)) {            // torch: NaN beats everything, first NaN wins
                        int first = -1;
                        #pragma unroll
                        for (int k = NCH - 1; k >= 0; --k)
                            if (z[k] != z[k]) first = k;
                        if (first >= 0) { idx[r] = c0 + first; m[r] = INFINITY; s[r] = NAN; nanlock |= 1u << r; }
                    }
                    if (SOFTMAX) {
                        if (m[r] > -INFINITY) {
                            const float m2 = __fmul_rn(m[r], kLog2e);
                            float a = 0.f;
                            #pragma unroll
                            for (int k = 0; k < NCH; ++k) a += ex2_approx(__fmaf_rn(z[k], kLog2e, -m2));
                            s[r] += a;
                        }
                        const int yk = lab[r] - c0;
                        if (yk >= 0 && yk < nvalid) {             // label logit, recomputed with the same ops
                            const LT* q0 = S + yk * plane + r0 * p.box_w;
                            const LT* q1 = S + yk * plane + r1 * p.box_w;
                            const float t0 = lerp_pinned(Logit<LT>::ld(q0 + j0r), lw0, Logit<LT>::ld(q0 + j1r), lw1);
                            const float t1 = lerp_pinned(Logit<LT>::ld(q1 + j0r), lw0, Logit<LT>::ld(q1 + j1r), lw1);
                            zy[r] = Logit<LT>::rnd(lerp_pinned(t0, lh0[r], t1, lh1[r]));
                        }
                    }
                }
            }
        }

        if (c.chunk == p.n_chunks - 1) {
            // ---- pass epilogue: per-pixel outputs, histogram, loss sums ----
            #pragma unroll
            for (int r = 0; r < kFwdG; ++r) {
                if (r < nrows) {                                   // warp-uniform (a warp shares gy)
                    const bool live = colok;
                    const long long px = ((long long)c.b * p.ay.out + (Yb + r)) * p.ax.out + X;
                    if (SOFTMAX && live) {
                        // s was accumulated as sum 2^(z*log2e - rn(m*log2e)); undo the rounding of m*log2e
                        const float m2 = __fmul_rn(m[r], kLog2e);
                        const float corr = __fmaf_rn(m[r], kLog2e, -m2);
                        const float lg = (log2f(s[r]) - corr) * kLn2;
                        const float lse = m[r] + lg;
                        float l = 0.f;
                        if (lab[r] >= 0) {
                            const float wy = p.class_w ? __ldg(p.class_w + lab[r]) : 1.f;
                            l = wy * (lg - (zy[r] - m[r]));
                            acc_loss += (double)l; acc_w += (double)wy; ++cnt_valid;
                            if (l > p.thresh) { acc_hard += (double)l; ++cnt_hard; }
                        }
                        if (p.loss_px) p.loss_px[px] = l;
                        if (p.lse_px) p.lse_px[px] = lse;
                    }
                    if (want_hist) {
                        const int key = (live && lab[r] >= 0) ? lab[r] * p.C + idx[r] : -1;   // per-thread runs (see sf_forward_fast.cu)
                        if (key != run_key) {
                            if (run_key >= 0) atomicAdd(&hist_s[run_key], run_cnt);
                            run_key = key; run_cnt = 0;
                        }
                        ++run_cnt;
                    }
                }
            }
        }

        __syncthreads();                                           // everyone is done with this stage
        if (p.use_tma && tid == 0 && n + kFwdStages < n_items)
            issue_tma<LT, NCH>(p, &tmap, decode_item<NCH>(p, n + kFwdStages), stages + stage * p.stage_bytes, &full[stage]);
    }

    // ---- CTA epilogue ----
    if (want_hist) {
        if (run_key >= 0) atomicAdd(&hist_s[run_key], run_cnt);
        __syncthreads();
        for (int k = tid; k < p.C * p.C; k += kFwdThreads) {
            const int v = hist_s[k];
            if (v) atomicAdd(p.hist + k, (unsigned long long)v);
        }
    }
    if (SOFTMAX) {
        const double r0d = block_sum<kFwdThreads>(acc_loss, red_scratch);
        const double r1d = block_sum<kFwdThreads>(acc_w, red_scratch);
        const double r2d = block_sum<kFwdThreads>(acc_hard, red_scratch);
        const double r3d = block_sum<kFwdThreads>((double)cnt_valid, red_scratch);
        const double r4d = block_sum<kFwdThreads>((double)cnt_hard, red_scratch);
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
int pick_nch(int C) {
    if (C == 19) return 19;
    const int pad16 = (C + 15) / 16 * 16, pad8 = (C + 7) / 8 * 8;
    return pad8 < pad16 ? 8 : 16;
}

template <typename LT, int NCH, bool SOFTMAX>
static int launch_one(const CUtensorMap& tmap, const FwdParams& p, int grid, size_t smem, cudaStream_t st) {
    auto k = fwd_kernel<LT, NCH, SOFTMAX>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    k<<<grid, kFwdThreads, smem, st>>>(tmap, p);
    return (int)cudaGetLastError();
}

template <typename LT, bool SOFTMAX>
static int launch_nch(int nch, const CUtensorMap& tmap, const FwdParams& p, int grid, size_t smem, cudaStream_t st) {
    switch (nch) {
        case 19: return launch_one<LT, 19, SOFTMAX>(tmap, p, grid, smem, st);
        case 16: return launch_one<LT, 16, SOFTMAX>(tmap, p, grid, smem, st);
        default: return launch_one<LT, 8, SOFTMAX>(tmap, p, grid, smem, st);
    }
}

int setup_generic_forward(const FwdArgs& a, int nch, size_t aux_bytes, FwdParams* pp, CUtensorMap* tmap_out, int* grid_out,
                          size_t* smem_out) {
    const int esz = a.ldtype == SF_F32 ? 4 : 2;
    FwdParams p{};
    p.logits = a.logits; p.labels = a.labels; p.class_w = a.class_w;
    p.loss_px = a.loss_px; p.lse_px = a.lse_px; p.hist = a.hist; p.ws = reinterpret_cast<Workspace*>(a.ws);
    p.ay = Axis{a.h, a.H, (float)a.h / (float)a.H, a.h == a.H};
    p.ax = Axis{a.w, a.W, (float)a.w / (float)a.W, a.w == a.W};
    p.B = a.B; p.C = a.C; p.ignore = a.ignore; p.label_u8 = a.ydtype == SF_U8;
    p.thresh = a.thresh;

    p.n_chunks = (a.C + nch - 1) / nch;
    p.tiles_x = (a.W + kFwdTW - 1) / kFwdTW;
    const int n_groups = p.ay.ident ? a.H : a.h + 1;               // u = -1 .. h-1 (or 0 .. H-1)
    p.tiles_y = (n_groups + kFwdNGY - 1) / kFwdNGY;
    const long long total = (long long)a.B * p.tiles_x * p.tiles_y;
    if (total >= (1LL << 31)) return SF_ERR_SHAPE;
    p.total_tiles = (int)total;

    // rows per group -> passes; box width = widest source span of a tile, rounded to 16 bytes
    int max_rows = 1;
    for (int u = p.ay.ident ? 0 : -1; u < a.h; ++u)
        max_rows = std::max(max_rows, first_dst_of(p.ay, u + 1) - first_dst_of(p.ay, u));
    p.n_passes = (max_rows + kFwdG - 1) / kFwdG;
    const int align = 16 / esz;
    p.xalign = align;
    int span = 1;
    for (int t = 0; t < p.tiles_x; ++t) {
        const int xa = t * kFwdTW, xb = std::min(xa + kFwdTW, a.W) - 1;
        const int jb = src_i0(p.ax, xb);
        span = std::max(span, jb + (jb < a.w - 1 && !p.ax.ident ? 1 : 0) - floor_al(src_i0(p.ax, xa), align) + 1);
    }
    p.box_w = (span + align - 1) / align * align;
    if (p.box_w > 256) return SF_ERR_SHAPE;
    p.stage_bytes = ((nch * kFwdBH * p.box_w * esz) + 127) & ~127;

    const size_t hist_bytes = aux_bytes;
    p.hist_bytes = (int)hist_bytes;
    const size_t smem = (size_t)kFwdStages * p.stage_bytes + hist_bytes + kFwdStages * 8 + 16;
    if (smem > 227 * 1024) return SF_ERR_TOO_MANY_CLASSES;

    CUtensorMap tmap{};
    p.use_tma = 0;
    // TMA needs 16-byte aligned rows and (observed on B200, driver 580) a box no wider than the tensor row;
    // tiny or odd-pitch maps take the plain cooperative loader instead
    if (!a.force_no_tma && ((size_t)a.w * esz) % 16 == 0 && ((uintptr_t)a.logits % 16) == 0 && p.box_w <= a.w && kFwdBH <= a.h) {
        const int rc = make_logits_tmap(&tmap, a.logits, a.ldtype, a.B * a.C, a.h, a.w, nch, kFwdBH, p.box_w);
        if (rc != 0) return rc;
        p.use_tma = 1;
    }

    *grid_out = (int)std::min<long long>(total, std::min(kMaxCtas, 2 * sm_count()));
    *smem_out = smem;
    *tmap_out = tmap;
    *pp = p;
    return 0;
}

int launch_forward(const FwdArgs& a) {
    if (!a.logits || !a.labels || a.B <= 0 || a.C <= 0 || a.h <= 0 || a.w <= 0 || a.H <= 0 || a.W <= 0) return SF_ERR_ARG;
    if (a.ldtype < SF_F32 || a.ldtype > SF_F16 || a.ydtype < SF_I64 || a.ydtype > SF_U8) return SF_ERR_ARG;
    if (a.softmax && !a.ws) return SF_ERR_ARG;
    const int esz = a.ldtype == SF_F32 ? 4 : 2;
    if (((uintptr_t)a.logits % esz) || (a.ydtype == SF_I64 && ((uintptr_t)a.labels % 8))) return SF_ERR_ALIGN;
    if ((long long)a.B * a.H * a.W >= (1LL << 40) || (long long)a.B * a.C >= (1LL << 31)) return SF_ERR_SHAPE;
    if (!env_no_fast()) {
        int rc = launch_forward_direct(a);                         // logits already at label resolution (sf_direct.cu)
        if (rc != -1000) return rc;
        rc = launch_forward_fast(a);                               // integer-scale specialisation (sf_forward_fast.cu)
        if (rc != -1000) return rc;
        rc = launch_forward_chunk(a);                              // integer scale, > 24 classes (sf_forward_chunk.cu)
        if (rc != -1000) return rc;
    }

    const int nch = pick_nch(a.C);
    const size_t hist_bytes = a.hist ? (((size_t)a.C * a.C * 4 + 15) & ~(size_t)15) : 16;
    FwdParams p{};
    CUtensorMap tmap{};
    int grid = 0;
    size_t smem = 0;
    const int rc = setup_generic_forward(a, nch, hist_bytes, &p, &tmap, &grid, &smem);
    if (rc != 0) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a.stream);
    if (a.ldtype == SF_F32)
        return a.softmax ? launch_nch<float, true>(nch, tmap, p, grid, smem, st) : launch_nch<float, false>(nch, tmap, p, grid, smem, st);
    if (a.ldtype == SF_BF16)
        return a.softmax ? launch_nch<__nv_bfloat16, true>(nch, tmap, p, grid, smem, st) : launch_nch<__nv_bfloat16, false>(nch, tmap, p, grid, smem, st);
    return a.softmax ? launch_nch<__half, true>(nch, tmap, p, grid, smem, st) : launch_nch<__half, false>(nch, tmap, p, grid, smem, st);
}

}  // namespace sf
