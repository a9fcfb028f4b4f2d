// sf_direct.cu -- same-size path: logits already at label resolution (h == H, w == W).
//
// This is what an UNMODIFIED engine.py hands to the loss / metrics (models/build_models.py:65 has already upsampled):
// no interpolation is involved, so the kernels are plain streaming passes over [B,C,H,W] and are HBM-bound
// (C*e bytes per pixel in, plus C*e out for the gradient).  Replaces util/losses.py:32,50,55 (+ autograd) and
// util/metrics.py:25-27 / engine.py:90 for that case.  Same loss arithmetic as the fused kernels.
//   * one thread = 4 consecutive pixels, 128-bit loads (fp32) / 64-bit (bf16, fp16) along W for every class plane;
//   * classes stream in chunks of 8 with an online softmax (running max, rescaled sum, first-max index, label logit);
//     a NaN logit wins the argmax at its first occurrence, like torch.argmax;
//   * confusion counts: shared-memory C x C int32 fed by per-thread runs, one 64-bit flush per CTA;
//   * backward: g = coef * (2^(z*log2e - lse*log2e) - [c == y]) written straight to grad (overwrites every element).
#include "sf_common.cuh"
#include "sf_dice_bins.cuh"
#include "sf_launch.h"

namespace sf {

constexpr int kDirThreads = 256, kDirVec = 4, kDirCh = 8;

template <typename LT> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void st(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[4]) {
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[4]) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        uint2 t; t.x = *reinterpret_cast<const unsigned*>(&a); t.y = *reinterpret_cast<const unsigned*>(&b);
        *reinterpret_cast<uint2*>(p) = t;
    }
};
template <> struct Vec4<__half> {
    static __device__ __forceinline__ void ld(const __half* p, float (&v)[4]) {
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
    static __device__ __forceinline__ void st(__half* p, const float (&v)[4]) {
        const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
        uint2 t; t.x = *reinterpret_cast<const unsigned*>(&a); t.y = *reinterpret_cast<const unsigned*>(&b);
        *reinterpret_cast<uint2*>(p) = t;
    }
};

struct DirParams {
    const void* logits; const void* labels; const float* class_w;
    float* loss_px; float* lse_px; unsigned long long* hist; Workspace* ws;
    const float* loss_in; const float* lse_in; const float* grad_out; void* grad;
    DiceArgs dice;              // dice walks (stats / coef) and, with coef_a/coef_b/tb set, the CE + dice backward
    long long plane;            // H*W (multiple of 4)
    long long n_quads;          // B*H*W / 4
    int C, ignore, label_u8;
    float thresh;
};

template <typename LT, bool SOFTMAX>
__global__ void __launch_bounds__(kDirThreads)
fwd_direct_kernel(const DirParams p) {
    extern __shared__ int hist_s[];
    __shared__ double red_scratch[kDirThreads / 32];
    __shared__ float cw_s[256];
    const int tid = threadIdx.x, C = p.C;
    const bool want_hist = p.hist != nullptr;
    if (want_hist)
        for (int k = tid; k < C * C; k += kDirThreads) hist_s[k] = 0;
    cw_s[tid] = (p.class_w && tid < C) ? __ldg(p.class_w + tid) : 1.0f;
    __syncthreads();

    const LT* logits = reinterpret_cast<const LT*>(p.logits);
    double acc_loss = 0.0, acc_w = 0.0, acc_hard = 0.0;
    unsigned cnt_valid = 0, cnt_hard = 0;
    int run_key = -1, run_cnt = 0;
    const long long stride = (long long)gridDim.x * kDirThreads;
    const long long n_round = p.n_quads;

    for (long long qd = (long long)blockIdx.x * kDirThreads + tid; qd < n_round; qd += stride) {
        const bool live = qd < p.n_quads;
        const long long px = (live ? qd : 0) * kDirVec;              // first pixel of the quad (flat over B*H*W)
        const long long b = px / p.plane, off = px - b * p.plane;
        const LT* base = logits + b * C * p.plane + off;
        int lab[kDirVec];
        #pragma unroll
        for (int v = 0; v < kDirVec; ++v) {
            const long long y = live ? ld_label(p.labels, p.label_u8, px + v) : -1;
            lab[v] = (y != p.ignore && y >= 0 && y < C) ? (int)y : -1;
        }
        float m[kDirVec], s[kDirVec], zy[kDirVec];
        int idx[kDirVec], nanidx[kDirVec];
        #pragma unroll
        for (int v = 0; v < kDirVec; ++v) { m[v] = -INFINITY; s[v] = 0.f; zy[v] = 0.f; idx[v] = 0; nanidx[v] = -1; }

        for (int c0 = 0; c0 < C; c0 += kDirCh) {
            float z[kDirCh][kDirVec];
            #pragma unroll
            for (int k = 0; k < kDirCh; ++k) {
                if (c0 + k < C) Vec4<LT>::ld(base + (long long)(c0 + k) * p.plane, z[k]);
                else { z[k][0] = z[k][1] = z[k][2] = z[k][3] = -INFINITY; }
            }
            #pragma unroll
            for (int v = 0; v < kDirVec; ++v) {
                float cm = z[0][v]; int ci = 0;
                #pragma unroll
                for (int k = 1; k < kDirCh; ++k)
                    if (z[k][v] > cm) { cm = z[k][v]; ci = k; }
                #pragma unroll
                for (int k = 0; k < kDirCh; ++k)
                    if (z[k][v] != z[k][v] && nanidx[v] < 0) nanidx[v] = c0 + k;          // first NaN over all classes
                if (cm > m[v]) {
                    if (SOFTMAX) s[v] *= ex2_approx(__fsub_rn(__fmul_rn(m[v], kLog2e), __fmul_rn(cm, kLog2e)));
                    m[v] = cm; idx[v] = c0 + ci;
                }
                if (SOFTMAX) {
                    if (m[v] > -INFINITY) {
                        const float m2 = __fmul_rn(m[v], kLog2e);
                        float a = 0.f;
                        #pragma unroll
                        for (int k = 0; k < kDirCh; ++k) a += ex2_approx(__fmaf_rn(z[k][v], kLog2e, -m2));
                        s[v] += a;
                    }
                    #pragma unroll
                    for (int k = 0; k < kDirCh; ++k)
                        if (c0 + k == lab[v]) zy[v] = z[k][v];
                }
            }
        }

        #pragma unroll
        for (int v = 0; v < kDirVec; ++v) {
            if (nanidx[v] >= 0) { idx[v] = nanidx[v]; s[v] = NAN; }  // torch.argmax: the first NaN wins
            if (SOFTMAX && live) {
                const float m2 = __fmul_rn(m[v], kLog2e);
                const float corr = __fmaf_rn(m[v], kLog2e, -m2);
                const float lg = (log2f(s[v]) - corr) * kLn2;
                float l = 0.f;
                if (lab[v] >= 0) {
                    const float wy = cw_s[lab[v]];
                    l = wy * (lg - (zy[v] - m[v]));
                    acc_loss += (double)l; acc_w += (double)wy; ++cnt_valid;
                    if (l > p.thresh) { acc_hard += (double)l; ++cnt_hard; }
                }
                if (p.loss_px) p.loss_px[px + v] = l;
                if (p.lse_px) p.lse_px[px + v] = m[v] + lg;
            }
            if (want_hist) {
                const int key = (live && lab[v] >= 0) ? lab[v] * C + idx[v] : -1;   // counted as per-thread runs (see sf_forward_fast.cu)
                if (key != run_key) {
                    if (run_key >= 0) atomicAdd(&hist_s[run_key], run_cnt);
                    run_key = key; run_cnt = 0;
                }
                ++run_cnt;
            }
        }
    }

    if (want_hist) {
        if (run_key >= 0) atomicAdd(&hist_s[run_key], run_cnt);
        __syncthreads();
        for (int k = tid; k < C * C; k += kDirThreads) {
            const int v = hist_s[k];
            if (v) atomicAdd(p.hist + k, (unsigned long long)v);
        }
    }
    if (SOFTMAX) {
        const double r0d = block_sum<kDirThreads>(acc_loss, red_scratch);
        const double r1d = block_sum<kDirThreads>(acc_w, red_scratch);
        const double r2d = block_sum<kDirThreads>(acc_hard, red_scratch);
        const double r3d = block_sum<kDirThreads>((double)cnt_valid, red_scratch);
        const double r4d = block_sum<kDirThreads>((double)cnt_hard, red_scratch);
        if (tid == 0) {
            double* o = p.ws->partials[blockIdx.x];
            o[0] = r0d; o[1] = r1d; o[2] = r2d; o[3] = r3d; o[4] = r4d;
            if (blockIdx.x == 0) p.ws->n_partials = gridDim.x;
        }
    }
}

// DICE: d/dz_k = g [p_k (A + tb[b][k]) - [k == y] Bc] with the per-pixel A, Bc of the coef walk (engine.criterion).
template <typename LT, bool DICE>
__global__ void __launch_bounds__(kDirThreads)
bwd_direct_kernel(const DirParams p) {
    __shared__ float cw_s[256];
    const int tid = threadIdx.x, C = p.C;
    cw_s[tid] = (p.class_w && tid < C) ? __ldg(p.class_w + tid) : 1.0f;
    __syncthreads();
    const sf_loss_state* stp = &p.ws->state;                        // DICE: never dereferenced
    const int branch = DICE ? 0 : stp->branch;
    const float kth = DICE ? 0.f : stp->kth, tie = DICE ? 0.f : stp->tie_share;
    const float gscale = (p.grad_out ? __ldg(p.grad_out) : 1.0f) * (DICE ? 1.0f : stp->inv_denom);
    const LT* logits = reinterpret_cast<const LT*>(p.logits);
    LT* grad = reinterpret_cast<LT*>(p.grad);
    const long long stride = (long long)gridDim.x * kDirThreads;
    for (long long qd = (long long)blockIdx.x * kDirThreads + tid; qd < p.n_quads; qd += stride) {
        const long long px = qd * kDirVec;
        const long long b = px / p.plane, off = px - b * p.plane;
        const long long base = b * C * p.plane + off;
        float coef[kDirVec], cf[kDirVec], lse2[kDirVec], uf[kDirVec], cbv[kDirVec];
        int lab[kDirVec];
        const float* tbb = DICE ? p.dice.tb + b * C : nullptr;
        #pragma unroll
        for (int v = 0; v < kDirVec; ++v) {
            const long long y = ld_label(p.labels, p.label_u8, px + v);
            coef[v] = 0.f; cf[v] = 0.f; lse2[v] = 0.f; lab[v] = -1; uf[v] = 0.f; cbv[v] = 0.f;
            if (DICE) {
                if (y != p.ignore && y >= 0 && y < C) {
                    const float lse = __ldg(p.lse_in + px + v);
                    lse2[v] = __fmul_rn(lse, kLog2e);
                    const float pf = __fmaf_rn(-__fmaf_rn(lse, kLog2e, -lse2[v]), kLn2, 1.0f);
                    coef[v] = 1.f;                                   // marks the pixel live
                    cf[v] = __ldg(p.dice.coef_a + px + v) * gscale * pf;
                    uf[v] = gscale * pf;
                    cbv[v] = __ldg(p.dice.coef_b + px + v) * gscale;
                    lab[v] = (int)y;
                }
            } else if (y != p.ignore && y >= 0 && y < C) {
                float keep = 1.f;
                if (branch != 0) { const float l = __ldg(p.loss_in + px + v); keep = branch == 3 ? l : (l > kth ? 1.f : (l == kth ? tie : 0.f)); }
                if (keep != 0.f) {
                    const float lse = __ldg(p.lse_in + px + v);
                    coef[v] = keep * cw_s[(int)y] * gscale;
                    lse2[v] = __fmul_rn(lse, kLog2e);
                    cf[v] = coef[v] * __fmaf_rn(-__fmaf_rn(lse, kLog2e, -lse2[v]), kLn2, 1.0f);
                    lab[v] = (int)y;
                }
            }
        }
        #pragma unroll 4
        for (int c = 0; c < C; ++c) {
            float z[kDirVec], g[kDirVec];
            Vec4<LT>::ld(logits + base + (long long)c * p.plane, z);
            const float tbc = DICE ? __ldg(tbb + c) : 0.f;
            #pragma unroll
            for (int v = 0; v < kDirVec; ++v) {
                const float pc = coef[v] != 0.f ? ex2_approx(__fmaf_rn(z[v], kLog2e, -lse2[v])) : 0.f;
                if (DICE) g[v] = __fmaf_rn(pc, __fmaf_rn(tbc, uf[v], cf[v]), c == lab[v] ? -cbv[v] : 0.f);
                else g[v] = __fmaf_rn(pc, cf[v], c == lab[v] ? -coef[v] : 0.f);
            }
            Vec4<LT>::st(grad + base + (long long)c * p.plane, g);
        }
    }
}

// Dice walks of engine.criterion on same-size logits (see sf_dice.cu for the scheme).  One CTA column per image
// (blockIdx.y = b), so the shared bins belong to one image; everything else streams like the kernels above.
// MODE DICE_STATS: per class sum p_c (2^40 fixed point, REDUX over the warp, 64-bit shared/global integer atomics),
// sum p_y [y == c] (2^26 per pixel) and the label counts.  MODE DICE_COEF: G = sum_j tb_j p_j -> coef_a, coef_b.
template <typename LT, int MODE>
__global__ void __launch_bounds__(kDirThreads)
dice_direct_kernel(const DirParams p) {
    extern __shared__ unsigned bin_words[];                          // STATS: DiceBins
    const DiceBins bins{bin_words, p.C};
    __shared__ float cw_s[256], ta_s[256], tb_s[256];
    const int tid = threadIdx.x, C = p.C, lane = tid & 31;
    const int b = blockIdx.y;
    if (MODE == DICE_STATS) bins.zero(tid, kDirThreads);
    cw_s[tid] = (p.class_w && tid < C) ? __ldg(p.class_w + tid) : 1.0f;
    if (MODE == DICE_COEF) {
        ta_s[tid] = tid < C ? __ldg(p.dice.ta + (long long)b * C + tid) : 0.f;
        tb_s[tid] = tid < C ? __ldg(p.dice.tb + (long long)b * C + tid) : 0.f;
    }
    __syncthreads();
    const float inv_denom = (MODE == DICE_COEF && p.ws) ? p.ws->state.inv_denom : 0.f;
    const LT* logits = reinterpret_cast<const LT*>(p.logits) + (long long)b * C * p.plane;
    const long long q_img = p.plane / kDirVec;
    const long long stride = (long long)gridDim.x * kDirThreads;
    const long long n_round = (q_img + 31) / 32 * 32;                // whole warps stay in the loop (match_any / redux)

    int iters = 0;
    for (long long qd = (long long)blockIdx.x * kDirThreads + tid; qd < n_round; qd += stride) {
        if (MODE == DICE_STATS && ++iters >= kDiceFlushAdds / 32) { bins.flush(p.dice.stats, b, tid, kDirThreads); iters = 0; }   // 8 warps x 4 adds
        const bool live = qd < q_img;
        const long long off = (live ? qd : 0) * kDirVec;
        const long long px = (long long)b * p.plane + off;
        int lab[kDirVec];
        float lse2[kDirVec], pcf[kDirVec], py[kDirVec], G[kDirVec];
        #pragma unroll
        for (int v = 0; v < kDirVec; ++v) {
            const long long y = live ? ld_label(p.labels, p.label_u8, px + v) : -1;
            lab[v] = (y != p.ignore && y >= 0 && y < C) ? (int)y : -1;
            lse2[v] = 0.f; pcf[v] = 0.f; py[v] = 0.f; G[v] = 0.f;
            if (lab[v] >= 0) {
                const float lse = __ldg(p.dice.lse_in + px + v);
                lse2[v] = __fmul_rn(lse, kLog2e);
                pcf[v] = __fmaf_rn(-__fmaf_rn(lse, kLog2e, -lse2[v]), kLn2, 1.0f);
            }
        }
        const bool any_valid = __any_sync(0xffffffffu, lab[0] >= 0 || lab[1] >= 0 || lab[2] >= 0 || lab[3] >= 0);
        if (any_valid) {
            #pragma unroll 4
            for (int c = 0; c < C; ++c) {
                float z[kDirVec];
                Vec4<LT>::ld(logits + off + (long long)c * p.plane, z);
                float sum4 = 0.f;
                #pragma unroll
                for (int v = 0; v < kDirVec; ++v) {
                    const float pk = lab[v] >= 0 ? ex2_approx(__fmaf_rn(z[v], kLog2e, -lse2[v])) * pcf[v] : 0.f;
                    if (c == lab[v]) py[v] = pk;
                    if (MODE == DICE_STATS) sum4 += pk; else G[v] = fmaf(tb_s[c], pk, G[v]);
                }
                if (MODE == DICE_STATS) {
                    const unsigned long long f = __float2ull_rn(sum4 * 1099511627776.f);       // <= 4 * 2^40
                    const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)(f & 0xFFFFFull));
                    const unsigned hi = __reduce_add_sync(0xffffffffu, (unsigned)(f >> 20));
                    if (lane == 0) bins.add_sum(c, ((unsigned long long)hi << 20) + lo);
                }
            }
        }
        #pragma unroll
        for (int v = 0; v < kDirVec; ++v) {
            if (MODE == DICE_STATS) {
                const int key = lab[v];
                const unsigned peers = __match_any_sync(0xffffffffu, key);
                const unsigned sum = __reduce_add_sync(peers, key >= 0 ? __float2uint_rn(py[v] * 67108864.f) : 0u);
                if (key >= 0 && lane == __ffs(peers) - 1) bins.add_inter(key, sum, __popc(peers));
            } else if (live) {
                float ca = 0.f, cb = 0.f;
                if (lab[v] >= 0) {
                    cb = fmaf(ta_s[lab[v]], py[v], inv_denom * cw_s[lab[v]]);
                    ca = cb - G[v];
                }
                p.dice.coef_a[px + v] = ca;
                p.dice.coef_b[px + v] = cb;
            }
        }
    }
    if (MODE == DICE_STATS) bins.flush(p.dice.stats, b, tid, kDirThreads);
}

static bool direct_ok(int h, int w, int H, int W, int C, const void* logits, int ldtype, const void* grad) {
    if (h != H || w != W || C > 256) return false;
    if (((long long)H * W) % kDirVec) return false;
    const int esz = ldtype == SF_F32 ? 4 : 2;
    if (((uintptr_t)logits % (esz * kDirVec)) || (grad && ((uintptr_t)grad % (esz * kDirVec)))) return false;
    return true;
}

int launch_forward_direct(const FwdArgs& a) {
    if (!direct_ok(a.h, a.w, a.H, a.W, a.C, a.logits, a.ldtype, nullptr)) return -1000;
    const size_t smem = a.hist ? (size_t)a.C * a.C * 4 : 0;
    if (smem > 200 * 1024) return -1000;
    DirParams p{};
    p.logits = a.logits; p.labels = a.labels; p.class_w = a.class_w; p.loss_px = a.loss_px; p.lse_px = a.lse_px;
    p.hist = a.hist; p.ws = reinterpret_cast<Workspace*>(a.ws);
    p.plane = (long long)a.H * a.W; p.n_quads = (long long)a.B * p.plane / kDirVec;
    p.C = a.C; p.ignore = a.ignore; p.label_u8 = a.ydtype == SF_U8; p.thresh = a.thresh;
    const long long want = (p.n_quads + kDirThreads - 1) / kDirThreads;
    const int per_sm = smem > 100 * 1024 ? 1 : (smem > 48 * 1024 ? 2 : 4);
    const int grid = (int)std::min<long long>(want, std::min(kMaxCtas, per_sm * sm_count()));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a.stream);
#define SF_LAUNCH(LT, SM)                                                                                           \
    {                                                                                                               \
        auto k = fwd_direct_kernel<LT, SM>;                                                                         \
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
        if (e != cudaSuccess) return (int)e;                                                                        \
        k<<<grid, kDirThreads, smem, st>>>(p);                                                                      \
    }
    if (a.ldtype == SF_F32) { if (a.softmax) SF_LAUNCH(float, true) else SF_LAUNCH(float, false) }
    else if (a.ldtype == SF_BF16) { if (a.softmax) SF_LAUNCH(__nv_bfloat16, true) else SF_LAUNCH(__nv_bfloat16, false) }
    else { if (a.softmax) SF_LAUNCH(__half, true) else SF_LAUNCH(__half, false) }
#undef SF_LAUNCH
    return (int)cudaGetLastError();
}

int launch_dice_direct(const FwdArgs& a) {
    if (!direct_ok(a.h, a.w, a.H, a.W, a.C, a.logits, a.ldtype, nullptr) || a.B > 65535) return -1000;
    DirParams p{};
    p.logits = a.logits; p.labels = a.labels; p.class_w = a.class_w; p.ws = reinterpret_cast<Workspace*>(a.ws);
    p.dice = a.dice;
    p.plane = (long long)a.H * a.W; p.n_quads = (long long)a.B * p.plane / kDirVec;
    p.C = a.C; p.ignore = a.ignore; p.label_u8 = a.ydtype == SF_U8;
    const long long want = (p.plane / kDirVec + kDirThreads - 1) / kDirThreads;
    const int gx = (int)std::max<long long>(1, std::min<long long>(want, (8LL * sm_count() + a.B - 1) / a.B));
    const dim3 grid(gx, a.B);
    const size_t smem = a.dice_mode == DICE_STATS ? (size_t)3 * a.C * 8 : 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a.stream);
#define SF_LAUNCH(LT)                                                                                               \
    {                                                                                                               \
        if (a.dice_mode == DICE_STATS) dice_direct_kernel<LT, DICE_STATS><<<grid, kDirThreads, smem, st>>>(p);      \
        else dice_direct_kernel<LT, DICE_COEF><<<grid, kDirThreads, smem, st>>>(p);                                 \
    }
    if (a.ldtype == SF_F32) SF_LAUNCH(float) else if (a.ldtype == SF_BF16) SF_LAUNCH(__nv_bfloat16) else SF_LAUNCH(__half)
#undef SF_LAUNCH
    return (int)cudaGetLastError();
}

int launch_backward_direct(const BwdArgs& a) {
    if (!direct_ok(a.h, a.w, a.H, a.W, a.C, a.logits, a.ldtype, a.grad_logits)) return -1000;
    DirParams p{};
    p.dice.coef_a = const_cast<float*>(a.coef_a); p.dice.coef_b = const_cast<float*>(a.coef_b); p.dice.tb = a.dice_tb;
    p.logits = a.logits; p.labels = a.labels; p.class_w = a.class_w; p.loss_in = a.loss_px; p.lse_in = a.lse_px;
    p.grad_out = a.grad_out; p.ws = const_cast<Workspace*>(reinterpret_cast<const Workspace*>(a.ws)); p.grad = a.grad_logits;
    p.plane = (long long)a.H * a.W; p.n_quads = (long long)a.B * p.plane / kDirVec;
    p.C = a.C; p.ignore = a.ignore; p.label_u8 = a.ydtype == SF_U8;
    const long long want = (p.n_quads + kDirThreads - 1) / kDirThreads;
    const int grid = (int)std::min<long long>(want, 8LL * sm_count());
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a.stream);
    if (a.coef_a) {
        if (a.ldtype == SF_F32) bwd_direct_kernel<float, true><<<grid, kDirThreads, 0, st>>>(p);
        else if (a.ldtype == SF_BF16) bwd_direct_kernel<__nv_bfloat16, true><<<grid, kDirThreads, 0, st>>>(p);
        else bwd_direct_kernel<__half, true><<<grid, kDirThreads, 0, st>>>(p);
    } else {
        if (a.ldtype == SF_F32) bwd_direct_kernel<float, false><<<grid, kDirThreads, 0, st>>>(p);
        else if (a.ldtype == SF_BF16) bwd_direct_kernel<__nv_bfloat16, false><<<grid, kDirThreads, 0, st>>>(p);
        else bwd_direct_kernel<__half, false><<<grid, kDirThreads, 0, st>>>(p);
    }
    return (int)cudaGetLastError();
}

}  // namespace sf
