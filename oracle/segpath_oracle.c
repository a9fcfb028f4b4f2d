/*
 * segpath_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of Segmentation_Factory's per-pixel head-to-metric path.  It is the
 * checker for the CUDA kernels in segmentation_factory_b200/csrc and the `cpu_baseline`
 * leg of bench.py.  Nothing under segmentation_factory_b200/ may import, link or call it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against
 * golden vectors produced by the reference's own code (tests/golden/make_golden.py imports
 * /root/reference/util and torch's F.interpolate).  The reference ships no tests of its own
 * for this path, so those vectors are the pin.
 *
 * Reference lines followed (relative to the reference checkout):
 *   models/build_models.py:65   F.interpolate(bilinear, align_corners=False)   -> interp_*()
 *   util/losses.py:28-41        CrossEntropy (weighted mean over non-ignored)  -> orc_path_forward + orc_ce_reduce
 *   util/losses.py:44-66        OhemCrossEntropy (thresh / top-k fallback)     -> orc_ohem_reduce
 *   util/metrics.py:24-27       Metrics.update (argmax -> bincount)            -> orc_path_forward (hist)
 *   util/utils.py:99-109        ConfusionMatrix.update                         -> orc_hist_pairs
 *   autograd of the above                                                      -> orc_path_backward
 *   engine.py:10-15 criterion: CE + dice_loss (util/losses.py:126-177)         -> orc_dice_stats, orc_dice_backward
 *
 * The arithmetic lives in PyTorch ATen (reference pin torch==2.0.1; this image 2.11.0).
 * Bilinear rule, established by differential probing of torch CPU (generic NCHW kernel,
 * taken whenever H_out + W_out > 128):
 *   scale = (float)in / out;  r = fmaf(dst + 0.5f, scale, -0.5f); r = max(r, 0)
 *   i0 = min(floor(r), in-1); l1 = r - i0; l0 = 1 - l1; i1 = i0 + (i0 < in-1)
 *   in == out in a dimension: i0 = i1 = dst, l0 = 1, l1 = 0
 *   bf16/fp16 inputs: l1 is rounded to the input dtype, l0 = rnd(1 - rnd(l1))
 *   t_k  = fmaf(v[k][j0], lw0, rn(v[k][j1] * lw1))      for k in {i0, i1}
 *   out  = fmaf(t_i0, lh0, rn(t_i1 * lh1)), then rounded to the input dtype
 * Softmax / CE are evaluated in double and rounded to float at the end (a tighter truth than
 * the reference's fp32; gates are 1e-5 relative).
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { ORC_F32 = 0, ORC_BF16 = 1, ORC_F16 = 2 };
enum { ORC_I64 = 0, ORC_U8 = 1 };

/* ---------- low-precision helpers (round-to-nearest-even, like torch) ---------- */
static inline float bf16_to_f32(uint16_t b) {
    uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f;
}
static inline uint16_t f32_to_bf16(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   /* quiet NaN */
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline float f16_to_f32(uint16_t h) {
    uint32_t s = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu, u;
    if (e == 0) {
        if (m == 0) u = s;
        else { int k = 0; while (!(m & 0x400u)) { m <<= 1; ++k; } m &= 0x3ffu; u = s | ((uint32_t)(113 - k) << 23) | (m << 13); }
    } else if (e == 31) u = s | 0x7f800000u | (m << 13);
    else u = s | ((e + 112u) << 23) | (m << 13);
    float f; memcpy(&f, &u, 4); return f;
}
static inline uint16_t f32_to_f16(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    uint32_t s = (u >> 16) & 0x8000u; int32_t e = (int32_t)((u >> 23) & 0xffu) - 127 + 15; uint32_t m = u & 0x7fffffu;
    if (((u >> 23) & 0xffu) == 0xffu) return (uint16_t)(s | 0x7c00u | (m ? 0x200u : 0));
    if (e >= 31) return (uint16_t)(s | 0x7c00u);
    if (e <= 0) {
        if (e < -10) return (uint16_t)s;
        m |= 0x800000u; uint32_t shift = (uint32_t)(14 - e);
        uint32_t r = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (r & 1u))) ++r;
        return (uint16_t)(s | r);
    }
    uint32_t r = ((uint32_t)e << 10) | (m >> 13), rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) ++r;
    return (uint16_t)(s | r);
}
static inline float rnd_dtype(float v, int dt) {
    if (dt == ORC_BF16) return bf16_to_f32(f32_to_bf16(v));
    if (dt == ORC_F16)  return f16_to_f32(f32_to_f16(v));
    return v;
}
static inline float load_logit(const void* p, int dt, int64_t idx) {
    if (dt == ORC_F32)  return ((const float*)p)[idx];
    if (dt == ORC_BF16) return bf16_to_f32(((const uint16_t*)p)[idx]);
    return f16_to_f32(((const uint16_t*)p)[idx]);
}
static inline int64_t load_label(const void* p, int dt, int64_t idx) {
    return dt == ORC_I64 ? ((const int64_t*)p)[idx] : (int64_t)((const uint8_t*)p)[idx];
}

/* ---------- bilinear index/lambda tables (build_models.py:65 -> ATen) ---------- */
typedef struct { int32_t i0, i1; float l0, l1; } orc_tap;

static void make_taps(int in, int out, int dt, orc_tap* t) {
    const float scale = (float)in / (float)out;
    for (int d = 0; d < out; ++d) {
        if (in == out) { t[d].i0 = t[d].i1 = d; t[d].l0 = 1.0f; t[d].l1 = 0.0f; continue; }
        float r = fmaf((float)d + 0.5f, scale, -0.5f);
        if (r < 0.0f) r = 0.0f;
        int i0 = (int)floorf(r);
        if (i0 > in - 1) i0 = in - 1;
        float l1 = r - (float)i0;
        if (l1 < 0.0f) l1 = 0.0f;
        if (l1 > 1.0f) l1 = 1.0f;
        l1 = rnd_dtype(l1, dt);
        float l0 = rnd_dtype(1.0f - l1, dt);
        t[d].i0 = i0; t[d].i1 = i0 + (i0 < in - 1 ? 1 : 0); t[d].l0 = l0; t[d].l1 = l1;
    }
}

/* one interpolated logit, already rounded to the logits' dtype */
static inline float interp_one(const void* logits, int dt, int64_t plane, int w,
                               const orc_tap* ty, const orc_tap* tx) {
    const float v00 = load_logit(logits, dt, plane + (int64_t)ty->i0 * w + tx->i0);
    const float v01 = load_logit(logits, dt, plane + (int64_t)ty->i0 * w + tx->i1);
    const float v10 = load_logit(logits, dt, plane + (int64_t)ty->i1 * w + tx->i0);
    const float v11 = load_logit(logits, dt, plane + (int64_t)ty->i1 * w + tx->i1);
    const float m0 = v01 * tx->l1, m1 = v11 * tx->l1;        /* rn() products */
    const float t0 = fmaf(v00, tx->l0, m0);
    const float t1 = fmaf(v10, tx->l0, m1);
    const float m2 = t1 * ty->l1;
    return rnd_dtype(fmaf(t0, ty->l0, m2), dt);
}

/* materialised upsample: out[B,C,H,W] fp32 holding dtype-rounded values (test use only) */
int orc_upsample(const void* logits, int ldtype, int B, int C, int h, int w, int H, int W, float* out) {
    orc_tap* ty = (orc_tap*)malloc(sizeof(orc_tap) * (size_t)H);
    orc_tap* tx = (orc_tap*)malloc(sizeof(orc_tap) * (size_t)W);
    if (!ty || !tx) return -1;
    make_taps(h, H, ldtype, ty); make_taps(w, W, ldtype, tx);
    #pragma omp parallel for schedule(static)
    for (int64_t bc = 0; bc < (int64_t)B * C; ++bc)
        for (int Y = 0; Y < H; ++Y)
            for (int X = 0; X < W; ++X)
                out[(bc * H + Y) * W + X] = interp_one(logits, ldtype, bc * h * w, w, &ty[Y], &tx[X]);
    free(ty); free(tx);
    return 0;
}

/*
 * Forward over every label pixel: interpolate C logits, first-max argmax (NaN wins, first NaN
 * wins: metrics.py:25), log-softmax, per-pixel CE with ignore (losses.py:32,50,55), GT x pred
 * histogram (metrics.py:26-27, rows = GT).  Labels outside [0,C) are treated as ignored.
 *   loss_px [B*H*W]  w_y * (lse - z_y), 0 for ignored            (nullable)
 *   lse_px  [B*H*W]  max + log(sum exp)                          (nullable)
 *   arg_px  [B*H*W]  int32 argmax                                (nullable)
 *   hist    [C*C]    int64, accumulated into                     (nullable)
 *   sums    [4]      double: sum(loss_px), sum(w_y over valid), n_valid, n(loss_px > thresh); overwritten (nullable)
 */
int orc_path_forward(const void* logits, int ldtype, const void* labels, int ydtype,
                     int B, int C, int h, int w, int H, int W, int ignore,
                     const float* class_w, float thresh,
                     float* loss_px, float* lse_px, int32_t* arg_px, int64_t* hist, double* sums) {
    orc_tap* ty = (orc_tap*)malloc(sizeof(orc_tap) * (size_t)H);
    orc_tap* tx = (orc_tap*)malloc(sizeof(orc_tap) * (size_t)W);
    if (!ty || !tx) return -1;
    make_taps(h, H, ldtype, ty); make_taps(w, W, ldtype, tx);
    double s_loss = 0, s_w = 0, n_valid = 0, n_hard = 0;
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    int64_t* hpriv = hist ? (int64_t*)calloc((size_t)nthreads * C * C, sizeof(int64_t)) : NULL;
    if (hist && !hpriv) return -1;
    #pragma omp parallel reduction(+ : s_loss, s_w, n_valid, n_hard)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        float* z = (float*)malloc(sizeof(float) * (size_t)C);
        int64_t* hp = hpriv ? hpriv + (int64_t)tid * C * C : NULL;
        #pragma omp for schedule(static)
        for (int64_t row = 0; row < (int64_t)B * H; ++row) {
            const int b = (int)(row / H), Y = (int)(row % H);
            for (int X = 0; X < W; ++X) {
                const int64_t px = row * W + X;
                int best = 0; float bv = 0; int best_nan = 0;
                for (int c = 0; c < C; ++c) {
                    z[c] = interp_one(logits, ldtype, ((int64_t)b * C + c) * h * w, w, &ty[Y], &tx[X]);
                    if (c == 0) { best = 0; bv = z[0]; best_nan = isnan(z[0]); }
                    else if (!best_nan && (isnan(z[c]) || z[c] > bv)) { best = c; bv = z[c]; best_nan = isnan(z[c]); }
                }
                double se = 0;
                for (int c = 0; c < C; ++c) se += exp((double)z[c] - (double)bv);
                const double lg = log(se);
                const int64_t y = load_label(labels, ydtype, px);
                const int valid = (y != ignore) && y >= 0 && y < C;
                float l = 0.0f;
                if (valid) {
                    const double wy = class_w ? (double)class_w[y] : 1.0;
                    l = (float)(wy * (lg - ((double)z[y] - (double)bv)));
                    s_loss += l; s_w += wy; n_valid += 1;
                    if (l > thresh) n_hard += 1;
                    if (hp) hp[y * C + best] += 1;
                }
                if (loss_px) loss_px[px] = l;
                if (lse_px) lse_px[px] = (float)((double)bv + lg);
                if (arg_px) arg_px[px] = best;
            }
        }
        free(z);
    }
    if (hist) {
        for (int t = 0; t < nthreads; ++t)
            for (int k = 0; k < C * C; ++k) hist[k] += hpriv[(int64_t)t * C * C + k];
        free(hpriv);
    }
    if (sums) { sums[0] = s_loss; sums[1] = s_w; sums[2] = n_valid; sums[3] = n_hard; }
    free(ty); free(tx);
    return 0;
}

/* CrossEntropy reduction (losses.py:32,36): sum(w_y * l) / sum(w_y); NaN when nothing is valid */
float orc_ce_reduce(const double* sums) { return (float)(sums[0] / sums[1]); }

static int cmp_desc(const void* a, const void* b) {
    const float x = *(const float*)a, y = *(const float*)b;
    return (x < y) - (x > y);
}

/*
 * OhemCrossEntropy._forward (losses.py:52-61) on the per-pixel loss vector (ignored pixels are
 * zeros that stay in the vector).  n_min = n_valid / 16.  Outputs: mean loss, the value every
 * kept pixel must exceed-or-tie (`kth`; equals thresh on the threshold branch), the number of
 * kept pixels K, and branch (0 = threshold, 1 = top-k).
 */
int orc_ohem_reduce(const float* loss_px, int64_t n, int64_t n_valid, float thresh,
                    float* loss_out, float* kth, int64_t* kept, int* branch) {
    const int64_t n_min = n_valid / 16;
    int64_t n_hard = 0; double s = 0;
    for (int64_t i = 0; i < n; ++i) if (loss_px[i] > thresh) { ++n_hard; s += loss_px[i]; }
    if (n_hard >= n_min) {
        *branch = 0; *kept = n_hard; *kth = thresh;
        *loss_out = n_hard ? (float)(s / (double)n_hard) : NAN;      /* mean(empty) = nan */
        return 0;
    }
    float* tmp = (float*)malloc(sizeof(float) * (size_t)n);
    if (!tmp) return -1;
    memcpy(tmp, loss_px, sizeof(float) * (size_t)n);
    qsort(tmp, (size_t)n, sizeof(float), cmp_desc);
    s = 0;
    for (int64_t i = 0; i < n_min; ++i) s += tmp[i];
    *branch = 1; *kept = n_min; *kth = tmp[n_min - 1];
    *loss_out = (float)(s / (double)n_min);
    free(tmp);
    return 0;
}

/*
 * Gradient of the scalar loss w.r.t. the LOW-RES logits (autograd of interpolate + CE/OHEM).
 *   coef_px[px] : d loss / d loss_px  (0 for ignored / not kept; 1/K, 1/sum_w ... times grad_out)
 *   d loss_px / d z_c = w_y * (softmax_c - [c == y]);  then the transposed bilinear with the same taps.
 * Low-precision logits: the dtype rounding is treated as identity (straight-through), as autograd does.
 * grad [B,C,h,w] fp32 is overwritten.
 */
int orc_path_backward(const void* logits, int ldtype, const void* labels, int ydtype,
                      int B, int C, int h, int w, int H, int W, int ignore,
                      const float* class_w, const float* lse_px, const float* coef_px, float* grad) {
    orc_tap* ty = (orc_tap*)malloc(sizeof(orc_tap) * (size_t)H);
    orc_tap* tx = (orc_tap*)malloc(sizeof(orc_tap) * (size_t)W);
    if (!ty || !tx) return -1;
    make_taps(h, H, ldtype, ty); make_taps(w, W, ldtype, tx);
    #pragma omp parallel
    {
        double* acc = (double*)malloc(sizeof(double) * (size_t)h * w);
        #pragma omp for schedule(dynamic, 1)
        for (int64_t bc = 0; bc < (int64_t)B * C; ++bc) {
            const int b = (int)(bc / C), c = (int)(bc % C);
            memset(acc, 0, sizeof(double) * (size_t)h * w);
            for (int Y = 0; Y < H; ++Y)
                for (int X = 0; X < W; ++X) {
                    const int64_t px = ((int64_t)b * H + Y) * W + X;
                    const double cf = coef_px[px];
                    if (cf == 0.0) continue;
                    const int64_t y = load_label(labels, ydtype, px);
                    if (y == ignore || y < 0 || y >= C) continue;
                    const float zc = interp_one(logits, ldtype, bc * h * w, w, &ty[Y], &tx[X]);
                    const double p = exp((double)zc - (double)lse_px[px]);
                    const double wy = class_w ? (double)class_w[y] : 1.0;
                    const double g = cf * wy * (p - (c == y ? 1.0 : 0.0));
                    acc[ty[Y].i0 * w + tx[X].i0] += g * ty[Y].l0 * tx[X].l0;
                    acc[ty[Y].i0 * w + tx[X].i1] += g * ty[Y].l0 * tx[X].l1;
                    acc[ty[Y].i1 * w + tx[X].i0] += g * ty[Y].l1 * tx[X].l0;
                    acc[ty[Y].i1 * w + tx[X].i1] += g * ty[Y].l1 * tx[X].l1;
                }
            for (int k = 0; k < h * w; ++k) grad[bc * h * w + k] = (float)acc[k];
        }
        free(acc);
    }
    free(ty); free(tx);
    return 0;
}

/*
 * Dice term of engine.criterion (engine.py:10-15 -> util/losses.py:126-177, multiclass=True):
 *   p = softmax(upsampled logits); per image b and class c over the pixels whose label != ignore (build_target
 *   paints every channel of an ignored pixel with ignore_index, dice_coeff masks them out):
 *     inter = sum p_c [y == c];  sets = sum p_c + sum [y == c];  if sets == 0: sets = 2 inter
 *     D_bc = (2 inter + eps) / (sets + eps);   loss = 1 - mean_bc D_bc
 * stats [B][C][3] double = {sum p_c, inter, count[y == c]}; overwritten.
 */
int orc_dice_stats(const void* logits, int ldtype, const void* labels, int ydtype,
                   int B, int C, int h, int w, int H, int W, int ignore,
                   const float* lse_px, double* stats) {
    orc_tap* ty = (orc_tap*)malloc(sizeof(orc_tap) * (size_t)H);
    orc_tap* tx = (orc_tap*)malloc(sizeof(orc_tap) * (size_t)W);
    if (!ty || !tx) return -1;
    make_taps(h, H, ldtype, ty); make_taps(w, W, ldtype, tx);
    #pragma omp parallel for schedule(dynamic, 1)
    for (int64_t bc = 0; bc < (int64_t)B * C; ++bc) {
        const int b = (int)(bc / C), c = (int)(bc % C);
        double sp = 0.0, si = 0.0, sn = 0.0;
        for (int Y = 0; Y < H; ++Y)
            for (int X = 0; X < W; ++X) {
                const int64_t px = ((int64_t)b * H + Y) * W + X;
                const int64_t y = load_label(labels, ydtype, px);
                if (y == ignore || y < 0 || y >= C) continue;
                const float zc = interp_one(logits, ldtype, bc * h * w, w, &ty[Y], &tx[X]);
                const double p = exp((double)zc - (double)lse_px[px]);
                sp += p;
                if (y == c) { si += p; sn += 1.0; }
            }
        stats[bc * 3 + 0] = sp; stats[bc * 3 + 1] = si; stats[bc * 3 + 2] = sn;
    }
    free(ty); free(tx);
    return 0;
}

/*
 * Backward of  ce_scale * sum_px w_y (lse - z_y)  +  (1 - mean_bc D_bc)  through softmax and the bilinear resize.
 *   d(-mean D)/d p_c(px) = -(ta_bc [y == c] - tb_bc),   ta = 2 / (BC (sets+eps)),  tb = D / (BC (sets+eps))
 *   through softmax:  d/d z_k = p_k (A + tb_k) - [k == y] Bc,   Bc = ce_scale w_y + ta_y p_y,   A = Bc - sum_j tb_j p_j
 * ta, tb: [B][C] double (0 where the reference's sets == 0 branch makes D constant).
 */
int orc_dice_backward(const void* logits, int ldtype, const void* labels, int ydtype,
                      int B, int C, int h, int w, int H, int W, int ignore,
                      const float* class_w, const float* lse_px, double ce_scale,
                      const double* ta, const double* tb, float* grad) {
    orc_tap* ty = (orc_tap*)malloc(sizeof(orc_tap) * (size_t)H);
    orc_tap* tx = (orc_tap*)malloc(sizeof(orc_tap) * (size_t)W);
    const int64_t n = (int64_t)B * H * W;
    double* cA = (double*)malloc(sizeof(double) * (size_t)n);
    double* cB = (double*)malloc(sizeof(double) * (size_t)n);
    if (!ty || !tx || !cA || !cB) return -1;
    make_taps(h, H, ldtype, ty); make_taps(w, W, ldtype, tx);
    #pragma omp parallel for schedule(static)
    for (int64_t px = 0; px < n; ++px) {
        cA[px] = 0.0; cB[px] = 0.0;
        const int64_t y = load_label(labels, ydtype, px);
        if (y == ignore || y < 0 || y >= C) continue;
        const int b = (int)(px / ((int64_t)H * W)), Y = (int)(px / W % H), X = (int)(px % W);
        double G = 0.0, py = 0.0;
        for (int c = 0; c < C; ++c) {
            const float zc = interp_one(logits, ldtype, ((int64_t)b * C + c) * h * w, w, &ty[Y], &tx[X]);
            const double p = exp((double)zc - (double)lse_px[px]);
            G += tb[b * C + c] * p;
            if (c == y) py = p;
        }
        const double wy = class_w ? (double)class_w[y] : 1.0;
        cB[px] = ce_scale * wy + ta[b * C + y] * py;
        cA[px] = cB[px] - G;
    }
    #pragma omp parallel
    {
        double* acc = (double*)malloc(sizeof(double) * (size_t)h * w);
        #pragma omp for schedule(dynamic, 1)
        for (int64_t bc = 0; bc < (int64_t)B * C; ++bc) {
            const int b = (int)(bc / C), c = (int)(bc % C);
            memset(acc, 0, sizeof(double) * (size_t)h * w);
            for (int Y = 0; Y < H; ++Y)
                for (int X = 0; X < W; ++X) {
                    const int64_t px = ((int64_t)b * H + Y) * W + X;
                    const int64_t y = load_label(labels, ydtype, px);
                    if (y == ignore || y < 0 || y >= C) continue;
                    const float zc = interp_one(logits, ldtype, bc * h * w, w, &ty[Y], &tx[X]);
                    const double p = exp((double)zc - (double)lse_px[px]);
                    const double g = p * (cA[px] + tb[bc]) - (c == y ? cB[px] : 0.0);
                    acc[ty[Y].i0 * w + tx[X].i0] += g * ty[Y].l0 * tx[X].l0;
                    acc[ty[Y].i0 * w + tx[X].i1] += g * ty[Y].l0 * tx[X].l1;
                    acc[ty[Y].i1 * w + tx[X].i0] += g * ty[Y].l1 * tx[X].l0;
                    acc[ty[Y].i1 * w + tx[X].i1] += g * ty[Y].l1 * tx[X].l1;
                }
            for (int k = 0; k < h * w; ++k) grad[bc * h * w + k] = (float)acc[k];
        }
        free(acc);
    }
    free(ty); free(tx); free(cA); free(cB);
    return 0;
}

/* ConfusionMatrix.update (utils.py:99-109): k = (a>=0)&(a<n); mat[n*a[k]+b[k]] += 1.  b must be in [0,n). */
int orc_hist_pairs(const int64_t* a, const int64_t* b, int64_t n, int C, int64_t* mat) {
    for (int64_t i = 0; i < n; ++i)
        if (a[i] >= 0 && a[i] < C) {
            if (b[i] < 0 || b[i] >= C) return -2;
            mat[a[i] * C + b[i]] += 1;
        }
    return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Benchmark helper: torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU baseline must still use every
 * host core it is given. */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
