/*
 * segfactory.h -- C ABI of libsegfactory.so (sm_100a).
 *
 * The reference (jiaowoguanren0615/Segmentation_Factory) has no FFI for this path: the per-pixel
 * head-to-metric work is a chain of PyTorch ATen calls made from Python.  Each entry point below
 * replaces one such chain; the Python mirror of the reference API (segmentation_factory_b200/)
 * binds them with ctypes.  INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless it says "host"; the caller owns all memory, the
 *     library allocates nothing and keeps no mutable global state (re-entrant, any host thread);
 *   - `stream` is a cudaStream_t passed as void*; nothing here synchronises with the host;
 *   - tensors are dense: logits [B,C,h,w] (NCHW), labels [B,H,W]; H,W may differ from h,w, the
 *     bilinear upsample (align_corners=False) of models/build_models.py:65 happens in-kernel;
 *   - return 0 on success, a negative sf_status, or a positive cudaError_t from the launch;
 *   - labels equal to `ignore`, negative, or >= C take no part in loss, gradient or histogram; int64 labels must fit int32
 *     (the integer-scale kernels read their low 32 bits: every class id, 255, -100 and -1 keep their meaning).
 */
#ifndef SEGFACTORY_H_
#define SEGFACTORY_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SF_VERSION 200            /* 0.2.0 */

enum sf_dtype       { SF_F32 = 0, SF_BF16 = 1, SF_F16 = 2 };     /* logits / gradient element type */
enum sf_label_dtype { SF_I64 = 0, SF_U8 = 1 };                   /* datasets/*.py: int64, Kvasir uint8 */

enum sf_status {
    SF_OK              =  0,
    SF_ERR_ARG         = -1,     /* null pointer, non-positive extent, unknown dtype           */
    SF_ERR_ALIGN       = -2,     /* a pointer is not aligned to its element size               */
    SF_ERR_TOO_MANY_CLASSES = -3,/* C*C int32 histogram does not fit in shared memory (C > 230) */
    SF_ERR_WORKSPACE   = -4,     /* workspace smaller than sf_workspace_bytes()                */
    SF_ERR_DRIVER      = -5,     /* cuTensorMapEncodeTiled unavailable / failed                */
    SF_ERR_SHAPE       = -6,     /* extents overflow the kernels' 32-bit tile arithmetic       */
    SF_ERR_UNSUPPORTED = -7      /* sf_ce_onepass: shape outside the one-pass kernels -- use
                                    sf_ce_forward / sf_loss_finalize / sf_ce_backward instead     */
};

/* Device-resident result of sf_loss_finalize (first bytes of the workspace). */
typedef struct sf_loss_state {
    float    loss;        /* the scalar the reference's forward() returns                       */
    float    kth;         /* pixels with loss_px >  kth are kept with weight 1                  */
    float    tie_share;   /* pixels with loss_px == kth are kept with this weight (top-k ties)  */
    float    inv_denom;   /* 1/K (OHEM) or 1/sum(w_y) (CE): d loss / d loss_px of a kept pixel  */
    int32_t  branch;      /* 0 = CE, 1 = OHEM threshold branch, 2 = OHEM top-k branch, 3 = focal */
    int32_t  reserved;
    uint64_t n_valid;     /* non-ignored pixels                                                 */
    uint64_t n_hard;      /* pixels with loss_px > thresh                                       */
    uint64_t n_kept;      /* K                                                                  */
    double   sum_kept;    /* sum of kept per-pixel losses                                       */
    double   sum_w;       /* sum of class weights over valid pixels                             */
    float    focal_alpha; /* branch 3: FocalLoss alpha ...                                      */
    float    focal_gamma; /* ... and gamma; d loss / d ce_px = a[(1-pt)^g + g ce pt (1-pt)^(g-1)] */
} sf_loss_state;

int         sf_version(void);
const char* sf_strerror(int code);

/* Bytes of zero-initialised-once device scratch sf_ce_forward / sf_loss_finalize / sf_ce_backward share. */
size_t sf_workspace_bytes(void);

/*
 * Replaces engine.py:90-91 -- outputs.argmax(1) twice + Metrics.update (util/metrics.py:24-27) +
 * ConfusionMatrix.update (util/utils.py:99-109) -- and the upsample feeding them.
 * hist[C*C] (rows = ground truth, cols = prediction) is ACCUMULATED into.
 */
int sf_eval_hist(const void* logits, int ldtype, const void* labels, int ydtype,
                 int B, int C, int h, int w, int H, int W, int ignore,
                 unsigned long long* hist, void* stream);

/*
 * Replaces F.interpolate + nn.CrossEntropyLoss(weight, ignore_index, reduction='none')
 * (util/losses.py:32,50,55) and, when hist != NULL, the argmax/bincount of Metrics.update.
 *   class_w   [C] fp32 or NULL
 *   ohem_thresh  -log(0.7) style threshold; pass a negative value for plain CrossEntropy
 *   loss_px   [B*H*W] fp32, w_y*(lse - z_y), 0 where ignored   (NULL allowed for plain CE)
 *   lse_px    [B*H*W] fp32 log-sum-exp, consumed by sf_ce_backward (NULL when no gradient is needed)
 *   hist      [C*C] accumulated (NULL to skip the argmax/histogram)
 *   workspace >= sf_workspace_bytes(); per-CTA partial sums are left there for sf_loss_finalize
 */
int sf_ce_forward(const void* logits, int ldtype, const void* labels, int ydtype,
                  int B, int C, int h, int w, int H, int W, int ignore,
                  const float* class_w, float ohem_thresh,
                  float* loss_px, float* lse_px, unsigned long long* hist,
                  void* workspace, size_t workspace_bytes, void* stream);

/*
 * Replaces the reductions of util/losses.py:36 (weighted mean) and :54-61 (n_min, threshold test,
 * topk fallback, mean) without host synchronisation.  Reads the partial sums sf_ce_forward left in
 * `workspace`; if fewer than n_valid/16 pixels exceed ohem_thresh it radix-selects the n_min-th
 * largest loss from loss_px.  Writes sf_loss_state at workspace[0] and, when loss_out != NULL, the
 * scalar loss to *loss_out (device fp32) so the caller can hand it to autograd without a copy.
 */
int sf_loss_finalize(const float* loss_px, long long n_px, float ohem_thresh,
                     void* workspace, size_t workspace_bytes, float* loss_out, void* stream);

/*
 * Replaces util/losses.py:9-25 (FocalLoss: alpha * (1 - exp(-ce))**gamma * ce, mean over ALL pixels or sum) on the
 * per-pixel cross-entropy sf_ce_forward wrote to loss_px (call it with ohem_thresh < 0 and class_w = NULL).
 * loss_px is OVERWRITTEN with d focal / d ce per pixel, which sf_ce_backward (sf_loss_state.branch == 3) then uses as
 * the pixel's weight -- no transcendental work is added to the backward kernels.
 */
int sf_focal_finalize(float* loss_px, long long n_px, float alpha, float gamma, int size_average,
                      void* workspace, size_t workspace_bytes, float* loss_out, void* stream);

/*
 * Replaces autograd's nll_loss2d_backward + _log_softmax_backward_data + upsample_bilinear2d_backward.
 * Gather-form: every low-res gradient element is produced by exactly one CTA, no global atomics.
 *   grad_out     device fp32 scalar (d L / d loss, e.g. the GradScaler's scale) or NULL for 1.0
 *   grad_logits  [B,C,h,w] in the logits' dtype, overwritten
 */
int sf_ce_backward(const void* logits, int ldtype, const void* labels, int ydtype,
                   int B, int C, int h, int w, int H, int W, int ignore,
                   const float* class_w, const float* loss_px, const float* lse_px,
                   const float* grad_out, const void* workspace,
                   void* grad_logits, void* stream);

/*
 * One-pass form of sf_ce_forward + sf_ce_backward for training steps (util/losses.py:32-36, 52-57 + autograd): ONE walk
 * over the label pixels computes the per-pixel loss, the confusion counts AND the softmax-minus-onehot gradient, because
 * the keep rule of plain CE (label != ignore) and of the OHEM threshold branch (loss > thresh) is pixel-local; only the
 * 1/sum(w) or 1/K scale is global and is applied by sf_ce_onepass_backward.  Call order for one step:
 *   sf_ce_onepass -> sf_loss_finalize -> [sf_ce_onepass_backward when the gradient is wanted].
 *   loss_px   [B*H*W] fp32, required when ohem_thresh >= 0 (the top-k fallback selects from it), else nullable
 *   hist      [C*C] accumulated, nullable
 *   grad_acc  [B*C*h*w] fp32 scratch: unscaled gradient sums, overwritten
 * Returns SF_ERR_UNSUPPORTED for shapes outside the one-pass kernels (H == S*h, W == S*w, S in {4, 8}, C in {2, 9, 19, 21},
 * 16-byte aligned rows); sf_ce_onepass_supported() tells beforehand.  Nothing is launched in that case.
 */
int sf_ce_onepass_supported(int ldtype, const void* logits, int C, int h, int w, int H, int W);
int sf_ce_onepass(const void* logits, int ldtype, const void* labels, int ydtype,
                  int B, int C, int h, int w, int H, int W, int ignore,
                  const float* class_w, float ohem_thresh,
                  float* loss_px, unsigned long long* hist, float* grad_acc,
                  void* workspace, size_t workspace_bytes, void* stream);

/*
 * Completes the backward of sf_ce_onepass after sf_loss_finalize, without host synchronisation:
 *   branch 0 / 1 (CE, OHEM threshold): grad_logits = grad_acc * grad_out / denom (one streaming kernel);
 *   branch 2 (OHEM top-k fallback, util/losses.py:58-59): the speculative sums are void -- a second walk keeps the pixels
 *   with loss_px > k-th value (ties share, see sf_loss_state) and writes grad_logits directly.
 * One launch either way: the CTAs read the device-side branch and either scale their slice or re-walk their tile.
 */
int sf_ce_onepass_backward(const void* logits, int ldtype, const void* labels, int ydtype,
                           int B, int C, int h, int w, int H, int W, int ignore,
                           const float* class_w, const float* loss_px, const float* grad_acc,
                           const float* grad_out, const void* workspace,
                           void* grad_logits, void* stream);

/*
 * Dice term of engine.criterion (engine.py:10-15: loss = F.cross_entropy(...) + dice_loss(softmax(x), build_target(...),
 * multiclass=True, ignore_index); util/losses.py:126-177), fused with the CE path above.  Call order for one step:
 *   sf_ce_forward(ohem_thresh < 0, lse_px != NULL) -> sf_loss_finalize -> sf_dice_stats -> sf_dice_finalize
 *   [-> sf_dice_coef -> sf_dice_backward when a gradient is wanted].
 * dice_ws is caller-owned device scratch of sf_dice_workspace_bytes(B, C) bytes; its first bytes are an sf_dice_state.
 */
typedef struct sf_dice_state {
    float   dice_loss;    /* 1 - mean over (image, class) of the dice coefficient              */
    float   total_loss;   /* CE (sf_loss_state.loss of `workspace`, 0 if NULL) + dice_loss      */
    int32_t B, C;
    float   epsilon;
    int32_t reserved[11];
} sf_dice_state;

size_t sf_dice_workspace_bytes(int B, int C);

/* Per image and class over the non-ignored pixels: sum softmax_c, sum softmax_c [y == c], count [y == c]
 * (the three sums dice_coeff, util/losses.py:139-158, takes with masked gathers, dot and sum).  lse_px from sf_ce_forward. */
int sf_dice_stats(const void* logits, int ldtype, const void* labels, int ydtype,
                  int B, int C, int h, int w, int H, int W, int ignore,
                  const float* lse_px, void* dice_ws, size_t dice_ws_bytes, void* stream);

/* sf_ce_forward (plain CE, lse_px stored) and sf_dice_stats in one call: a single fused launch where the integer-scale
 * kernel covers the shape (the exponentials of the log-sum-exp are reused for the softmax sums), else the two launches. */
int sf_ce_dice_forward(const void* logits, int ldtype, const void* labels, int ydtype,
                       int B, int C, int h, int w, int H, int W, int ignore,
                       const float* class_w, float* lse_px, void* workspace, size_t workspace_bytes,
                       void* dice_ws, size_t dice_ws_bytes, void* stream);

/* D = (2 inter + eps) / (sets + eps), sets == 0 -> sets = 2 inter (util/losses.py:152-155); dice loss = 1 - mean D
 * (:161-177).  Writes sf_dice_state and, when loss_out != NULL, CE + dice to *loss_out (device fp32). */
int sf_dice_finalize(int B, int C, float epsilon, const void* workspace, void* dice_ws, size_t dice_ws_bytes,
                     float* loss_out, void* stream);

/* Per-pixel scalars of the fused CE + dice gradient  d L / d z_k = p_k (A + tb[b][k]) - [k == y] Bc :
 * coef_b_px = w_y / sum(w) + ta[b][y] p_y,  coef_a_px = coef_b_px - sum_j tb[b][j] p_j;  both [B*H*W] fp32, 0 where ignored.
 * workspace NULL: dice gradient only. */
int sf_dice_coef(const void* logits, int ldtype, const void* labels, int ydtype,
                 int B, int C, int h, int w, int H, int W, int ignore,
                 const float* class_w, const float* lse_px, const void* workspace,
                 const void* dice_ws, size_t dice_ws_bytes, float* coef_a_px, float* coef_b_px, void* stream);

/* Replaces autograd of engine.criterion through softmax, cross_entropy and the bilinear upsample; same gather form and
 * conventions as sf_ce_backward. */
int sf_dice_backward(const void* logits, int ldtype, const void* labels, int ydtype,
                     int B, int C, int h, int w, int H, int W, int ignore,
                     const float* lse_px, const float* coef_a_px, const float* coef_b_px,
                     const void* dice_ws, size_t dice_ws_bytes, const float* grad_out,
                     void* grad_logits, void* stream);

/*
 * Label ingest.  Replaces the datasets' per-sample numpy gathers -- Cityscapes.encode_target
 * (datasets/cityscapes.py:92,130-131: id_to_train_id[np.array(target)]) and COCOStuff.encode (datasets/coco_stuff.py:66-69,
 * 98-100: label_map[label]) -- and the int64 collate behind them: the annotation bytes are copied to the device as they are
 * and mapped there.  out[i] = lut[raw[i]];  lut is a DEVICE table of 256 bytes;  out_dtype SF_U8 (what the kernels above
 * read at 1 B/px) or SF_I64 (the reference's dtype).
 */
int sf_label_lut(const unsigned char* raw, long long n, const unsigned char* lut, void* out, int out_dtype, void* stream);

/*
 * Replaces ConfusionMatrix.update(a, b) (util/utils.py:99-109) when the caller already holds the
 * flattened label / prediction index vectors (the unmodified engine.py:90 call).
 *   a [n] labels (int64 or uint8), b [n] int64 predictions in [0,C); mat[C*C] accumulated.
 */
int sf_hist_pairs(const void* a, int adtype, const long long* b, long long n, int C,
                  unsigned long long* mat, void* stream);

/*
 * Replaces the three tensor updates that follow every histogram in the reference -- `self.hist += bincount(...)`
 * (util/metrics.py:27, float32 += int64) and `self.mat += bincount(...)` (util/utils.py:109) -- with one launch:
 * hist_f32[k] = fl32(hist_f32[k] + fl32(counts[k])), hist_i64[k] += counts[k], mat[k] += counts[k], counts[k] = 0.
 * Any of the three destinations may be NULL.  `counts` is the per-call matrix sf_eval_hist just filled.
 */
int sf_accumulate_counts(unsigned long long* counts, float* hist_f32, long long* hist_i64, long long* mat,
                         int n, void* stream);

/*
 * engine.py:90-91 for one batch in ONE call: sf_eval_hist into `counts` (device scratch of C*C, zero on entry, zero again on
 * return) followed by sf_accumulate_counts into the caller's accumulators.  One crossing of the language boundary: at
 * validation batch size 1 the host side of two calls costs more than the 17 us of GPU work per image.  Where the integer-scale
 * kernel covers the shape (C in {2, 9, 19, 21}, scale 4 / 8) it is also ONE launch: the kernel's last CTA (atomic ticket kept
 * in bits 48..63 of counts[0] while the kernel runs -- a call counts fewer than 2^40 pixels) performs the accumulation.
 */
int sf_eval_update(const void* logits, int ldtype, const void* labels, int ydtype,
                   int B, int C, int h, int w, int H, int W, int ignore,
                   unsigned long long* counts, float* hist_f32, long long* hist_i64, long long* mat, void* stream);

/*
 * The decode head's 1x1 classifier, the producer of the low-resolution logits: replaces
 * nn.Conv2d(embed_dim, num_classes, 1) -- SegFormerHead.linear_pred (models/heads/segformer.py:39,57), UPerHead.conv_seg
 * (models/heads/upernet.py:32,50), FPNHead.conv_seg (models/heads/fpn.py:23,37) -- forward only:
 *     logits[b][c][i][j] = bias[c] + sum_k weight[c][k] * x[b][k][i][j]
 * x [B,K,h,w] and weight [C,K] in bf16 or fp16 (in_dtype; what autocast feeds the conv), bias [C] fp32 or NULL, logits
 * [B,C,h,w] in out_dtype (SF_F32 / SF_BF16 / SF_F16) -- the NCHW layout and dtypes sf_ce_onepass / sf_eval_hist read, so
 * nothing is converted in between.  tcgen05 tensor cores with the accumulator in tensor memory, operands staged by TMA.
 * Needs h*w % 8 == 0, K % 8 == 0, C <= 256 and 16-byte aligned x / weight; otherwise SF_ERR_UNSUPPORTED (fp32 inputs included:
 * keep torch's conv for those).
 */
int sf_head_classifier(const void* x, const void* weight, const float* bias, void* logits, int in_dtype, int out_dtype,
                       int B, int C, int K, int h, int w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEGFACTORY_H_ */
