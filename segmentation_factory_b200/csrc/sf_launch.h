// sf_launch.h -- host-side launcher interfaces shared by the translation units of libsegfactory.so.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>

#include <cstdlib>

#include <algorithm>

namespace sf {

// Dice walks (sf_dice.cu): the forward tile walk with a different per-pixel body.
enum { DICE_OFF = 0, DICE_STATS = 1, DICE_COEF = 2 };
struct DiceArgs {
    const float* lse_in;                           // [B*H*W] from sf_ce_forward
    unsigned long long* stats;                     // [B][C][3]  STATS: accumulated into
    const float* ta;                               // [B][C]     COEF
    const float* tb;                               // [B][C]     COEF
    float* coef_a;                                 // [B*H*W]    COEF
    float* coef_b;
};

struct FwdArgs {
    const void* logits; int ldtype;
    const void* labels; int ydtype;
    int B, C, h, w, H, W, ignore;
    const float* class_w;
    float thresh;
    float* loss_px; float* lse_px;
    unsigned long long* hist;
    void* ws;
    void* stream;
    bool softmax;          // false: argmax + histogram only (sf_eval_hist)
    bool force_no_tma;     // SF_NO_TMA=1 in the environment (debug / A-B measurement)
    int dice_mode;         // DICE_OFF: the CE / argmax forward
    DiceArgs dice;
    // sf_eval_update: a kernel that can do so folds `hist_f32 += hist; hist_i64 += hist; mat += hist; hist = 0` into its last
    // CTA and sets *folded_out = 1 (host side, at launch); otherwise the caller runs launch_accumulate_counts afterwards
    float* fold_f32; long long* fold_i64; long long* fold_mat; int* folded_out;
};
int launch_forward(const FwdArgs& a);
int launch_forward_fast(const FwdArgs& a);    // -1000: shape not covered by the integer-scale specialisation
int launch_forward_direct(const FwdArgs& a);  // h == H, w == W: streaming kernels, no interpolation; -1000: not covered
int launch_dice_direct(const FwdArgs& a);     // dice walks on same-size logits (a.dice_mode, a.dice); -1000: not covered
int launch_forward_chunk(const FwdArgs& a);   // integer scale, many classes (16-class chunks); -1000: not covered

struct BwdArgs {
    const void* logits; int ldtype;
    const void* labels; int ydtype;
    int B, C, h, w, H, W, ignore;
    const float* class_w;
    const float* loss_px; const float* lse_px;
    const float* grad_out;
    const void* ws;
    const float* coef_a; const float* coef_b; const float* dice_tb;   // non-null: CE + dice gradient (sf_dice.cu)
    void* grad_logits;
    void* stream;
    bool force_no_tma;
};
int launch_backward(const BwdArgs& a);
int launch_backward_direct(const BwdArgs& a);
int launch_backward_fast(const BwdArgs& a);   // -1000: shape not covered

// One-pass forward + backward (sf_onepass.cu): integer-scale heads with C <= 24.
struct OnepassArgs {
    const void* logits; int ldtype;
    const void* labels; int ydtype;
    int B, C, h, w, H, W, ignore;
    const float* class_w;
    float thresh;                  // < 0: plain CE
    float* loss_px;                // fused walk: written (required under OHEM); backward walk: read
    unsigned long long* hist;      // fused walk only, nullable
    float* grad_acc;               // fused walk: [B,C,h,w] fp32, unscaled (softmax - onehot) * keep * w_y sums
    const float* grad_out;         // backward walk
    void* grad_logits;             // backward walk: [B,C,h,w] in the logits' dtype, final
    void* ws;
    void* stream;
    int mode;                      // 0 = fused walk, 1 = backward walk (top-k fallback)
    int only_branch2;              // backward walk: no-op unless sf_loss_state.branch == 2
};
bool onepass_supported(int ldtype, const void* logits, int C, int h, int w, int H, int W);
int launch_onepass(const OnepassArgs& a);      // SF_ERR_UNSUPPORTED: shape not covered
int launch_onepass_scale(const float* acc, void* out, int ldtype, long long n, const float* grad_out, const void* ws,
                         void* stream);

// dice term of engine.criterion (sf_dice.cu)
size_t dice_ws_bytes(int B, int C);
int launch_dice_stats(const FwdArgs& a, const float* lse_px, void* dice_ws);
int launch_ce_dice_forward(const FwdArgs& a, void* dice_ws);   // a: as for sf_ce_forward (lse_px required)
int launch_dice_finalize(int B, int C, float eps, const void* ws, void* dice_ws, float* loss_out, void* stream);
int launch_dice_coef(const FwdArgs& a, const float* lse_px, const void* dice_ws, float* coef_a, float* coef_b);
const float* dice_tb_table(const void* dice_ws, int B, int C);

int launch_finalize(const float* loss_px, long long n_px, float thresh, void* ws, float* loss_out, void* stream);
int launch_focal(float* loss_px, long long n_px, float alpha, float gamma, int size_average, void* ws,
                 float* loss_out, void* stream);
int launch_accumulate_counts(unsigned long long* counts, float* hist_f32, long long* hist_i64, long long* mat, int n, void* stream);
int launch_label_lut(const unsigned char* raw, long long n, const unsigned char* lut, void* out, int out_dtype, void* stream);
int launch_hist_pairs(const void* a, int adtype, const long long* b, long long n, int C,
                      unsigned long long* mat, void* stream);

// 3-D tiled tensor map over logits viewed as [planes = B*C][h][w]; box = [box_c][box_h][box_w]
int make_logits_tmap(CUtensorMap* map, const void* logits, int ldtype, int planes, int h, int w,
                     int box_c, int box_h, int box_w);
// Tiled tensor map of rank 2 or 3 over 2-byte elements (bf16 / fp16): dims / box innermost first, strides in bytes for
// dimensions 1.., optional 128-byte swizzle (sf_head_gemm.cu)
int make_tmap_generic(CUtensorMap* map, const void* base, int dtype, int rank, const unsigned long long* dims,
                      const unsigned long long* strides, const unsigned* box, int swizzle128);
int launch_head_classifier(const void* x, const void* w, const float* bias, void* out, int in_dtype, int out_dtype,
                           int B, int C, int K, long long N, void* stream);
int sm_count();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) costs ~1.5 us per call: do it once per kernel instantiation and device
// (`slot` is a function-local static of the launcher; a larger limit set earlier also covers a smaller request).
struct SmemLimitOnce { int bytes[64] = {}; };
template <typename K>
inline cudaError_t set_smem_limit_once(K kernel, size_t smem, SmemLimitOnce& slot) {
    int dev = 0;
    cudaGetDevice(&dev);
    int& have = slot.bytes[dev & 63];
    if (have >= (int)smem) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) have = (int)smem;
    return e;
}
// SF_NO_XCHG=1 (read per launch): the gather-form kernels recompute the halo row group of a tile instead of exchanging the
// boundary row with the tile above through two float atomics (A/B measurement, tests)
inline bool env_no_xchg() {
    const char* e = std::getenv("SF_NO_XCHG");
    return e && e[0] == '1';
}
bool env_no_tma();
bool env_no_fast();   // SF_NO_FAST=1: always take the generic kernels (A/B measurement, tests)

}  // namespace sf
