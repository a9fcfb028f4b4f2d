// sf_api.cu -- the extern "C" surface declared in include/segfactory.h, plus the TMA descriptor helper.
#include <cstdlib>
#include <cstring>

#include "sf_common.cuh"
#include "sf_launch.h"

namespace sf {

int sm_count() {
    static int cached[64] = {};                                      // per device; the attribute query is a driver round trip
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cached[dev & 63] > 0) return cached[dev & 63];
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cached[dev & 63] = n;
    return n;
}

bool env_no_tma() {
    const char* e = std::getenv("SF_NO_TMA");
    return e && e[0] == '1';
}
bool env_no_fast() {
    const char* e = std::getenv("SF_NO_FAST");
    return e && e[0] == '1';
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is not linked: the encoder is fetched through the runtime (no -lcuda, works with any driver >= 12.0)
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
        if (q != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

int make_logits_tmap(CUtensorMap* map, const void* logits, int ldtype, int planes, int h, int w,
                     int box_c, int box_h, int box_w) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return SF_ERR_DRIVER;
    // The encoder is a driver-API call: it needs the primary context current on THIS host thread.  A thread that has
    // only ever been handed work (autograd's backward thread) may not have it bound yet -- seen under compute-sanitizer
    // as CUDA_ERROR_INVALID_CONTEXT.  Bind it once per thread with a runtime call.
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
    const cuuint64_t esz = ldtype == SF_F32 ? 4 : 2;
    const CUtensorMapDataType dt = ldtype == SF_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : ldtype == SF_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                     : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)planes};
    const cuuint64_t gstr[2] = {(cuuint64_t)w * esz, (cuuint64_t)w * h * esz};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_c};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, dt, 3, const_cast<void*>(logits), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_ERROR_INVALID_CONTEXT || r == CUDA_ERROR_NOT_INITIALIZED) {
        // The encoder is a driver-API call and needs the primary context current on THIS host thread; a thread that has
        // only been handed work (autograd's backward thread) may not have it yet.  cudaFree(0) binds it; retry once.
        cudaFree(nullptr);
        r = enc(map, dt, 3, const_cast<void*>(logits), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    return r == CUDA_SUCCESS ? 0 : SF_ERR_DRIVER;
}

int make_tmap_generic(CUtensorMap* map, const void* base, int dtype, int rank, const unsigned long long* dims,
                      const unsigned long long* strides, const unsigned* box, int swizzle128) {
    EncodeTiledFn enc = encode_fn();
    if (!enc || rank < 2 || rank > 3) return SF_ERR_DRIVER;
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
    const CUtensorMapDataType dt = dtype == SF_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype == SF_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    cuuint64_t gdim[3]; cuuint64_t gstr[2]; cuuint32_t bx[3]; cuuint32_t estr[3] = {1, 1, 1};
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
    for (int i = 0; i < rank - 1; ++i) gstr[i] = strides[i];
    const CUresult r = enc(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : SF_ERR_DRIVER;
}

}  // namespace sf

extern "C" {

int sf_version(void) { return SF_VERSION; }

const char* sf_strerror(int code) {
    switch (code) {
        case SF_OK: return "ok";
        case SF_ERR_ARG: return "invalid argument (null pointer, non-positive extent or unknown dtype)";
        case SF_ERR_ALIGN: return "pointer not aligned to its element size";
        case SF_ERR_TOO_MANY_CLASSES: return "C*C int32 histogram does not fit in shared memory";
        case SF_ERR_WORKSPACE: return "workspace smaller than sf_workspace_bytes()";
        case SF_ERR_DRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
        case SF_ERR_SHAPE: return "extents overflow the kernels' tile arithmetic";
        case SF_ERR_UNSUPPORTED: return "shape outside the one-pass kernels (use sf_ce_forward / sf_ce_backward)";
        default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "unknown segfactory error";
    }
}

size_t sf_workspace_bytes(void) { return sizeof(sf::Workspace); }

int sf_eval_hist(const void* logits, int ldtype, const void* labels, int ydtype, int B, int C, int h, int w, int H,
                 int W, int ignore, unsigned long long* hist, void* stream) {
    if (!hist) return SF_ERR_ARG;
    sf::FwdArgs a{};
    a.logits = logits; a.ldtype = ldtype; a.labels = labels; a.ydtype = ydtype;
    a.B = B; a.C = C; a.h = h; a.w = w; a.H = H; a.W = W; a.ignore = ignore;
    a.hist = hist; a.stream = stream; a.softmax = false; a.force_no_tma = sf::env_no_tma();
    return sf::launch_forward(a);
}

int sf_ce_forward(const void* logits, int ldtype, const void* labels, int ydtype, int B, int C, int h, int w, int H,
                  int W, int ignore, const float* class_w, float ohem_thresh, float* loss_px, float* lse_px,
                  unsigned long long* hist, void* workspace, size_t workspace_bytes, void* stream) {
    if (!workspace) return SF_ERR_ARG;
    if (workspace_bytes < sizeof(sf::Workspace)) return SF_ERR_WORKSPACE;
    if (ohem_thresh >= 0.f && !loss_px) return SF_ERR_ARG;          // the top-k fallback needs the loss vector
    sf::FwdArgs a{};
    a.logits = logits; a.ldtype = ldtype; a.labels = labels; a.ydtype = ydtype;
    a.B = B; a.C = C; a.h = h; a.w = w; a.H = H; a.W = W; a.ignore = ignore;
    a.class_w = class_w; a.thresh = ohem_thresh; a.loss_px = loss_px; a.lse_px = lse_px;
    a.hist = hist; a.ws = workspace; a.stream = stream; a.softmax = true; a.force_no_tma = sf::env_no_tma();
    return sf::launch_forward(a);
}

int sf_loss_finalize(const float* loss_px, long long n_px, float ohem_thresh, void* workspace, size_t workspace_bytes,
                     float* loss_out, void* stream) {
    if (!workspace || n_px <= 0) return SF_ERR_ARG;
    if (workspace_bytes < sizeof(sf::Workspace)) return SF_ERR_WORKSPACE;
    if (ohem_thresh >= 0.f && !loss_px) return SF_ERR_ARG;
    return sf::launch_finalize(loss_px, n_px, ohem_thresh, workspace, loss_out, stream);
}

int sf_focal_finalize(float* loss_px, long long n_px, float alpha, float gamma, int size_average, void* workspace,
                      size_t workspace_bytes, float* loss_out, void* stream) {
    if (!workspace || !loss_px || n_px <= 0) return SF_ERR_ARG;
    if (workspace_bytes < sizeof(sf::Workspace)) return SF_ERR_WORKSPACE;
    return sf::launch_focal(loss_px, n_px, alpha, gamma, size_average, workspace, loss_out, stream);
}

int sf_ce_backward(const void* logits, int ldtype, const void* labels, int ydtype, int B, int C, int h, int w, int H,
                   int W, int ignore, const float* class_w, const float* loss_px, const float* lse_px,
                   const float* grad_out, const void* workspace, void* grad_logits, void* stream) {
    if (!workspace || !grad_logits || !lse_px) return SF_ERR_ARG;
    sf::BwdArgs a{};
    a.logits = logits; a.ldtype = ldtype; a.labels = labels; a.ydtype = ydtype;
    a.B = B; a.C = C; a.h = h; a.w = w; a.H = H; a.W = W; a.ignore = ignore;
    a.class_w = class_w; a.loss_px = loss_px; a.lse_px = lse_px; a.grad_out = grad_out; a.ws = workspace;
    a.grad_logits = grad_logits; a.stream = stream; a.force_no_tma = sf::env_no_tma();
    return sf::launch_backward(a);
}

int sf_ce_onepass_supported(int ldtype, const void* logits, int C, int h, int w, int H, int W) {
    return sf::onepass_supported(ldtype, logits, C, h, w, H, W) ? 1 : 0;
}

static int onepass_check(const void* logits, int ldtype, const void* labels, int ydtype, int B, int C, int h, int w, int H,
                         int W) {
    if (!logits || !labels || B <= 0 || C <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return SF_ERR_ARG;
    if (ldtype < SF_F32 || ldtype > SF_F16 || ydtype < SF_I64 || ydtype > SF_U8) return SF_ERR_ARG;
    if (ydtype == SF_I64 && ((uintptr_t)labels % 8)) return SF_ERR_ALIGN;
    if ((long long)B * H * W >= (1LL << 40) || (long long)B * C >= (1LL << 31)) return SF_ERR_SHAPE;
    return SF_OK;
}

int sf_ce_onepass(const void* logits, int ldtype, const void* labels, int ydtype, int B, int C, int h, int w, int H,
                  int W, int ignore, const float* class_w, float ohem_thresh, float* loss_px, unsigned long long* hist,
                  float* grad_acc, void* workspace, size_t workspace_bytes, void* stream) {
    if (!workspace || !grad_acc) return SF_ERR_ARG;
    if (workspace_bytes < sizeof(sf::Workspace)) return SF_ERR_WORKSPACE;
    if (ohem_thresh >= 0.f && !loss_px) return SF_ERR_ARG;          // the top-k fallback needs the loss vector
    const int rc = onepass_check(logits, ldtype, labels, ydtype, B, C, h, w, H, W);
    if (rc != SF_OK) return rc;
    sf::OnepassArgs a{};
    a.logits = logits; a.ldtype = ldtype; a.labels = labels; a.ydtype = ydtype;
    a.B = B; a.C = C; a.h = h; a.w = w; a.H = H; a.W = W; a.ignore = ignore;
    a.class_w = class_w; a.thresh = ohem_thresh; a.loss_px = loss_px; a.hist = hist; a.grad_acc = grad_acc;
    a.ws = workspace; a.stream = stream; a.mode = 0;
    return sf::launch_onepass(a);
}

int sf_ce_onepass_backward(const void* logits, int ldtype, const void* labels, int ydtype, int B, int C, int h, int w,
                           int H, int W, int ignore, const float* class_w, const float* loss_px, const float* grad_acc,
                           const float* grad_out, const void* workspace, void* grad_logits, void* stream) {
    if (!workspace || !grad_acc || !grad_logits) return SF_ERR_ARG;
    int rc = onepass_check(logits, ldtype, labels, ydtype, B, C, h, w, H, W);
    if (rc != SF_OK) return rc;
    if (loss_px) {
        // OHEM: one launch whose CTAs either scale their slice of grad_acc (branch 1) or re-walk their tile (branch 2)
        sf::OnepassArgs a{};
        a.logits = logits; a.ldtype = ldtype; a.labels = labels; a.ydtype = ydtype;
        a.B = B; a.C = C; a.h = h; a.w = w; a.H = H; a.W = W; a.ignore = ignore;
        a.class_w = class_w; a.thresh = 0.f; a.loss_px = const_cast<float*>(loss_px); a.grad_out = grad_out;
        a.grad_acc = const_cast<float*>(grad_acc); a.grad_logits = grad_logits; a.ws = const_cast<void*>(workspace);
        a.stream = stream; a.mode = 1; a.only_branch2 = 1;
        return sf::launch_onepass(a);
    }
    // plain CE keeps no loss vector and has no top-k branch: the scale alone
    return sf::launch_onepass_scale(grad_acc, grad_logits, ldtype, (long long)B * C * h * w, grad_out, workspace, stream);
}

size_t sf_dice_workspace_bytes(int B, int C) { return (B > 0 && C > 0) ? sf::dice_ws_bytes(B, C) : 0; }

static sf::FwdArgs dice_args(const void* logits, int ldtype, const void* labels, int ydtype, int B, int C, int h, int w,
                             int H, int W, int ignore, void* stream) {
    sf::FwdArgs a{};
    a.logits = logits; a.ldtype = ldtype; a.labels = labels; a.ydtype = ydtype;
    a.B = B; a.C = C; a.h = h; a.w = w; a.H = H; a.W = W; a.ignore = ignore;
    a.stream = stream; a.softmax = false; a.force_no_tma = sf::env_no_tma();
    return a;
}

int sf_dice_stats(const void* logits, int ldtype, const void* labels, int ydtype, int B, int C, int h, int w, int H,
                  int W, int ignore, const float* lse_px, void* dice_ws, size_t dice_ws_bytes, void* stream) {
    if (!lse_px || !dice_ws || B <= 0 || C <= 0) return SF_ERR_ARG;
    if (dice_ws_bytes < sf::dice_ws_bytes(B, C)) return SF_ERR_WORKSPACE;
    if ((uintptr_t)dice_ws % 8) return SF_ERR_ALIGN;
    return sf::launch_dice_stats(dice_args(logits, ldtype, labels, ydtype, B, C, h, w, H, W, ignore, stream), lse_px, dice_ws);
}

int sf_ce_dice_forward(const void* logits, int ldtype, const void* labels, int ydtype, int B, int C, int h, int w, int H,
                       int W, int ignore, const float* class_w, float* lse_px, void* workspace, size_t workspace_bytes,
                       void* dice_ws, size_t dice_ws_bytes, void* stream) {
    if (!workspace || !lse_px || !dice_ws || B <= 0 || C <= 0) return SF_ERR_ARG;
    if (workspace_bytes < sizeof(sf::Workspace) || dice_ws_bytes < sf::dice_ws_bytes(B, C)) return SF_ERR_WORKSPACE;
    if ((uintptr_t)dice_ws % 8) return SF_ERR_ALIGN;
    sf::FwdArgs a = dice_args(logits, ldtype, labels, ydtype, B, C, h, w, H, W, ignore, stream);
    a.class_w = class_w; a.thresh = -1.0f; a.lse_px = lse_px; a.ws = workspace; a.softmax = true;
    return sf::launch_ce_dice_forward(a, dice_ws);
}

int sf_dice_finalize(int B, int C, float epsilon, const void* workspace, void* dice_ws, size_t dice_ws_bytes,
                     float* loss_out, void* stream) {
    if (!dice_ws || B <= 0 || C <= 0 || (long long)B * C >= (1LL << 24)) return SF_ERR_ARG;
    if (dice_ws_bytes < sf::dice_ws_bytes(B, C)) return SF_ERR_WORKSPACE;
    return sf::launch_dice_finalize(B, C, epsilon, workspace, dice_ws, loss_out, stream);
}

int sf_dice_coef(const void* logits, int ldtype, const void* labels, int ydtype, int B, int C, int h, int w, int H,
                 int W, int ignore, const float* class_w, const float* lse_px, const void* workspace,
                 const void* dice_ws, size_t dice_ws_bytes, float* coef_a_px, float* coef_b_px, void* stream) {
    if (!lse_px || !dice_ws || !coef_a_px || !coef_b_px || B <= 0 || C <= 0) return SF_ERR_ARG;
    if (dice_ws_bytes < sf::dice_ws_bytes(B, C)) return SF_ERR_WORKSPACE;
    sf::FwdArgs a = dice_args(logits, ldtype, labels, ydtype, B, C, h, w, H, W, ignore, stream);
    a.class_w = class_w; a.ws = const_cast<void*>(workspace);
    return sf::launch_dice_coef(a, lse_px, dice_ws, coef_a_px, coef_b_px);
}

int sf_dice_backward(const void* logits, int ldtype, const void* labels, int ydtype, int B, int C, int h, int w, int H,
                     int W, int ignore, const float* lse_px, const float* coef_a_px, const float* coef_b_px,
                     const void* dice_ws, size_t dice_ws_bytes, const float* grad_out, void* grad_logits, void* stream) {
    if (!lse_px || !dice_ws || !coef_a_px || !coef_b_px || !grad_logits || B <= 0 || C <= 0) return SF_ERR_ARG;
    if (dice_ws_bytes < sf::dice_ws_bytes(B, C)) return SF_ERR_WORKSPACE;
    sf::BwdArgs a{};
    a.logits = logits; a.ldtype = ldtype; a.labels = labels; a.ydtype = ydtype;
    a.B = B; a.C = C; a.h = h; a.w = w; a.H = H; a.W = W; a.ignore = ignore;
    a.lse_px = lse_px; a.grad_out = grad_out; a.coef_a = coef_a_px; a.coef_b = coef_b_px;
    a.dice_tb = sf::dice_tb_table(dice_ws, B, C);
    a.grad_logits = grad_logits; a.stream = stream; a.force_no_tma = sf::env_no_tma();
    return sf::launch_backward(a);
}

int sf_label_lut(const unsigned char* raw, long long n, const unsigned char* lut, void* out, int out_dtype, void* stream) {
    if (n < 0 || (out_dtype != SF_I64 && out_dtype != SF_U8)) return SF_ERR_ARG;
    if (n == 0) return SF_OK;                                       // an empty tensor has no storage to point at
    if (!raw || !lut || !out) return SF_ERR_ARG;
    if (out_dtype == SF_I64 && ((uintptr_t)out % 8)) return SF_ERR_ALIGN;
    return sf::launch_label_lut(raw, n, lut, out, out_dtype, stream);
}

int sf_accumulate_counts(unsigned long long* counts, float* hist_f32, long long* hist_i64, long long* mat, int n,
                         void* stream) {
    if (!counts || n <= 0) return SF_ERR_ARG;
    return sf::launch_accumulate_counts(counts, hist_f32, hist_i64, mat, n, stream);
}

int sf_eval_update(const void* logits, int ldtype, const void* labels, int ydtype, int B, int C, int h, int w, int H, int W,
                   int ignore, unsigned long long* counts, float* hist_f32, long long* hist_i64, long long* mat, void* stream) {
    if (!counts) return SF_ERR_ARG;
    sf::FwdArgs a{};
    a.logits = logits; a.ldtype = ldtype; a.labels = labels; a.ydtype = ydtype;
    a.B = B; a.C = C; a.h = h; a.w = w; a.H = H; a.W = W; a.ignore = ignore;
    a.hist = counts; a.stream = stream; a.softmax = false; a.force_no_tma = sf::env_no_tma();
    int folded = 0;                                                 // set when the kernel's last CTA did the accumulation itself
    a.fold_f32 = hist_f32; a.fold_i64 = hist_i64; a.fold_mat = mat; a.folded_out = &folded;
    const int rc = sf::launch_forward(a);
    if (rc != SF_OK || folded) return rc;
    return sf::launch_accumulate_counts(counts, hist_f32, hist_i64, mat, C * C, stream);
}

int sf_head_classifier(const void* x, const void* weight, const float* bias, void* logits, int in_dtype, int out_dtype,
                       int B, int C, int K, int h, int w, void* stream) {
    if (h <= 0 || w <= 0) return SF_ERR_ARG;
    return sf::launch_head_classifier(x, weight, bias, logits, in_dtype, out_dtype, B, C, K, (long long)h * w, stream);
}

int sf_hist_pairs(const void* a, int adtype, const long long* b, long long n, int C, unsigned long long* mat,
                  void* stream) {
    if (!a || !b || !mat || n < 0 || C <= 0) return SF_ERR_ARG;
    if (adtype < SF_I64 || adtype > SF_U8) return SF_ERR_ARG;
    if (n == 0) return SF_OK;
    return sf::launch_hist_pairs(a, adtype, b, n, C, mat, stream);
}

}  // extern "C"
