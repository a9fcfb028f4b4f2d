// sf_onepass.cu -- dispatcher of the one-pass forward+backward walk (kernel: sf_onepass.cuh) and the plain-CE scale kernel.
#include "sf_onepass.cuh"

namespace sf {

int onepass_dispatch_f32(const OnepassArgs& a, int s);
int onepass_dispatch_bf16(const OnepassArgs& a, int s);
int onepass_dispatch_f16(const OnepassArgs& a, int s);
int onepass_chunk_dispatch_f32(const OnepassArgs& a, int s);    // many classes (sf_onepass_chunk.cuh)
int onepass_chunk_dispatch_bf16(const OnepassArgs& a, int s);
int onepass_chunk_dispatch_f16(const OnepassArgs& a, int s);

// The many-class walk is OPT-IN (SF_ONEPASS_CHUNK=1, read per call): at BASELINE configs[2] / [3] it executes fewer
// instructions than the forward + backward kernel pair (561 M vs 620 M warp instructions) but issues them more slowly
// (12 warps per SM, a CTA barrier per chunk: 1.05 + 0.055 ms vs 0.906 ms, profiles/r02_oc_*), so the default route for
// C > 24 stays sf_forward_chunk.cu + sf_backward_fast.cu.
static bool env_onepass_chunk() {
    const char* e = std::getenv("SF_ONEPASS_CHUNK");
    return e && e[0] == '1';
}

// grad = acc * (grad_out * inv_denom): plain CE (no loss vector, no top-k branch) completes its backward with this alone
template <typename LT>
__global__ void __launch_bounds__(256)
onepass_scale_kernel(const float* __restrict__ acc, LT* __restrict__ out, long long n, const float* __restrict__ grad_out,
                     const Workspace* __restrict__ ws) {
    scale_slice<LT, 16>(acc, out, n, grad_out, &ws->state, (long long)blockIdx.x * blockDim.x + threadIdx.x,
                        (long long)gridDim.x * blockDim.x);
}

bool onepass_supported(int ldtype, const void* logits, int C, int h, int w, int H, int W) {
    if (env_no_tma() || env_no_fast() || h <= 0 || w <= 0 || H % h || W % w) return false;
    const int s = H / h;
    if (s != W / w || (s != 4 && s != 8)) return false;
    const bool many = C > 24 && C <= 255 && env_onepass_chunk();                // 16-class chunks (sf_onepass_chunk.cuh)
    if (C != 19 && C != 21 && C != 9 && C != 2 && !many) return false;              // train_gpu.py:48-49 class counts that fit the registers
    const int esz = ldtype == SF_F32 ? 4 : 2;
    if (((size_t)w * esz) % 16 || ((uintptr_t)logits % 16)) return false;
    const int xal = 16 / esz;
    if (many) return w >= round_up(xal - 1 + (128 / s - 1) + 2, xal) && h >= 2;     // the narrow tile's two-row box
    const int bw = round_up(xal - 1 + (kOpThreads / s - 1) + 2, xal);
    return w >= bw && h >= 8 + 2;                                   // the 8-row tile's box
}

// SF_ERR_UNSUPPORTED when the shape is not covered (the caller then takes the forward / finalize / backward route).
int launch_onepass(const OnepassArgs& a) {
    if (!onepass_supported(a.ldtype, a.logits, a.C, a.h, a.w, a.H, a.W)) return SF_ERR_UNSUPPORTED;
    const int s = a.H / a.h;
    if (a.C > 24) {
        if (a.ldtype == SF_F32) return onepass_chunk_dispatch_f32(a, s);
        if (a.ldtype == SF_BF16) return onepass_chunk_dispatch_bf16(a, s);
        return onepass_chunk_dispatch_f16(a, s);
    }
    if (a.ldtype == SF_F32) return onepass_dispatch_f32(a, s);
    if (a.ldtype == SF_BF16) return onepass_dispatch_bf16(a, s);
    return onepass_dispatch_f16(a, s);
}

int launch_onepass_scale(const float* acc, void* out, int ldtype, long long n, const float* grad_out, const void* ws,
                         void* stream) {
    const int grid = (int)std::min<long long>(std::max<long long>((n / 4 + 255) / 256, 1), 8LL * sm_count());
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const Workspace* w = reinterpret_cast<const Workspace*>(ws);
    if (ldtype == SF_F32) onepass_scale_kernel<float><<<grid, 256, 0, st>>>(acc, reinterpret_cast<float*>(out), n, grad_out, w);
    else if (ldtype == SF_BF16) onepass_scale_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(acc, reinterpret_cast<__nv_bfloat16*>(out), n, grad_out, w);
    else onepass_scale_kernel<__half><<<grid, 256, 0, st>>>(acc, reinterpret_cast<__half*>(out), n, grad_out, w);
    return (int)cudaGetLastError();
}

}  // namespace sf
