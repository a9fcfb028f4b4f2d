// sf_onepass.cu -- dispatcher of the one-pass forward+backward walk (kernel: sf_onepass.cuh) and the plain-CE scale kernel.
#include "sf_onepass.cuh"

namespace sf {

int onepass_dispatch_f32(const OnepassArgs& a, int s);
int onepass_dispatch_bf16(const OnepassArgs& a, int s);
int onepass_dispatch_f16(const OnepassArgs& a, int s);

// grad = acc * (grad_out * inv_denom): plain CE (no loss vector, no top-k branch) completes its backward with this alone
template <typename LT>
__global__ void __launch_bounds__(256)
onepass_scale_kernel(const float* __restrict__ acc, LT* __restrict__ out, long long n, const float* __restrict__ grad_out,
                     const Workspace* __restrict__ ws) {
    if (ws->state.branch == 2) return;
    const float s = (grad_out ? __ldg(grad_out) : 1.0f) * ws->state.inv_denom;
    scale_slice<LT>(acc, out, n, s, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

bool onepass_supported(int ldtype, const void* logits, int C, int h, int w, int H, int W) {
    if (env_no_tma() || env_no_fast() || h <= 0 || w <= 0 || H % h || W % w) return false;
    const int s = H / h;
    if (s != W / w || (s != 4 && s != 8)) return false;
    if (C != 19 && C != 21 && C != 9 && C != 2) return false;        // train_gpu.py:48-49 class counts that fit the registers
    const int esz = ldtype == SF_F32 ? 4 : 2;
    if (((size_t)w * esz) % 16 || ((uintptr_t)logits % 16)) return false;
    const int xal = 16 / esz;
    const int bw = round_up(xal - 1 + (kOpThreads / s - 1) + 2, xal);
    return w >= bw && h >= 8 + 2;                                   // the 8-row tile's box
}

// SF_ERR_UNSUPPORTED when the shape is not covered (the caller then takes the forward / finalize / backward route).
int launch_onepass(const OnepassArgs& a) {
    if (!onepass_supported(a.ldtype, a.logits, a.C, a.h, a.w, a.H, a.W)) return SF_ERR_UNSUPPORTED;
    const int s = a.H / a.h;
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
