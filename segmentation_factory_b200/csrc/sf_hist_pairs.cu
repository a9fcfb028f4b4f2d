// sf_hist_pairs.cu -- ConfusionMatrix.update(a, b) for callers that already hold index vectors.
//
// Replaces util/utils.py:99-109: k = (a >= 0) & (a < n); mat += bincount(n * a[k] + b[k], minlength=n*n).
// Grid-stride pass with warp-aggregated atomics into a shared-memory C x C histogram when it fits,
// else straight to global; one 64-bit flush per CTA.
#include "sf_common.cuh"
#include "sf_launch.h"

namespace sf {

constexpr int kPairThreads = 256;

template <bool SMEM>
__global__ void __launch_bounds__(kPairThreads)
hist_pairs_kernel(const void* __restrict__ a, int a_u8, const long long* __restrict__ b, long long n, int C,
                  unsigned long long* __restrict__ mat) {
    extern __shared__ int sh[];
    if (SMEM) {
        for (int k = threadIdx.x; k < C * C; k += kPairThreads) sh[k] = 0;
        __syncthreads();
    }
    const long long stride = (long long)gridDim.x * kPairThreads;
    const long long n_round = (n + 31) / 32 * 32;                    // keep whole warps in the loop for match_any
    for (long long i = (long long)blockIdx.x * kPairThreads + threadIdx.x; i < n_round; i += stride) {
        int key = -1;
        if (i < n) {
            const long long av = ld_label(a, a_u8, i);
            const long long bv = __ldg(b + i);
            if (av >= 0 && av < C && bv >= 0 && bv < C) key = (int)(av * C + bv);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) {
            if (SMEM) atomicAdd(&sh[key], __popc(peers));
            else atomicAdd(mat + key, (unsigned long long)__popc(peers));
        }
    }
    if (SMEM) {
        __syncthreads();
        for (int k = threadIdx.x; k < C * C; k += kPairThreads)
            if (sh[k]) atomicAdd(mat + k, (unsigned long long)sh[k]);
    }
}

__global__ void accumulate_counts_kernel(unsigned long long* __restrict__ counts, float* __restrict__ hist_f32,
                                         long long* __restrict__ hist_i64, long long* __restrict__ mat, int n) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const unsigned long long c = counts[k];
        if (hist_f32) hist_f32[k] = __fadd_rn(hist_f32[k], __ull2float_rn(c));   // torch: float32 += int64 casts, then adds in fp32
        if (hist_i64) hist_i64[k] += (long long)c;
        if (mat) mat[k] += (long long)c;
        counts[k] = 0ull;
    }
}

int launch_accumulate_counts(unsigned long long* counts, float* hist_f32, long long* hist_i64, long long* mat, int n, void* stream) {
    const int grid = std::min((n + 255) / 256, 64);
    accumulate_counts_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(counts, hist_f32, hist_i64, mat, n);
    return (int)cudaGetLastError();
}

int launch_hist_pairs(const void* a, int adtype, const long long* b, long long n, int C,
                      unsigned long long* mat, void* stream) {
    const size_t smem = (size_t)C * C * 4;
    const long long want = (n + kPairThreads - 1) / kPairThreads;
    const int grid = (int)std::min<long long>(want, 2LL * sm_count());
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (smem <= 200 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(hist_pairs_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        hist_pairs_kernel<true><<<grid, kPairThreads, smem, st>>>(a, adtype == SF_U8, b, n, C, mat);
    } else {
        hist_pairs_kernel<false><<<grid, kPairThreads, 0, st>>>(a, adtype == SF_U8, b, n, C, mat);
    }
    return (int)cudaGetLastError();
}

}  // namespace sf
