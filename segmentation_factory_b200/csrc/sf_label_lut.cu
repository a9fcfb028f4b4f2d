// sf_label_lut.cu -- label ingest: raw uint8 annotation ids -> training ids through a 256-entry table.
//
// Replaces the per-sample numpy gathers of the datasets (datasets/cityscapes.py:92,130-131 `id_to_train_id[np.array(target)]`,
// datasets/coco_stuff.py:66-69,98-100 `label_map[label]`) and the int64 collate behind them: the annotation PNG's bytes go to
// the device as they are (1 B/px instead of 8 B/px over PCIe) and are mapped there.  Output uint8 (what the loss / metric
// kernels read at 1 B/px) or int64 (the reference's dtype).  Pure streaming, table in shared memory, every warp-level
// store instruction covers contiguous bytes.  Bound: HBM, 2 B/px (uint8 out) or 9 B/px (int64 out).
#include "sf_common.cuh"
#include "sf_launch.h"

namespace sf {

constexpr int kLutThreads = 256;

// uint8 out: 16 pixels per thread (one 128-bit load, one 128-bit store).
__global__ void __launch_bounds__(kLutThreads)
label_lut_u8_kernel(const uint4* __restrict__ raw, long long n16, const unsigned char* __restrict__ lut, uint4* __restrict__ out) {
    __shared__ unsigned char lut_s[256];
    lut_s[threadIdx.x] = __ldg(lut + threadIdx.x);                  // kLutThreads == 256
    __syncthreads();
    auto map16 = [&](const uint4 v) -> uint4 {
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
        unsigned r[4];
        #pragma unroll
        for (int k = 0; k < 4; ++k)
            r[k] = (unsigned)lut_s[w[k] & 0xffu] | ((unsigned)lut_s[(w[k] >> 8) & 0xffu] << 8) |
                   ((unsigned)lut_s[(w[k] >> 16) & 0xffu] << 16) | ((unsigned)lut_s[w[k] >> 24] << 24);
        return make_uint4(r[0], r[1], r[2], r[3]);
    };
    const long long stride = (long long)gridDim.x * kLutThreads;
    long long i = (long long)blockIdx.x * kLutThreads + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {                 // four independent 128-bit loads in flight
        uint4 v[4];
        #pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = __ldg(raw + i + k * stride);
        #pragma unroll
        for (int k = 0; k < 4; ++k) out[i + k * stride] = map16(v[k]);
    }
    for (; i < n16; i += stride) out[i] = map16(__ldg(raw + i));
}

// int64 out: the 8 B/px store stream is what matters, so a thread owns 2 pixels = one 16-byte store and a warp's store
// instruction covers 512 contiguous bytes (16 pixels per thread would put the lanes 128 B apart: measured 1.4 TB/s).
// Four independent pairs per iteration keep enough stores in flight.
__global__ void __launch_bounds__(kLutThreads)
label_lut_i64_kernel(const unsigned short* __restrict__ raw, long long n2, const unsigned char* __restrict__ lut,
                     longlong2* __restrict__ out) {
    __shared__ unsigned char lut_s[256];
    lut_s[threadIdx.x] = __ldg(lut + threadIdx.x);
    __syncthreads();
    const long long stride = (long long)gridDim.x * kLutThreads;
    long long i = (long long)blockIdx.x * kLutThreads + threadIdx.x;
    for (; i + 3 * stride < n2; i += 4 * stride) {
        unsigned v[4];
        #pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = __ldg(raw + i + k * stride);
        #pragma unroll
        for (int k = 0; k < 4; ++k) out[i + k * stride] = make_longlong2(lut_s[v[k] & 0xffu], lut_s[v[k] >> 8]);
    }
    for (; i < n2; i += stride) {
        const unsigned v = __ldg(raw + i);
        out[i] = make_longlong2(lut_s[v & 0xffu], lut_s[v >> 8]);
    }
}

template <bool OUT64>
__global__ void __launch_bounds__(kLutThreads)
label_lut_scalar_kernel(const unsigned char* __restrict__ raw, long long first, long long n, const unsigned char* __restrict__ lut,
                        void* __restrict__ out) {
    __shared__ unsigned char lut_s[256];
    lut_s[threadIdx.x] = __ldg(lut + threadIdx.x);
    __syncthreads();
    for (long long i = first + (long long)blockIdx.x * kLutThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kLutThreads) {
        const unsigned char t = lut_s[__ldg(raw + i)];
        if (OUT64) reinterpret_cast<long long*>(out)[i] = t; else reinterpret_cast<unsigned char*>(out)[i] = t;
    }
}

int launch_label_lut(const unsigned char* raw, long long n, const unsigned char* lut, void* out, int out_dtype, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool out64 = out_dtype == SF_I64;
    const int max_grid = 8 * sm_count();
    long long done = 0;
    if (out64) {
        if (((uintptr_t)raw % 2) == 0 && ((uintptr_t)out % 16) == 0 && n >= 2) {
            const long long n2 = n / 2;
            const int grid = (int)std::min<long long>((n2 + kLutThreads - 1) / kLutThreads, max_grid);
            label_lut_i64_kernel<<<grid, kLutThreads, 0, st>>>(reinterpret_cast<const unsigned short*>(raw), n2, lut,
                                                               reinterpret_cast<longlong2*>(out));
            done = n2 * 2;
        }
    } else if (((uintptr_t)raw % 16) == 0 && ((uintptr_t)out % 16) == 0 && n >= 16) {
        const long long n16 = n / 16;
        const int grid = (int)std::min<long long>((n16 + kLutThreads - 1) / kLutThreads, max_grid);
        label_lut_u8_kernel<<<grid, kLutThreads, 0, st>>>(reinterpret_cast<const uint4*>(raw), n16, lut, reinterpret_cast<uint4*>(out));
        done = n16 * 16;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    if (done < n) {                                                  // unaligned views and tails
        const int grid = (int)std::min<long long>((n - done + kLutThreads - 1) / kLutThreads, max_grid);
        if (out64) label_lut_scalar_kernel<true><<<grid, kLutThreads, 0, st>>>(raw, done, n, lut, out);
        else label_lut_scalar_kernel<false><<<grid, kLutThreads, 0, st>>>(raw, done, n, lut, out);
    }
    return (int)cudaGetLastError();
}

}  // namespace sf
