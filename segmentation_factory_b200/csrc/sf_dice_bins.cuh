// sf_dice_bins.cuh -- shared-memory accumulators of the dice sums of ONE image (sf_dice.cu).
//
// Integer accumulation keeps the sums independent of arrival order (bitwise reproducible).  64-bit shared atomics
// compile to a compare-and-swap spin loop (ATOMS.CAST.SPIN.64; measured: 4.4 tries per add with 16 warps on 16 bins),
// so every quantity is kept as 16-bit limbs in 32-bit words and every update is a fire-and-forget ATOMS.ADD:
//   sum p_c   (2^40 fixed point, <= 2^48 per warp-level add)  -> limbs S0, S1, S2
//   sum p_y   (2^26 fixed point, <= 2^31 per warp-level add)  -> limbs I0, I1
//   count                                                      -> N
// A limb word overflows after 2^16 adds: callers flush (-> 64-bit global atomics, native RED.64) at least every
// kDiceFlushAdds warp-level adds per bin.
#pragma once

#include "sf_common.cuh"

namespace sf {

constexpr int kDiceBinWords = 6;                   // 32-bit words per class: S0 S1 S2 I0 I1 N
constexpr int kDiceFlushAdds = 32768;

struct DiceBins {
    unsigned* w;                                   // [6][C]
    int C;
    __device__ __forceinline__ void zero(int tid, int nthreads) const {
        for (int k = tid; k < kDiceBinWords * C; k += nthreads) w[k] = 0u;
    }
    __device__ __forceinline__ void add_sum(int c, unsigned long long tot) const {
        atomicAdd(&w[c], (unsigned)(tot & 0xFFFFull));
        atomicAdd(&w[C + c], (unsigned)((tot >> 16) & 0xFFFFull));
        const unsigned top = (unsigned)(tot >> 32);
        if (top) atomicAdd(&w[2 * C + c], top);
    }
    __device__ __forceinline__ void add_inter(int c, unsigned sum, unsigned cnt) const {
        atomicAdd(&w[3 * C + c], sum & 0xFFFFu);
        atomicAdd(&w[4 * C + c], sum >> 16);
        atomicAdd(&w[5 * C + c], cnt);
    }
    // CTA-uniform: bins of image b -> stats[b][c][3] (64-bit global atomics), bins = 0
    __device__ __forceinline__ void flush(unsigned long long* stats, int b, int tid, int nthreads) const {
        __syncthreads();
        for (int c = tid; c < C; c += nthreads) {
            const unsigned long long s = (unsigned long long)w[c] + ((unsigned long long)w[C + c] << 16) +
                                         ((unsigned long long)w[2 * C + c] << 32);
            const unsigned long long i = (unsigned long long)w[3 * C + c] + ((unsigned long long)w[4 * C + c] << 16);
            const unsigned long long n = w[5 * C + c];
            unsigned long long* o = stats + ((long long)b * C + c) * 3;
            if (s) atomicAdd(o, s);
            if (i) atomicAdd(o + 1, i);
            if (n) atomicAdd(o + 2, n);
            #pragma unroll
            for (int k = 0; k < kDiceBinWords; ++k) w[k * C + c] = 0u;
        }
        __syncthreads();
    }
};

}  // namespace sf
