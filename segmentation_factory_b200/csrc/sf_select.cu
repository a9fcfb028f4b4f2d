// sf_select.cu -- loss reduction + OHEM k-th-largest radix select, entirely on device.
//
// Replaces util/losses.py:36 (weighted mean) and util/losses.py:54-61:
//     n_min = labels[labels != ignore].numel() // 16          (host sync in the reference)
//     loss_hard = loss[loss > thresh]                          (host sync)
//     if loss_hard.numel() < n_min: loss_hard, _ = loss.topk(n_min)   (python branch on a device count)
//     return mean(loss_hard)
// One cooperative launch: every CTA reduces the per-CTA partial sums sf_ce_forward left in the workspace
// (fixed order -> identical decision in every CTA, bitwise reproducible loss).  Only when fewer than n_min
// pixels exceed the threshold does the grid continue into a 3-pass (11+11+10 bit) radix select of the
// n_min-th largest per-pixel loss over order-preserving uint32 keys; the last pass also sums the losses above
// the prefix, and the ones inside it follow exactly from the last histogram (one float value per bin).  Ties at the k-th value share the remaining slots equally (tie_share); torch.topk breaks
// them arbitrarily, the loss value is identical either way.
#include <cooperative_groups.h>
#include <cstdlib>

#include "sf_common.cuh"
#include "sf_launch.h"

namespace cg = cooperative_groups;

namespace sf {

constexpr int kSelThreads = 1024;              // one CTA per SM: as many loads in flight as 2 x 512 with half the CTAs to launch

__device__ __forceinline__ unsigned key_of(float v) {           // total order: -NaN < -inf < ... < +inf < +NaN
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float val_of(unsigned key) {
    return __uint_as_float((key & 0x80000000u) ? (key & 0x7fffffffu) : ~key);
}

struct BinPick { unsigned bin; unsigned long long k_rem; unsigned long long cnt; };

// Find the bin holding the k-th largest key (bins scanned from the top).  All threads must call.
__device__ BinPick pick_bin(const unsigned* hist, int nbins, unsigned long long k,
                            unsigned long long* warp_tot, BinPick* out) {
    const int per = nbins / kSelThreads;                          // 2 (2048 bins) or 1 (1024 bins)
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    unsigned long long c = 0;
    for (int q = 0; q < per; ++q) c += __ldcg(hist + (nbins - 1 - (t * per + q)));   // written by atomics this launch
    unsigned long long incl = c;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    unsigned long long base = 0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    const unsigned long long excl = base + incl - c;
    if (excl < k && k <= excl + c) {
        unsigned long long cum = excl;
        for (int q = 0; q < per; ++q) {
            const int b = nbins - 1 - (t * per + q);
            const unsigned long long h = __ldcg(hist + b);
            if (k <= cum + h) { out->bin = (unsigned)b; out->k_rem = k - cum; out->cnt = h; break; }
            cum += h;
        }
    }
    __syncthreads();
    return *out;
}

// One counting pass over the loss vector: histogram of the next digit of the keys that match `prefix`.
// SUM_ABOVE (last pass only): also returns this thread's sum of the values whose key is above the prefix, so that no
// separate summing pass is needed (the values inside the prefix are recovered exactly from the last histogram).
// Four independent 128-bit loads are kept in flight per thread: with one CTA per SM a single load per iteration
// leaves the L2 -> SM path latency-bound.
template <int SHIFT, int NBINS, int PREFIX_SHIFT, bool SUM_ABOVE>
__device__ double radix_pass(const float* __restrict__ v, long long n, unsigned prefix, unsigned* sh, unsigned* gh) {
    for (int b = threadIdx.x; b < NBINS; b += kSelThreads) sh[b] = 0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * kSelThreads;
    const long long gtid = (long long)blockIdx.x * kSelThreads + threadIdx.x;
    double above = 0.0;
    auto add = [&](float x) {
        const unsigned key = key_of(x);
        if (PREFIX_SHIFT >= 32) { atomicAdd(&sh[(key >> SHIFT) & (NBINS - 1)], 1u); return; }
        const unsigned hi = key >> (PREFIX_SHIFT & 31);
        if (hi == prefix) atomicAdd(&sh[(key >> SHIFT) & (NBINS - 1)], 1u);
        else if (SUM_ABOVE && hi > prefix) above += (double)x;
    };
    if ((reinterpret_cast<uintptr_t>(v) & 15) == 0) {
        const long long n4 = n >> 2;
        const float4* v4 = reinterpret_cast<const float4*>(v);
        long long i = gtid;
        for (; i + 3 * stride < n4; i += 4 * stride) {
            float4 q[4];
            #pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = __ldg(v4 + i + u * stride);
            #pragma unroll
            for (int u = 0; u < 4; ++u) { add(q[u].x); add(q[u].y); add(q[u].z); add(q[u].w); }
        }
        for (; i < n4; i += stride) {
            const float4 q = __ldg(v4 + i);
            add(q.x); add(q.y); add(q.z); add(q.w);
        }
        for (long long j = (n4 << 2) + gtid; j < n; j += stride) add(__ldg(v + j));
    } else {
        for (long long i = gtid; i < n; i += stride) add(__ldg(v + i));
    }
    __syncthreads();
    for (int b = threadIdx.x; b < NBINS; b += kSelThreads)
        if (sh[b]) atomicAdd(&gh[b], sh[b]);
    return above;
}

__global__ void __launch_bounds__(kSelThreads, 1)
finalize_kernel(const float* __restrict__ loss_px, long long n, float thresh, Workspace* ws, float* loss_out) {
    __shared__ double red[kSelThreads / 32];
    __shared__ double tot[5];
    __shared__ double red5[32][5];
    __shared__ unsigned sh_hist[kRadixBins];
    __shared__ unsigned long long warp_tot[kSelThreads / 32];
    __shared__ BinPick pick;

    // ---- phase 0: fixed-order reduction of the forward kernel's per-CTA partials (all five sums behind one barrier
    //      pair: this prologue is the whole kernel on the CE / threshold branches) ----
    if (ws->n_partials == kFxMarker) {                              // one-pass kernel: order-free fixed-point totals
        if (threadIdx.x == 0) {
            const unsigned int fl = __ldcg(&ws->fx_flags);
            tot[0] = fx_get(ws->fx_loss, fl, 0); tot[1] = fx_get(ws->fx_w, fl, 1); tot[2] = fx_get(ws->fx_hard, fl, 2);
            tot[3] = (double)__ldcg(&ws->fx_n_valid); tot[4] = (double)__ldcg(&ws->fx_n_hard);
        }
        __syncthreads();
    } else {
        const int np = (int)ws->n_partials;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int nw = (np + 31) / 32;                              // warps that hold partials (<= 32)
        if (warp < nw) {
            double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            for (int i = threadIdx.x; i < np; i += 32 * nw)
                #pragma unroll
                for (int f = 0; f < 5; ++f) a[f] += ws->partials[i][f];
            #pragma unroll
            for (int f = 0; f < 5; ++f) {
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) a[f] += __shfl_down_sync(0xffffffffu, a[f], o);
                if (lane == 0) red5[warp][f] = a[f];
            }
        }
        __syncthreads();
        if (threadIdx.x < 5) {
            double r = 0.0;
            for (int w = 0; w < nw; ++w) r += red5[w][threadIdx.x];
            tot[threadIdx.x] = r;
        }
        __syncthreads();
    }
    const double sum_loss = tot[0], sum_w = tot[1], sum_hard = tot[2];
    const unsigned long long n_valid = (unsigned long long)tot[3], n_hard = (unsigned long long)tot[4];
    const bool writer = blockIdx.x == 0 && threadIdx.x == 0;
    sf_loss_state st;
    st.reserved = 0; st.focal_alpha = 0.f; st.focal_gamma = 0.f; st.n_valid = n_valid; st.n_hard = n_hard; st.sum_w = sum_w; st.tie_share = 0.f;

    if (thresh < 0.f) {                                             // CrossEntropy: sum(w_y l) / sum(w_y)
        if (writer) {
            st.loss = (float)(sum_loss / sum_w); st.kth = -INFINITY; st.inv_denom = (float)(1.0 / sum_w);
            st.branch = 0; st.n_kept = n_valid; st.sum_kept = sum_loss;
            ws->state = st;
            if (loss_out) *loss_out = st.loss;
        }
        return;
    }
    const unsigned long long n_min = n_valid / 16;
    if (n_hard >= n_min) {                                          // threshold branch (mean of empty = NaN)
        if (writer) {
            st.loss = (float)(sum_hard / (double)n_hard); st.kth = thresh; st.inv_denom = (float)(1.0 / (double)n_hard);
            st.branch = 1; st.n_kept = n_hard; st.sum_kept = sum_hard;
            ws->state = st;
            if (loss_out) *loss_out = st.loss;
        }
        return;
    }

    // ---- top-k branch: k = n_min >= 1 ----
    cg::grid_group grid = cg::this_grid();
    const unsigned long long k = n_min;
    {
        unsigned* rh = &ws->radix_hist[0][0];
        for (long long i = (long long)blockIdx.x * kSelThreads + threadIdx.x; i < 3 * kRadixBins;
             i += (long long)gridDim.x * kSelThreads) rh[i] = 0;
    }
    grid.sync();
    radix_pass<21, 2048, 32, false>(loss_px, n, 0u, sh_hist, ws->radix_hist[0]);
    grid.sync();
    const BinPick p1 = pick_bin(ws->radix_hist[0], 2048, k, warp_tot, &pick);
    radix_pass<10, 2048, 21, false>(loss_px, n, p1.bin, sh_hist, ws->radix_hist[1]);
    grid.sync();
    const BinPick p2 = pick_bin(ws->radix_hist[1], 2048, p1.k_rem, warp_tot, &pick);
    const unsigned prefix22 = (p1.bin << 11) | p2.bin;
    // last digit + the sum of every loss whose 22-bit prefix is above the k-th value's
    const double a = radix_pass<0, 1024, 10, true>(loss_px, n, prefix22, sh_hist, ws->radix_hist[2]);
    const double cta_sum = block_sum<kSelThreads>(a, red);
    if (threadIdx.x == 0) ws->sel_partials[blockIdx.x] = cta_sum;
    grid.sync();
    const BinPick p3 = pick_bin(ws->radix_hist[2], 1024, p2.k_rem, warp_tot, &pick);
    const unsigned kth_key = (prefix22 << 10) | p3.bin;
    const float kth = val_of(kth_key);
    if (blockIdx.x == 0) {
        // losses strictly above the k-th value = those above the prefix (per-CTA partials) + those inside the prefix with a
        // larger last digit: a bin of the last histogram is ONE float value, so count * value is exact
        double b = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += kSelThreads) b += __ldcg(&ws->sel_partials[i]);
        for (int bin = (int)p3.bin + 1 + threadIdx.x; bin < 1024; bin += kSelThreads) {
            const unsigned c = __ldcg(&ws->radix_hist[2][bin]);
            if (c) b += (double)c * (double)val_of((prefix22 << 10) | (unsigned)bin);
        }
        const double sum_gt = block_sum<kSelThreads>(b, red);
        if (threadIdx.x == 0) {
            const unsigned long long take = p3.k_rem;               // ties kept (1 .. p3.cnt)
            const double sum_kept = sum_gt + (double)take * (double)kth;
            st.loss = (float)(sum_kept / (double)k); st.kth = kth; st.tie_share = (float)((double)take / (double)p3.cnt);
            st.inv_denom = (float)(1.0 / (double)k); st.branch = 2; st.n_kept = k; st.sum_kept = sum_kept;
            ws->state = st;
            if (loss_out) *loss_out = st.loss;
        }
    }
}

// FocalLoss reduction (util/losses.py:17-25) over the per-pixel CE: one persistent pass, per-CTA partials in a fixed
// order, the last CTA to finish (atomic ticket) sums them and writes the state -- deterministic, single launch.
__global__ void __launch_bounds__(kSelThreads)
focal_kernel(float* __restrict__ ce, long long n, float alpha, float gamma, int size_average, Workspace* ws,
             float* loss_out) {
    __shared__ double red[kSelThreads / 32];
    __shared__ bool last;
    double a = 0.0;
    const long long stride = (long long)gridDim.x * kSelThreads;
    for (long long i = (long long)blockIdx.x * kSelThreads + threadIdx.x; i < n; i += stride) {
        float d;
        a += (double)focal_value(ce[i], alpha, gamma, &d);
        ce[i] = d;                                                   // loss_px now holds d focal / d ce: the backward's per-pixel weight
    }
    const double cta = block_sum<kSelThreads>(a, red);
    if (threadIdx.x == 0) {
        ws->sel_partials[blockIdx.x] = cta;
        __threadfence();
        last = atomicAdd(&ws->radix_hist[0][0], 1u) == gridDim.x - 1;   // ticket counter, reset below
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double b = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kSelThreads) b += __ldcg(&ws->sel_partials[i]);
    const double tot = block_sum<kSelThreads>(b, red);
    if (threadIdx.x == 0) {
        ws->radix_hist[0][0] = 0;
        sf_loss_state st;
        st.reserved = 0; st.n_valid = 0; st.n_hard = 0; st.n_kept = (unsigned long long)n; st.sum_w = 0.0;
        const double denom = size_average ? (double)n : 1.0;
        st.loss = (float)(tot / denom); st.kth = 0.f; st.tie_share = 0.f; st.inv_denom = (float)(1.0 / denom);
        st.branch = 3; st.focal_alpha = alpha; st.focal_gamma = gamma; st.sum_kept = tot;
        ws->state = st;
        if (loss_out) *loss_out = st.loss;
    }
}

int launch_focal(float* loss_px, long long n_px, float alpha, float gamma, int size_average, void* ws,
                 float* loss_out, void* stream) {
    Workspace* w = reinterpret_cast<Workspace*>(ws);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(&w->radix_hist[0][0], 0, sizeof(unsigned), st);   // ticket counter
    if (e != cudaSuccess) return (int)e;
    const int grid = (int)std::min<long long>((n_px + kSelThreads - 1) / kSelThreads, std::min(2 * sm_count(), kMaxCtas));
    focal_kernel<<<grid, kSelThreads, 0, st>>>(loss_px, n_px, alpha, gamma, size_average, w, loss_out);
    return (int)cudaGetLastError();
}

int launch_finalize(const float* loss_px, long long n_px, float thresh, void* ws, float* loss_out, void* stream) {
    Workspace* w = reinterpret_cast<Workspace*>(ws);
    int grid = 1;
    if (thresh >= 0.f) {                                             // cooperative: every CTA must be resident
        static int per_sm = 0;
        if (per_sm == 0) {
            int v = 1;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, finalize_kernel, kSelThreads, 0) != cudaSuccess || v < 1) v = 1;
            per_sm = std::min(v, 1);
            if (const char* e = std::getenv("SF_DEBUG_SELECT_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(e)));   // experiment knob
        }
        grid = std::min(per_sm * sm_count(), kMaxCtas);
    }
    void* args[] = {(void*)&loss_px, (void*)&n_px, (void*)&thresh, (void*)&w, (void*)&loss_out};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)finalize_kernel, dim3(grid), dim3(kSelThreads), args, 0,
                                                reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return (int)e;
    return (int)cudaGetLastError();
}

}  // namespace sf
