// sf_head_gemm.cu -- the decode head's 1x1 classifier on the 5th-generation tensor cores.
//
// Replaces nn.Conv2d(embed_dim, num_classes, 1) at the end of every head -- SegFormerHead.linear_pred
// (models/heads/segformer.py:39,57), UPerHead.conv_seg (models/heads/upernet.py:32,50), FPNHead.conv_seg
// (models/heads/fpn.py:23,37) -- the producer of the low-resolution logits the loss / metric kernels consume:
//     logits[b][c][n] = bias[c] + sum_k W[c][k] * X[b][k][n],        n = the h*w positions (contiguous in NCHW)
// The op is a skinny GEMM (19..172 output channels) whose cost is streaming X once: 2*K bytes per position in, 2..4*C
// bytes out, so it is HBM-bound and judged on bytes; the contraction runs on tcgen05 so that the FMA
// pipe (which a CUDA-core version would load to about the same time as HBM) stays out of the way.
//
// Mapping: D[128 positions x Cpad] (fp32, TMEM) += A[128 x 16] * B[16 x Cpad] per tcgen05.mma (kind::f16, M = 128).
//   A = X^T: MN-major -- for a fixed k the 128 positions are contiguous in global memory.  TMA (SWIZZLE_128B) drops a
//       [64 k][64 positions] box as 64 rows of 128 B; that IS the canonical MN-major SW128 layout (8-row groups 1024 B
//       apart = SBO, the second 64-position half 8192 B further = LBO).
//   B = W^T: K-major -- W is [C][K] with K contiguous.  TMA boxes of [8 | 16 | 32 classes][64 k] are SW128 atoms of 8 rows
//       (SBO = 1024 B); rows past C are zero-filled by TMA (every box starts inside the tensor) or, below 16 classes, zeroed
//       once at start, so Cpad = the next multiple of 16 and the padded classes cost nothing but tensor-core time.
// Warp roles (320 threads, one persistent CTA per SM): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warps 2..9 = epilogue (two warps per 32-lane TMEM quarter, each taking half of the class columns: tcgen05.ld, + bias,
// convert, stage [16 classes][32 positions] in shared memory, one TMA store per group -- the tensor map clips the ragged
// last tile and the classes past C).  Two pipelines: shared-memory stages (full / empty mbarriers, tcgen05.commit frees a
// stage) and a double-buffered TMEM accumulator (tmem_full / tmem_empty), so the epilogue of tile i overlaps the loads and
// MMAs of tile i + 1.  When W (cpad x K) fits in 64 KB it is loaded ONCE per CTA and stays resident (the usual case:
// 19..21 classes); otherwise its 64-k slices travel with the X stages.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "sf_common.cuh"
#include "sf_launch.h"

namespace sf {

constexpr int kHgThreads = 320;
constexpr int kHgEpiWarps = 8;
constexpr int kHgBM = 128;                     // positions per tile (= UMMA M)
constexpr int kHgBK = 64;                      // k per stage: 64 bf16 = one 128-byte swizzle row
constexpr int kHgUK = 16;                      // k per tcgen05.mma (kind::f16)
constexpr int kHgXStage = kHgBK * kHgBM * 2;   // 16 KB: two [64 k][64 positions] boxes

struct HgParams {
    const float* bias;                          // [C] or null
    void* out;                                  // [B][C][N]
    int B, C, K, N;                             // N = h*w
    int cpad;                                   // C rounded up to 16 (UMMA N)
    int stages;
    int tiles_per_img, total_tiles;
    int w_resident;                             // W stays in shared memory for the whole kernel
    int wbox;                                   // class rows per W TMA box (8, 16 or 32)
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// shared memory -> global, tile mode; completion is tracked by the issuing thread's bulk async-groups
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {       // arrives on `bar` when every MMA issued so far has finished
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread (thread = its warp's TMEM lane)
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    #pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor): start address, leading / stride byte
// offsets in 16-byte units, version 1 (Blackwell), SWIZZLE_128B
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                     // version_ = 1
    d |= (uint64_t)2 << 61;                     // layout_type_ = SWIZZLE_128B
    return d;
}

template <typename OT> __device__ __forceinline__ void hg_store(OT* p, float v);
template <> __device__ __forceinline__ void hg_store<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void hg_store<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void hg_store<__half>(__half* p, float v) { *p = __float2half_rn(v); }

template <bool BF16, typename OT>
__global__ void __launch_bounds__(kHgThreads, 1)
head_gemm_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_o, const HgParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // carve (all 1024-byte aligned, SWIZZLE_128B atoms): resident W [nkb][cpad rows x 128 B] | [stages] x { X: 16 KB (+ a W slice) }
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int nkb = (p.K + kHgBK - 1) / kHgBK;
    const int w_slice = p.cpad * 128;                              // one 64-k slice of W
    const int w_res_bytes = p.w_resident ? nkb * w_slice : 0;
    const int stage_bytes = kHgXStage + (p.w_resident ? 0 : w_slice);
    unsigned char* stage0 = smem + w_res_bytes;
    __shared__ uint64_t full_bar[12], empty_bar[12], tmem_full[2], tmem_empty[2], w_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float bias_s[256];
    __shared__ __align__(128) OT out_stage[kHgEpiWarps][2][16 * 32];   // per epilogue warp: two [16 classes][32 positions] boxes

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // Class rows covered by TMA boxes; every box starts inside the tensor, so TMA zero-fills the rows past C.  Only with
    // fewer than 16 classes (8-row boxes) is there a whole 8-row group [wrows, cpad) no box writes: zeroed once below.
    // (Fewer, larger boxes matter when W is streamed: 19 eight-row boxes per stage at 150 classes made the single
    // producer thread the bottleneck.)
    const int wrows = (p.C + p.wbox - 1) / p.wbox * p.wbox;

    if (tid == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_o);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], kHgEpiWarps); }
        mbar_init(&w_bar, 1);
        mbar_fence_init();
    }
    for (int c = tid; c < 256; c += kHgThreads) bias_s[c] = (p.bias && c < p.C) ? __ldg(p.bias + c) : 0.f;
    // zero the class rows no TMA box ever writes (whole 8-row groups past C)
    {
        const int n_slices = p.w_resident ? nkb : p.stages;
        for (int s = 0; s < n_slices; ++s) {
            unsigned char* wz = p.w_resident ? smem + s * w_slice : stage0 + s * stage_bytes + kHgXStage;
            for (int i = tid * 16; i < (p.cpad - wrows) * 128; i += kHgThreads * 16)
                *reinterpret_cast<uint4*>(wz + wrows * 128 + i) = make_uint4(0, 0, 0, 0);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy zeros visible to the tensor-core (async) proxy
    const uint32_t tmem_cols = 2 * p.cpad <= 32 ? 32 : 2 * p.cpad <= 64 ? 64 : 2 * p.cpad <= 128 ? 128 : 2 * p.cpad <= 256 ? 256 : 512;
    if (warp == 2) {                             // one warp allocates (and later frees) the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    int n_tiles = 0;
    if ((int)blockIdx.x < p.total_tiles) n_tiles = (p.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            if (p.w_resident) {
                mbar_expect_tx(&w_bar, (uint32_t)(nkb * wrows * 128));
                for (int kb = 0; kb < nkb; ++kb)
                    for (int r = 0; r < wrows; r += p.wbox) tma_load_2d(smem + kb * w_slice + r * 128, &tmap_w, &w_bar, kb * kHgBK, r);
            }
            int stage = 0; uint32_t phase = 0;
            for (int t = 0; t < n_tiles; ++t) {
                const int tile = blockIdx.x + t * gridDim.x;
                const int b = tile / p.tiles_per_img, n0 = (tile % p.tiles_per_img) * kHgBM;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    unsigned char* xs = stage0 + stage * stage_bytes;
                    mbar_expect_tx(&full_bar[stage], (uint32_t)(kHgXStage + (p.w_resident ? 0 : wrows * 128)));
                    tma_load_3d(xs, &tmap_x, &full_bar[stage], n0, kb * kHgBK, b);
                    tma_load_3d(xs + kHgXStage / 2, &tmap_x, &full_bar[stage], n0 + 64, kb * kHgBK, b);
                    if (!p.w_resident)
                        for (int r = 0; r < wrows; r += p.wbox) tma_load_2d(xs + kHgXStage + r * 128, &tmap_w, &full_bar[stage], kb * kHgBK, r);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16 | F16, A MN-major, B K-major, N = cpad, M = 128
        const uint32_t idesc = (1u << 4) | ((BF16 ? 1u : 0u) << 7) | ((BF16 ? 1u : 0u) << 10) | (1u << 15) | (0u << 16) |
                               ((uint32_t)(p.cpad >> 3) << 17) | ((uint32_t)(kHgBM >> 4) << 24);
        if (p.w_resident) mbar_wait(&w_bar, 0);
        int stage = 0; uint32_t phase = 0;
        for (int t = 0; t < n_tiles; ++t) {
            const int acc = t & 1;
            const uint32_t acc_phase = (uint32_t)((t >> 1) & 1);
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);          // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.cpad);
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t xa = smem_u32(stage0 + stage * stage_bytes);
                    const uint32_t wa = p.w_resident ? smem_u32(smem + kb * w_slice) : xa + kHgXStage;
                    #pragma unroll
                    for (int k = 0; k < kHgBK / kHgUK; ++k) {
                        // A: two 8-row k groups per MMA (2048 B per step); B: 32 bytes along the swizzled 128-byte rows
                        const uint64_t da = smem_desc_sw128(xa + k * 2048, kHgXStage / 2, 1024);
                        const uint64_t db = smem_desc_sw128(wa + k * 32, 16, 1024);
                        tc_mma_f16(tmem_d, da, db, idesc, (kb | k) != 0);
                    }
                    tc_commit(&empty_bar[stage]);                    // stage reusable once these MMAs have read it
                    if (kb == nkb - 1) tc_commit(&tmem_full[acc]);   // accumulator complete
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue: warps 2..9; warp w reads TMEM lanes 32 * (w % 4) .. + 32 and every second 16-column group =====
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        int sb = 0;                                                 // staging buffer of the next group
        for (int t = 0; t < n_tiles; ++t) {
            const int tile = blockIdx.x + t * gridDim.x;
            const int b = tile / p.tiles_per_img, n0 = (tile % p.tiles_per_img) * kHgBM;
            const int acc = t & 1;
            mbar_wait(&tmem_full[acc], (uint32_t)((t >> 1) & 1));
            tc_fence_after();
            for (int c0 = half * 16; c0 < p.C; c0 += 32) {
                float v[16];
                tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.cpad + c0), v);
                // the store issued two groups ago has finished READING this staging buffer
                if (lane == 0) tma_store_wait_read<1>();
                __syncwarp();
                OT* sbuf = out_stage[warp - 2][sb];
                #pragma unroll
                for (int i = 0; i < 16; ++i) hg_store<OT>(sbuf + i * 32 + lane, v[i] + bias_s[c0 + i]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes visible to the TMA (async proxy)
                __syncwarp();
                if (lane == 0) {
                    tma_store_3d(&tmap_o, sbuf, n0 + q * 32, c0, b);       // rows n >= N and classes >= C are clipped by the map
                    tma_store_commit();
                }
                sb ^= 1;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (lane == 0) tma_store_wait_all();                        // shared memory must outlive the last stores
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
int launch_head_classifier(const void* x, const void* w, const float* bias, void* out, int in_dtype, int out_dtype,
                           int B, int C, int K, long long N, void* stream) {
    if (!x || !w || !out || B <= 0 || C <= 0 || K <= 0 || N <= 0) return SF_ERR_ARG;
    if (in_dtype != SF_BF16 && in_dtype != SF_F16) return SF_ERR_UNSUPPORTED;       // tensor-core operands: bf16 / fp16
    if (out_dtype < SF_F32 || out_dtype > SF_F16) return SF_ERR_ARG;
    if (C > 256) return SF_ERR_TOO_MANY_CLASSES;                                   // UMMA N <= 256
    if ((N % 8) || (K % 8) || ((uintptr_t)x % 16) || ((uintptr_t)w % 16)) return SF_ERR_UNSUPPORTED;   // TMA strides: 16-byte multiples
    if (N >= (1LL << 31) || (long long)B * K >= (1LL << 31)) return SF_ERR_SHAPE;
    HgParams p{};
    p.bias = bias; p.out = out; p.B = B; p.C = C; p.K = K; p.N = (int)N;
    p.cpad = (C + 15) / 16 * 16;
    p.wbox = (C >= 32 && p.cpad % 32 == 0) ? 32 : (C >= 16 ? 16 : 8);
    const int nkb = (K + kHgBK - 1) / kHgBK;
    const int w_all = nkb * p.cpad * 128;
    p.w_resident = w_all <= 64 * 1024;
    const int stage_bytes = kHgXStage + (p.w_resident ? 0 : p.cpad * 128);
    const int w_res_bytes = p.w_resident ? w_all : 0;
    if ((uintptr_t)out % 16) return SF_ERR_UNSUPPORTED;                             // TMA store: 16-byte aligned rows
    const int osz = out_dtype == SF_F32 ? 4 : 2;
    const int out_stage_bytes = kHgEpiWarps * 2 * 16 * 32 * osz;                   // static shared memory of the epilogue
    p.stages = std::min(12, std::max(2, (208 * 1024 - out_stage_bytes - w_res_bytes) / stage_bytes));
    p.tiles_per_img = (int)((N + kHgBM - 1) / kHgBM);
    const long long total = (long long)B * p.tiles_per_img;
    if (total >= (1LL << 31)) return SF_ERR_SHAPE;
    p.total_tiles = (int)total;
    CUtensorMap tx{}, tw{}, to{};
    {
        const unsigned long long dims[3] = {(unsigned long long)N, (unsigned long long)K, (unsigned long long)B};
        const unsigned long long strides[2] = {(unsigned long long)N * 2, (unsigned long long)N * K * 2};
        const unsigned box[3] = {64, (unsigned)kHgBK, 1};
        int rc = make_tmap_generic(&tx, x, in_dtype, 3, dims, strides, box, /*swizzle128=*/1);
        if (rc != 0) return rc;
        const unsigned long long dw[2] = {(unsigned long long)K, (unsigned long long)C};
        const unsigned long long sw[1] = {(unsigned long long)K * 2};
        const unsigned bw[2] = {(unsigned)kHgBK, (unsigned)p.wbox};
        rc = make_tmap_generic(&tw, w, in_dtype, 2, dw, sw, bw, 1);
        if (rc != 0) return rc;
        // logits [B][C][N]: the epilogue stores [16 classes][32 positions] boxes
        const unsigned long long dout[3] = {(unsigned long long)N, (unsigned long long)C, (unsigned long long)B};
        const unsigned long long sout[2] = {(unsigned long long)N * osz, (unsigned long long)N * C * osz};
        const unsigned bo[3] = {32, 16, 1};
        rc = make_tmap_generic(&to, out, out_dtype, 3, dout, sout, bo, /*swizzle128=*/0);
        if (rc != 0) return rc;
    }
    const size_t smem = (size_t)w_res_bytes + (size_t)p.stages * stage_bytes + 1024;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = (int)std::min<long long>(total, sm_count());
#define SF_HG_LAUNCH(BF, OT)                                                                     \
    do {                                                                                         \
        static SmemLimitOnce once;                                                               \
        const cudaError_t e = set_smem_limit_once(head_gemm_kernel<BF, OT>, smem, once);         \
        if (e != cudaSuccess) return (int)e;                                                     \
        head_gemm_kernel<BF, OT><<<grid, kHgThreads, smem, st>>>(tx, tw, to, p);                     \
    } while (0)
    if (in_dtype == SF_BF16) {
        if (out_dtype == SF_F32) SF_HG_LAUNCH(true, float);
        else if (out_dtype == SF_BF16) SF_HG_LAUNCH(true, __nv_bfloat16);
        else SF_HG_LAUNCH(true, __half);
    } else {
        if (out_dtype == SF_F32) SF_HG_LAUNCH(false, float);
        else if (out_dtype == SF_BF16) SF_HG_LAUNCH(false, __nv_bfloat16);
        else SF_HG_LAUNCH(false, __half);
    }
#undef SF_HG_LAUNCH
    return (int)cudaGetLastError();
}

}  // namespace sf
