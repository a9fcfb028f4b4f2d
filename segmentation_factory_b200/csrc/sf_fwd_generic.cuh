// sf_fwd_generic.cuh -- work decomposition of the generic (any scale, any class count) forward walk, shared by
// sf_forward.cu (CE / argmax / histogram) and sf_dice.cu (dice statistics and coefficients of engine.criterion).
// See sf_forward.cu's header comment for the tile / pass / chunk item scheme.
#pragma once

#include "sf_common.cuh"
#include "sf_launch.h"

namespace sf {

constexpr int kFwdTW = 64;                 // output columns per tile (= threads in x)
constexpr int kFwdNGY = 4;                 // row groups per tile     (= threads in y)
constexpr int kFwdThreads = kFwdTW * kFwdNGY;
constexpr int kFwdG = 4;                   // rows of a group handled per pass (register state per thread)
constexpr int kFwdStages = 3;
constexpr int kFwdBH = kFwdNGY + 1;        // source rows per box

struct FwdParams {
    const void* logits;
    const void* labels;
    const float* class_w;
    float* loss_px;
    float* lse_px;
    unsigned long long* hist;
    Workspace* ws;
    Axis ay, ax;
    int B, C, ignore, label_u8;
    float thresh;
    int tiles_x, tiles_y, total_tiles;
    int n_chunks, n_passes, box_w;
    int use_tma;
    int xalign;                            // box x-origin granularity in elements: TMA needs 16-byte aligned inner coordinates
    int stage_bytes;                       // bytes of one smem stage (128-aligned)
    int hist_bytes;                        // bytes reserved for the smem histogram (16-aligned)
};

struct ItemCoord { int b, ua, Xa, pass, chunk; };

__host__ __device__ __forceinline__ int floor_al(int x, int a) { return x & ~(a - 1); }   // a is a power of two

template <int NCH>
__device__ __forceinline__ ItemCoord decode_item(const FwdParams& p, int n) {
    const int ipt = p.n_passes * p.n_chunks;
    const int tile = blockIdx.x + (n / ipt) * gridDim.x;
    const int rem = n % ipt;
    ItemCoord c;
    c.pass = rem / p.n_chunks;
    c.chunk = rem % p.n_chunks;
    const int per_img = p.tiles_x * p.tiles_y;
    c.b = tile / per_img;
    const int r = tile % per_img;
    c.ua = (p.ay.ident ? 0 : -1) + (r / p.tiles_x) * kFwdNGY;
    c.Xa = (r % p.tiles_x) * kFwdTW;
    return c;
}

template <typename LT, int NCH>
__device__ __forceinline__ void issue_tma(const FwdParams& p, const CUtensorMap* tmap, const ItemCoord& c,
                                          unsigned char* stage, uint64_t* bar) {
    const int jA = floor_al(src_i0(p.ax, c.Xa), p.xalign);
    mbar_expect_tx(bar, (uint32_t)(NCH * kFwdBH * p.box_w * (int)sizeof(LT)));
    tma_load_3d(stage, tmap, bar, jA, c.ua, c.b * p.C + c.chunk * NCH);
}

// Fallback loader for shapes TMA cannot describe (row pitch not a multiple of 16 bytes).
template <typename LT, int NCH>
__device__ __forceinline__ void load_plain(const FwdParams& p, const ItemCoord& c, LT* stage) {
    const int jA = floor_al(src_i0(p.ax, c.Xa), p.xalign);
    const LT* src = reinterpret_cast<const LT*>(p.logits);
    const int n = NCH * kFwdBH * p.box_w;
    for (int e = threadIdx.x; e < n; e += kFwdThreads) {
        const int col = e % p.box_w, row = (e / p.box_w) % kFwdBH, k = e / (p.box_w * kFwdBH);
        const int ch = c.chunk * NCH + k, i = c.ua + row, j = jA + col;
        LT v = LT(0.0f);
        if (ch < p.C && i >= 0 && i < p.ay.in && j >= 0 && j < p.ax.in)
            v = src[(((long long)c.b * p.C + ch) * p.ay.in + i) * p.ax.in + j];
        stage[e] = v;
    }
}

// Host: tile grid, box width, smem layout and the TMA descriptor for a walk with nch-class chunks.  aux_bytes is the
// kernel's own shared-memory area between the stages and the mbarriers (histogram / dice bins).
int setup_generic_forward(const FwdArgs& a, int nch, size_t aux_bytes, FwdParams* p, CUtensorMap* tmap, int* grid,
                          size_t* smem);
int pick_nch(int C);

}  // namespace sf
