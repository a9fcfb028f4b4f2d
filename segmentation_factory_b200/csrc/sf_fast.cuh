// sf_fast.cuh -- compile-time geometry shared by the integer-scale fast kernels.
//
// The fast path covers what the reference's 1/4- and 1/8-resolution heads produce (SegFormer, UperNet, FPN,
// EfficientViT-Seg; models/heads/*.py): H == S*h, W == S*w with S a power of two, all C channels held by one
// thread at a time.  For such S the ATen bilinear taps are periodic and exactly representable:
//     u(d)   = floor((d + 0.5)/S - 0.5) = (d + S/2)/S - 1          (integer arithmetic, -1 for the clamped rim)
//     phase  = (d + S/2) % S,   l1 = (phase + 0.5)/S  (0 when u < 0),   l0 = 1 - l1
//     i0 = max(u, 0),  i1 = min(i0 + 1, in - 1)
// which is bit-identical to the float formula in sf_common.cuh::make_tap (scale = 1/S is a power of two, so
// (d + 0.5) * scale - 0.5 is exact), also after rounding to bf16/fp16 (l1 has at most 5 significant bits).
// Everything else -- strides, box extents, loop trip counts -- is a constexpr of <LT, C, S>.
//
// Arithmetic is issued as packed fp32x2 (FFMA2/FMUL2/FADD2, sm_100) over channel pairs: the lanes are IEEE
// round-to-nearest like the scalar ops, so results stay bit-exact while the FMA-pipe instruction count halves.
#pragma once

#include "sf_common.cuh"

namespace sf {

__host__ __device__ constexpr int round_up(int x, int a) { return (x + a - 1) / a * a; }

template <int S>
struct ScaleTap {
    static_assert(S == 2 || S == 4 || S == 8, "fast path scale must be 2, 4 or 8");
    // taps of destination index d along an axis of `in` source elements
    static __device__ __forceinline__ void get(int d, int in, int& i0, int& i1, float& l0, float& l1) {
        const int u = (d + S / 2) / S - 1;
        const int ph = (d + S / 2) % S;
        i0 = max(u, 0);
        i1 = min(i0 + 1, in - 1);
        l1 = u < 0 ? 0.0f : ((float)ph + 0.5f) * (1.0f / S);
        l0 = 1.0f - l1;
    }
    // rows of group u (u = -1 .. in-1): [first, first + count)
    static __device__ __forceinline__ int first(int u) { return u < 0 ? 0 : u * S + S / 2; }
    static __device__ __forceinline__ int count(int u, int in) { return (u < 0 || u == in - 1) ? S / 2 : S; }
};

__device__ __forceinline__ float2 lerp2_pinned(float2 a, float2 la, float2 b, float2 lb) {
    return __ffma2_rn(a, la, __fmul2_rn(b, lb));
}
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }

template <typename LT>
__device__ __forceinline__ float2 rnd2(float2 v) { return v; }
template <>
__device__ __forceinline__ float2 rnd2<__nv_bfloat16>(float2 v) {
    return __bfloat1622float2(__float22bfloat162_rn(v));
}
template <>
__device__ __forceinline__ float2 rnd2<__half>(float2 v) {
    return __half22float2(__float22half2_rn(v));
}

}  // namespace sf
