// sf_onepass_bf16.cu -- instantiates the one-pass walk (sf_onepass.cuh) for __nv_bfloat16 logits (one translation unit per
// element type keeps the parallel build short).
#include "sf_onepass.cuh"

namespace sf {
int onepass_dispatch_bf16(const OnepassArgs& a, int s) { return op_classes<__nv_bfloat16>(a, s); }
}  // namespace sf
