// sf_onepass_f16.cu -- instantiates the one-pass walk (sf_onepass.cuh) for __half logits (one translation unit per
// element type keeps the parallel build short).
#include "sf_onepass.cuh"

namespace sf {
int onepass_dispatch_f16(const OnepassArgs& a, int s) { return op_classes<__half>(a, s); }
}  // namespace sf
