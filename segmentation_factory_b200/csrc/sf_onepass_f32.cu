// sf_onepass_f32.cu -- instantiates the one-pass walk (sf_onepass.cuh) for float logits (one translation unit per
// element type keeps the parallel build short).
#include "sf_onepass.cuh"

namespace sf {
int onepass_dispatch_f32(const OnepassArgs& a, int s) { return op_classes<float>(a, s); }
}  // namespace sf
