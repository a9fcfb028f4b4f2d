// sf_onepass_chunk_f32.cu -- float instantiations of the many-class one-pass walk (kernel: sf_onepass_chunk.cuh).
#include "sf_onepass_chunk.cuh"

namespace sf {
int onepass_chunk_dispatch_f32(const OnepassArgs& a, int s) { return oc_dispatch<float>(a, s); }
}  // namespace sf
