// sf_onepass_chunk_f16.cu -- __half instantiations of the many-class one-pass walk (kernel: sf_onepass_chunk.cuh).
#include "sf_onepass_chunk.cuh"

namespace sf {
int onepass_chunk_dispatch_f16(const OnepassArgs& a, int s) { return oc_dispatch<__half>(a, s); }
}  // namespace sf
