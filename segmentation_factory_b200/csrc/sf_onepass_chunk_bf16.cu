// sf_onepass_chunk_bf16.cu -- __nv_bfloat16 instantiations of the many-class one-pass walk (kernel: sf_onepass_chunk.cuh).
#include "sf_onepass_chunk.cuh"

namespace sf {
int onepass_chunk_dispatch_bf16(const OnepassArgs& a, int s) { return oc_dispatch<__nv_bfloat16>(a, s); }
}  // namespace sf
