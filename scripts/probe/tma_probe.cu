// Minimal TMA probe: loads a [NCH][BH][BW] box at (x0,y0,z0) and dumps it. Build: see scripts/probe/Makefile
#include <cstdio>
#include <vector>
#include "../../segmentation_factory_b200/csrc/sf_common.cuh"
#include "../../segmentation_factory_b200/csrc/sf_launch.h"
using namespace sf;

__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int n, int x0, int y0, int z0, int mode) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* S = reinterpret_cast<float*>(smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + ((n * 4 + 127) & ~127));
    if (threadIdx.x == 0) {
        if (mode & 1) tma_prefetch_desc(&tmap);
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, n * 4);
        tma_load_3d(S, &tmap, bar, x0, y0, z0);
    }
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, 0)) { if (++spins > (1u << 22)) { if (threadIdx.x == 0) printf("TIMEOUT\n"); return; } }
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = S[i];
}

int main(int argc, char** argv) {
    int planes = 38, h = 32, w = 64, bc = 19, bh = 5, bw = 20;
    std::vector<float> hsrc((size_t)planes * h * w);
    for (size_t i = 0; i < hsrc.size(); ++i) hsrc[i] = (float)i;
    float *d, *o; int n = bc * bh * bw;
    cudaMalloc(&d, hsrc.size() * 4); cudaMalloc(&o, n * 4);
    cudaMemcpy(d, hsrc.data(), hsrc.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap m{};
    int rc = make_logits_tmap(&m, d, SF_F32, planes, h, w, bc, bh, bw);
    printf("encode rc=%d\n", rc);
    int x0 = atoi(argv[1]), y0 = atoi(argv[2]), z0 = atoi(argv[3]), mode = atoi(argv[4]);
    size_t smem = ((n * 4 + 127) & ~127) + 64;
    probe<<<1, 128, smem>>>(m, o, n, x0, y0, z0, mode);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> r(n);
    cudaMemcpy(r.data(), o, n * 4, cudaMemcpyDeviceToHost);
    printf("mode %d (%d,%d,%d): err=%s  S[0..3]=%g %g %g %g  S[bw]=%g S[bh*bw]=%g\n", mode, x0, y0, z0, cudaGetErrorString(e),
           r[0], r[1], r[2], r[3], r[bw], r[bh * bw]);
    return 0;
}
