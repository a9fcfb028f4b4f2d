// Do the FMA and MUFU pipes overlap for one warp / for 4 warps per scheduler?  Body per iteration and per chain:
// FMUL2, FFMA2, FFMA2, 2x MUFU.EX2, FFMA2, FFMA2 (the backward's channel-pair body).  ILP = independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, bool MUFU>
__global__ void k(float* out, int iters, long long* cyc) {
    float2 t[ILP], b[ILP], a0[ILP], a1[ILP];
    for (int i = 0; i < ILP; ++i) { t[i] = make_float2(threadIdx.x * 0.001f + i, 0.5f); b[i] = make_float2(0.3f, 0.2f * i); a0[i] = a1[i] = make_float2(0, 0); }
    const float2 l0 = make_float2(0.625f, 0.625f), l1 = make_float2(0.375f, 0.375f), l2e = make_float2(1.44f, 1.44f), nl = make_float2(-3.f, -3.f), v0 = make_float2(0.1f, 0.1f), v1 = make_float2(0.2f, 0.2f);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int i = 0; i < ILP; ++i) {
            float2 z = __ffma2_rn(t[i], l0, __fmul2_rn(b[i], l1));
            float2 e = __ffma2_rn(z, l2e, nl);
            float2 pc;
            if (MUFU) { asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(pc.x) : "f"(e.x)); asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(pc.y) : "f"(e.y)); }
            else pc = e;
            a0[i] = __ffma2_rn(pc, v0, a0[i]);
            a1[i] = __ffma2_rn(pc, v1, a1[i]);
            t[i].x += 1e-3f;
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < ILP; ++i) s += a0[i].x + a0[i].y + a1[i].x + a1[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, bool MUFU> void run(int warps) {
    float* out; long long* cyc; cudaMalloc(&out, 4 * 2048); cudaMalloc(&cyc, 8);
    int iters = 2000;
    for (int r = 0; r < 2; ++r) { k<ILP, MUFU><<<1, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize(); }
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ILP=%d mufu=%d warps/SMSP=%d: %.1f cycles per pair-body per SMSP (FMA-only floor %.1f, MUFU floor %.1f)\n", ILP, (int)MUFU, warps / 4,
           (double)h / (iters * ILP * (warps / 4.0)), 5 * 2.14 + 1, MUFU ? 16.0 : 0.0);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<3, true>(4); run<3, true>(16); run<3, false>(16); run<6, true>(16); run<10, true>(16); run<10, true>(24); run<10, true>(32); run<10, false>(16);
    return 0;
}
