// Micro-probe: per-SMSP issue rate of FFMA, FFMA2 (packed fp32x2), MUFU.EX2, FSETP+SEL with 1..8 warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int ILP>
__global__ void k(float* out, int iters, long long* cyc) {
    float a[ILP]; float2 a2[ILP]; int idx[ILP];
    for (int i = 0; i < ILP; ++i) { a[i] = threadIdx.x * 0.001f + i; a2[i] = make_float2(a[i], a[i] + 1.f); idx[i] = i; }
    const float b = 1.0001f, c = 0.5f; const float2 b2 = make_float2(b, b), c2 = make_float2(c, c);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) a[i] = __fmaf_rn(a[i], b, c);
            if (MODE == 1) a2[i] = __ffma2_rn(a2[i], b2, c2);
            if (MODE == 2) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a[i])); a[i] = y; }
            if (MODE == 3) { idx[i] = (a[i] > c + it) ? it : idx[i]; a[i] += 1.0f; }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < ILP; ++i) s += a[i] + a2[i].x + a2[i].y + idx[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE, int ILP> void run(const char* name, int warps) {
    float* out; long long* cyc; cudaMalloc(&out, 4 * 1024 * 148); cudaMalloc(&cyc, 8);
    int iters = 2000;
    k<MODE, ILP><<<1, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
    k<MODE, ILP><<<1, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_smsp_warps = warps / 4.0;
    printf("%-8s ILP=%d warps/SM=%2d: %.2f cycles per warp-instr per SMSP (issue interval), %.1f cyc per instr per warp\n", name, ILP, warps,
           (double)h / (iters * ILP * per_smsp_warps), (double)h / (iters * ILP));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {4, 8, 16, 32}) { run<0, 8>("FFMA", w); run<1, 8>("FFMA2", w); run<2, 8>("MUFU.EX2", w); run<3, 8>("FSETP+SEL+FADD", w); }
    run<0, 1>("FFMA", 4); run<1, 1>("FFMA2", 4); run<2, 1>("MUFU.EX2", 4);
    run<0, 2>("FFMA", 4); run<1, 2>("FFMA2", 4); run<2, 2>("MUFU.EX2", 4); run<1, 4>("FFMA2", 4); run<2, 4>("MUFU.EX2", 4);
    return 0;
}
