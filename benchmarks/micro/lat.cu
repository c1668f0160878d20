// Developer micro-benchmark: latency / throughput of the instruction classes the ICP step is made of (sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int warps_mode) {
    double a = out[threadIdx.x], b = 1.0000001, c = 0.5;
    float f = (float)a;
    long long t0, t1;
    // 1. dependent DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) a = fma(a, b, c);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // 2. dependent un-fused mul/add
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 128; ++i) a = __dadd_rn(__dmul_rn(a, b), c);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // 3. dependent division
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) a = __ddiv_rn(c, a) + 1.5;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // 4. independent DFMA (8 chains)
    double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        x0 = fma(x0, b, c); x1 = fma(x1, b, c); x2 = fma(x2, b, c); x3 = fma(x3, b, c);
        x4 = fma(x4, b, c); x5 = fma(x5, b, c); x6 = fma(x6, b, c); x7 = fma(x7, b, c);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    a = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    // 5. float -> double conversions, independent
    float g0 = f, g1 = f + 1, g2 = f + 2, g3 = f + 3;
    double s = 0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        s += (double)g0; g0 += 1.f; s += (double)g1; g1 += 1.f; s += (double)g2; g2 += 1.f; s += (double)g3; g3 += 1.f;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    a += s;
    // 6. shuffle chain (double = 2 SHFL)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) a = __shfl_xor_sync(0xffffffffu, a, 1) + 1.0;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // 7. sqrt chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) a = sqrt(a + 2.0);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    out[threadIdx.x + blockIdx.x * blockDim.x] = a;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 148 * 1024); cudaMemset(out, 0, 8 * 148 * 1024);
    cudaMalloc(&cyc, 64);
    const char* names[7] = {"256 dependent DFMA", "128 dependent DMUL+DADD", "32 dependent DDIV+DADD", "64x8 independent DFMA",
                            "256 F2F.F64.F32 (+DADD+FADD)", "64 double SHFL+DADD", "32 DSQRT+DADD"};
    for (int threads : {32, 128, 640}) {
        for (int rep = 0; rep < 2; ++rep) k<<<1, threads>>>(out, cyc, 0);
        long long h[8];
        cudaMemcpy(h, cyc, 56, cudaMemcpyDeviceToHost);
        printf("threads/SM = %d\n", threads);
        for (int i = 0; i < 7; ++i) printf("  %-34s %6lld cycles\n", names[i], h[i]);
    }
    return 0;
}
