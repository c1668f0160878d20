// Developer micro-benchmark: SM clock seen by a short kernel (clock64 vs globaltimer), cold and after a warm-up burst.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* res, int iters) {
    double a = out[threadIdx.x];
    unsigned long long g0, g1;
    long long c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    for (int i = 0; i < iters; ++i) a = fma(a, 1.0000001, 0.5);
    long long c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (threadIdx.x == 0 && blockIdx.x == 0) { res[0] = c1 - c0; res[1] = (long long)(g1 - g0); }
    out[threadIdx.x] = a;
}
int main() {
    double* out; long long* res; long long h[2];
    cudaMalloc(&out, 8 * 1024); cudaMemset(out, 0, 8 * 1024); cudaMalloc(&res, 16);
    for (int rep = 0; rep < 6; ++rep) {
        k<<<148, 640>>>(out, res, 20000);
        cudaMemcpy(h, res, 16, cudaMemcpyDeviceToHost);
        printf("rep %d: %lld cycles in %lld ns -> %.0f MHz (8 cycles per dependent DFMA: %.2f)\n", rep, h[0], h[1], 1e3 * h[0] / h[1], (double)h[0] / 20000);
    }
    // many tiny kernels back to back, like an ICP chain
    for (int rep = 0; rep < 3; ++rep) {
        for (int i = 0; i < 200; ++i) k<<<148, 640>>>(out, res, 500);
        cudaMemcpy(h, res, 16, cudaMemcpyDeviceToHost);
        printf("tiny kernels: %lld cycles in %lld ns -> %.0f MHz\n", h[0], h[1], 1e3 * h[0] / h[1]);
    }
    return 0;
}
