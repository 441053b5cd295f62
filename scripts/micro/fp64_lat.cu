// Microbenchmark: dependent-chain latency and per-SM throughput of FP64 DFMA / DADD / division on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat_kernel(double* out, long long* cyc, int iters, double a, double b) {
    double x = a + threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) { x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); }
    long long t1 = clock64();
    double y = a + threadIdx.x * 1e-9;
    for (int i = 0; i < iters; i++) { y = y / b + a; }
    long long t2 = clock64();
    float z = (float)a;
    for (int i = 0; i < iters; i++) { z = fmaf(z, (float)b, (float)a); z = fmaf(z, (float)b, (float)a); z = fmaf(z, (float)b, (float)a); z = fmaf(z, (float)b, (float)a); }
    long long t3 = clock64();
    if (threadIdx.x == 0) { cyc[blockIdx.x * 3 + 0] = t1 - t0; cyc[blockIdx.x * 3 + 1] = t2 - t1; cyc[blockIdx.x * 3 + 2] = t3 - t2; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + y + z;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 1 << 16);
    const int iters = 20000;
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        lat_kernel<<<1, warps * 32>>>(out, cyc, iters, 1.000001, 0.999999);
        cudaDeviceSynchronize();
        long long h[3]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("warps/SM %2d: DFMA chain %.1f cyc/op | (div+add) chain %.1f cyc/iter | FFMA chain %.1f cyc/op\n", warps,
               (double)h[0] / (4.0 * iters), (double)h[1] / iters, (double)h[2] / (4.0 * iters));
    }
    return 0;
}
