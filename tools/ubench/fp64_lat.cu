// fp64 latency / throughput probe on sm_100a: dependent DFMA chain, independent DFMA streams, rsqrt/sqrt/div chains, LDS+DFMA chain.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(double* out, long long* cyc, int iters) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double a = 1.000001 + threadIdx.x * 1e-9, b = 0.999999, c = 1e-7;
    long long t0, t1;
    // 1. dependent DFMA chain
    t0 = clock64();
    for (int i = 0; i < iters; i++) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // 2. 8 independent chains
    double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
    t0 = clock64();
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, b, c); x1 = fma(x1, b, c); x2 = fma(x2, b, c); x3 = fma(x3, b, c);
        x4 = fma(x4, b, c); x5 = fma(x5, b, c); x6 = fma(x6, b, c); x7 = fma(x7, b, c);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    a = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    // 3. rsqrt chain
    t0 = clock64();
    for (int i = 0; i < iters; i++) { a = rsqrt(a + 1.5); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // 4. sqrt chain
    t0 = clock64();
    for (int i = 0; i < iters; i++) { a = sqrt(a + 1.5); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // 5. div chain
    t0 = clock64();
    for (int i = 0; i < iters; i++) { a = 1.7 / (a + 1.5); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // 6. LDS + DFMA dependent on accumulator only (address independent)
    t0 = clock64();
    int idx = threadIdx.x & 31;
    for (int i = 0; i < iters; i++) { a = fma(sm[(idx + 4 * i) & 1023], sm[(idx + 4 * i + 1) & 1023], a); a = fma(sm[(idx + 4 * i + 2) & 1023], sm[(idx + 4 * i + 3) & 1023], a); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // 7. __syncthreads cost
    t0 = clock64();
    for (int i = 0; i < iters; i++) { __syncthreads(); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 << 20); cudaMalloc(&cyc, 64);
    for (int threads : {32, 256, 1024}) {
        for (int blocks : {1, 148 * 2}) {
            const int iters = 2000;
            k_lat<<<blocks, threads>>>(out, cyc, iters);
            k_lat<<<blocks, threads>>>(out, cyc, iters);
            cudaDeviceSynchronize();
            long long h[8]; cudaMemcpy(h, cyc, 56, cudaMemcpyDeviceToHost);
            printf("threads %4d blocks %3d | dep DFMA %.1f cyc/op | 8-indep DFMA %.2f cyc/op/thread | rsqrt %.0f | sqrt %.0f | div %.0f | LDS+DFMA pair %.1f | bar %.0f\n", threads, blocks,
                   h[0] / (4.0 * iters), h[1] / (8.0 * iters), (double)h[2] / iters, (double)h[3] / iters, (double)h[4] / iters, h[5] / (2.0 * iters), (double)h[6] / iters);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
