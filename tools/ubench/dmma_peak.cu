// fp64 tensor-pipe (DMMA.8x8x4) issue-rate probe on sm_100a: register-only mma.sync loops, no memory traffic.
// Prints TFLOP/s for 1..16 warps per SM and 8/16/32 independent accumulator pairs per warp: the ceiling our GEMM tiles and
// cuBLAS DGEMM are measured against.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void k_dmma(double* out, int iters) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
void run(double* out, int warps, int ctas_per_sm) {
    const int iters = 4000, blocks = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_dmma<NACC><<<blocks, warps * 32>>>(out, iters);
    cudaEventRecord(e0);
    k_dmma<NACC><<<blocks, warps * 32>>>(out, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 512.0 * NACC * iters * warps * blocks;
    printf("acc %2d warps/CTA %2d CTAs/SM %d : %.3f ms  %.2f TFLOP/s\n", NACC, warps, ctas_per_sm, ms, flops / (ms * 1e-3) / 1e12);
}
int main() {
    double* out;
    cudaMalloc(&out, 64 << 20);
    for (int warps : {1, 2, 4, 8, 16}) {
        run<8>(out, warps, 1);
        run<16>(out, warps, 1);
        run<32>(out, warps, 1);
    }
    run<32>(out, 4, 2);
    run<32>(out, 8, 2);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
