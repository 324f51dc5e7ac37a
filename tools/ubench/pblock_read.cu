// HBM ceiling for the support-scoring access pattern (6 x 6 fp64 blocks P[y_j, y_p] of a column-major covariance): pure loads, no
// math.  Tells how much of the gap between k_ransac_support and the streaming-copy peak is the ACCESS PATTERN (scattered column
// segments) rather than the kernel.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pblock_read pblock_read.cu
//   variants:  stream   : grid-stride sequential read of the same number of bytes
//              tile JxH : CTA reads H hypotheses x 6 columns x (6 J) rows (J matches), rows coalesced down the columns,
//                         CTAs ordered J-tile fastest (like the kernel); L loads in flight per thread
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void k_stream(const double* __restrict__ P, size_t n, double* out) {
    double s = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * 4) {
        double a = P[i], b = i + stride < n ? P[i + stride] : 0, c = i + 2 * stride < n ? P[i + 2 * stride] : 0, d = i + 3 * stride < n ? P[i + 3 * stride] : 0;
        s += a + b + c + d;
    }
    if (s == 1.2345) out[0] = s;
}
// rows per CTA = 6 * J; thread handles rows tid, tid + T, ...; H hypotheses per CTA, HL of them in flight at once (6 * HL loads)
template <int J, int H, int HL, int T>
__global__ void __launch_bounds__(T) k_tile(const double* __restrict__ P, int ld, int m, const int* __restrict__ hyp_off, int nhyp, double* out) {
    const int J0 = blockIdx.x * J;
    const int h0 = blockIdx.y * H;
    double s = 0;
    for (int k = threadIdx.x; k < 6 * J; k += T) {
        const int jj = J0 + k / 6;
        if (jj >= m) continue;
        const int row = 13 + 6 * jj + k % 6;
#pragma unroll 1
        for (int g = 0; g < H; g += HL) {
            double v[HL][6];
#pragma unroll
            for (int q = 0; q < HL; q++) {
                const int hh = h0 + g + q;
                const double* col = P + (size_t)hyp_off[hh < nhyp ? hh : 0] * ld + row;
#pragma unroll
                for (int c = 0; c < 6; c++) v[q][c] = col[(size_t)c * ld];
            }
#pragma unroll
            for (int q = 0; q < HL; q++)
#pragma unroll
                for (int c = 0; c < 6; c++) s += v[q][c];
        }
    }
    if (s == 1.2345) out[0] = s;
}
template <int J, int H, int HL, int T>
void run_tile(const char* name, const double* P, int ld, int m, const int* hyp_off, int nhyp, double* out) {
    dim3 grid((m + J - 1) / J, (nhyp + H - 1) / H);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_tile<J, H, HL, T><<<grid, T>>>(P, ld, m, hyp_off, nhyp, out);
    cudaEventRecord(e0);
    k_tile<J, H, HL, T><<<grid, T>>>(P, ld, m, hyp_off, nhyp, out);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = 288.0 * m * nhyp;
    printf("%-28s grid %5d x %5d  %.3f ms  %.0f GB/s  (%s)\n", name, grid.x, grid.y, ms, bytes / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main(int argc, char** argv) {
    const int m = 5000, n = 13 + 6 * m, ld = (n + 15) / 16 * 16;
    const int nhyp = argc > 1 ? atoi(argv[1]) : 5000;
    double* P;
    if (cudaMalloc(&P, (size_t)ld * n * 8) != cudaSuccess) return 1;
    cudaMemset(P, 0, (size_t)ld * n * 8);
    int* hoff_h = (int*)malloc(sizeof(int) * nhyp);
    unsigned long long st = 99;
    for (int i = 0; i < nhyp; i++) {  // distinct hypotheses in order (dedupe sweep) or random draws (brute force)
        st = st * 6364136223846793005ULL + 1442695040888963407ULL;
        const int feat = nhyp <= m ? i : (int)((st >> 33) % m);
        hoff_h[i] = 13 + 6 * feat;
    }
    int* hoff;
    cudaMalloc(&hoff, sizeof(int) * nhyp);
    cudaMemcpy(hoff, hoff_h, sizeof(int) * nhyp, cudaMemcpyHostToDevice);
    double* out;
    cudaMalloc(&out, 64);
    {
        const size_t nelem = (size_t)36 * m * nhyp < (size_t)ld * n ? (size_t)36 * m * nhyp : (size_t)ld * n;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        k_stream<<<148 * 8, 256>>>(P, nelem, out);
        cudaEventRecord(e0);
        k_stream<<<148 * 8, 256>>>(P, nelem, out);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%-28s %.3f ms  %.0f GB/s\n", "stream", ms, nelem * 8.0 / (ms * 1e-3) / 1e9);
    }
    run_tile<64, 8, 2, 256>("tile 64x8 L12 T256 (v1)", P, ld, m, hoff, nhyp, out);
    run_tile<64, 8, 4, 256>("tile 64x8 L24 T256", P, ld, m, hoff, nhyp, out);
    run_tile<64, 8, 8, 128>("tile 64x8 L48 T128", P, ld, m, hoff, nhyp, out);
    run_tile<32, 6, 6, 192>("tile 32x6 L36 T192 (v2)", P, ld, m, hoff, nhyp, out);
    run_tile<128, 4, 4, 256>("tile 128x4 L24 T256", P, ld, m, hoff, nhyp, out);
    run_tile<256, 2, 2, 256>("tile 256x2 L12 T256", P, ld, m, hoff, nhyp, out);
    run_tile<256, 4, 4, 512>("tile 256x4 L24 T512", P, ld, m, hoff, nhyp, out);
    run_tile<512, 2, 2, 512>("tile 512x2 L12 T512", P, ld, m, hoff, nhyp, out);
    run_tile<1024, 1, 1, 1024>("tile 1024x1 L6 T1024", P, ld, m, hoff, nhyp, out);
    run_tile<1024, 2, 2, 1024>("tile 1024x2 L12 T1024", P, ld, m, hoff, nhyp, out);
    return 0;
}
