// debug probe: three formulations of the FAST-9 strength at given pixels of a raw 8-bit image file (w h x y ...)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
__device__ int strength_a(const unsigned char* img, int stride) {
    const int v = img[0];
    constexpr int RX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    constexpr int RY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = v - (int)img[RY[k] * stride + RX[k]];
    int best = -256;
#pragma unroll
    for (int s0 = 0; s0 < 16; s0++) {
        int a = 256, bb = 256;
#pragma unroll
        for (int j = 0; j < 9; j++) {
            a = min(a, d[(s0 + j) & 15]);
            bb = min(bb, -d[(s0 + j) & 15]);
        }
        best = max(best, max(a, bb));
    }
    return best;
}
__device__ int strength_b(const unsigned char* img, int stride) {
    const int v = img[0];
    const int RX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    const int RY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    int d[16];
#pragma unroll 1
    for (int k = 0; k < 16; k++) d[k] = v - (int)img[RY[k] * stride + RX[k]];
    int best = -256;
#pragma unroll 1
    for (int s0 = 0; s0 < 16; s0++) {
        int a = 256, bb = 256;
#pragma unroll 1
        for (int j = 0; j < 9; j++) {
            const int t = d[(s0 + j) & 15];
            if (t < a) a = t;
            if (-t < bb) bb = -t;
        }
        const int m = a > bb ? a : bb;
        if (m > best) best = m;
    }
    return best;
}
__global__ void k(const unsigned char* image, int stride, int x, int y, int* out) {
    const unsigned char* img = image + (size_t)y * stride + x;
    out[0] = strength_a(img, stride);
    out[1] = strength_b(img, stride);
}
int main(int argc, char** argv) {
    const int w = 61, h = 41;
    std::vector<unsigned char> im(w * h);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(im.data(), 1, w * h, f) != (size_t)(w * h)) return 1;
    unsigned char* d;
    int* o;
    cudaMalloc(&d, w * h);
    cudaMalloc(&o, 16);
    cudaMemcpy(d, im.data(), w * h, cudaMemcpyHostToDevice);
    const int pts[3][2] = {{56, 37}, {3, 3}, {57, 8}};
    for (auto& p : pts) {
        k<<<1, 1>>>(d, w, p[0], p[1], o);
        int r[2];
        cudaMemcpy(r, o, 8, cudaMemcpyDeviceToHost);
        printf("(%d,%d): unrolled min/max %d   plain loops %d   [%s]\n", p[0], p[1], r[0], r[1], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
