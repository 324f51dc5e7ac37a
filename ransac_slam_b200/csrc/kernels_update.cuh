// kernels_update.cuh -- (d) joint EKF update of ExtendKF::update (src/ExtendKF.cpp:597-639) for the low- / high-innovation
// inlier sets (src/ExtendKF.cpp:559-596, 640-678), re-formulated for the GPU:
//
//     W = P H^T                      (gather product over the 7 + 6 structurally non-zero columns of each H_i)
//     S = H W + I                    (k x k, k = 2m)
//     S = L L^T                      (blocked right-looking Cholesky, NB = 64, trailing updates on fp64 tensor cores)
//     V = W L^-T ,  y = L^-1 nu      (blocked TRSM; nu rides along as row n of W)
//     x+ = x + V y
//     P+ = P - V V^T                 (SYRK on DMMA tiles, lower triangle + mirrored store => exactly symmetric, which is what
//                                     the reference's 0.5*(P + P^T) produces from its K S K^T form)
//     quaternion normalisation + Jnorm fix-up of rows/cols 3..6 (with the reference's pow(s,-3/2) -> s^-1 quirk, Q4)
//
// In exact arithmetic this equals the reference's K = P H^T S^-1, P - K S K^T.  All fp64.  P stays resident in HBM and is
// updated in place.  m (and therefore k) is data dependent and lives in device memory; grids are sized for the worst case and
// surplus CTAs exit immediately, so the whole frame is a fixed launch sequence (CUDA-graph friendly, no host round trip).
#pragma once
#include "common.cuh"

namespace rslam {

// ---- U1: ordered list of the inlier set + innovation ----------------------------------------------------------------
// which = 0: low_innovation_inlier, prior state x_k_km1 (:559-596); which = 1: high_innovation_inlier, state x_k_k (:640-678)
__global__ void __launch_bounds__(256) k_upd_gather(DevFilter* Fs, int which) {
    DevFilter& F = Fs[blockIdx.y];
    __shared__ int s_scan[256];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const unsigned char* flag = which == 0 ? F.li : F.hi;
    for (int base = 0; base < F.N; base += 256) {
        const int i = base + threadIdx.x;
        const int a = (i < F.N && flag[i]) ? 1 : 0;
        s_scan[threadIdx.x] = a;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {
            int v = 0;
            if (threadIdx.x >= o) v = s_scan[threadIdx.x - o];
            __syncthreads();
            s_scan[threadIdx.x] += v;
            __syncthreads();
        }
        if (a) {
            const int t = s_base + s_scan[threadIdx.x] - 1;
            F.upd_list[t] = i;
            // innovation z - h rides along as row n of W
            F.W[F.n + (size_t)(2 * t) * F.ldw] = F.z[2 * i] - F.h[2 * i];
            F.W[F.n + (size_t)(2 * t + 1) * F.ldw] = F.z[2 * i + 1] - F.h[2 * i + 1];
        }
        __syncthreads();
        if (threadIdx.x == 255) s_base += s_scan[255];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        F.ctl[CTL_M] = s_base;
        F.ctl[CTL_K] = 2 * s_base;
    }
}

// ---- U2: W = P H^T.  Thread per state row r, CTA handles a chunk of measurements. ---------------------------------------
constexpr int kWChunk = 32;  // measurements (features) per CTA in y
__global__ void __launch_bounds__(256) k_upd_W(DevFilter* Fs) {
    DevFilter& F = Fs[blockIdx.z];
    const int m = F.ctl[CTL_M];
    const int t0 = blockIdx.y * kWChunk;
    if (t0 >= m) return;
    const int t1 = min(m, t0 + kWChunk);
    __shared__ double sH[kWChunk][26];
    __shared__ int sOff[kWChunk], sFs[kWChunk];
    for (int e = threadIdx.x; e < (t1 - t0) * 26; e += blockDim.x) {
        const int tt = e / 26, c = e % 26;
        const int f = F.upd_list[t0 + tt];
        sH[tt][c] = c < 14 ? F.Hc[14 * f + c] : F.Hf[12 * f + (c - 14)];
    }
    for (int e = threadIdx.x; e < (t1 - t0); e += blockDim.x) {
        const int f = F.upd_list[t0 + e];
        sOff[e] = F.foff[f];
        sFs[e] = F.ftype[f] == 0 ? 6 : 3;
    }
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= F.n) return;
    const int ld = F.ldp;
    const double* P = F.P;
    double pc[7];
#pragma unroll
    for (int c = 0; c < 7; c++) pc[c] = P[r + (size_t)c * ld];
    for (int tt = 0; tt < t1 - t0; tt++) {
        double w0 = 0, w1 = 0;
#pragma unroll
        for (int c = 0; c < 7; c++) {
            w0 += pc[c] * sH[tt][c];
            w1 += pc[c] * sH[tt][7 + c];
        }
        const int off = sOff[tt], fs = sFs[tt];
        for (int c = 0; c < fs; c++) {
            const double p = P[r + (size_t)(off + c) * ld];
            w0 += p * sH[tt][14 + c];
            w1 += p * sH[tt][20 + c];
        }
        F.W[r + (size_t)(2 * (t0 + tt)) * F.ldw] = w0;
        F.W[r + (size_t)(2 * (t0 + tt) + 1) * F.ldw] = w1;
    }
}

// ---- U3: S = H W + I (lower block triangle only). Thread per (ti >= tj) measurement pair. ------------------------------
__global__ void __launch_bounds__(256) k_upd_S(DevFilter* Fs) {
    DevFilter& F = Fs[blockIdx.z];
    const int m = F.ctl[CTL_M];
    const int tj = blockIdx.y * 16 + (threadIdx.x >> 4);
    const int ti = blockIdx.x * 16 + (threadIdx.x & 15);
    if (ti >= m || tj >= m || ti < tj) return;
    const int f = F.upd_list[ti];
    const int off = F.foff[f], fs = F.ftype[f] == 0 ? 6 : 3;
    const double* Hc = F.Hc + 14 * f;
    const double* Hf = F.Hf + 12 * f;
    const double* w0 = F.W + (size_t)(2 * tj) * F.ldw;
    const double* w1 = w0 + F.ldw;
    double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
#pragma unroll
    for (int c = 0; c < 7; c++) {
        const double a = w0[c], b = w1[c];
        s00 += Hc[c] * a;
        s01 += Hc[c] * b;
        s10 += Hc[7 + c] * a;
        s11 += Hc[7 + c] * b;
    }
    for (int c = 0; c < fs; c++) {
        const double a = w0[off + c], b = w1[off + c];
        s00 += Hf[c] * a;
        s01 += Hf[c] * b;
        s10 += Hf[6 + c] * a;
        s11 += Hf[6 + c] * b;
    }
    if (ti == tj) {
        s00 += 1.0;  // R = I (src/ExtendKF.cpp:594,676)
        s11 += 1.0;
        // keep the diagonal block exactly symmetric: use the lower entry for both
        s01 = s10;
    }
    double* S = F.Sm;
    const int ld = F.lds;
    S[(2 * ti) + (size_t)(2 * tj) * ld] = s00;
    S[(2 * ti + 1) + (size_t)(2 * tj) * ld] = s10;
    S[(2 * ti) + (size_t)(2 * tj + 1) * ld] = s01;
    S[(2 * ti + 1) + (size_t)(2 * tj + 1) * ld] = s11;
}

// ---- shared helper: solve X * L^T = A for one row held in registers (forward substitution), L (NB x NB, lower) in smem ----
// sL is padded to identity beyond the active width.
__device__ __forceinline__ void row_solve_LT(double* xr, const double (*sL)[kNB + 1]) {
#pragma unroll
    for (int c = 0; c < kNB; c++) {
        double s = xr[c];
#pragma unroll
        for (int t = 0; t < c; t++) s -= xr[t] * sL[c][t];
        xr[c] = s / sL[c][c];
    }
}

// factor the NB x NB diagonal block held in smem (unblocked right-looking Cholesky, all threads of the CTA)
__device__ __forceinline__ void smem_potrf(double (*sL)[kNB + 1], int w) {
    for (int j = 0; j < w; j++) {
        __syncthreads();
        const double d = sqrt(sL[j][j]);
        __syncthreads();
        for (int i = j + threadIdx.x; i < w; i += blockDim.x) sL[i][j] = (i == j) ? d : sL[i][j] / d;
        __syncthreads();
        const int rem = w - j - 1;
        for (int e = threadIdx.x; e < rem * rem; e += blockDim.x) {
            const int i = j + 1 + e / rem, c = j + 1 + e % rem;
            if (i >= c) sL[i][c] -= sL[i][j] * sL[c][j];
        }
    }
    __syncthreads();
}

// ---- U4a: Cholesky panel at block column `step`: factor the diagonal block (every CTA redundantly, CTA 0 stores it) and
//           solve the sub-diagonal blocks  L[rb, step] = S[rb, step] * L_jj^-T.  grid.x = row blocks below (and incl.) diagonal
__global__ void __launch_bounds__(128) k_chol_panel(DevFilter* Fs, int step) {
    DevFilter& F = Fs[blockIdx.y];
    const int kk = F.ctl[CTL_K];
    const int j0 = kNB * step;
    if (j0 >= kk) return;
    const int r0 = j0 + kNB * blockIdx.x;
    if (r0 >= kk) return;
    const int w = min(kNB, kk - j0);
    __shared__ double sL[kNB][kNB + 1];
    double* S = F.Sm;
    const int ld = F.lds;
    for (int e = threadIdx.x; e < kNB * kNB; e += blockDim.x) {
        const int i = e % kNB, c = e / kNB;
        double v = (i == c) ? 1.0 : 0.0;
        if (i < w && c < w && i >= c) v = S[(j0 + i) + (size_t)(j0 + c) * ld];
        sL[i][c] = v;
    }
    smem_potrf(sL, w);
    if (blockIdx.x == 0) {
        for (int e = threadIdx.x; e < kNB * kNB; e += blockDim.x) {
            const int i = e % kNB, c = e / kNB;
            if (i < w && c < w) S[(j0 + i) + (size_t)(j0 + c) * ld] = (i >= c) ? sL[i][c] : 0.0;
        }
        return;
    }
    const int r = r0 + threadIdx.x;
    if (threadIdx.x >= kNB || r >= kk) return;
    double xr[kNB];
#pragma unroll
    for (int c = 0; c < kNB; c++) xr[c] = (c < w) ? S[r + (size_t)(j0 + c) * ld] : 0.0;
    row_solve_LT(xr, sL);
#pragma unroll
    for (int c = 0; c < kNB; c++)
        if (c < w) S[r + (size_t)(j0 + c) * ld] = xr[c];
}

// ---- U5a: TRSM panel at block column `step`:  V[:, step] = W[:, step] * L_jj^-T  for the n+1 rows of W --------------------
__global__ void __launch_bounds__(128) k_trsm_panel(DevFilter* Fs, int step) {
    DevFilter& F = Fs[blockIdx.y];
    const int kk = F.ctl[CTL_K];
    const int j0 = kNB * step;
    if (j0 >= kk) return;
    const int w = min(kNB, kk - j0);
    __shared__ double sL[kNB][kNB + 1];
    const double* S = F.Sm;
    const int ld = F.lds;
    for (int e = threadIdx.x; e < kNB * kNB; e += blockDim.x) {
        const int i = e % kNB, c = e / kNB;
        double v = (i == c) ? 1.0 : 0.0;
        if (i < w && c < w && i >= c) v = S[(j0 + i) + (size_t)(j0 + c) * ld];
        sL[i][c] = v;
    }
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > F.n) return;  // rows 0..n (row n = innovation)
    double* Wp = F.W;
    const int ldw = F.ldw;
    double xr[kNB];
#pragma unroll
    for (int c = 0; c < kNB; c++) xr[c] = (c < w) ? Wp[r + (size_t)(j0 + c) * ldw] : 0.0;
    row_solve_LT(xr, sL);
#pragma unroll
    for (int c = 0; c < kNB; c++)
        if (c < w) Wp[r + (size_t)(j0 + c) * ldw] = xr[c];
}

// ---- fp64 tensor-core GEMM:  C -= A * B^T  (A: M x K, B: N x K, C: M x N, all column-major) ------------------------------
// DMMA m8n8k4 (mma.sync.aligned.m8n8k4.row.col.f64): tcgen05 has no fp64 kind, so mma.sync DMMA is the fp64 tensor path on
// sm_100a.  CTA tile 128 x 128 x 16, 8 warps (2 x 4), warp tile 64 x 32, 3-stage cp.async pipeline, padded smem (+4 doubles per
// k-row) so that the fragment loads are bank-conflict free.
constexpr int GBM = 128, GBN = 128, GBK = 16, GSTAGES = 3, GPAD = 4;
constexpr int GLDS = GBM + GPAD;  // 132 doubles per k-row
constexpr int kGemmSmemBytes = GSTAGES * 2 * GBK * GLDS * (int)sizeof(double);

enum GemmMode { GEMM_SYRK_P = 0, GEMM_CHOL_TRAIL = 1, GEMM_TRSM_TRAIL = 2 };

struct GemmProb {
    const double* A;
    const double* B;
    double* C;
    int lda, ldb, ldc, M, N, K;
    bool lower, mirror;
};

__device__ __forceinline__ bool gemm_setup(const DevFilter& F, int mode, int step, GemmProb& g) {
    const int kk = F.ctl[CTL_K];
    if (mode == GEMM_SYRK_P) {
        if (kk <= 0) return false;
        g.A = g.B = F.W;
        g.lda = g.ldb = F.ldw;
        g.C = F.P;
        g.ldc = F.ldp;
        g.M = g.N = F.n;
        g.K = kk;
        g.lower = g.mirror = true;
        return true;
    }
    const int o = kNB * (step + 1);
    if (kk <= o) return false;
    if (mode == GEMM_CHOL_TRAIL) {
        g.A = g.B = F.Sm + o + (size_t)(o - kNB) * F.lds;
        g.lda = g.ldb = F.lds;
        g.C = F.Sm + o + (size_t)o * F.lds;
        g.ldc = F.lds;
        g.M = g.N = kk - o;
        g.K = kNB;
        g.lower = true;
        g.mirror = false;
        return true;
    }
    // GEMM_TRSM_TRAIL: W[:, o:] -= V[:, o-NB:o] * L[o:, o-NB:o]^T
    g.A = F.W + (size_t)(o - kNB) * F.ldw;
    g.lda = F.ldw;
    g.B = F.Sm + o + (size_t)(o - kNB) * F.lds;
    g.ldb = F.lds;
    g.C = F.W + (size_t)o * F.ldw;
    g.ldc = F.ldw;
    g.M = F.n + 1;
    g.N = kk - o;
    g.K = kNB;
    g.lower = g.mirror = false;
    return true;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// tiles_x: number of tile columns in the launch (for full problems); lower-triangular problems enumerate (ti >= tj) linearly.
__global__ void __launch_bounds__(256, 1) k_gemm_dmma(DevFilter* Fs, int mode, int step) {
    const DevFilter& F = Fs[blockIdx.z];
    GemmProb g;
    if (!gemm_setup(F, mode, step, g)) return;
    int ti, tj;
    const int tm = (g.M + GBM - 1) / GBM, tn = (g.N + GBN - 1) / GBN;
    if (g.lower) {
        const long long t = blockIdx.x;
        if (t >= (long long)tm * (tm + 1) / 2) return;
        int i = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
        while ((long long)i * (i + 1) / 2 > t) i--;
        while ((long long)(i + 1) * (i + 2) / 2 <= t) i++;
        ti = i;
        tj = (int)(t - (long long)i * (i + 1) / 2);
    } else {
        if ((long long)blockIdx.x >= (long long)tm * tn) return;
        ti = blockIdx.x % tm;
        tj = blockIdx.x / tm;
    }
    extern __shared__ __align__(16) double gsm[];
    double* As = gsm;                               // [stage][GBK][GLDS]
    double* Bs = gsm + GSTAGES * GBK * GLDS;        // [stage][GBK][GLDS]
    const int m0 = ti * GBM, n0 = tj * GBN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm0 = (warp & 1) * 64, wn0 = (warp >> 1) * 32;
    const int nkt = (g.K + GBK - 1) / GBK;

    auto load_stage = [&](int stage, int kt) {
        const int k0 = kt * GBK;
        // A tile: GBK columns x 64 16-byte chunks; 1024 chunks / 256 threads = 4 each (same for B)
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int ch = tid + it * 256;
            const int kc = ch >> 6, r2 = ch & 63;
            const int row = m0 + 2 * r2, col = k0 + kc;
            const bool v = (row < g.M) && (col < g.K);
            const double* src = v ? (g.A + row + (size_t)col * g.lda) : g.A;
            cp_async16(&As[(stage * GBK + kc) * GLDS + 2 * r2], src, v);
        }
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int ch = tid + it * 256;
            const int kc = ch >> 6, r2 = ch & 63;
            const int row = n0 + 2 * r2, col = k0 + kc;
            const bool v = (row < g.N) && (col < g.K);
            const double* src = v ? (g.B + row + (size_t)col * g.ldb) : g.B;
            cp_async16(&Bs[(stage * GBK + kc) * GLDS + 2 * r2], src, v);
        }
    };

    double acc[8][4][2];
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

#pragma unroll
    for (int s = 0; s < GSTAGES - 1; s++) {
        if (s < nkt) load_stage(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < nkt; kt++) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + GSTAGES - 1;
            if (nk < nkt) load_stage(nk % GSTAGES, nk);
            cp_async_commit();
        }
        const double* as = As + (kt % GSTAGES) * GBK * GLDS;
        const double* bs = Bs + (kt % GSTAGES) * GBK * GLDS;
#pragma unroll
        for (int ks = 0; ks < GBK / 4; ks++) {
            const int krow = ks * 4 + (lane & 3);
            double af[8], bf[4];
#pragma unroll
            for (int mt = 0; mt < 8; mt++) af[mt] = as[krow * GLDS + wm0 + mt * 8 + (lane >> 2)];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) bf[nt] = bs[krow * GLDS + wn0 + nt * 8 + (lane >> 2)];
#pragma unroll
            for (int mt = 0; mt < 8; mt++)
#pragma unroll
                for (int nt = 0; nt < 4; nt++) dmma8x8x4(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        }
    }
    cp_async_wait<0>();
    // epilogue: C -= acc ; lower: only row >= col ; mirror: also store the transposed element
#pragma unroll
    for (int mt = 0; mt < 8; mt++) {
        const int row = m0 + wm0 + mt * 8 + (lane >> 2);
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            const int col = n0 + wn0 + nt * 8 + 2 * (lane & 3);
            if (row >= g.M) continue;
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int cc = col + e;
                if (cc >= g.N) continue;
                if (g.lower && row < cc) continue;
                double* cp = g.C + row + (size_t)cc * g.ldc;
                const double v = *cp - acc[mt][nt][e];
                *cp = v;
                if (g.mirror && row != cc) g.C[cc + (size_t)row * g.ldc] = v;
            }
        }
    }
}

// ---- U7: x+ = x + V y  (y = row n of W after the TRSM).  Thread per state row. ------------------------------------------
__global__ void __launch_bounds__(128) k_upd_x(DevFilter* Fs, int which) {
    DevFilter& F = Fs[blockIdx.y];
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= F.n) return;
    const int kk = F.ctl[CTL_K];
    const double* x0 = which == 0 ? F.x_km1 : F.x_kk;
    const double* W = F.W;
    const int ldw = F.ldw;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int c = 0;
    for (; c + 3 < kk; c += 4) {
        s0 += W[r + (size_t)c * ldw] * W[F.n + (size_t)c * ldw];
        s1 += W[r + (size_t)(c + 1) * ldw] * W[F.n + (size_t)(c + 1) * ldw];
        s2 += W[r + (size_t)(c + 2) * ldw] * W[F.n + (size_t)(c + 2) * ldw];
        s3 += W[r + (size_t)(c + 3) * ldw] * W[F.n + (size_t)(c + 3) * ldw];
    }
    for (; c < kk; c++) s0 += W[r + (size_t)c * ldw] * W[F.n + (size_t)c * ldw];
    F.x_kk[r] = x0[r] + ((s0 + s1) + (s2 + s3));
}

// ---- U9: quaternion normalisation and its Jacobian applied to P (src/ExtendKF.cpp:611-634).  One CTA per filter. -----------
__global__ void __launch_bounds__(256) k_upd_jnorm(DevFilter* Fs, ParDev par) {
    DevFilter& F = Fs[blockIdx.y];
    if (F.ctl[CTL_K] <= 0) return;  // no measurements: x_k_k = x, p_k_k = P untouched (:635-638)
    __shared__ double sJ[16];
    __shared__ double sB[16], sT[16];
    const int ld = F.ldp;
    double* P = F.P;
    const double r = F.x_kk[3], x = F.x_kk[4], y = F.x_kk[5], z = F.x_kk[6];
    __syncthreads();  // everybody has read the un-normalised quaternion
    if (threadIdx.x == 0) {
        const double s = r * r + x * x + y * y + z * z;
        const double scale = (par.quirks & RSLAM_Q4_JNORM_INT_EXPONENT) ? 1.0 / s : 1.0 / (s * sqrt(s));
        const double tv[16] = {x * x + y * y + z * z, -r * x, -r * y, -r * z, -x * r, r * r + y * y + z * z, -x * y, -x * z,
                               -y * r, -y * x, r * r + x * x + z * z, -y * z, -z * r, -z * x, -z * y, r * r + x * x + y * y};
        for (int e = 0; e < 16; e++) sJ[e] = scale * tv[e];
        const double nrm = sqrt(s);
        F.x_kk[3] = r / nrm;
        F.x_kk[4] = x / nrm;
        F.x_kk[5] = y / nrm;
        F.x_kk[6] = z / nrm;
    }
    if (threadIdx.x < 16) sB[threadIdx.x] = P[(3 + threadIdx.x / 4) + (size_t)(3 + threadIdx.x % 4) * ld];
    __syncthreads();
    if (threadIdx.x < 16) {  // (Jn * B)
        const int i = threadIdx.x / 4, j = threadIdx.x % 4;
        double s = 0;
        for (int l = 0; l < 4; l++) s += sJ[i * 4 + l] * sB[l * 4 + j];
        sT[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x < 16) {  // (Jn * B) * Jn^T
        const int i = threadIdx.x / 4, j = threadIdx.x % 4;
        double s = 0;
        for (int l = 0; l < 4; l++) s += sT[i * 4 + l] * sJ[j * 4 + l];
        if (i >= j) {  // lower triangle is authoritative; mirror so that P stays exactly symmetric
            P[(3 + i) + (size_t)(3 + j) * ld] = s;
            P[(3 + j) + (size_t)(3 + i) * ld] = s;
        }
    }
    for (int c = threadIdx.x; c < F.n; c += blockDim.x) {
        if (c >= 3 && c < 7) continue;
        double v[4], o[4];
#pragma unroll
        for (int l = 0; l < 4; l++) v[l] = P[(3 + l) + (size_t)c * ld];
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = sJ[i * 4] * v[0] + sJ[i * 4 + 1] * v[1] + sJ[i * 4 + 2] * v[2] + sJ[i * 4 + 3] * v[3];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            P[(3 + i) + (size_t)c * ld] = o[i];
            P[c + (size_t)(3 + i) * ld] = o[i];
        }
    }
}

}  // namespace rslam
