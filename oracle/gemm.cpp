// oracle/gemm.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Dense double GEMM + PartialPivLU inverse for the CPU oracle.  The reference links Eigen3 (not present in
// this image, version unpinned: CMakeLists.txt:25) and uses its dense product kernel and MatrixXd::inverse().
// This file restates those two primitives so that the oracle is (a) numerically equivalent up to summation
// order and (b) not a straw-man when it is timed as the CPU baseline: a packed, cache-blocked kernel that GCC
// vectorises (-O3 -march=native), optionally threaded with OpenMP over row panels (the reference itself is
// single-threaded: no -fopenmp in CMakeLists.txt; thread count defaults to 1 and is set explicitly by bench.py).
#include "omat.h"
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

static int g_threads = 1;
void set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int get_threads() { return g_threads; }

namespace {
constexpr int MR = 8, NR = 6, MC = 96, KC = 256, NC = 3072;

inline double opA(bool t, const double* A, int lda, int i, int k) { return t ? A[(size_t)k + (size_t)i * lda] : A[(size_t)i + (size_t)k * lda]; }
inline double opB(bool t, const double* B, int ldb, int k, int j) { return t ? B[(size_t)j + (size_t)k * ldb] : B[(size_t)k + (size_t)j * ldb]; }

// micro kernel: acc[NR][MR] += sum_k a[k*MR+i] * b[k*NR+j]
#if defined(__AVX2__) && defined(__FMA__)
}  // namespace
}  // namespace orc
#include <immintrin.h>
namespace orc {
namespace {
// 8 x 6 register tile: 12 ymm accumulators, two loads of A and one broadcast of B per k (what a BLAS / Eigen GEBP kernel does)
inline void micro(int kc, const double* __restrict a, const double* __restrict b, double* __restrict C, int ldc, int mr, int nr,
                  bool first) {
    __m256d acc[NR][2];
    for (int j = 0; j < NR; j++) acc[j][0] = acc[j][1] = _mm256_setzero_pd();
    for (int k = 0; k < kc; k++) {
        const __m256d a0 = _mm256_loadu_pd(a + (size_t)k * MR), a1 = _mm256_loadu_pd(a + (size_t)k * MR + 4);
        const double* bk = b + (size_t)k * NR;
#pragma GCC unroll 6
        for (int j = 0; j < NR; j++) {
            const __m256d bj = _mm256_broadcast_sd(bk + j);
            acc[j][0] = _mm256_fmadd_pd(a0, bj, acc[j][0]);
            acc[j][1] = _mm256_fmadd_pd(a1, bj, acc[j][1]);
        }
    }
    double tmp[NR][MR];
    for (int j = 0; j < NR; j++) {
        _mm256_storeu_pd(&tmp[j][0], acc[j][0]);
        _mm256_storeu_pd(&tmp[j][4], acc[j][1]);
    }
    for (int j = 0; j < nr; j++)
        for (int i = 0; i < mr; i++) {
            double* c = &C[(size_t)i + (size_t)j * ldc];
            *c = first ? tmp[j][i] : *c + tmp[j][i];
        }
}
#else
inline void micro(int kc, const double* __restrict a, const double* __restrict b, double* __restrict C, int ldc, int mr, int nr,
                  bool first) {
    double acc[NR][MR];
    for (int j = 0; j < NR; j++)
        for (int i = 0; i < MR; i++) acc[j][i] = 0.0;
    for (int k = 0; k < kc; k++) {
        const double* ak = a + (size_t)k * MR;
        const double* bk = b + (size_t)k * NR;
        for (int j = 0; j < NR; j++) {
            double bj = bk[j];
            for (int i = 0; i < MR; i++) acc[j][i] += ak[i] * bj;
        }
    }
    for (int j = 0; j < nr; j++)
        for (int i = 0; i < mr; i++) {
            double* c = &C[(size_t)i + (size_t)j * ldc];
            *c = first ? acc[j][i] : *c + acc[j][i];
        }
}
#endif

void gemm_small_rows(bool tA, bool tB, int M, int N, int K, const double* A, int lda, const double* B, int ldb, double* C,
                     int ldc) {
    // few rows (H_i * P, (H_i P) * H_i^T ...): one pass over B
    for (int j = 0; j < N; j++) {
        for (int i = 0; i < M; i++) {
            double s = 0.0;
            if (!tB && !tA) {
                const double* b = &B[(size_t)j * ldb];
                const double* a = &A[i];
                for (int k = 0; k < K; k++) s += a[(size_t)k * lda] * b[k];
            } else {
                for (int k = 0; k < K; k++) s += opA(tA, A, lda, i, k) * opB(tB, B, ldb, k, j);
            }
            C[(size_t)i + (size_t)j * ldc] = s;
        }
    }
}

void gemm_small_cols(bool tA, bool tB, int M, int N, int K, const double* A, int lda, const double* B, int ldb, double* C,
                     int ldc) {
    // few columns (P * H_i^T ...): axpy over the columns of A
    for (int j = 0; j < N; j++) {
        double* c = &C[(size_t)j * ldc];
        for (int i = 0; i < M; i++) c[i] = 0.0;
        if (!tA) {
            for (int k = 0; k < K; k++) {
                double bkj = opB(tB, B, ldb, k, j);
                const double* a = &A[(size_t)k * lda];
                for (int i = 0; i < M; i++) c[i] += a[i] * bkj;
            }
        } else {
            for (int i = 0; i < M; i++) {
                double s = 0.0;
                const double* a = &A[(size_t)i * lda];
                for (int k = 0; k < K; k++) s += a[k] * opB(tB, B, ldb, k, j);
                c[i] = s;
            }
        }
    }
}
}  // namespace

void dgemm(bool tA, bool tB, int M, int N, int K, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
    if (M <= 0 || N <= 0) return;
    if (K <= 0) {
        for (int j = 0; j < N; j++)
            for (int i = 0; i < M; i++) C[(size_t)i + (size_t)j * ldc] = 0.0;
        return;
    }
    if (M <= 4) return gemm_small_rows(tA, tB, M, N, K, A, lda, B, ldb, C, ldc);
    if (N <= 4) return gemm_small_cols(tA, tB, M, N, K, A, lda, B, ldb, C, ldc);

    std::vector<double> Bp((size_t)KC * (((NC + NR - 1) / NR) * NR));
    for (int jc = 0; jc < N; jc += NC) {
        int nc = std::min(NC, N - jc);
        for (int pc = 0; pc < K; pc += KC) {
            int kc = std::min(KC, K - pc);
            bool first = (pc == 0);
            // pack B block (kc x nc) into NR-wide column panels
            int npan = (nc + NR - 1) / NR;
#pragma omp parallel for num_threads(g_threads) schedule(static) if (g_threads > 1 && (size_t)kc * nc > 65536)
            for (int p = 0; p < npan; p++) {
                double* dst = &Bp[(size_t)p * kc * NR];
                for (int k = 0; k < kc; k++)
                    for (int j = 0; j < NR; j++) {
                        int jj = p * NR + j;
                        dst[(size_t)k * NR + j] = jj < nc ? opB(tB, B, ldb, pc + k, jc + jj) : 0.0;
                    }
            }
            int nblk = (M + MC - 1) / MC;
#pragma omp parallel num_threads(g_threads) if (g_threads > 1 && (size_t)M * nc * kc > 262144)
            {
                std::vector<double> Ap((size_t)MC * KC);
#pragma omp for schedule(dynamic, 1)
                for (int ib = 0; ib < nblk; ib++) {
                    int ic = ib * MC;
                    int mc = std::min(MC, M - ic);
                    int mpan = (mc + MR - 1) / MR;
                    for (int p = 0; p < mpan; p++) {
                        double* dst = &Ap[(size_t)p * kc * MR];
                        for (int k = 0; k < kc; k++)
                            for (int i = 0; i < MR; i++) {
                                int ii = p * MR + i;
                                dst[(size_t)k * MR + i] = ii < mc ? opA(tA, A, lda, ic + ii, pc + k) : 0.0;
                            }
                    }
                    for (int jp = 0; jp < npan; jp++) {
                        int nr = std::min(NR, nc - jp * NR);
                        for (int ip = 0; ip < mpan; ip++) {
                            int mr = std::min(MR, mc - ip * MR);
                            micro(kc, &Ap[(size_t)ip * kc * MR], &Bp[(size_t)jp * kc * NR],
                                  &C[(size_t)(ic + ip * MR) + (size_t)(jc + jp * NR) * ldc], ldc, mr, nr, first);
                        }
                    }
                }
            }
        }
    }
}

// PartialPivLU + inverse (Eigen: MatrixXd::inverse() for dynamic sizes; used at src/Tracking.cpp:322,421,591 and
// src/ExtendKF.cpp:603).
Mat lu_inverse(const Mat& A) {
    assert(A.r == A.c);
    const int n = A.r;
    Mat LU = A;
    std::vector<int> perm(n);
    for (int i = 0; i < n; i++) perm[i] = i;
    for (int k = 0; k < n; k++) {
        int piv = k;
        double best = std::fabs(LU(k, k));
        for (int i = k + 1; i < n; i++) {
            double v = std::fabs(LU(i, k));
            if (v > best) {
                best = v;
                piv = i;
            }
        }
        if (piv != k) {
            for (int j = 0; j < n; j++) std::swap(LU(k, j), LU(piv, j));
            std::swap(perm[k], perm[piv]);
        }
        double d = LU(k, k);
        for (int i = k + 1; i < n; i++) LU(i, k) /= d;
        const double* lk = &LU.a[(size_t)k * n];
#pragma omp parallel for num_threads(g_threads) schedule(static) if (g_threads > 1 && (n - k) > 256)
        for (int j = k + 1; j < n; j++) {
            double* cj = &LU.a[(size_t)j * n];
            double ukj = cj[k];
            for (int i = k + 1; i < n; i++) cj[i] -= lk[i] * ukj;
        }
    }
    Mat X(n, n);
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 8) if (g_threads > 1 && n > 256)
    for (int j = 0; j < n; j++) {
        double* x = &X.a[(size_t)j * n];
        // b = P * e_j
        for (int i = 0; i < n; i++) x[i] = (perm[i] == j) ? 1.0 : 0.0;
        // forward substitution, unit lower (column oriented)
        for (int k = 0; k < n; k++) {
            double xk = x[k];
            if (xk != 0.0) {
                const double* lk = &LU.a[(size_t)k * n];
                for (int i = k + 1; i < n; i++) x[i] -= lk[i] * xk;
            }
        }
        // back substitution, upper
        for (int k = n - 1; k >= 0; k--) {
            const double* uk = &LU.a[(size_t)k * n];
            x[k] /= uk[k];
            double xk = x[k];
            for (int i = 0; i < k; i++) x[i] -= uk[i] * xk;
        }
    }
    return X;
}

}  // namespace orc
