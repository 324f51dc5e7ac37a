// oracle/omat.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Minimal owned dense-matrix type used by the CPU oracle.  It restates the few
// Eigen semantics the reference relies on (Eigen is not available in this image):
//   * column-major storage (Eigen default; reference uses MatrixXd everywhere)
//   * left-to-right evaluation of product chains  (A*B*C == (A*B)*C)
//   * dynamic-size inverse() == PartialPivLU (row pivoting, first max wins)
//   * fixed-size 2x2 / 3x3 / 4x4 inverse() == closed-form cofactor expansion
//
// Nothing in the shipped product (ransac_slam_b200/) includes this header.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>

namespace orc {

void dgemm(bool tA, bool tB, int M, int N, int K, const double* A, int lda, const double* B, int ldb, double* C,
           int ldc);  // C = op(A)*op(B), see gemm.cpp
void set_threads(int n);
int get_threads();

struct Mat {
    int r = 0, c = 0;
    std::vector<double> a;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
    double& operator()(int i, int j) { return a[(size_t)i + (size_t)j * r]; }
    double operator()(int i, int j) const { return a[(size_t)i + (size_t)j * r]; }
    double& operator[](int i) { return a[i]; }
    double operator[](int i) const { return a[i]; }
    int rows() const { return r; }
    int cols() const { return c; }
    size_t size() const { return a.size(); }
    double* data() { return a.data(); }
    const double* data() const { return a.data(); }
    void resize(int r_, int c_) {
        r = r_;
        c = c_;
        a.assign((size_t)r_ * c_, 0.0);
    }
    static Mat Zero(int r_, int c_) { return Mat(r_, c_); }
    static Mat Identity(int n) {
        Mat m(n, n);
        for (int i = 0; i < n; i++) m(i, i) = 1.0;
        return m;
    }
    Mat block(int i0, int j0, int nr, int nc) const {
        Mat m(nr, nc);
        for (int j = 0; j < nc; j++)
            std::memcpy(&m.a[(size_t)j * nr], &a[(size_t)i0 + (size_t)(j0 + j) * r], sizeof(double) * nr);
        return m;
    }
    void set_block(int i0, int j0, const Mat& m) {
        for (int j = 0; j < m.c; j++)
            std::memcpy(&a[(size_t)i0 + (size_t)(j0 + j) * r], &m.a[(size_t)j * m.r], sizeof(double) * m.r);
    }
    Mat t() const {
        Mat m(c, r);
        for (int j = 0; j < c; j++)
            for (int i = 0; i < r; i++) m(j, i) = (*this)(i, j);
        return m;
    }
};

inline Mat operator*(const Mat& A, const Mat& B) {
    assert(A.c == B.r);
    Mat C(A.r, B.c);
    if (A.r && B.c && A.c) dgemm(false, false, A.r, B.c, A.c, A.data(), A.r, B.data(), B.r, C.data(), C.r);
    return C;
}
// A * B^T without materialising the transpose (Eigen does the same through its product kernel)
inline Mat mul_nt(const Mat& A, const Mat& B) {
    assert(A.c == B.c);
    Mat C(A.r, B.r);
    if (A.r && B.r && A.c) dgemm(false, true, A.r, B.r, A.c, A.data(), A.r, B.data(), B.r, C.data(), C.r);
    return C;
}
inline Mat operator+(const Mat& A, const Mat& B) {
    assert(A.r == B.r && A.c == B.c);
    Mat C(A.r, A.c);
    for (size_t i = 0; i < A.a.size(); i++) C.a[i] = A.a[i] + B.a[i];
    return C;
}
inline Mat operator-(const Mat& A, const Mat& B) {
    assert(A.r == B.r && A.c == B.c);
    Mat C(A.r, A.c);
    for (size_t i = 0; i < A.a.size(); i++) C.a[i] = A.a[i] - B.a[i];
    return C;
}
inline Mat operator*(const Mat& A, double s) {
    Mat C(A.r, A.c);
    for (size_t i = 0; i < A.a.size(); i++) C.a[i] = A.a[i] * s;
    return C;
}
inline Mat operator*(double s, const Mat& A) {
    Mat C(A.r, A.c);
    for (size_t i = 0; i < A.a.size(); i++) C.a[i] = s * A.a[i];
    return C;
}
inline Mat neg(const Mat& A) {
    Mat C(A.r, A.c);
    for (size_t i = 0; i < A.a.size(); i++) C.a[i] = -A.a[i];
    return C;
}

// Eigen fixed-size inverse, 2x2: inverse = [d -b; -c a] * (1/det)   (Eigen/src/LU/InverseImpl.h, compute_inverse_size2_helper)
inline Mat inv2_fixed(const Mat& m) {
    assert(m.r == 2 && m.c == 2);
    double det = m(0, 0) * m(1, 1) - m(1, 0) * m(0, 1);
    double invdet = 1.0 / det;
    Mat o(2, 2);
    o(0, 0) = m(1, 1) * invdet;
    o(1, 0) = -m(1, 0) * invdet;
    o(0, 1) = -m(0, 1) * invdet;
    o(1, 1) = m(0, 0) * invdet;
    return o;
}
// Eigen fixed-size inverse, 3x3: cofactors, det from first column of cofactors . first row (compute_inverse_size3_helper)
inline double cof3(const Mat& m, int i, int j) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
}
inline Mat inv3_fixed(const Mat& m) {
    assert(m.r == 3 && m.c == 3);
    double c00 = cof3(m, 0, 0), c10 = cof3(m, 1, 0), c20 = cof3(m, 2, 0);
    double det = c00 * m(0, 0) + c10 * m(1, 0) + c20 * m(2, 0);
    double invdet = 1.0 / det;
    Mat o(3, 3);
    // inverse(i,j) = cofactor(j,i) / det
    o(0, 0) = c00 * invdet;
    o(0, 1) = c10 * invdet;
    o(0, 2) = c20 * invdet;
    o(1, 0) = cof3(m, 0, 1) * invdet;
    o(1, 1) = cof3(m, 1, 1) * invdet;
    o(1, 2) = cof3(m, 2, 1) * invdet;
    o(2, 0) = cof3(m, 0, 2) * invdet;
    o(2, 1) = cof3(m, 1, 2) * invdet;
    o(2, 2) = cof3(m, 2, 2) * invdet;
    return o;
}
// 4x4 closed form (cofactor expansion by 2x2 sub-determinants; Eigen's size-4 helper uses the same
// adjugate/determinant formulation, SSE-vectorised for double)
inline Mat inv4_fixed(const Mat& m) {
    assert(m.r == 4 && m.c == 4);
    const double a00 = m(0, 0), a01 = m(0, 1), a02 = m(0, 2), a03 = m(0, 3);
    const double a10 = m(1, 0), a11 = m(1, 1), a12 = m(1, 2), a13 = m(1, 3);
    const double a20 = m(2, 0), a21 = m(2, 1), a22 = m(2, 2), a23 = m(2, 3);
    const double a30 = m(3, 0), a31 = m(3, 1), a32 = m(3, 2), a33 = m(3, 3);
    double s0 = a00 * a11 - a10 * a01, s1 = a00 * a12 - a10 * a02, s2 = a00 * a13 - a10 * a03;
    double s3 = a01 * a12 - a11 * a02, s4 = a01 * a13 - a11 * a03, s5 = a02 * a13 - a12 * a03;
    double c5 = a22 * a33 - a32 * a23, c4 = a21 * a33 - a31 * a23, c3 = a21 * a32 - a31 * a22;
    double c2 = a20 * a33 - a30 * a23, c1 = a20 * a32 - a30 * a22, c0 = a20 * a31 - a30 * a21;
    double det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
    double id = 1.0 / det;
    Mat o(4, 4);
    o(0, 0) = (a11 * c5 - a12 * c4 + a13 * c3) * id;
    o(0, 1) = (-a01 * c5 + a02 * c4 - a03 * c3) * id;
    o(0, 2) = (a31 * s5 - a32 * s4 + a33 * s3) * id;
    o(0, 3) = (-a21 * s5 + a22 * s4 - a23 * s3) * id;
    o(1, 0) = (-a10 * c5 + a12 * c2 - a13 * c1) * id;
    o(1, 1) = (a00 * c5 - a02 * c2 + a03 * c1) * id;
    o(1, 2) = (-a30 * s5 + a32 * s2 - a33 * s1) * id;
    o(1, 3) = (a20 * s5 - a22 * s2 + a23 * s1) * id;
    o(2, 0) = (a10 * c4 - a11 * c2 + a13 * c0) * id;
    o(2, 1) = (-a00 * c4 + a01 * c2 - a03 * c0) * id;
    o(2, 2) = (a30 * s4 - a31 * s2 + a33 * s0) * id;
    o(2, 3) = (-a20 * s4 + a21 * s2 - a23 * s0) * id;
    o(3, 0) = (-a10 * c3 + a11 * c1 - a12 * c0) * id;
    o(3, 1) = (a00 * c3 - a01 * c1 + a02 * c0) * id;
    o(3, 2) = (-a30 * s3 + a31 * s1 - a32 * s0) * id;
    o(3, 3) = (a20 * s3 - a21 * s1 + a22 * s0) * id;
    return o;
}

// Eigen dynamic-size MatrixXd::inverse(): PartialPivLU, then solve against the identity.
// Right-looking Gaussian elimination with row pivoting (largest |a_ik|, first wins), unit-lower L.
Mat lu_inverse(const Mat& A);

}  // namespace orc
