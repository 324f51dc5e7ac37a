// linalg_shim.h -- the few Eigen / OpenCV types that appear in the reference's class interfaces.
// With real Eigen / OpenCV installed the genuine headers are used; this image has neither, so minimal stand-ins with the same
// names, storage order (column-major matrices, row-major 8-bit cv::Mat) and accessors are provided: construction, resize,
// rows / cols / size, (i) and (i, j) element access, data().  The host classes use nothing else, so the same source builds against
// either.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
#else
namespace Eigen {
const int Dynamic = -1;
template <class T, int R, int C>
class Matrix {  // column-major, like Eigen's default; R / C == Dynamic: set at run time
  public:
    Matrix() : r_(R > 0 ? R : 0), c_(C > 0 ? C : 0), v_((size_t)r_ * c_, T(0)) {}
    explicit Matrix(int n) : r_(C == 1 ? n : (R > 0 ? R : 1)), c_(C == 1 ? 1 : n), v_((size_t)r_ * c_, T(0)) {}
    Matrix(int r, int c) : r_(r), c_(c), v_((size_t)r * c, T(0)) {}
    void resize(int n) {
        if (C == 1) {
            r_ = n;
            c_ = 1;
        } else {
            r_ = R > 0 ? R : 1;
            c_ = n;
        }
        v_.assign((size_t)r_ * c_, T(0));
    }
    void resize(int r, int c) {
        r_ = r;
        c_ = c;
        v_.assign((size_t)r * c, T(0));
    }
    int rows() const { return r_; }
    int cols() const { return c_; }
    int size() const { return r_ * c_; }
    T& operator()(int i) { return v_[i]; }
    T operator()(int i) const { return v_[i]; }
    T& operator[](int i) { return v_[i]; }
    T operator[](int i) const { return v_[i]; }
    T& operator()(int i, int j) { return v_[(size_t)i + (size_t)j * r_]; }
    T operator()(int i, int j) const { return v_[(size_t)i + (size_t)j * r_]; }
    T* data() { return v_.data(); }
    const T* data() const { return v_.data(); }

  private:
    int r_, c_;
    std::vector<T> v_;
};
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<double, 1, Dynamic> RowVectorXd;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 1, 2> RowVector2d;
typedef Matrix<double, 1, 4> RowVector4d;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
}  // namespace Eigen
#endif

#if __has_include(<opencv2/core/core.hpp>)
#include <opencv2/core/core.hpp>
#else
namespace cv {
class Mat {  // 8-bit single-channel, row-major, borrowed or owned
  public:
    Mat() {}
    Mat(int r, int c, const uint8_t* borrowed, size_t stride) : rows(r), cols(c), data(const_cast<uint8_t*>(borrowed)), step(stride) {}
    Mat(int r, int c) : rows(r), cols(c), own_((size_t)r * c, 0) {
        data = own_.data();
        step = (size_t)c;
    }
    int channels() const { return 1; }
    int rows = 0, cols = 0;
    uint8_t* data = nullptr;
    size_t step = 0;

  private:
    std::vector<uint8_t> own_;
};
}  // namespace cv
#endif
