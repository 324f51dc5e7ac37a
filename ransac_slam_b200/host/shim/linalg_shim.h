// linalg_shim.h -- the few Eigen / OpenCV types that appear in the reference's class interfaces.
// With real Eigen / OpenCV installed the genuine headers are used; this image has neither, so minimal stand-ins with the same
// names, storage order (column-major MatrixXd, row-major 8-bit cv::Mat) and accessors are provided.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
#else
namespace Eigen {
class VectorXd {
  public:
    VectorXd() {}
    explicit VectorXd(int n) : v_(n, 0.0) {}
    void resize(int n) { v_.assign(n, 0.0); }
    int rows() const { return (int)v_.size(); }
    int size() const { return (int)v_.size(); }
    double& operator()(int i) { return v_[i]; }
    double operator()(int i) const { return v_[i]; }
    double& operator[](int i) { return v_[i]; }
    double operator[](int i) const { return v_[i]; }
    double* data() { return v_.data(); }
    const double* data() const { return v_.data(); }

  private:
    std::vector<double> v_;
};
class MatrixXd {  // column-major, like Eigen's default
  public:
    MatrixXd() {}
    MatrixXd(int r, int c) : r_(r), c_(c), v_((size_t)r * c, 0.0) {}
    void resize(int r, int c) {
        r_ = r;
        c_ = c;
        v_.assign((size_t)r * c, 0.0);
    }
    int rows() const { return r_; }
    int cols() const { return c_; }
    double& operator()(int i, int j) { return v_[(size_t)i + (size_t)j * r_]; }
    double operator()(int i, int j) const { return v_[(size_t)i + (size_t)j * r_]; }
    double* data() { return v_.data(); }
    const double* data() const { return v_.data(); }

  private:
    int r_ = 0, c_ = 0;
    std::vector<double> v_;
};
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
