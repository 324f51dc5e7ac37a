// mini_cv.h -- stand-in for the handful of OpenCV (3.2, pinned by the reference's CMakeLists.txt:28-29) entry points that the
// reference's hot-path sources call.  TEST INFRASTRUCTURE (oracle/): see mini_eigen.h for why it exists.  Restated from OpenCV's
// documented behaviour; the three numerically relevant primitives are pinned against cv2 4.13 outputs by the fixtures in
// tests/golden/ (cv_fixtures.npz, fast_fixtures.npz):
//   cv::remap(CV_32F, INTER_LINEAR, BORDER_CONSTANT 0): map coordinates quantised to 1/32 px, float weights, float accumulation;
//   cv::calcCovarMatrix(CV_32F samples, CV_COVAR_NORMAL | CV_COVAR_ROWS) -> CV_64F scatter matrix about the double mean, unscaled;
//   cv::FAST(img, kps, threshold, nonmax = true), TYPE_9_16, score = largest threshold for which the pixel is still a corner.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

typedef unsigned char uchar;
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_8U
#define CV_32FC1 CV_32F
#define CV_64FC1 CV_64F
#define CV_COVAR_SCRAMBLED 0
#define CV_COVAR_NORMAL 1
#define CV_COVAR_USE_AVG 2
#define CV_COVAR_SCALE 4
#define CV_COVAR_ROWS 8
#define CV_COVAR_COLS 16
#define CV_BGR2GRAY 6
#define CV_GRAY2BGR 8

namespace cv {
[[noreturn]] inline void shim_fail(const char* what) {
    std::cerr << "mini_cv: " << what << std::endl;
    std::abort();
}
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { BORDER_CONSTANT = 0 };

struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}  // the reference passes doubles: truncation toward zero happens at the call (Q13)
    int size() const { return end - start; }
};
template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
};
typedef Point_<int> Point;
typedef Point_<float> Point2f;
struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {}
    static Scalar all(double v) { return Scalar(v, v, v, v); }
};
struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0) : pt(x, y), size(s), angle(a), response(r), octave(0), class_id(-1) {}
};

inline size_t elem_size(int type) {
    switch (type) {
        case CV_8U: case CV_8S: return 1;
        case CV_16U: case CV_16S: return 2;
        case CV_32S: case CV_32F: return 4;
        case CV_64F: return 8;
    }
    shim_fail("unsupported matrix type");
}
template <class T> struct type_of;
template <> struct type_of<uchar> { enum { value = CV_8U }; };
template <> struct type_of<int> { enum { value = CV_32S }; };
template <> struct type_of<float> { enum { value = CV_32F }; };
template <> struct type_of<double> { enum { value = CV_64F }; };

// single-channel 2-D matrix header over a shared buffer (copying the header shares the pixels, like cv::Mat)
class Mat {
  public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    size_t step = 0;  // bytes per row
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* borrowed, size_t step_bytes) : rows(r), cols(c), data((uchar*)borrowed), step(step_bytes), type_(type) {}
    template <class T> explicit Mat(const std::vector<T>& v) {  // n x 1 column, copied (the reference never outlives the vector anyway)
        create((int)v.size(), 1, type_of<T>::value);
        for (size_t i = 0; i < v.size(); i++) at<T>((int)i, 0) = v[i];
    }
    void create(int r, int c, int type) {
        rows = r;
        cols = c;
        type_ = type;
        step = (size_t)c * elem_size(type);
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * step + 16, (uchar)0);
        data = buf_->data();
    }
    int type() const { return type_; }
    int depth() const { return type_; }
    int channels() const { return 1; }
    bool empty() const { return rows == 0 || cols == 0 || data == nullptr; }
    size_t elemSize() const { return elem_size(type_); }
    // like OpenCV's release build, at<T>() does not check T against type(): the reference's toCvMat_i writes ints into a CV_32F header
    template <class T> T& at(int i, int j) { return *(T*)(data + (size_t)i * step + (size_t)j * sizeof(T)); }
    template <class T> const T& at(int i, int j) const { return *(const T*)(data + (size_t)i * step + (size_t)j * sizeof(T)); }
    Mat operator()(const Range& rr, const Range& cr) const {
        if (rr.start < 0 || cr.start < 0 || rr.end > rows || cr.end > cols || rr.end < rr.start || cr.end < cr.start) shim_fail("ROI out of range");
        Mat m(*this);
        m.rows = rr.size();
        m.cols = cr.size();
        m.data = data + (size_t)rr.start * step + (size_t)cr.start * elem_size(type_);
        return m;
    }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int i = 0; i < rows; i++) std::memcpy(m.data + (size_t)i * m.step, data + (size_t)i * step, (size_t)cols * elem_size(type_));
        return m;
    }
    Mat t() const {
        Mat m(cols, rows, type_);
        size_t es = elem_size(type_);
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++) std::memcpy(m.data + (size_t)j * m.step + (size_t)i * es, data + (size_t)i * step + (size_t)j * es, es);
        return m;
    }

  private:
    int type_ = CV_8U;
    std::shared_ptr<std::vector<uchar>> buf_;
};
inline Mat operator/(const Mat& a, double s) {
    Mat m = a.clone();
    for (int i = 0; i < m.rows; i++)
        for (int j = 0; j < m.cols; j++) switch (m.type()) {
                case CV_64F: m.at<double>(i, j) = m.at<double>(i, j) / s; break;
                case CV_32F: m.at<float>(i, j) = (float)(m.at<float>(i, j) / s); break;
                default: shim_fail("Mat / scalar: unsupported type");
            }
    return m;
}
inline void repeat(const Mat& src, int ny, int nx, Mat& dst) {
    Mat out(src.rows * ny, src.cols * nx, src.type());
    size_t es = src.elemSize();
    for (int i = 0; i < out.rows; i++)
        for (int j = 0; j < out.cols; j++)
            std::memcpy(out.data + (size_t)i * out.step + (size_t)j * es, src.data + (size_t)(i % src.rows) * src.step + (size_t)(j % src.cols) * es, es);
    dst = out;
}
inline void cvtColor(const Mat&, Mat&, int) { shim_fail("cvtColor: colour images are outside the replayed path"); }

// ---- cv::remap, CV_32F source and maps, INTER_LINEAR, BORDER_CONSTANT -----------------------------------------------------------
inline void remap(const Mat& src, Mat& dst, const Mat& map1, const Mat& map2, int interpolation, int borderMode = BORDER_CONSTANT,
                  const Scalar& borderValue = Scalar()) {
    if (interpolation != INTER_LINEAR || borderMode != BORDER_CONSTANT) shim_fail("remap: only INTER_LINEAR + BORDER_CONSTANT");
    if (src.type() != CV_32F || map1.type() != CV_32F || map2.type() != CV_32F) shim_fail("remap: CV_32F only");
    const float bv = (float)borderValue.val[0];
    Mat out(map1.rows, map1.cols, CV_32F);
    for (int i = 0; i < out.rows; i++)
        for (int j = 0; j < out.cols; j++) {
            const float mx = map1.at<float>(i, j), my = map2.at<float>(i, j);
            // fixed-point map with 5 fractional bits, round-half-to-even (cvRound)
            const int sx = (int)std::nearbyint((double)(mx * 32.0f)), sy = (int)std::nearbyint((double)(my * 32.0f));
            const int ix = sx >> 5, iy = sy >> 5;
            const float fx = (float)(sx & 31) / 32.0f, fy = (float)(sy & 31) / 32.0f;
            const float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx, w10 = fy * (1.f - fx), w11 = fy * fx;
            auto tap = [&](int y, int x) -> float { return (x < 0 || y < 0 || x >= src.cols || y >= src.rows) ? bv : src.at<float>(y, x); };
            out.at<float>(i, j) = tap(iy, ix) * w00 + tap(iy, ix + 1) * w01 + tap(iy + 1, ix) * w10 + tap(iy + 1, ix + 1) * w11;
        }
    dst = out;
}

// ---- cv::calcCovarMatrix(samples, covar, mean, CV_COVAR_NORMAL | CV_COVAR_ROWS, ctype = CV_64F) -----------------------------------
inline void calcCovarMatrix(const Mat& samples, Mat& covar, Mat& mean, int flags, int ctype = CV_64F) {
    if (!(flags & CV_COVAR_NORMAL) || !(flags & CV_COVAR_ROWS) || (flags & (CV_COVAR_USE_AVG | CV_COVAR_SCALE)) || ctype != CV_64F)
        shim_fail("calcCovarMatrix: only CV_COVAR_NORMAL | CV_COVAR_ROWS with a CV_64F result");
    if (samples.type() != CV_32F) shim_fail("calcCovarMatrix: CV_32F samples only");
    const int ns = samples.rows, nv = samples.cols;
    Mat mu(1, nv, CV_64F), cov(nv, nv, CV_64F);
    for (int j = 0; j < nv; j++) {
        double s = 0;
        for (int p = 0; p < ns; p++) s += (double)samples.at<float>(p, j);
        mu.at<double>(0, j) = s * (1.0 / ns);
    }
    std::vector<double> D((size_t)ns * nv);
    for (int p = 0; p < ns; p++)
        for (int j = 0; j < nv; j++) D[(size_t)j * ns + p] = (double)samples.at<float>(p, j) - mu.at<double>(0, j);
    for (int a = 0; a < nv; a++)
        for (int b = a; b < nv; b++) {
            double s = 0;
            const double *da = &D[(size_t)a * ns], *db = &D[(size_t)b * ns];
            for (int p = 0; p < ns; p++) s += da[p] * db[p];
            cov.at<double>(a, b) = s;
            cov.at<double>(b, a) = s;
        }
    covar = cov;
    mean = mu;
}

// ---- cv::FAST, TYPE_9_16 ------------------------------------------------------------------------------------------------------
inline int fast9_score(const Mat& img, int x, int y) {
    static const int ring[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3}, {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};
    const int v = img.at<uchar>(y, x);
    int d[25];
    for (int k = 0; k < 25; k++) d[k] = v - (int)img.at<uchar>(y + ring[k % 16][1], x + ring[k % 16][0]);
    int best = -256;
    for (int s = 0; s < 16; s++) {
        int lo = 256, hi = 256;
        for (int k = 0; k < 9; k++) {
            lo = std::min(lo, d[s + k]);
            hi = std::min(hi, -d[s + k]);
        }
        best = std::max(best, std::max(lo, hi));
    }
    return best;  // the pixel is a corner for every threshold t < best
}
inline void FAST(const Mat& img, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true) {
    keypoints.clear();
    if (img.type() != CV_8U) shim_fail("FAST: 8-bit images only");
    threshold = std::min(std::max(threshold, 0), 255);
    const int R = img.rows, C = img.cols;
    if (R < 7 || C < 7) return;
    std::vector<int> score((size_t)R * C, 0);
    for (int y = 3; y < R - 3; y++)
        for (int x = 3; x < C - 3; x++) {
            const int s = fast9_score(img, x, y);
            if (s > threshold) score[(size_t)y * C + x] = nonmaxSuppression ? s - 1 : 1;
        }
    for (int y = 3; y < R - 3; y++)
        for (int x = 3; x < C - 3; x++) {
            const int s = score[(size_t)y * C + x];
            if (!s) continue;
            bool keep = true;
            if (nonmaxSuppression)
                for (int dy = -1; dy <= 1 && keep; dy++)
                    for (int dx = -1; dx <= 1; dx++)
                        if ((dx || dy) && !(s > score[(size_t)(y + dy) * C + (x + dx)])) {
                            keep = false;
                            break;
                        }
            if (keep) keypoints.push_back(KeyPoint((float)x, (float)y, 7.f, -1.f, (float)s));
        }
}

// ---- cv::FileStorage on the reference's flat "key: value" YAML -------------------------------------------------------------------
class FileNode {
  public:
    FileNode() {}
    explicit FileNode(const std::string& s) : s_(s) {}
    operator double() const { return s_.empty() ? 0.0 : std::atof(s_.c_str()); }
    operator float() const { return (float)(double)(*this); }
    operator int() const { return (int)std::nearbyint((double)(*this)); }
    operator std::string() const { return s_; }
    bool empty() const { return s_.empty(); }

  private:
    std::string s_;
};
class FileStorage {
  public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string& path, int flags) { open(path, flags); }
    bool open(const std::string& path, int) {
        kv_.clear();
        opened_ = false;
        std::ifstream f(path.c_str());
        if (!f) return false;
        std::string line;
        while (std::getline(f, line)) {
            size_t hash = line.find('#');
            if (hash != std::string::npos) line = line.substr(0, hash);
            if (line.empty() || line[0] == '%') continue;
            size_t colon = line.find(':');
            if (colon == std::string::npos) continue;
            auto trim = [](std::string s) {
                size_t a = s.find_first_not_of(" \t\r\n\""), b = s.find_last_not_of(" \t\r\n\"");
                return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
            };
            kv_[trim(line.substr(0, colon))] = trim(line.substr(colon + 1));
        }
        opened_ = true;
        return true;
    }
    bool isOpened() const { return opened_; }
    void release() {
        kv_.clear();
        opened_ = false;
    }
    FileNode operator[](const char* key) const {
        auto it = kv_.find(key);
        return it == kv_.end() ? FileNode() : FileNode(it->second);
    }
    FileNode operator[](const std::string& key) const { return (*this)[key.c_str()]; }

  private:
    std::map<std::string, std::string> kv_;
    bool opened_ = false;
};
// drawing primitives used only by the reference's image overlay (src/System.cpp:199-227): accepted, nothing is drawn
struct Size {
    int width = 0, height = 0;
    Size() {}
    Size(double w, double h) : width((int)w), height((int)h) {}
};
inline void line(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) {}
inline void ellipse(Mat&, Point, Size, double, double, double, const Scalar&, int = 1, int = 8, int = 0) {}
}  // namespace cv
typedef cv::Scalar CvScalar;
inline cv::Scalar cvScalarAll(double v) { return cv::Scalar::all(v); }
