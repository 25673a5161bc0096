// Stand-in for the few OpenCV declarations visgeom's include/ocv.h and src/calibration/corner_detector.cpp name.
// TEST INFRASTRUCTURE ONLY (see oracle/shim/Eigen/Eigen): lets `make -C oracle ref` compile the reference's checkerboard
// detector where it lies under /root/reference.  Functional: Mat_<T> as a dense row-major array, cv::GaussianBlur for
// 8-bit single-channel images (OpenCV's bit-exact fixed-point path, restated in oracle/corner_oracle.c and pinned bit
// for bit on cv2 fixtures: tests/golden/corner_response.npz).  Display / file / drawing calls (imshow, imwrite, line,
// waitKey, resize: reached with DEBUG only) do nothing.
#ifndef VISGEOM_ORACLE_OPENCV_SHIM
#define VISGEOM_ORACLE_OPENCV_SHIM
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

extern "C" void vgo_gaussian_blur_u8(const uint8_t *src, int width, int height, int n, double sigma, uint8_t *dst);

namespace cv {
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Point { int x, y; Point() : x(0), y(0) {} Point(double x_, double y_) : x((int)x_), y((int)y_) {} };
struct Point2d { double x, y; Point2d() : x(0), y(0) {} Point2d(double x_, double y_) : x(x_), y(y_) {} };
struct Rect { int x, y, width, height; };
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; } };
struct Vec3b { unsigned char v[3]; };
enum { INTER_NEAREST = 0 };

class Mat {
public:
    int rows, cols;
    unsigned char *data;
    Mat() : rows(0), cols(0), data(nullptr), esz_(1) {}
    bool empty() const { return rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    int channels() const { return 1; }
protected:
    std::shared_ptr<std::vector<unsigned char> > store_;
    size_t esz_;
    void alloc(int r, int c, size_t esz)
    {
        if (r == rows && c == cols && esz == esz_ && store_) return;        // cv::Mat::create keeps a fitting buffer
        store_.reset(new std::vector<unsigned char>((size_t)r * c * esz));
        rows = r; cols = c; esz_ = esz; data = store_->data();
    }
};
typedef Mat MatND;

template <typename T> class Mat_ : public Mat {
public:
    Mat_() { esz_ = sizeof(T); }
    Mat_(int r, int c) { esz_ = sizeof(T); alloc(r, c, sizeof(T)); }
    explicit Mat_(Size s) { esz_ = sizeof(T); alloc(s.height, s.width, sizeof(T)); }
    void create(Size s) { alloc(s.height, s.width, sizeof(T)); }
    void create(int r, int c) { alloc(r, c, sizeof(T)); }
    T &operator()(int v, int u) { return reinterpret_cast<T *>(data)[(size_t)v * cols + u]; }
    const T &operator()(int v, int u) const { return reinterpret_cast<const T *>(data)[(size_t)v * cols + u]; }
    Mat_ &setTo(double s) { T *p = begin(); for (size_t i = 0, n = (size_t)rows * cols; i < n; i++) p[i] = (T)s; return *this; }
    void copyTo(Mat_ &o) const { o.create(rows, cols); if (data) std::memcpy(o.data, data, (size_t)rows * cols * sizeof(T)); }
    T *begin() { return reinterpret_cast<T *>(data); }
    T *end() { return begin() + (size_t)rows * cols; }
    // arithmetic that only the DEBUG branches reach: returns a copy
    Mat_ operator/(double) const { return *this; }
    Mat_ operator-(double) const { return *this; }
    friend Mat_ operator-(double, const Mat_ &m) { return m; }
    friend Mat_ operator*(double, const Mat_ &m) { return m; }
};

// cv::GaussianBlur(src, dst, Size(n, n), sigma, sigma): 8-bit images take OpenCV's bit-exact fixed-point path
inline void GaussianBlur(const Mat_<uint8_t> &src, Mat_<uint8_t> &dst, Size k, double sx, double)
{
    dst.create(src.rows, src.cols);
    vgo_gaussian_blur_u8(src.data, src.cols, src.rows, k.width, sx, dst.data);
}
inline void GaussianBlur(const Mat_<float> &src, Mat_<float> &dst, Size, double, double) { src.copyTo(dst); }   // DEBUG only

inline void imshow(const std::string &, const Mat &) {}
inline bool imwrite(const std::string &, const Mat &) { return true; }
inline Mat imread(const std::string &, int = 1) { return Mat(); }
inline int waitKey(int = 0) { return 0; }
inline void line(Mat &, Point, Point, const Scalar &, int = 1, int = 8, int = 0) {}
inline void circle(Mat &, Point, int, const Scalar &, int = 1, int = 8, int = 0) {}
template <typename A, typename B> inline void resize(const A &src, B &dst, Size, double = 0, double = 0, int = 0) { src.copyTo(dst); }
template <typename T> inline void swap(T &a, T &b) { T t = a; a = b; b = t; }
inline void setMouseCallback(const std::string &, void (*)(int, int, int, int, void *), void * = nullptr) {}
class BRISK;
class KeyPoint;
class DMatch;
class BFMatcher;
}  // namespace cv
#endif
