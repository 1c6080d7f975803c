// oracle/cvshim/opencv2/opencv.hpp -- TEST INFRASTRUCTURE ONLY.
//
// A minimal stand-in for the slice of the OpenCV C++ API that the reference's hot-path sources use
// (src/modelHandler.{hpp,cpp}, src/convertRoutine.{hpp,cpp}), so that THOSE FILES compile unmodified, where they
// lie under /root/reference, into oracle/_ref/libw2x_reference.so (recipe: oracle/Makefile).  The image has no
// OpenCV C++ headers or libraries; with this shim the reference's own control flow (JSON loading through picojson,
// thread partition, layer loop, padding, block-split arithmetic, crop and stitch) runs as written, while the
// arithmetic of the six cv:: calls it makes is restated here in plain fp32 (the real OpenCV arithmetic is covered by
// oracle/ref_cv2.py, which drives cv2's kernels with a restated control flow).  Semantics restated:
//   cv::Mat        reference-counted header over fp32 rows; copies are shallow; operator()(Range,Range) / rowRange /
//                  colRange are views; copyTo() writes INTO a destination of matching size (so ROI views receive
//                  data, src/convertRoutine.cpp:161) and reallocates otherwise
//   cv::UMat       a thin wrapper sharing the Mat's storage (getUMat / getMat)
//   cv::filter2D   correlation, anchor = centre, BORDER_REPLICATE, fp32 accumulate from delta in row-major tap order
//   cv::add / max / min / scaleAdd   element-wise fp32; double scalars are rounded to float first (OpenCV converts
//                  the scalar to the array depth: bias -> (float)bias, 0.1 -> 0.1f)
//   cv::copyMakeBorder   BORDER_REPLICATE
// For src/main.cpp (the CLI around the hot path) the image plumbing calls are adapters onto the repo's own host
// restatements, which tests/test_cli.py pins against cv2: imread / imwrite -> host/imageio.hpp; convertTo, cvtColor
// (RGB2YUV / YUV2RGB), resize (NEAREST / LINEAR / CUBIC) -> host/imgproc.hpp; split / merge are plain copies.  Those are
// only compiled in when W2X_CVSHIM_WITH_IMGPROC is defined (the reference CLI build, oracle/Makefile).
// Nothing outside tests/, oracle/ and bench.py's CPU-baseline leg may use this.
#ifndef W2X_ORACLE_CVSHIM_OPENCV_HPP_
#define W2X_ORACLE_CVSHIM_OPENCV_HPP_

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <memory>
#include <ostream>
#include <string>
#include <vector>

#ifdef W2X_CVSHIM_WITH_IMGPROC
#include "imageio.hpp"
#include "imgproc.hpp"
#endif

#define CV_8U 0
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#ifndef CV_Assert
#define CV_Assert(expr) assert(expr)   // opencv2/core/base.hpp
#endif
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)

namespace cv {

struct Size {
    int width = 0, height = 0;
    Size() {}
    Size(int w, int h) : width(w), height(h) {}
    bool operator==(const Size &o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size &o) const { return !(*this == o); }
};
struct Point {
    int x = 0, y = 0;
    Point() {}
    Point(int x_, int y_) : x(x_), y(y_) {}
};
struct Range {
    int start = 0, end = 0;
    Range() {}
    Range(int s, int e) : start(s), end(e) {}
};
enum { ACCESS_READ = 1 << 24, ACCESS_WRITE = 1 << 25, ACCESS_RW = 3 << 24 };
enum BorderTypes { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1 };

class UMat;

class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;                 // bytes between rows
    unsigned char *data = nullptr;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    static Mat zeros(int r, int c, int type) {
        Mat m(r, c, type);
        for (int y = 0; y < r; y++) std::memset(m.data + (size_t)y * m.step, 0, (size_t)c * m.elemSize());
        return m;
    }
    static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }

    void create(int r, int c, int type) {
        assert((type & 7) == CV_32F || (type & 7) == CV_8U);
        if (data && rows == r && cols == c && type_ == type) return;   // same shape: keep the storage (views included)
        type_ = type;
        const size_t bytes = (size_t)std::max(r, 0) * std::max(c, 0) * elemSize() + 16;
        owner_ = std::shared_ptr<float>(new float[(bytes + 3) / 4], std::default_delete<float[]>());
        data = reinterpret_cast<unsigned char *>(owner_.get());
        rows = r;
        cols = c;
        step = (size_t)c * elemSize();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }

    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { return (size_t)channels() * (depth() == CV_8U ? 1 : 4); }
    template <typename T> T &at(int r, int c) { return *reinterpret_cast<T *>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T &at(int r, int c) const { return *reinterpret_cast<const T *>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data + (size_t)r * step); }
    template <typename T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + (size_t)r * step); }

    Mat operator()(const Range &rr, const Range &cr) const {
        Mat v(*this);
        v.data = data + (size_t)rr.start * step + (size_t)cr.start * elemSize();
        v.rows = rr.end - rr.start;
        v.cols = cr.end - cr.start;
        return v;
    }
    Mat rowRange(int a, int b) const { return (*this)(Range(a, b), Range(0, cols)); }
    Mat rowRange(const Range &r) const { return rowRange(r.start, r.end); }
    Mat colRange(int a, int b) const { return (*this)(Range(0, rows), Range(a, b)); }
    Mat colRange(const Range &r) const { return colRange(r.start, r.end); }

    void copyTo(Mat &dst) const {
        if (dst.data == data && dst.rows == rows && dst.cols == cols && dst.step == step && dst.type_ == type_) return;
        dst.create(rows, cols, type_);
        for (int y = 0; y < rows; y++) std::memmove(dst.data + (size_t)y * dst.step, data + (size_t)y * step, (size_t)cols * elemSize());
    }
    Mat clone() const {
        Mat m;
        copyTo(m);
        return m;
    }
    inline UMat getUMat(int flags) const;
#ifdef W2X_CVSHIM_WITH_IMGPROC
    // image.convertTo(image, CV_32F, 1/255.) and image.convertTo(image, CV_8U, 255.) (src/main.cpp:75,172); dst may be *this
    void convertTo(Mat &dst, int rtype, double alpha = 1.0, double beta = 0.0) const {
        assert(beta == 0.0 && channels() == 3 && step == (size_t)cols * elemSize());
        (void)beta;
        if ((rtype & 7) == CV_32F && depth() == CV_8U) {
            Mat out(rows, cols, CV_32FC3);
            const float a = static_cast<float>(alpha);
            const unsigned char *s = data;
            float *d = out.ptr<float>();
            for (size_t i = 0; i < (size_t)rows * cols * 3; i++) d[i] = static_cast<float>(s[i]) * a;
            dst = out;
        } else if ((rtype & 7) == CV_8U && depth() == CV_32F) {
            assert(alpha == 255.0);
            w2ximg::Image3f im(cols, rows);
            std::memcpy(im.data.data(), data, im.data.size() * sizeof(float));
            std::vector<uint8_t> u8 = w2ximg::to_u8(im);
            Mat out(rows, cols, CV_8UC3);
            std::memcpy(out.data, u8.data(), u8.size());
            dst = out;
        } else {
            assert(!"convertTo: unsupported conversion in the shim");
        }
    }
#endif

private:
    std::shared_ptr<float> owner_;
    int type_ = CV_32FC1;
};

class UMat {
public:
    Mat m;
    UMat() {}
    UMat(Size s, int type) : m(s, type) {}
    UMat(Size s, int type, double value) : m(s, type) {
        const float v = static_cast<float>(value);
        for (int y = 0; y < m.rows; y++) {
            float *p = m.ptr<float>(y);
            for (int x = 0; x < m.cols; x++) p[x] = v;
        }
    }
    explicit UMat(const Mat &shared) : m(shared) {}
    Size size() const { return m.size(); }
    Mat getMat(int) const { return m; }
};
inline UMat Mat::getUMat(int) const { return UMat(*this); }

// dst(y,x) = delta + sum_{ky,kx} K(ky,kx) * src(clamp(y+ky-ay), clamp(x+kx-ax)), anchor (-1,-1) = kernel centre
inline void filter2D(const UMat &src, UMat &dst, int /*ddepth*/, const UMat &kernel, Point anchor = Point(-1, -1), double delta = 0.0,
                     int borderType = BORDER_REPLICATE) {
    assert(borderType == BORDER_REPLICATE);
    (void)borderType;
    const Mat &s = src.m, &k = kernel.m;
    const int ax = anchor.x < 0 ? k.cols / 2 : anchor.x, ay = anchor.y < 0 ? k.rows / 2 : anchor.y;
    // replicate-padded copy once, then every tap is a contiguous (vectorisable) row sweep; per pixel the taps are still
    // accumulated in row-major order starting from delta, one multiply and one add each (-ffp-contract=off)
    const int pw = s.cols + k.cols - 1, ph = s.rows + k.rows - 1;
    std::vector<float> pad((size_t)pw * ph);
    for (int y = 0; y < ph; y++) {
        const float *row = s.ptr<float>(std::min(std::max(y - ay, 0), s.rows - 1));
        float *p = &pad[(size_t)y * pw];
        for (int x = 0; x < pw; x++) p[x] = row[std::min(std::max(x - ax, 0), s.cols - 1)];
    }
    dst.m.create(s.rows, s.cols, CV_32FC1);              // (the padded copy makes in-place filtering safe)
    const float d0 = static_cast<float>(delta);
    for (int y = 0; y < s.rows; y++) {
        float *o = dst.m.ptr<float>(y);
        for (int x = 0; x < s.cols; x++) o[x] = d0;
        for (int ky = 0; ky < k.rows; ky++) {
            const float *kr = k.ptr<float>(ky);
            for (int kx = 0; kx < k.cols; kx++) {
                const float kv = kr[kx];
                const float *p = &pad[(size_t)(y + ky) * pw + kx];
                for (int x = 0; x < s.cols; x++) o[x] = o[x] + kv * p[x];
            }
        }
    }
}

inline void add(const UMat &a, const UMat &b, UMat &dst) {
    dst.m.create(a.m.rows, a.m.cols, CV_32FC1);
    for (int y = 0; y < a.m.rows; y++) {
        const float *pa = a.m.ptr<float>(y), *pb = b.m.ptr<float>(y);
        float *pd = dst.m.ptr<float>(y);
        for (int x = 0; x < a.m.cols; x++) pd[x] = pa[x] + pb[x];
    }
}
inline void add(const UMat &a, double scalar, UMat &dst) {
    const float s = static_cast<float>(scalar);
    dst.m.create(a.m.rows, a.m.cols, CV_32FC1);
    for (int y = 0; y < a.m.rows; y++) {
        const float *pa = a.m.ptr<float>(y);
        float *pd = dst.m.ptr<float>(y);
        for (int x = 0; x < a.m.cols; x++) pd[x] = pa[x] + s;
    }
}
inline void max(const UMat &a, double scalar, UMat &dst) {
    const float s = static_cast<float>(scalar);
    dst.m.create(a.m.rows, a.m.cols, CV_32FC1);
    for (int y = 0; y < a.m.rows; y++) {
        const float *pa = a.m.ptr<float>(y);
        float *pd = dst.m.ptr<float>(y);
        for (int x = 0; x < a.m.cols; x++) pd[x] = pa[x] > s ? pa[x] : s;
    }
}
inline void min(const UMat &a, double scalar, UMat &dst) {
    const float s = static_cast<float>(scalar);
    dst.m.create(a.m.rows, a.m.cols, CV_32FC1);
    for (int y = 0; y < a.m.rows; y++) {
        const float *pa = a.m.ptr<float>(y);
        float *pd = dst.m.ptr<float>(y);
        for (int x = 0; x < a.m.cols; x++) pd[x] = pa[x] < s ? pa[x] : s;
    }
}
// dst = src1 * alpha + src2
inline void scaleAdd(const UMat &src1, double alpha, const UMat &src2, UMat &dst) {
    const float a = static_cast<float>(alpha);
    dst.m.create(src1.m.rows, src1.m.cols, CV_32FC1);
    for (int y = 0; y < src1.m.rows; y++) {
        const float *p1 = src1.m.ptr<float>(y), *p2 = src2.m.ptr<float>(y);
        float *pd = dst.m.ptr<float>(y);
        for (int x = 0; x < src1.m.cols; x++) pd[x] = p1[x] * a + p2[x];
    }
}

inline void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int borderType) {
    assert(borderType == BORDER_REPLICATE);
    (void)borderType;
    Mat out(src.rows + top + bottom, src.cols + left + right, CV_32FC1);
    for (int y = 0; y < out.rows; y++) {
        const float *s = src.ptr<float>(std::min(std::max(y - top, 0), src.rows - 1));
        float *o = out.ptr<float>(y);
        for (int x = 0; x < out.cols; x++) o[x] = s[std::min(std::max(x - left, 0), src.cols - 1)];
    }
    dst = out;
}

#ifdef W2X_CVSHIM_WITH_IMGPROC
enum ImreadModes { IMREAD_COLOR = 1 };
enum ColorConversionCodes { COLOR_RGB2YUV = 83, COLOR_YUV2RGB = 85 };
enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2 };

namespace shim_detail {
inline w2ximg::Image3f to_image(const Mat &m) {
    assert(m.type() == CV_32FC3);
    w2ximg::Image3f im(m.cols, m.rows);
    for (int y = 0; y < m.rows; y++) std::memcpy(&im.data[(size_t)y * m.cols * 3], m.ptr<float>(y), (size_t)m.cols * 3 * sizeof(float));
    return im;
}
inline Mat from_image(const w2ximg::Image3f &im) {
    Mat m(im.height, im.width, CV_32FC3);
    std::memcpy(m.data, im.data.data(), im.data.size() * sizeof(float));
    return m;
}
}  // namespace shim_detail

// cv::imread(path, IMREAD_COLOR): 8-bit B,G,R; an empty Mat when the file cannot be read (src/main.cpp:74)
inline Mat imread(const std::string &path, int /*flags*/ = IMREAD_COLOR) {
    w2xio::Image8 im = w2xio::imread(path);
    if (im.empty()) return Mat();
    Mat m(im.height, im.width, CV_8UC3);
    std::memcpy(m.data, im.bgr.data(), im.bgr.size());
    return m;
}
inline bool imwrite(const std::string &path, const Mat &m) {
    assert(m.type() == CV_8UC3 && m.step == (size_t)m.cols * 3);
    return w2xio::imwrite(path, m.data, m.cols, m.rows);
}
inline void cvtColor(const Mat &src, Mat &dst, int code) {
    w2ximg::Image3f im = shim_detail::to_image(src);
    if (code == COLOR_RGB2YUV) w2ximg::rgb2yuv(im);
    else if (code == COLOR_YUV2RGB) w2ximg::yuv2rgb(im);
    else assert(!"cvtColor: unsupported code in the shim");
    dst = shim_detail::from_image(im);
}
inline void resize(const Mat &src, Mat &dst, Size dsize, double /*fx*/ = 0, double /*fy*/ = 0, int interpolation = INTER_LINEAR) {
    const w2ximg::Interp ip = interpolation == INTER_NEAREST ? w2ximg::NEAREST : interpolation == INTER_CUBIC ? w2ximg::CUBIC : w2ximg::LINEAR;
    dst = shim_detail::from_image(w2ximg::resize(shim_detail::to_image(src), dsize.width, dsize.height, ip));
}
inline void split(const Mat &src, std::vector<Mat> &planes) {
    assert(src.type() == CV_32FC3);
    planes.clear();
    for (int c = 0; c < 3; c++) {
        Mat p(src.rows, src.cols, CV_32FC1);
        for (int y = 0; y < src.rows; y++) {
            const float *s = src.ptr<float>(y);
            float *d = p.ptr<float>(y);
            for (int x = 0; x < src.cols; x++) d[x] = s[3 * x + c];
        }
        planes.push_back(p);
    }
}
inline void merge(const std::vector<Mat> &planes, Mat &dst) {
    assert(planes.size() == 3);
    Mat out(planes[0].rows, planes[0].cols, CV_32FC3);
    for (int c = 0; c < 3; c++)
        for (int y = 0; y < out.rows; y++) {
            const float *s = planes[(size_t)c].ptr<float>(y);
            float *d = out.ptr<float>(y);
            for (int x = 0; x < out.cols; x++) d[3 * x + c] = s[x];
        }
    dst = out;
}
#endif  // W2X_CVSHIM_WITH_IMGPROC

// std::cout << mat (Model::printWeightMatrix, debugging only): "[a, b, c;\n d, e, f]"
inline std::ostream &operator<<(std::ostream &os, const Mat &m) {
    os << "[";
    for (int y = 0; y < m.rows; y++) {
        for (int x = 0; x < m.cols; x++) os << m.at<float>(y, x) << (x + 1 < m.cols ? ", " : "");
        os << (y + 1 < m.rows ? ";\n " : "");
    }
    return os << "]";
}

}  // namespace cv
#endif
