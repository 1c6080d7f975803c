// imgproc.hpp -- the image plumbing AROUND the conv hot path, as the reference's main.cpp does it with OpenCV
// (src/main.cpp:74-76, 91-98, 132-146, 158-167, 171-172).  This image has no OpenCV C++, so the handful of
// OpenCV calls the CLI needs are restated here on the CPU (they are not the hot path) and pinned against cv2
// in tests/test_cli.py:
//   Mat::convertTo(CV_32F, 1/255)  ·  cvtColor(COLOR_RGB2YUV / COLOR_YUV2RGB) on 3-channel float data
//   resize(INTER_NEAREST | INTER_CUBIC (a = -0.75) | INTER_LINEAR)  ·  convertTo(CV_8U, 255) (round-half-even, saturate)
#ifndef W2X_IMGPROC_HPP_
#define W2X_IMGPROC_HPP_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <utility>
#include <vector>

namespace w2ximg {

// Row-parallel helper: the per-pixel arithmetic below does not depend on the split, so results are bit-identical to a
// single-threaded sweep.  set_threads(n) bounds the workers (the CLI passes its -j value; default = hardware threads, max 32).
inline int &thread_limit() {
    static int n = 0;
    return n;
}
inline void set_threads(int n) { thread_limit() = n; }
template <typename F>
inline void parallel_rows(int rows, long work_per_row, F &&fn) {   // fn(row_begin, row_end)
    int n = thread_limit() > 0 ? thread_limit() : (int)std::thread::hardware_concurrency();
    n = std::max(1, std::min(n, 32));
    if ((long)rows * work_per_row < (1L << 18)) n = 1;             // small images: not worth the thread launches
    n = std::min(n, std::max(rows, 1));
    if (n == 1) { fn(0, rows); return; }
    std::vector<std::thread> ts;
    for (int t = 0; t < n; t++) {
        const int r0 = (int)((long)rows * t / n), r1 = (int)((long)rows * (t + 1) / n);
        ts.emplace_back([&fn, r0, r1] { fn(r0, r1); });
    }
    for (auto &th : ts) th.join();
}

// std::vector without the zero fill of resize()/the sizing constructor: the images here are 100 MB and every element is
// written by the (row-parallel) routine that creates them, so the fill would be a serial extra pass over fresh pages
template <typename T>
struct no_init_alloc : std::allocator<T> {
    template <typename U> struct rebind { using other = no_init_alloc<U>; };
    template <typename U> void construct(U *p) noexcept { ::new (static_cast<void *>(p)) U; }
    template <typename U, typename... A> void construct(U *p, A &&...a) { ::new (static_cast<void *>(p)) U(std::forward<A>(a)...); }
};
using FloatBuf = std::vector<float, no_init_alloc<float>>;

#if defined(__GNUC__) && defined(__x86_64__) && !defined(W2X_NO_TARGET_CLONES)
#define W2X_SIMD_CLONES __attribute__((target_clones("avx2", "default")))   // runtime dispatch; no FMA: the arithmetic stays the scalar code's
#else
#define W2X_SIMD_CLONES
#endif

// interleaved 3-channel float image (what cv::Mat CV_32FC3 holds), channel order as loaded (B,G,R for imread)
struct Image3f {
    int width = 0, height = 0;
    FloatBuf data;   // [h][w][3]
    Image3f() {}
    Image3f(int w, int h) : width(w), height(h), data((size_t)w * h * 3) {}
    float *px(int y, int x) { return &data[((size_t)y * width + x) * 3]; }
    const float *px(int y, int x) const { return &data[((size_t)y * width + x) * 3]; }
};

// image.convertTo(image, CV_32F, 1.0/255.0)   (src/main.cpp:75): dst = (float)src * (float)(1/255.0)
inline Image3f from_u8(const uint8_t *bgr, int w, int h) {
    Image3f out(w, h);
    const float a = (float)(1.0 / 255.0);
    const size_t row = (size_t)w * 3;
    parallel_rows(h, (long)row, [&](int y0, int y1) {
        for (size_t i = (size_t)y0 * row; i < (size_t)y1 * row; i++) out.data[i] = (float)bgr[i] * a;
    });
    return out;
}

// image.convertTo(image, CV_8U, 255.0)   (src/main.cpp:172): saturate_cast<uchar>(cvRound(v * 255)) , round half to even
W2X_SIMD_CLONES inline void to_u8_span(const float *src, uint8_t *dst, size_t n) {
    for (size_t i = 0; i < n; i++) {
        float v = src[i] * 255.0f;
        v = v > 0.f ? v : 0.f;                 // saturate first (NaN -> 0, as saturate_cast<uchar>(cvRound(NaN)) gives) ...
        v = v < 255.f ? v : 255.f;
        const float t = v + 8388608.0f;        // ... then round to nearest, ties to even (cvRound): at 2^23 the float grid is the integers
        uint32_t bits;
        std::memcpy(&bits, &t, 4);
        dst[i] = (uint8_t)(bits & 0xFFu);
    }
}
inline std::vector<uint8_t> to_u8(const Image3f &im) {
    std::vector<uint8_t> out;
    out.resize(im.data.size());
    const size_t row = (size_t)im.width * 3;
    parallel_rows(im.height, (long)row, [&](int y0, int y1) { to_u8_span(im.data.data() + (size_t)y0 * row, out.data() + (size_t)y0 * row, (size_t)(y1 - y0) * row); });
    return out;
}

// cv::cvtColor(image, image, cv::COLOR_RGB2YUV) on float data (src/main.cpp:76).  The reference feeds BGR data into the
// RGB code; channel 0 is simply treated as "R".  Y = .299 c0 + .587 c1 + .114 c2 ; U = (c2 - Y)*.492 + .5 ; V = (c0 - Y)*.877 + .5
inline void rgb2yuv(Image3f &im) {
    const size_t row = (size_t)im.width * 3;
    parallel_rows(im.height, (long)row * 4, [&](int y0, int y1) {
    for (size_t i = (size_t)y0 * row; i < (size_t)y1 * row; i += 3) {
        float c0 = im.data[i], c1 = im.data[i + 1], c2 = im.data[i + 2];
        float Y = c0 * 0.299f + c1 * 0.587f + c2 * 0.114f;
        float U = (c2 - Y) * 0.492f + 0.5f;
        float V = (c0 - Y) * 0.877f + 0.5f;
        im.data[i] = Y; im.data[i + 1] = U; im.data[i + 2] = V;
    }
    });
}

// cv::cvtColor(image, image, cv::COLOR_YUV2RGB) (src/main.cpp:171):
// c2 = Y + 2.032 (U-.5) ; c1 = Y - 0.395 (U-.5) - 0.581 (V-.5) ; c0 = Y + 1.140 (V-.5)
inline void yuv2rgb(Image3f &im) {
    const size_t row = (size_t)im.width * 3;
    parallel_rows(im.height, (long)row * 4, [&](int y0, int y1) {
    for (size_t i = (size_t)y0 * row; i < (size_t)y1 * row; i += 3) {
        float Y = im.data[i], U = im.data[i + 1], V = im.data[i + 2];
        float c2 = Y + (U - 0.5f) * 2.032f;
        float c1 = Y + (V - 0.5f) * -0.581f + (U - 0.5f) * -0.395f;
        float c0 = Y + (V - 0.5f) * 1.14f;
        im.data[i] = c0; im.data[i + 1] = c1; im.data[i + 2] = c2;
    }
    });
}

// cv::split / cv::merge of one channel
inline void get_channel(const Image3f &im, int c, float *plane, size_t stride_floats) {
    parallel_rows(im.height, (long)im.width * 2, [&](int y0, int y1) {
        for (int y = y0; y < y1; y++) {
            const float *s = im.px(y, 0) + c;
            float *d = plane + (size_t)y * stride_floats;
            for (int x = 0; x < im.width; x++) d[x] = s[(size_t)x * 3];
        }
    });
}
inline std::vector<float> channel(const Image3f &im, int c) {
    std::vector<float> out((size_t)im.width * im.height);
    get_channel(im, c, out.data(), (size_t)im.width);
    return out;
}
inline void set_channel(Image3f &im, int c, const float *plane, size_t stride_floats) {
    parallel_rows(im.height, (long)im.width * 2, [&](int y0, int y1) {
        for (int y = y0; y < y1; y++) {
            float *d = im.px(y, 0) + c;
            const float *s = plane + (size_t)y * stride_floats;
            for (int x = 0; x < im.width; x++) d[(size_t)x * 3] = s[x];
        }
    });
}

enum Interp { NEAREST, LINEAR, CUBIC };

namespace detail {
// vertical pass of one output row: dst[i] = ((0 + r0[i] w0) + r1[i] w1) + ... in this order (the order the tests pin against cv2)
template <int NTAP>
W2X_SIMD_CLONES inline void vpass_row(float *dst, const float *const *rows, const float *w, size_t n) {
    if (NTAP == 4) {
        const float *r0 = rows[0], *r1 = rows[1], *r2 = rows[2], *r3 = rows[3];
        const float w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
        for (size_t i = 0; i < n; i++) dst[i] = ((r0[i] * w0 + r1[i] * w1) + r2[i] * w2) + r3[i] * w3;
    } else {
        const float *r0 = rows[0], *r1 = rows[1];
        const float w0 = w[0], w1 = w[1];
        for (size_t i = 0; i < n; i++) dst[i] = r0[i] * w0 + r1[i] * w1;
    }
}
// horizontal pass of one source row: dw output pixels x 3 channels, taps gathered through xi (element offsets, already x3)
template <int NTAP>
inline void hpass_row(float *dst, const float *src, const int *xi, const float *xw, int dw) {
    for (int x = 0; x < dw; x++, xi += NTAP, xw += NTAP, dst += 3) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int k = 0; k < NTAP; k++) {
            const float *s = src + xi[k];
            const float wk = xw[k];
            a0 += s[0] * wk; a1 += s[1] * wk; a2 += s[2] * wk;
        }
        dst[0] = a0; dst[1] = a1; dst[2] = a2;
    }
}
}  // namespace detail


// Channel c of cv::resize(src, Size(dw, dh), 0, 0, INTER_NEAREST) written straight into a plane (what src/main.cpp:135-139 obtains
// with resize + split): sx = min(floor(dx * scale), w - 1), scale = src/dst in double.
inline void resize_nearest_channel(const Image3f &src, int c, int dw, int dh, float *plane, size_t stride_floats) {
    const double sx = (double)src.width / dw, sy = (double)src.height / dh;
    std::vector<int> xo((size_t)dw);
    for (int x = 0; x < dw; x++) xo[(size_t)x] = 3 * std::min((int)std::floor(x * sx), src.width - 1) + c;
    parallel_rows(dh, (long)dw, [&](int y0, int y1) {
        for (int y = y0; y < y1; y++) {
            const float *s = src.px(std::min((int)std::floor(y * sy), src.height - 1), 0);
            float *d = plane + (size_t)y * stride_floats;
            for (int x = 0; x < dw; x++) d[x] = s[xo[(size_t)x]];
        }
    });
}

// cv::resize(src, dst, Size(dw, dh), 0, 0, interp) for CV_32FC3 (src/main.cpp:135,144,166).
// OpenCV conventions: scale = src/dst (double); nearest: sx = floor(dx*scale); linear/cubic: fx = (dx+0.5)*scale-0.5,
// taps clamped to the image (replicate); bicubic kernel A = -0.75; horizontal pass then vertical pass in float.
inline Image3f resize(const Image3f &src, int dw, int dh, Interp interp) {
    Image3f dst(dw, dh);
    const double sx = (double)src.width / dw, sy = (double)src.height / dh;
    if (interp == NEAREST) {
        std::vector<int> xo((size_t)dw);
        for (int x = 0; x < dw; x++) xo[(size_t)x] = 3 * std::min((int)std::floor(x * sx), src.width - 1);
        parallel_rows(dh, (long)dw * 3, [&](int y0, int y1) {
            for (int y = y0; y < y1; y++) {
                const float *srow = src.px(std::min((int)std::floor(y * sy), src.height - 1), 0);
                float *d = dst.px(y, 0);
                for (int x = 0; x < dw; x++, d += 3) {
                    const float *s = srow + xo[(size_t)x];
                    d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
                }
            }
        });
        return dst;
    }
    const int ntap = interp == CUBIC ? 4 : 2;
    auto coeffs = [&](double scale, int d, int limit, int *idx, float *w) {
        // source coordinate in double, fraction rounded to float once (what the OpenCV 4.x the tests pin against does; taking
        // the fraction of a float coordinate loses ~2e-5 on planes wider than a few hundred pixels)
        const double fd = (d + 0.5) * scale - 0.5;
        int s = (int)std::floor(fd);
        float f = (float)(fd - (double)s);
        if (interp == LINEAR) {
            if (s < 0) { f = 0.f; s = 0; }
            if (s >= limit - 1) { f = 0.f; s = limit - 1; }
            idx[0] = s; idx[1] = std::min(s + 1, limit - 1);
            w[0] = 1.f - f; w[1] = f;
        } else {
            const float A = -0.75f;
            w[0] = ((A * (f + 1) - 5 * A) * (f + 1) + 8 * A) * (f + 1) - 4 * A;
            w[1] = ((A + 2) * f - (A + 3)) * f * f + 1;
            w[2] = ((A + 2) * (1 - f) - (A + 3)) * (1 - f) * (1 - f) + 1;
            w[3] = 1.f - w[0] - w[1] - w[2];
            for (int k = 0; k < 4; k++) idx[k] = std::min(std::max(s - 1 + k, 0), limit - 1);
        }
    };
    // horizontal pass: rows of src -> tmp (src.height x dw)
    std::vector<int> xi((size_t)dw * ntap);
    std::vector<float> xw((size_t)dw * ntap);
    for (int x = 0; x < dw; x++) coeffs(sx, x, src.width, &xi[(size_t)x * ntap], &xw[(size_t)x * ntap]);
    for (auto &v : xi) v *= 3;                                                    // element offsets into an interleaved row
    FloatBuf tmp((size_t)src.height * dw * 3);
    parallel_rows(src.height, (long)dw * 3 * ntap, [&](int y0, int y1) {
        for (int y = y0; y < y1; y++) {
            float *t = &tmp[(size_t)y * dw * 3];
            if (ntap == 4) detail::hpass_row<4>(t, src.px(y, 0), xi.data(), xw.data(), dw);
            else detail::hpass_row<2>(t, src.px(y, 0), xi.data(), xw.data(), dw);
        }
    });
    // vertical pass
    parallel_rows(dh, (long)dw * 3 * ntap, [&](int y0, int y1) {
        for (int y = y0; y < y1; y++) {
            int yi[4];
            float yw[4];
            coeffs(sy, y, src.height, yi, yw);
            const float *rows[4];
            for (int k = 0; k < ntap; k++) rows[k] = &tmp[(size_t)yi[k] * dw * 3];
            if (ntap == 4) detail::vpass_row<4>(dst.px(y, 0), rows, yw, (size_t)dw * 3);
            else detail::vpass_row<2>(dst.px(y, 0), rows, yw, (size_t)dw * 3);
        }
    });
    return dst;
}

}  // namespace w2ximg
#endif
