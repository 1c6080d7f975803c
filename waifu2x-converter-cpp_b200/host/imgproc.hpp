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
#include <thread>
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

// interleaved 3-channel float image (what cv::Mat CV_32FC3 holds), channel order as loaded (B,G,R for imread)
struct Image3f {
    int width = 0, height = 0;
    std::vector<float> data;   // [h][w][3]
    Image3f() {}
    Image3f(int w, int h) : width(w), height(h), data((size_t)w * h * 3, 0.f) {}
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
inline std::vector<uint8_t> to_u8(const Image3f &im) {
    std::vector<uint8_t> out(im.data.size());
    const size_t row = (size_t)im.width * 3;
    parallel_rows(im.height, (long)row, [&](int y0, int y1) {
        for (size_t i = (size_t)y0 * row; i < (size_t)y1 * row; i++) {
            float v = im.data[i] * 255.0f;
            long r = std::lrintf(v);   // FE_TONEAREST: ties to even, as cvRound
            out[i] = (uint8_t)std::min(255L, std::max(0L, r));
        }
    });
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
inline std::vector<float> channel(const Image3f &im, int c) {
    std::vector<float> out((size_t)im.width * im.height);
    for (size_t i = 0; i < out.size(); i++) out[i] = im.data[i * 3 + c];
    return out;
}
inline void set_channel(Image3f &im, int c, const float *plane, size_t stride_floats) {
    for (int y = 0; y < im.height; y++)
        for (int x = 0; x < im.width; x++) im.px(y, x)[c] = plane[(size_t)y * stride_floats + x];
}

enum Interp { NEAREST, LINEAR, CUBIC };

// cv::resize(src, dst, Size(dw, dh), 0, 0, interp) for CV_32FC3 (src/main.cpp:135,144,166).
// OpenCV conventions: scale = src/dst (double); nearest: sx = floor(dx*scale); linear/cubic: fx = (dx+0.5)*scale-0.5,
// taps clamped to the image (replicate); bicubic kernel A = -0.75; horizontal pass then vertical pass in float.
inline Image3f resize(const Image3f &src, int dw, int dh, Interp interp) {
    Image3f dst(dw, dh);
    const double sx = (double)src.width / dw, sy = (double)src.height / dh;
    if (interp == NEAREST) {
        parallel_rows(dh, (long)dw * 3, [&](int y0, int y1) {
            for (int y = y0; y < y1; y++) {
                int yy = std::min((int)std::floor(y * sy), src.height - 1);
                for (int x = 0; x < dw; x++) {
                    int xx = std::min((int)std::floor(x * sx), src.width - 1);
                    const float *s = src.px(yy, xx);
                    float *d = dst.px(y, x);
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
    std::vector<float> tmp((size_t)src.height * dw * 3);
    parallel_rows(src.height, (long)dw * 3 * ntap, [&](int y0, int y1) {
        for (int y = y0; y < y1; y++)
            for (int x = 0; x < dw; x++)
                for (int c = 0; c < 3; c++) {
                    float acc = 0.f;
                    for (int k = 0; k < ntap; k++) acc += src.px(y, xi[(size_t)x * ntap + k])[c] * xw[(size_t)x * ntap + k];
                    tmp[((size_t)y * dw + x) * 3 + c] = acc;
                }
    });
    // vertical pass
    parallel_rows(dh, (long)dw * 3 * ntap, [&](int y0, int y1) {
        for (int y = y0; y < y1; y++) {
            int yi[4];
            float yw[4];
            coeffs(sy, y, src.height, yi, yw);
            for (int x = 0; x < dw; x++)
                for (int c = 0; c < 3; c++) {
                    float acc = 0.f;
                    for (int k = 0; k < ntap; k++) acc += tmp[((size_t)yi[k] * dw + x) * 3 + c] * yw[k];
                    dst.px(y, x)[c] = acc;
                }
        }
    });
    return dst;
}

}  // namespace w2ximg
#endif
