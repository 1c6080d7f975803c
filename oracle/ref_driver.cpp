// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C entry points around the REFERENCE'S OWN hot-path code (w2xc::modelUtility::generateModelFromJSON,
// w2xc::Model::filter, w2xc::convertWithModels), compiled from /root/reference/src/{modelHandler,convertRoutine}.cpp
// against the OpenCV API shim in oracle/cvshim (the image has no OpenCV C++).  Built by oracle/Makefile into
// oracle/_ref/libw2x_reference.so when /root/reference is present; used by tests/test_reference_build.py to pin the
// restated oracle (oracle/w2x_oracle.c) and the golden vectors against the reference's real control flow.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <streambuf>
#include <string>
#include <vector>

#include "convertRoutine.hpp"   // the reference's header (-I /root/reference/src), pulls in modelHandler.hpp

namespace {
struct NullBuf : std::streambuf {
    int overflow(int c) override { return c; }
};
struct Quiet {   // the reference reports progress on std::cout; keep test output readable
    NullBuf nb;
    std::streambuf *old;
    Quiet() : old(std::cout.rdbuf(&nb)) {}
    ~Quiet() { std::cout.rdbuf(old); }
};
struct Handle {
    std::vector<std::unique_ptr<w2xc::Model>> models;
};
cv::Mat wrap(const float *p, int w, int h, long stride_floats) {   // deep copy into a dense cv::Mat
    cv::Mat m(h, w, CV_32FC1);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) m.at<float>(y, x) = p[(long)y * stride_floats + x];
    return m;
}
}  // namespace

extern "C" {

// generateModelFromJSON (src/modelHandler.cpp:170-197): returns the number of layers, or -1
int w2xr_load(const char *path, void **out) {
    std::unique_ptr<Handle> h(new Handle);
    if (!w2xc::modelUtility::generateModelFromJSON(path, h->models)) return -1;
    const int n = (int)h->models.size();
    *out = h.release();
    return n;
}
void w2xr_free(void *h) { delete static_cast<Handle *>(h); }
int w2xr_layer_dims(void *h, int layer, int *n_in, int *n_out) {
    Handle *H = static_cast<Handle *>(h);
    if (layer < 0 || layer >= (int)H->models.size()) return -1;
    *n_in = H->models[(size_t)layer]->getNInputPlanes();
    *n_out = H->models[(size_t)layer]->getNOutputPlanes();
    return 0;
}
// modelUtility setters (src/modelHandler.cpp:199-220)
int w2xr_config(int n_job, int block_exp) {
    bool ok = true;
    if (n_job > 0) ok = w2xc::modelUtility::getInstance().setNumberOfJobs(n_job) && ok;
    if (block_exp >= 0) ok = w2xc::modelUtility::getInstance().setBlockSizeExp2Square(block_exp) && ok;
    return ok ? 0 : -1;
}
// convertWithModels (src/convertRoutine.cpp:21-51)
int w2xr_convert(void *h, const float *in, int w, int hgt, long in_stride_floats, float *out, long out_stride_floats, int block_splitting) {
    Handle *H = static_cast<Handle *>(h);
    Quiet q;
    cv::Mat src = wrap(in, w, hgt, in_stride_floats), dst;
    if (!w2xc::convertWithModels(src, dst, H->models, block_splitting != 0)) return -1;
    if (dst.size().width != w || dst.size().height != hgt) return -2;
    for (int y = 0; y < hgt; y++)
        for (int x = 0; x < w; x++) out[(long)y * out_stride_floats + x] = dst.at<float>(y, x);
    return 0;
}
// convertWithModels with the reference's progress output ("Iteration #k...", "start process block (c,r) ...",
// src/convertRoutine.cpp:67,133-134) captured into `log` (NUL-terminated, truncated to cap): the processing order is
// the observable trace of its block arithmetic.
int w2xr_convert_log(void *h, const float *in, int w, int hgt, long in_stride_floats, float *out, long out_stride_floats, int block_splitting,
                     char *log, int cap) {
    Handle *H = static_cast<Handle *>(h);
    std::ostringstream cap_os;
    std::streambuf *old = std::cout.rdbuf(cap_os.rdbuf());
    cv::Mat src = wrap(in, w, hgt, in_stride_floats), dst;
    const bool ok = w2xc::convertWithModels(src, dst, H->models, block_splitting != 0);
    std::cout.rdbuf(old);
    if (log && cap > 0) {
        const std::string t = cap_os.str();
        const size_t n = std::min(t.size(), (size_t)cap - 1);
        std::memcpy(log, t.data(), n);
        log[n] = 0;
    }
    if (!ok) return -1;
    if (dst.size().width != w || dst.size().height != hgt) return -2;
    for (int y = 0; y < hgt; y++)
        for (int x = 0; x < w; x++) out[(long)y * out_stride_floats + x] = dst.at<float>(y, x);
    return 0;
}
// Model::filter (src/modelHandler.cpp:26-72): planar [n][h][w] in and out
int w2xr_filter(void *h, int layer, const float *in, int w, int hgt, float *out) {
    Handle *H = static_cast<Handle *>(h);
    if (layer < 0 || layer >= (int)H->models.size()) return -1;
    w2xc::Model &M = *H->models[(size_t)layer];
    std::vector<cv::Mat> ip, op;
    for (int i = 0; i < M.getNInputPlanes(); i++) ip.push_back(wrap(in + (long)i * w * hgt, w, hgt, w));
    if (!M.filter(ip, op)) return -2;
    if ((int)op.size() != M.getNOutputPlanes()) return -3;
    for (int o = 0; o < (int)op.size(); o++)
        for (int y = 0; y < hgt; y++)
            for (int x = 0; x < w; x++) out[((long)o * hgt + y) * w + x] = op[(size_t)o].at<float>(y, x);
    return 0;
}

}  // extern "C"
