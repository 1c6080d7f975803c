// engine.cu -- GPU context, plane driver and the compute half of the C ABI.
//
// The plane driver re-creates w2xc::convertWithModels (reference src/convertRoutine.cpp:21-51),
// convertWithModelsBasic (:53-82) and convertWithModelsBlockSplit (:84-169) on top of the layer
// kernels: replicate-pad by nModel, run the layers, crop nModel.  A plane the reference would
// block-split is by default processed whole (every output pixel still sees exactly the operands
// and the operation order it sees inside its reference block, so the result is bit-identical);
// W2X_WALK_BLOCKS walks the reference's blocks literally.
//
// There is no CPU fallback anywhere in this file: without an sm_100 device every compute entry
// point fails with W2X_ERR_NO_DEVICE.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.h"
#include "w2x_internal.h"

using namespace w2x;

namespace {

struct DevModel {                       // device-resident copy of one model
    std::vector<float *> w;             // per layer [Cout][Cin][9] fp32
    std::vector<float *> b;             // per layer [Cout] fp32  ((float)bias, src/modelHandler.cpp:147)
    std::vector<std::vector<float>> b_host;   // the same on the host (the tcgen05 kernels take them as kernel parameters)
    std::vector<uint16_t *> pack;       // per layer tcgen05 operand image (nullptr if not eligible)
    std::vector<uint8_t *> pack8;       // same for the "f8" flavour (fp16 main product + e4m3 corrections)
    std::vector<uint8_t *> strip, strip8;   // row-strip kernel images of the narrow layers (nullptr otherwise), both flavours
    std::vector<float> out_scale;       // 1 / (wscale * ACT_SCALE)
    std::vector<float> last_w_t;        // HOST: last layer's weights transposed to [9][Cin] (fused last layer, passed as kernel parameters)
};

struct TimedSpan { int layer; cudaEvent_t e0, e1; };

}  // namespace

struct w2x_ctx {
    int device = 0;
    int num_sms = 0;
    int cc_major = 0, cc_minor = 0;
    int engine = W2X_ENGINE_AUTO;
    int walk = W2X_WALK_FUSED;
    bool fuse_last = true;             // fold the N->1 last layer into the preceding tcgen05 layer's epilogue
    int precision = W2X_PRECISION_F16_F8X2;   // default; W2X_PRECISION=f16x3 in the environment or w2x_ctx_set_precision() selects the 3 x fp16 scheme
    int strip = 1;                     // 1 = run the narrow layers (Cin, Cout <= 64) on the row-strip kernel; w2x_debug_set_strip(0) = 16x16-tile kernel
    int pair = 1;                      // 1 = run the 128-wide layers on CTA pairs (cta_group::2); w2x_debug_set_pair(0) = single-CTA kernels
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    size_t scratch_limit = (size_t)16 << 30;
    w2x_log_fn log = nullptr;
    void *log_user = nullptr;
    uint64_t launches = 0;
    bool timing = false;
    std::vector<TimedSpan> spans;
    std::vector<cudaEvent_t> event_pool;
    std::vector<std::string> layer_kernel;
    std::map<uint64_t, DevModel> models;
    // scratch
    void *buf[2] = {nullptr, nullptr};
    size_t buf_bytes[2] = {0, 0};
    float *pad_buf = nullptr;
    size_t pad_bytes = 0;
    float *io_buf[2] = {nullptr, nullptr};   // device staging for the host-buffer entry points
    size_t io_bytes[2] = {0, 0};
    bool tc_ready = false;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;   // host<->device copies of w2x_convert_plane overlap the compute stream
    cudaEvent_t ev_in[8] = {}, ev_done[8] = {};
    int host_bands = 0;                                   // 0 = automatic (up to 4 bands of >= 512 rows), 1 = no pipelining
    unsigned long long *prof_buf = nullptr;   // [16 layers][PROF_MAX_CTAS][PROF_WORDS], debug profile
};

namespace {

#define CU_CHECK(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(W2X_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(e__), __FILE__, __LINE__, \
                        cudaGetErrorString(e__));                                                   \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int ensure(void **p, size_t *have, size_t need) {
    if (*have >= need) return W2X_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *have = 0;
    cudaError_t e = cudaMalloc(p, need);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(W2X_ERR_NOMEM, "cudaMalloc of %zu bytes failed (%s)", need, cudaGetErrorString(e));
    }
    *have = need;
    return W2X_OK;
}

int get_dev_model(w2x_ctx *ctx, const w2x_model *m, DevModel **out) {
    auto it = ctx->models.find(m->uid);
    if (it != ctx->models.end()) {
        *out = &it->second;
        return W2X_OK;
    }
    DevModel dm;
    const size_t n = m->layers.size();
    dm.w.assign(n, nullptr);
    dm.b.assign(n, nullptr);
    dm.pack.assign(n, nullptr);
    dm.pack8.assign(n, nullptr);
    dm.strip.assign(n, nullptr);
    dm.strip8.assign(n, nullptr);
    dm.out_scale.assign(n, 1.f);
    for (size_t i = 0; i < n; i++) {
        const Layer &L = m->layers[i];
        CU_CHECK(cudaMalloc(&dm.w[i], L.w.size() * sizeof(float)));
        CU_CHECK(cudaMemcpyAsync(dm.w[i], L.w.data(), L.w.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        std::vector<float> bf(L.b.size());
        for (size_t k = 0; k < bf.size(); k++) bf[k] = static_cast<float>(L.b[k]);
        dm.b_host.push_back(bf);
        CU_CHECK(cudaMalloc(&dm.b[i], bf.size() * sizeof(float)));
        CU_CHECK(cudaMemcpyAsync(dm.b[i], bf.data(), bf.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        CU_CHECK(cudaStreamSynchronize(ctx->stream));   // bf goes out of scope
        const TcPack &P = m->tc[i];
        if (!P.bytes.empty()) {
            CU_CHECK(cudaMalloc(&dm.pack[i], P.bytes.size() * 2));
            CU_CHECK(cudaMemcpyAsync(dm.pack[i], P.bytes.data(), P.bytes.size() * 2, cudaMemcpyHostToDevice, ctx->stream));
            dm.out_scale[i] = 1.0f / (P.wscale * tc::ACT_SCALE);
            CU_CHECK(cudaMalloc(&dm.pack8[i], P.bytes8.size()));
            CU_CHECK(cudaMemcpyAsync(dm.pack8[i], P.bytes8.data(), P.bytes8.size(), cudaMemcpyHostToDevice, ctx->stream));
            if (!P.strip.empty()) {
                CU_CHECK(cudaMalloc(&dm.strip[i], P.strip.size()));
                CU_CHECK(cudaMemcpyAsync(dm.strip[i], P.strip.data(), P.strip.size(), cudaMemcpyHostToDevice, ctx->stream));
                CU_CHECK(cudaMalloc(&dm.strip8[i], P.strip8.size()));
                CU_CHECK(cudaMemcpyAsync(dm.strip8[i], P.strip8.data(), P.strip8.size(), cudaMemcpyHostToDevice, ctx->stream));
            }
        }
    }
    if (m->tc_eligible) {
        const Layer &L = m->layers.back();                // n_out == 1: w is [1][Cin][9]
        dm.last_w_t.assign((size_t)9 * L.n_in, 0.f);
        for (int c = 0; c < L.n_in; c++)
            for (int t = 0; t < 9; t++) dm.last_w_t[(size_t)t * L.n_in + c] = L.w[(size_t)c * 9 + t];
    }
    CU_CHECK(cudaStreamSynchronize(ctx->stream));
    auto res = ctx->models.emplace(m->uid, std::move(dm));
    *out = &res.first->second;
    return W2X_OK;
}

cudaEvent_t take_event(w2x_ctx *ctx) {
    if (!ctx->event_pool.empty()) {
        cudaEvent_t e = ctx->event_pool.back();
        ctx->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

struct LayerTimer {   // brackets one layer launch with events when timing is on
    w2x_ctx *ctx;
    TimedSpan span{};
    bool on;
    LayerTimer(w2x_ctx *c, int layer) : ctx(c), on(c->timing) {
        if (on) {
            span.layer = layer;
            span.e0 = take_event(c);
            span.e1 = take_event(c);
            cudaEventRecord(span.e0, c->stream);
        }
    }
    ~LayerTimer() {
        if (on) {
            cudaEventRecord(span.e1, ctx->stream);
            ctx->spans.push_back(span);
        }
    }
};

void note_kernel(w2x_ctx *ctx, int layer, const char *name) {
    if ((int)ctx->layer_kernel.size() <= layer) ctx->layer_kernel.resize((size_t)layer + 1);
    ctx->layer_kernel[(size_t)layer] = name;
}

void logf(w2x_ctx *ctx, const char *fmt, int a, int b = 0) {
    if (!ctx->log) return;
    char line[128];
    snprintf(line, sizeof line, fmt, a, b);
    ctx->log(line, ctx->log_user);
}

int pick_engine(w2x_ctx *ctx, const w2x_model *m) {
    int e = ctx->engine;
    if (e == W2X_ENGINE_AUTO) e = m->tc_eligible ? W2X_ENGINE_TC : W2X_ENGINE_FP32;
    if (e == W2X_ENGINE_TC && !m->tc_eligible) {
        fail(W2X_ERR_UNSUPPORTED, "tcgen05 engine needs a 1->{32,64,128}...->1 layer chain");
        return -1;
    }
    return e;
}

int ensure_tc(w2x_ctx *ctx) {
    if (ctx->tc_ready) return W2X_OK;
    CU_CHECK(tc::init_kernels());
    ctx->tc_ready = true;
    return W2X_OK;
}

// One tcgen05 layer `li` on frames of pw x ph: in -> out (or, fused with the last layer, -> per-pixel tap partials in `out`).
int launch_layer_tc(w2x_ctx *ctx, const w2x_model *m, DevModel *dm, int li, const __half *in, __half *out, int pw, int ph,
                    bool fused, bool profile) {
    const Layer &L = m->layers[(size_t)li];
    const int f8 = ctx->precision == W2X_PRECISION_F16_F8X2 ? 1 : 0;
    const bool strip = ctx->strip && !fused && tc::strip_supported(L.n_in, L.n_out);
    {
        LayerTimer t(ctx, li);
        CU_CHECK(tc::launch_tc_layer(in, f8 ? (const void *)dm->pack8[(size_t)li] : (const void *)dm->pack[(size_t)li],
                                     strip ? (f8 ? (const void *)dm->strip8[(size_t)li] : (const void *)dm->strip[(size_t)li]) : nullptr,
                                     dm->b_host[(size_t)li].data(), out, L.n_in, L.n_out, pw, ph, dm->out_scale[(size_t)li], f8,
                                     ctx->num_sms, ctx->stream,
                                     profile && ctx->prof_buf ? ctx->prof_buf + (size_t)li * tc::PROF_MAX_CTAS * tc::PROF_WORDS : nullptr,
                                     fused ? dm->last_w_t.data() : nullptr, fused ? reinterpret_cast<float *>(out) : nullptr, ctx->pair));
    }
    note_kernel(ctx, li, f8 ? (fused ? "tcgen05_f16+f8x2+last" : strip ? "tcgen05_f16+f8x2_strip" : "tcgen05_f16+f8x2")
                            : (fused ? "tcgen05_f16x3+last" : strip ? "tcgen05_f16x3_strip" : "tcgen05_f16x3"));
    ctx->launches++;
    return W2X_OK;
}

// ---- convertWithModelsBasic on an already padded ROI ------------------------------------------
// src: pw x ph fp32 region (row stride src_stride floats) that already contains the n-pixel ring.
// dst: receives the (pw-2n) x (ph-2n) interior.
int run_basic(w2x_ctx *ctx, const w2x_model *m, DevModel *dm, int engine, const float *src, long src_stride, int pw,
              int ph, float *dst, long dst_stride) {
    const int n = (int)m->layers.size();
    if (pw - 2 * n < 1 || ph - 2 * n < 1) return fail(W2X_ERR_ARG, "plane smaller than the model's receptive ring");
    int maxc = 1;
    for (auto &L : m->layers) maxc = std::max(maxc, std::max(L.n_in, L.n_out));
    if (engine == W2X_ENGINE_FP32) {
        const size_t need = (size_t)maxc * pw * ph * sizeof(float);
        for (int i = 0; i < 2; i++) {
            int rc = ensure(&ctx->buf[i], &ctx->buf_bytes[i], need);
            if (rc) return rc;
        }
        float *cur = static_cast<float *>(ctx->buf[0]), *nxt = static_cast<float *>(ctx->buf[1]);
        CU_CHECK(launch_copy2d(src, src_stride, cur, pw, pw, ph, ctx->stream));   // ROI -> dense plane
        ctx->launches++;
        for (int li = 0; li < n; li++) {
            const Layer &L = m->layers[(size_t)li];
            logf(ctx, "Iteration #%d...", li + 1);                                 // src/convertRoutine.cpp:67
            {
                LayerTimer t(ctx, li);
                CU_CHECK(launch_conv3x3_fp32(cur, nxt, dm->w[(size_t)li], dm->b[(size_t)li], L.n_in, L.n_out, pw, ph,
                                             ctx->stream));
            }
            note_kernel(ctx, li, "fp32_direct");
            ctx->launches++;
            std::swap(cur, nxt);
        }
        CU_CHECK(launch_crop(cur, pw - 2 * n, ph - 2 * n, n, dst, dst_stride, ctx->stream));
        ctx->launches++;
        return W2X_OK;
    }
    // ---- tcgen05 engine ----
    const int f8 = ctx->precision == W2X_PRECISION_F16_F8X2 ? 1 : 0;
    int rc = ensure_tc(ctx);
    if (rc) return rc;
    const size_t need = tc::act_bytes(maxc, pw, ph);
    for (int i = 0; i < 2; i++) {
        rc = ensure(&ctx->buf[i], &ctx->buf_bytes[i], need);
        if (rc) return rc;
    }
    __half *cur = static_cast<__half *>(ctx->buf[0]), *nxt = static_cast<__half *>(ctx->buf[1]);
    {
        const Layer &L = m->layers[0];
        logf(ctx, "Iteration #%d...", 1);
        LayerTimer t(ctx, 0);
        CU_CHECK(tc::launch_first(src, src_stride, pw, ph, L.w.data(), dm->b_host[0].data(), L.n_out, cur, ctx->stream, f8));
        note_kernel(ctx, 0, "first_1xN");
        ctx->launches++;
    }
    const bool fuse = ctx->fuse_last && n >= 3 && !dm->last_w_t.empty();
    for (int li = 1; li + 1 < n; li++) {
        logf(ctx, "Iteration #%d...", li + 1);
        rc = launch_layer_tc(ctx, m, dm, li, cur, nxt, pw, ph, fuse && li == n - 2, true);
        if (rc) return rc;
        std::swap(cur, nxt);
    }
    {
        const Layer &L = m->layers[(size_t)n - 1];
        logf(ctx, "Iteration #%d...", n);
        LayerTimer t(ctx, n - 1);
        if (fuse) {
            CU_CHECK(tc::launch_last_gather(reinterpret_cast<const float *>(cur), pw, ph, static_cast<float>(L.b[0]), n, dst,
                                            dst_stride, ctx->stream));
            note_kernel(ctx, n - 1, "last_gather");
        } else {
            CU_CHECK(tc::launch_last(cur, L.n_in, pw, ph, dm->w[(size_t)n - 1], static_cast<float>(L.b[0]), n, dst,
                                     dst_stride, ctx->stream, f8));
            note_kernel(ctx, n - 1, "last_Nx1");
        }
        ctx->launches++;
    }
    return W2X_OK;
}

// Whole plane (or row band) already available as a padded plane: cut it into horizontal bands that
// respect the scratch limit, each band re-reads n rows of context above and below.
int run_padded_plane(w2x_ctx *ctx, const w2x_model *m, DevModel *dm, int engine, const float *padp, int w, int h,
                     float *dst, long dst_stride) {
    const int n = (int)m->layers.size();
    const int pw = w + 2 * n;
    int maxc = 1;
    for (auto &L : m->layers) maxc = std::max(maxc, std::max(L.n_in, L.n_out));
    const size_t per_row = (size_t)maxc * pw * 4;   // both engines: 4 bytes per activation element
    long max_rows = (long)(ctx->scratch_limit / per_row) - 2 * n;
    if (max_rows < 16) max_rows = 16;
    int band = (int)std::min<long>(h, max_rows);
    for (int y0 = 0; y0 < h; y0 += band) {
        const int bh = std::min(band, h - y0);
        int rc = run_basic(ctx, m, dm, engine, padp + (long)y0 * pw, pw, pw, bh + 2 * n, dst + (long)y0 * dst_stride,
                           dst_stride);
        if (rc) return rc;
    }
    return W2X_OK;
}

int check_ctx(w2x_ctx *ctx) {
    if (!ctx) return fail(W2X_ERR_ARG, "NULL context");
    return W2X_OK;
}

int convert_device(w2x_ctx *ctx, const w2x_model *m, const float *d_in, int w, int h, size_t in_stride_bytes,
                   int rows_above, int rows_below, float *d_out, size_t out_stride_bytes, int block_splitting) {
    if (!m || !d_in || !d_out || w < 1 || h < 1) return fail(W2X_ERR_ARG, "w2x_convert_plane: bad argument");
    if (in_stride_bytes % 4 || out_stride_bytes % 4 || in_stride_bytes < (size_t)w * 4 || out_stride_bytes < (size_t)w * 4)
        return fail(W2X_ERR_ARG, "w2x_convert_plane: row strides must be multiples of 4 bytes and >= width*4");
    if (m->layers.front().n_in != 1 || m->layers.back().n_out != 1)
        return fail(W2X_ERR_ARG, "w2x_convert_plane: model must map 1 plane to 1 plane");
    DeviceGuard g(ctx->device);
    const int engine = pick_engine(ctx, m);
    if (engine < 0) return W2X_ERR_UNSUPPORTED;
    DevModel *dm = nullptr;
    int rc = get_dev_model(ctx, m, &dm);
    if (rc) return rc;
    const int n = (int)m->layers.size();
    const int pw = w + 2 * n, ph = h + 2 * n;
    rc = ensure(reinterpret_cast<void **>(&ctx->pad_buf), &ctx->pad_bytes, (size_t)pw * ph * sizeof(float));
    if (rc) return rc;
    // cv::copyMakeBorder(in, temp, n, n, n, n, BORDER_REPLICATE)   (src/convertRoutine.cpp:35, :96)
    CU_CHECK(launch_pad_replicate(d_in, w, h, (long)(in_stride_bytes / 4), n, std::min(rows_above, n),
                                  std::min(rows_below, n), ctx->pad_buf, ctx->stream));
    ctx->launches++;
    const long ostride = (long)(out_stride_bytes / 4);
    const bool split = block_splitting && w2x_requires_splitting(w, h);
    if (split && ctx->walk == W2X_WALK_BLOCKS) {
        // the literal block walk of convertWithModelsBlockSplit (src/convertRoutine.cpp:114-165)
        const Config &cfg = config();
        int nb = block_table(w, h, cfg.block_w, cfg.block_h, n, nullptr, 0, nullptr, nullptr);
        if (nb < 0) return fail(W2X_ERR_ARG, "block size too small for a %d-layer model", n);
        std::vector<int> tab((size_t)nb * 8);
        block_table(w, h, cfg.block_w, cfg.block_h, n, tab.data(), nb, nullptr, nullptr);
        for (int i = 0; i < nb; i++) {
            const int *t = &tab[(size_t)i * 8];
            logf(ctx, "start process block (%d,%d) ...", t[1], t[0]);              // :133-134 prints (c,r)
            const int bw_i = t[5] - t[4], bh_i = t[3] - t[2];
            if (t[6] + bh_i - 2 * n > h || t[7] + bw_i - 2 * n > w)
                return fail(W2X_ERR_ARG, "block (%d,%d) does not fit the output plane (non-square block size?)", t[1], t[0]);
            rc = run_basic(ctx, m, dm, engine, ctx->pad_buf + (long)t[2] * pw + t[4], pw, bw_i, bh_i,
                           d_out + (long)t[6] * ostride + t[7], ostride);
            if (rc) return rc;
        }
        return W2X_OK;
    }
    return run_padded_plane(ctx, m, dm, engine, ctx->pad_buf, w, h, d_out, ostride);
}

}  // namespace

// ================================================================================================
// Row-band session with per-layer halo exchange
// ================================================================================================
struct w2x_band {
    w2x_ctx *ctx = nullptr;
    const w2x_model *model = nullptr;
    DevModel *dm = nullptr;
    int width = 0, rows = 0, n = 0;
    bool up = false, down = false;
    int pt = 0, pb = 0;            // frame rows above / below the band: 1 (neighbour halo) or n (replicated image border)
    int pw = 0, hf = 0;            // frame width / height
    float *pad = nullptr;          // padded fp32 input frame
    __half *act[2] = {nullptr, nullptr};
    size_t act_bytes = 0;
    int cur = 0;                   // act[cur] holds the output of the last queued step
    int last_step = -1;
};

namespace {
int band_check(w2x_band *b) {
    if (!b || !b->ctx || !b->model) return fail(W2X_ERR_ARG, "NULL band session");
    return W2X_OK;
}
}  // namespace

extern "C" {

int w2x_band_create(w2x_ctx *ctx, const w2x_model *model, int width, int band_rows, int has_up, int has_down,
                    w2x_band **out_band) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    if (!model || !out_band || width < 1 || band_rows < 1) return fail(W2X_ERR_ARG, "w2x_band_create: bad argument");
    *out_band = nullptr;
    if (!model->tc_eligible || ctx->engine == W2X_ENGINE_FP32)
        return fail(W2X_ERR_UNSUPPORTED, "w2x_band_create: the per-layer halo mode needs the tcgen05 engine and a 1->{32,64,128}..->1 model");
    DeviceGuard g(ctx->device);
    int rc = ensure_tc(ctx);
    if (rc) return rc;
    auto b = std::make_unique<w2x_band>();
    b->ctx = ctx;
    b->model = model;
    rc = get_dev_model(ctx, model, &b->dm);
    if (rc) return rc;
    b->n = (int)model->layers.size();
    b->width = width;
    b->rows = band_rows;
    b->up = has_up != 0;
    b->down = has_down != 0;
    b->pt = b->up ? 1 : b->n;
    b->pb = b->down ? 1 : b->n;
    b->pw = width + 2 * b->n;
    b->hf = band_rows + b->pt + b->pb;
    int maxc = 1;
    for (auto &L : model->layers) maxc = std::max(maxc, std::max(L.n_in, L.n_out));
    b->act_bytes = tc::act_bytes(maxc, b->pw, b->hf);
    CU_CHECK(cudaMalloc(&b->pad, (size_t)b->pw * b->hf * sizeof(float)));
    for (int i = 0; i < 2; i++) CU_CHECK(cudaMalloc(&b->act[i], b->act_bytes));
    *out_band = b.release();
    return W2X_OK;
}

void w2x_band_destroy(w2x_band *band) {
    if (!band) return;
    if (band->ctx) {
        DeviceGuard g(band->ctx->device);
        cudaStreamSynchronize(band->ctx->stream);
        cudaFree(band->pad);
        cudaFree(band->act[0]);
        cudaFree(band->act[1]);
    }
    delete band;
}

int w2x_band_load(w2x_band *band, const float *d_in, size_t in_stride_bytes) {
    if (band_check(band)) return W2X_ERR_ARG;
    if (!d_in || in_stride_bytes % 4 || in_stride_bytes < (size_t)band->width * 4) return fail(W2X_ERR_ARG, "w2x_band_load: bad input");
    w2x_ctx *ctx = band->ctx;
    DeviceGuard g(ctx->device);
    const long stride = (long)(in_stride_bytes / 4);
    const float *band0 = d_in + (band->up ? stride : 0);
    CU_CHECK(launch_pad_replicate_xy(band0, band->width, band->rows, stride, band->n, band->pt, band->pb, band->up ? 1 : 0,
                                     band->down ? 1 : 0, band->pad, ctx->stream));
    ctx->launches++;
    band->last_step = -1;
    band->cur = 0;
    return W2X_OK;
}

int w2x_band_step(w2x_band *band, int step) {
    if (band_check(band)) return W2X_ERR_ARG;
    if (step != band->last_step + 1 || step < 0 || step > band->n - 2)
        return fail(W2X_ERR_ARG, "w2x_band_step: steps must run in order 0..%d (got %d after %d)", band->n - 2, step, band->last_step);
    w2x_ctx *ctx = band->ctx;
    DeviceGuard g(ctx->device);
    const w2x_model *m = band->model;
    DevModel *dm = band->dm;
    const Layer &L = m->layers[(size_t)step];
    const int f8 = ctx->precision == W2X_PRECISION_F16_F8X2 ? 1 : 0;
    if (step == 0) {
        LayerTimer t(ctx, 0);
        CU_CHECK(tc::launch_first(band->pad, band->pw, band->pw, band->hf, L.w.data(), dm->b_host[0].data(), L.n_out, band->act[0], ctx->stream, f8));
        band->cur = 0;
        note_kernel(ctx, 0, "first_1xN");
        ctx->launches++;
    } else {
        int rc = launch_layer_tc(ctx, m, dm, step, band->act[band->cur], band->act[band->cur ^ 1], band->pw, band->hf, step == band->n - 2, false);
        if (rc) return rc;
        band->cur ^= 1;
    }
    band->last_step = step;
    return W2X_OK;
}

int w2x_band_halo(w2x_band *band, int step, int *n_segments, void **send_up, void **recv_up, void **send_down,
                  void **recv_down, size_t *seg_bytes) {
    if (band_check(band)) return W2X_ERR_ARG;
    if (step != band->last_step || !n_segments || !send_up || !recv_up || !send_down || !recv_down || !seg_bytes)
        return fail(W2X_ERR_ARG, "w2x_band_halo: call it for the step that was queued last");
    const int n = band->n;
    char *base = reinterpret_cast<char *>(band->act[band->cur]);
    // every segment: `bytes` at  base + plane_off + row * pitch + in_row
    struct Seg { size_t plane_off, pitch, in_row; } seg[4];
    size_t bytes;
    int nseg;
    if (step == n - 2) {              // per-pixel tap partials [hf][pw][12] fp32
        bytes = (size_t)band->pw * 12 * sizeof(float);
        seg[0] = {0, bytes, 0};
        nseg = 1;
    } else {
        const size_t C = (size_t)band->model->layers[(size_t)step].n_out, px = (size_t)band->pw * C, fr = px * band->hf;
        if (band->ctx->precision == W2X_PRECISION_F16_F8X2) {   // frame [xh fp16][xh8][xl8]: four segments of pw*C bytes
            bytes = px;
            seg[0] = {0, 2 * px, 0};
            seg[1] = {0, 2 * px, px};
            seg[2] = {2 * fr, px, 0};
            seg[3] = {3 * fr, px, 0};
            nseg = 4;
        } else {                                                 // frame [hi fp16][lo fp16]: two segments of pw*C*2 bytes
            bytes = 2 * px;
            seg[0] = {0, 2 * px, 0};
            seg[1] = {2 * fr, 2 * px, 0};
            nseg = 2;
        }
    }
    *n_segments = nseg;
    *seg_bytes = bytes;
    for (int s = 0; s < 4; s++) {
        const bool on = s < nseg;
        char *pl = on ? base + seg[s].plane_off + seg[s].in_row : nullptr;
        const size_t pitch = on ? seg[s].pitch : 0;
        send_up[s] = on && band->up ? pl + pitch * 1 : nullptr;                                 // first owned row
        recv_up[s] = on && band->up ? pl : nullptr;                                            // halo row above
        send_down[s] = on && band->down ? pl + pitch * (size_t)(band->hf - 2) : nullptr;         // last owned row
        recv_down[s] = on && band->down ? pl + pitch * (size_t)(band->hf - 1) : nullptr;         // halo row below
    }
    return W2X_OK;
}

int w2x_band_finish(w2x_band *band, float *d_out, size_t out_stride_bytes) {
    if (band_check(band)) return W2X_ERR_ARG;
    if (band->last_step != band->n - 2) return fail(W2X_ERR_ARG, "w2x_band_finish: steps 0..%d must have run", band->n - 2);
    if (!d_out || out_stride_bytes % 4 || out_stride_bytes < (size_t)band->width * 4) return fail(W2X_ERR_ARG, "w2x_band_finish: bad output");
    w2x_ctx *ctx = band->ctx;
    DeviceGuard g(ctx->device);
    const Layer &L = band->model->layers.back();
    {
        LayerTimer t(ctx, band->n - 1);
        CU_CHECK(tc::launch_last_gather_xy(reinterpret_cast<const float *>(band->act[band->cur]), band->pw, band->hf,
                                           static_cast<float>(L.b[0]), band->n, band->pt, band->pb, d_out,
                                           (long)(out_stride_bytes / 4), ctx->stream));
    }
    note_kernel(ctx, band->n - 1, "last_gather");
    ctx->launches++;
    band->last_step = band->n - 1;
    return W2X_OK;
}

}  // extern "C"

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int w2x_ctx_create(int device, w2x_ctx **out_ctx) {
    if (!out_ctx) return fail(W2X_ERR_ARG, "w2x_ctx_create: NULL out pointer");
    *out_ctx = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(W2X_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= count) return fail(W2X_ERR_ARG, "w2x_ctx_create: device %d out of range [0,%d)", device, count);
    cudaDeviceProp prop;
    CU_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(W2X_ERR_NO_DEVICE, "device %d (%s) is sm_%d%d; this build carries sm_100a code only", device,
                    prop.name, prop.major, prop.minor);
    auto ctx = std::make_unique<w2x_ctx>();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    DeviceGuard g(device);
    CU_CHECK(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    if (const char *pe = std::getenv("W2X_PRECISION")) {
        if (!std::strcmp(pe, "f16x3")) ctx->precision = W2X_PRECISION_F16X3;
        else if (!std::strcmp(pe, "f16+f8x2") || !std::strcmp(pe, "f8")) ctx->precision = W2X_PRECISION_F16_F8X2;
    }
    CU_CHECK(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    CU_CHECK(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    for (int i = 0; i < 8; i++) {
        CU_CHECK(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
        CU_CHECK(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
    }
    *out_ctx = ctx.release();
    return W2X_OK;
}

void w2x_ctx_destroy(w2x_ctx *ctx) {
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->models) {
        for (auto p : kv.second.w) cudaFree(p);
        for (auto p : kv.second.b) cudaFree(p);
        for (auto p : kv.second.pack) cudaFree(p);
        for (auto p : kv.second.pack8) cudaFree(p);
        for (auto p : kv.second.strip) cudaFree(p);
        for (auto p : kv.second.strip8) cudaFree(p);
    }
    for (int i = 0; i < 2; i++) {
        cudaFree(ctx->buf[i]);
        cudaFree(ctx->io_buf[i]);
    }
    cudaFree(ctx->pad_buf);
    cudaFree(ctx->prof_buf);
    for (auto &s : ctx->spans) { cudaEventDestroy(s.e0); cudaEventDestroy(s.e1); }
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    for (int i = 0; i < 8; i++) {
        if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
        if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    }
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int w2x_ctx_set_engine(w2x_ctx *ctx, int engine) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    if (engine < W2X_ENGINE_AUTO || engine > W2X_ENGINE_TC) return fail(W2X_ERR_ARG, "unknown engine %d", engine);
    ctx->engine = engine;
    return W2X_OK;
}
int w2x_ctx_get_engine(const w2x_ctx *ctx) { return ctx ? ctx->engine : -1; }

int w2x_ctx_set_precision(w2x_ctx *ctx, int precision) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    if (precision != W2X_PRECISION_F16X3 && precision != W2X_PRECISION_F16_F8X2) return fail(W2X_ERR_ARG, "unknown precision mode %d", precision);
    ctx->precision = precision;
    return W2X_OK;
}
int w2x_ctx_get_precision(const w2x_ctx *ctx) { return ctx ? ctx->precision : -1; }

int w2x_ctx_set_stream(w2x_ctx *ctx, void *cuda_stream) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return W2X_OK;
}

int w2x_ctx_synchronize(w2x_ctx *ctx) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    DeviceGuard g(ctx->device);
    CU_CHECK(cudaStreamSynchronize(ctx->stream));
    return W2X_OK;
}

int w2x_ctx_set_log(w2x_ctx *ctx, w2x_log_fn fn, void *user) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    ctx->log = fn;
    ctx->log_user = user;
    return W2X_OK;
}

int w2x_ctx_set_block_walk(w2x_ctx *ctx, int mode) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    if (mode != W2X_WALK_FUSED && mode != W2X_WALK_BLOCKS) return fail(W2X_ERR_ARG, "unknown block walk mode %d", mode);
    ctx->walk = mode;
    return W2X_OK;
}

int w2x_ctx_set_scratch_limit(w2x_ctx *ctx, size_t bytes) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    ctx->scratch_limit = bytes ? bytes : (size_t)16 << 30;
    return W2X_OK;
}

// Probe hooks (not part of the stable ABI): per-role wait/work cycle counters of the tcgen05 layer kernels.
W2X_API int w2x_debug_tc_profile_enable(w2x_ctx *ctx, int on) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    DeviceGuard g(ctx->device);
    const size_t bytes = (size_t)16 * tc::PROF_MAX_CTAS * tc::PROF_WORDS * sizeof(unsigned long long);
    if (on) {
        if (!ctx->prof_buf) CU_CHECK(cudaMalloc(&ctx->prof_buf, bytes));
        CU_CHECK(cudaMemsetAsync(ctx->prof_buf, 0, bytes, ctx->stream));
    } else if (ctx->prof_buf) {
        CU_CHECK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->prof_buf);
        ctx->prof_buf = nullptr;
    }
    return W2X_OK;
}
// out[PROF_WORDS]: counters of `layer` summed over CTAs; *n_ctas = CTAs that ran.
W2X_API int w2x_debug_tc_profile_read(w2x_ctx *ctx, int layer, unsigned long long *out, int *n_ctas) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    if (!ctx->prof_buf || layer < 0 || layer >= 16 || !out) return fail(W2X_ERR_ARG, "profile not enabled or bad layer");
    DeviceGuard g(ctx->device);
    std::vector<unsigned long long> h((size_t)tc::PROF_MAX_CTAS * tc::PROF_WORDS);
    CU_CHECK(cudaStreamSynchronize(ctx->stream));
    CU_CHECK(cudaMemcpy(h.data(), ctx->prof_buf + (size_t)layer * h.size(), h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    int n = 0;
    for (int w = 0; w < tc::PROF_WORDS; w++) out[w] = 0;
    for (int c = 0; c < tc::PROF_MAX_CTAS; c++) {
        if (h[(size_t)c * tc::PROF_WORDS] == 0) continue;
        n++;
        for (int w = 0; w < tc::PROF_WORDS; w++) out[w] += h[(size_t)c * tc::PROF_WORDS + w];
    }
    if (n_ctas) *n_ctas = n;
    return W2X_OK;
}

// Probe switch (not part of the stable ABI): 1 = CTA-pair (cta_group::2) kernels for the 128-wide layers.
W2X_API int w2x_debug_set_pair(w2x_ctx *ctx, int on) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    ctx->pair = on != 0;
    return W2X_OK;
}

// Probe switch (not part of the stable ABI): 1 = row-strip kernel for the narrow layers (default), 0 = the 16x16-tile kernel.
W2X_API int w2x_debug_set_strip(w2x_ctx *ctx, int on) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    ctx->strip = on != 0;
    return W2X_OK;
}

// Probe switch (not part of the stable ABI): number of host-copy pipeline bands of w2x_convert_plane (0 auto, 1 off).
W2X_API int w2x_debug_set_host_bands(w2x_ctx *ctx, int bands) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    ctx->host_bands = bands < 0 ? 0 : bands;
    return W2X_OK;
}

// Probe switch (not part of the stable ABI): 1 = fold the last layer into the preceding tcgen05 layer (default), 0 = separate kernel.
W2X_API int w2x_debug_set_fuse_last(w2x_ctx *ctx, int on) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    ctx->fuse_last = on != 0;
    return W2X_OK;
}

int w2x_convert_plane_device(w2x_ctx *ctx, const w2x_model *model, const float *d_in, int width, int height,
                             size_t in_stride_bytes, float *d_out, size_t out_stride_bytes, int block_splitting) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    return convert_device(ctx, model, d_in, width, height, in_stride_bytes, 0, 0, d_out, out_stride_bytes, block_splitting);
}

int w2x_convert_band_device(w2x_ctx *ctx, const w2x_model *model, const float *d_in, int width, int band_height,
                            int rows_above, int rows_below, size_t in_stride_bytes, float *d_out, size_t out_stride_bytes) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    if (rows_above < 0 || rows_below < 0) return fail(W2X_ERR_ARG, "w2x_convert_band_device: negative halo");
    if (model && ((rows_above && rows_above < (int)model->layers.size()) || (rows_below && rows_below < (int)model->layers.size())))
        return fail(W2X_ERR_ARG, "w2x_convert_band_device: a halo must be 0 (image border) or >= the layer count (%zu)",
                    model->layers.size());
    if (!d_in) return fail(W2X_ERR_ARG, "w2x_convert_band_device: NULL input");
    const float *band0 = d_in + (size_t)rows_above * (in_stride_bytes / 4);
    return convert_device(ctx, model, band0, width, band_height, in_stride_bytes, rows_above, rows_below, d_out,
                          out_stride_bytes, 0);
}

int w2x_convert_plane(w2x_ctx *ctx, const w2x_model *model, const float *in, int width, int height, size_t in_stride_bytes,
                      float *out, size_t out_stride_bytes, int block_splitting) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    if (!in || !out || width < 1 || height < 1) return fail(W2X_ERR_ARG, "w2x_convert_plane: bad argument");
    if (in_stride_bytes < (size_t)width * 4 || out_stride_bytes < (size_t)width * 4)
        return fail(W2X_ERR_ARG, "w2x_convert_plane: row stride smaller than a row");
    DeviceGuard g(ctx->device);
    const size_t bytes = (size_t)width * height * sizeof(float);
    for (int i = 0; i < 2; i++) {
        int rc = ensure(reinterpret_cast<void **>(&ctx->io_buf[i]), &ctx->io_bytes[i], bytes);
        if (rc) return rc;
    }
    // Large planes are cut into row bands so that the upload of band i+1 and the download of band i-1 overlap the
    // layers of band i (copy engines + compute run concurrently); each band re-reads n real rows of context from its
    // neighbours, which keeps the result bit-identical to the single-pass path.
    const int n_model = model ? (int)model->layers.size() : 0;
    int nb = ctx->host_bands > 0 ? ctx->host_bands : std::min(4, height / 512);
    const bool literal_walk = block_splitting && ctx->walk == W2X_WALK_BLOCKS && w2x_requires_splitting(width, height);
    if (nb > 8) nb = 8;
    if (nb < 2 || literal_walk || ctx->log || !model || height / nb < 2 * n_model) {   // (a log sink wants the reference's exact line sequence)
        CU_CHECK(cudaMemcpy2DAsync(ctx->io_buf[0], (size_t)width * 4, in, in_stride_bytes, (size_t)width * 4, (size_t)height,
                                   cudaMemcpyHostToDevice, ctx->stream));
        int rc = convert_device(ctx, model, ctx->io_buf[0], width, height, (size_t)width * 4, 0, 0, ctx->io_buf[1],
                                (size_t)width * 4, block_splitting);
        if (rc) return rc;
        CU_CHECK(cudaMemcpy2DAsync(out, out_stride_bytes, ctx->io_buf[1], (size_t)width * 4, (size_t)width * 4, (size_t)height,
                                   cudaMemcpyDeviceToHost, ctx->stream));
        CU_CHECK(cudaStreamSynchronize(ctx->stream));
        return W2X_OK;
    }
    std::vector<int> r0((size_t)nb + 1);
    for (int i = 0; i <= nb; i++) r0[(size_t)i] = (int)((long)height * i / nb);
    for (int i = 0; i < nb; i++) {
        const int y = r0[(size_t)i], rows = r0[(size_t)i + 1] - y;
        CU_CHECK(cudaMemcpy2DAsync(ctx->io_buf[0] + (size_t)y * width, (size_t)width * 4, reinterpret_cast<const char *>(in) + (size_t)y * in_stride_bytes,
                                   in_stride_bytes, (size_t)width * 4, (size_t)rows, cudaMemcpyHostToDevice, ctx->copy_in));
        CU_CHECK(cudaEventRecord(ctx->ev_in[i], ctx->copy_in));
    }
    for (int i = 0; i < nb; i++) {
        const int y = r0[(size_t)i], rows = r0[(size_t)i + 1] - y;
        CU_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[std::min(i + 1, nb - 1)], 0));   // needs the first rows of the next band
        const int above = i > 0 ? n_model : 0, below = i + 1 < nb ? n_model : 0;
        int rc = convert_device(ctx, model, ctx->io_buf[0] + (size_t)y * width, width, rows, (size_t)width * 4, above, below,
                                ctx->io_buf[1] + (size_t)y * width, (size_t)width * 4, 0);
        if (rc) return rc;
        CU_CHECK(cudaEventRecord(ctx->ev_done[i], ctx->stream));
        CU_CHECK(cudaStreamWaitEvent(ctx->copy_out, ctx->ev_done[i], 0));
        CU_CHECK(cudaMemcpy2DAsync(reinterpret_cast<char *>(out) + (size_t)y * out_stride_bytes, out_stride_bytes, ctx->io_buf[1] + (size_t)y * width,
                                   (size_t)width * 4, (size_t)width * 4, (size_t)rows, cudaMemcpyDeviceToHost, ctx->copy_out));
    }
    CU_CHECK(cudaStreamSynchronize(ctx->copy_out));
    CU_CHECK(cudaStreamSynchronize(ctx->stream));
    return W2X_OK;
}

int w2x_filter_layer_device(w2x_ctx *ctx, const w2x_model *model, int layer, const float *d_in, float *d_out, int width,
                            int height) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    if (!model || layer < 0 || layer >= (int)model->layers.size() || !d_in || !d_out || width < 1 || height < 1)
        return fail(W2X_ERR_ARG, "w2x_filter_layer: bad argument");
    DeviceGuard g(ctx->device);
    DevModel *dm = nullptr;
    int rc = get_dev_model(ctx, model, &dm);
    if (rc) return rc;
    const Layer &L = model->layers[(size_t)layer];
    int engine = ctx->engine == W2X_ENGINE_AUTO ? W2X_ENGINE_FP32 : ctx->engine;
    if (engine == W2X_ENGINE_TC) {
        if (!tc::layer_supported(L.n_in, L.n_out))
            return fail(W2X_ERR_UNSUPPORTED, "tcgen05 engine does not support a %d->%d layer", L.n_in, L.n_out);
        rc = ensure_tc(ctx);
        if (rc) return rc;
        // Model::filter semantics (same size, BORDER_REPLICATE): stage a frame with a replicated ring
        // of one pixel, run the same-size tcgen05 layer on it, return the interior.
        const int pw = width + 2, ph = height + 2;
        rc = ensure(&ctx->buf[0], &ctx->buf_bytes[0], tc::act_bytes(L.n_in, pw, ph));
        if (rc) return rc;
        rc = ensure(&ctx->buf[1], &ctx->buf_bytes[1], tc::act_bytes(L.n_out, pw, ph));
        if (rc) return rc;
        __half *fin = static_cast<__half *>(ctx->buf[0]), *fout = static_cast<__half *>(ctx->buf[1]);
        const int f8 = ctx->precision == W2X_PRECISION_F16_F8X2 ? 1 : 0;
        CU_CHECK(tc::launch_planar_to_nhwc(d_in, L.n_in, width, height, fin, ctx->stream, f8));
        rc = launch_layer_tc(ctx, model, dm, layer, fin, fout, pw, ph, false, false);
        if (rc) return rc;
        CU_CHECK(tc::launch_nhwc_to_planar(fout, L.n_out, width, height, d_out, ctx->stream, f8));
        ctx->launches += 2;
        return W2X_OK;
    }
    {
        LayerTimer t(ctx, layer);
        CU_CHECK(launch_conv3x3_fp32(d_in, d_out, dm->w[(size_t)layer], dm->b[(size_t)layer], L.n_in, L.n_out, width, height,
                                     ctx->stream));
    }
    note_kernel(ctx, layer, "fp32_direct");
    ctx->launches++;
    return W2X_OK;
}

int w2x_filter_layer(w2x_ctx *ctx, const w2x_model *model, int layer, const float *const *in_planes, int n_in_planes,
                     float *const *out_planes, int n_out_planes, int width, int height, size_t in_stride_bytes,
                     size_t out_stride_bytes) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    if (!model || layer < 0 || layer >= (int)model->layers.size() || !in_planes || !out_planes || width < 1 || height < 1)
        return fail(W2X_ERR_ARG, "w2x_filter_layer: bad argument");
    const Layer &L = model->layers[(size_t)layer];
    if (n_in_planes != L.n_in)   // src/modelHandler.cpp:29-35
        return fail(W2X_ERR_ARG, "Error : Model-filter : \nnumber of input planes mismatch.\n%d,%d", n_in_planes, L.n_in);
    if (n_out_planes != L.n_out)
        return fail(W2X_ERR_ARG, "w2x_filter_layer: %d output planes supplied, layer produces %d", n_out_planes, L.n_out);
    if (in_stride_bytes < (size_t)width * 4 || out_stride_bytes < (size_t)width * 4)
        return fail(W2X_ERR_ARG, "w2x_filter_layer: row stride smaller than a row");
    DeviceGuard g(ctx->device);
    const size_t plane = (size_t)width * height * sizeof(float);
    int rc = ensure(reinterpret_cast<void **>(&ctx->io_buf[0]), &ctx->io_bytes[0], plane * (size_t)L.n_in);
    if (rc) return rc;
    rc = ensure(reinterpret_cast<void **>(&ctx->io_buf[1]), &ctx->io_bytes[1], plane * (size_t)L.n_out);
    if (rc) return rc;
    for (int i = 0; i < L.n_in; i++) {
        if (!in_planes[i]) return fail(W2X_ERR_ARG, "w2x_filter_layer: NULL input plane %d", i);
        CU_CHECK(cudaMemcpy2DAsync(reinterpret_cast<char *>(ctx->io_buf[0]) + plane * (size_t)i, (size_t)width * 4, in_planes[i],
                                   in_stride_bytes, (size_t)width * 4, (size_t)height, cudaMemcpyHostToDevice, ctx->stream));
    }
    rc = w2x_filter_layer_device(ctx, model, layer, ctx->io_buf[0], ctx->io_buf[1], width, height);
    if (rc) return rc;
    for (int i = 0; i < L.n_out; i++) {
        if (!out_planes[i]) return fail(W2X_ERR_ARG, "w2x_filter_layer: NULL output plane %d", i);
        CU_CHECK(cudaMemcpy2DAsync(out_planes[i], out_stride_bytes, reinterpret_cast<char *>(ctx->io_buf[1]) + plane * (size_t)i,
                                   (size_t)width * 4, (size_t)width * 4, (size_t)height, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU_CHECK(cudaStreamSynchronize(ctx->stream));
    return W2X_OK;
}

int w2x_ctx_launch_count(const w2x_ctx *ctx, uint64_t *n_launches) {
    if (!ctx || !n_launches) return fail(W2X_ERR_ARG, "w2x_ctx_launch_count: NULL argument");
    *n_launches = ctx->launches;
    return W2X_OK;
}

int w2x_ctx_set_timing(w2x_ctx *ctx, int enabled) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    ctx->timing = enabled != 0;
    return W2X_OK;
}

int w2x_ctx_layer_times(w2x_ctx *ctx, int max_layers, float *ms, int *launches, int *n_layers_out, int reset) {
    if (check_ctx(ctx)) return W2X_ERR_ARG;
    if (max_layers < 0 || (max_layers > 0 && (!ms || !launches))) return fail(W2X_ERR_ARG, "w2x_ctx_layer_times: bad argument");
    DeviceGuard g(ctx->device);
    CU_CHECK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < max_layers; i++) { ms[i] = 0.f; launches[i] = 0; }
    int top = 0;
    for (auto &s : ctx->spans) {
        float t = 0.f;
        cudaEventElapsedTime(&t, s.e0, s.e1);
        if (s.layer >= 0 && s.layer < max_layers) { ms[s.layer] += t; launches[s.layer]++; }
        top = std::max(top, s.layer + 1);
    }
    if (n_layers_out) *n_layers_out = top;
    if (reset) {
        for (auto &s : ctx->spans) { ctx->event_pool.push_back(s.e0); ctx->event_pool.push_back(s.e1); }
        ctx->spans.clear();
    }
    return W2X_OK;
}

const char *w2x_ctx_layer_kernel_name(const w2x_ctx *ctx, int layer) {
    if (!ctx || layer < 0 || layer >= (int)ctx->layer_kernel.size()) return "";
    return ctx->layer_kernel[(size_t)layer].c_str();
}

}  // extern "C"
