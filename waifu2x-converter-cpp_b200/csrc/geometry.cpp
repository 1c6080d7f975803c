// geometry.cpp -- process-wide configuration and the block-split index arithmetic (host only).
//
// Config mirrors the w2xc::modelUtility singleton (reference src/modelHandler.hpp:92-113,
// src/modelHandler.cpp:161-224): nJob = 4, blockSplittingSize = 512 x 512 by default.
// block_table reproduces the loop bounds of convertWithModelsBlockSplit
// (src/convertRoutine.cpp:100-131 for the input ROI, :143-155 for the output ROI), including the
// float ceil of :100-105 and the use of blockSize.height for the output column offset at :153-154.
#include <cmath>

#include "w2x_internal.h"

namespace w2x {

Config &config() {
    static Config c;
    return c;
}

int block_table(int w, int h, int bw, int bh, int n_model, int *table, int capacity, int *sc_out, int *sr_out) {
    if (w < 1 || h < 1 || n_model < 0 || bw - 2 * n_model < 1 || bh - 2 * n_model < 1) return -W2X_ERR_ARG;
    const unsigned n = (unsigned)n_model;
    const unsigned sc = static_cast<unsigned>(std::ceil(static_cast<float>(w) / static_cast<float>(bw - 2 * (int)n)));
    const unsigned sr = static_cast<unsigned>(std::ceil(static_cast<float>(h) / static_cast<float>(bh - 2 * (int)n)));
    if (sc_out) *sc_out = (int)sc;
    if (sr_out) *sr_out = (int)sr;
    const int pw = w + 2 * n_model, ph = h + 2 * n_model;
    int idx = 0;
    for (unsigned r = 0; r < sr; r++) {
        const int y0 = (int)(r * (unsigned)(bh - 2 * (int)n));
        const int y1 = (r == sr - 1) ? ph : y0 + bh;
        for (unsigned c = 0; c < sc; c++) {
            const int x0 = (int)(c * (unsigned)(bw - 2 * (int)n));
            const int x1 = (c == sc - 1) ? pw : x0 + bw;
            if (table && idx < capacity) {
                int *t = table + 8 * idx;
                t[0] = (int)r; t[1] = (int)c;
                t[2] = y0; t[3] = y1; t[4] = x0; t[5] = x1;
                t[6] = (int)(r * (unsigned)(bh - 2 * (int)n));
                t[7] = (int)(c * (unsigned)(bh - 2 * (int)n));   // blockSize.height, as the reference
            }
            idx++;
        }
    }
    return idx;
}

}  // namespace w2x

extern "C" {

int w2x_set_jobs(int n_job) {
    if (n_job < 1) return w2x::fail(W2X_ERR_ARG, "w2x_set_jobs: number of jobs must be >= 1");
    w2x::config().n_job = n_job;
    return W2X_OK;
}
int w2x_get_jobs(void) { return w2x::config().n_job; }

int w2x_set_block_size(int width, int height) {
    if (width < 0 || height < 0) return w2x::fail(W2X_ERR_ARG, "w2x_set_block_size: negative size");
    w2x::config().block_w = width;
    w2x::config().block_h = height;
    return W2X_OK;
}
int w2x_set_block_size_exp2_square(int exp) {
    if (exp < 0 || exp > 30) return w2x::fail(W2X_ERR_ARG, "w2x_set_block_size_exp2_square: bad exponent");
    int len = 1 << exp;
    w2x::config().block_w = len;
    w2x::config().block_h = len;
    return W2X_OK;
}
void w2x_get_block_size(int *width, int *height) {
    if (width) *width = w2x::config().block_w;
    if (height) *height = w2x::config().block_h;
}

int w2x_requires_splitting(int width, int height) {
    const w2x::Config &c = w2x::config();
    return (width * height) > c.block_w * c.block_h * 3 / 2 ? 1 : 0;  // int math, src/convertRoutine.cpp:25-26
}

int w2x_block_table(int width, int height, int n_model, int *table, int capacity, int *split_cols, int *split_rows) {
    const w2x::Config &c = w2x::config();
    int n = w2x::block_table(width, height, c.block_w, c.block_h, n_model, table, capacity, split_cols, split_rows);
    if (n < 0) {
        w2x::fail(W2X_ERR_ARG, "w2x_block_table: bad plane size, pad width or block size");
        return -W2X_ERR_ARG;
    }
    return n;
}

}  // extern "C"
