/*
 * w2x_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C CPU restatement of the convolution hot path of
 * WL-Amigo/waifu2x-converter-cpp, written from the reference's behaviour:
 *
 *   w2xo_filter()            <- Model::filter / Model::filterWorker
 *                               (reference src/modelHandler.cpp:26-72, :117-159)
 *   w2xo_convert_basic()     <- convertWithModelsBasic (src/convertRoutine.cpp:53-82)
 *   w2xo_convert()           <- convertWithModels / convertWithModelsBlockSplit
 *                               (src/convertRoutine.cpp:21-51, :84-169)
 *   w2xo_block_table()       <- the block index arithmetic of
 *                               src/convertRoutine.cpp:100-131, :143-155
 *   w2xo_pad_replicate()     <- cv::copyMakeBorder(..., BORDER_REPLICATE)
 *                               as called at src/convertRoutine.cpp:35, :96
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's shared object.  The product
 * library (libw2x_b200.so) never links or calls it.
 *
 * Parity status: the reference holds no golden vectors or assertions for this
 * path (its src/test.cpp has none), and it cannot be compiled here (no OpenCV
 * C++).  This restatement is therefore pinned against outputs of the
 * reference's own arithmetic backend (OpenCV, through Python cv2 running the
 * identical cv::filter2D / add / max / min / scaleAdd / copyMakeBorder calls,
 * see oracle/ref_cv2.py) committed under tests/golden/ -- AND against the
 * reference's own sources: src/modelHandler.cpp and src/convertRoutine.cpp compile
 * unmodified against the OpenCV API shim in oracle/cvshim into
 * oracle/_ref/libw2x_reference.so (oracle/Makefile); this restatement is bit-identical
 * to that library on every path (tests/test_reference_build.py).
 *
 * Arithmetic follows the reference op for op in fp32:
 *   per (o,i): tmp = sum over the 9 taps, row-major (ky,kx), starting from 0
 *              (cv::filter2D, correlation, anchor centre, BORDER_REPLICATE)
 *   acc += tmp           (i ascending, cv::add)
 *   acc += (float)bias   (cv::add with a double scalar on a CV_32F plane)
 *   out = min(acc,0) * 0.1f + max(acc,0)     (cv::max/min/scaleAdd)
 * Compile with -ffp-contract=off so no FMA contraction changes the order.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n_in;         /* nInputPlane  */
    int n_out;        /* nOutputPlane */
    int k;            /* kW == kH (3) */
    const float *w;   /* [n_out][n_in][k][k], already rounded double->float
                         (reference src/modelHandler.cpp:96-97) */
    const double *b;  /* [n_out], kept double (src/modelHandler.hpp:30) */
} w2xo_layer;

/* ---- cv::copyMakeBorder(BORDER_REPLICATE) ------------------------------ */
void w2xo_pad_replicate(const float *in, int w, int h, long in_stride /*floats*/,
                        int pad, float *out /* (h+2p) x (w+2p), dense */)
{
    int W = w + 2 * pad, H = h + 2 * pad;
    for (int y = 0; y < H; y++) {
        int sy = y - pad; if (sy < 0) sy = 0; if (sy > h - 1) sy = h - 1;
        const float *src = in + (long)sy * in_stride;
        float *dst = out + (long)y * W;
        for (int x = 0; x < W; x++) {
            int sx = x - pad; if (sx < 0) sx = 0; if (sx > w - 1) sx = w - 1;
            dst[x] = src[sx];
        }
    }
}

/* ---- Model::filterWorker ------------------------------------------------ */
typedef struct {
    const w2xo_layer *L;
    const float *in_pad;   /* n_in planes, each (h+2) x (w+2), replicate-padded by 1 */
    float *out;            /* n_out planes, each h x w dense */
    int w, h;
    int begin, count;      /* output planes [begin, begin+count) */
} worker_arg;

static void *filter_worker(void *p)
{
    worker_arg *a = (worker_arg *)p;
    const w2xo_layer *L = a->L;
    const int w = a->w, h = a->h, pw = w + 2;
    const long plane = (long)w * h, pplane = (long)pw * (h + 2);
    float *tmp = (float *)malloc(sizeof(float) * (size_t)w);
    for (int op = a->begin; op < a->begin + a->count; op++) {
        float *acc = a->out + plane * op;
        memset(acc, 0, sizeof(float) * (size_t)plane);              /* Mat::zeros, :131 */
        for (int ip = 0; ip < L->n_in; ip++) {                      /* :134 */
            const float *wk = L->w + ((long)op * L->n_in + ip) * 9;
            const float *src = a->in_pad + pplane * ip;
            for (int y = 0; y < h; y++) {
                const float *r0 = src + (long)y * pw, *r1 = r0 + pw, *r2 = r1 + pw;
                float *arow = acc + (long)y * w;
                for (int x = 0; x < w; x++) {                       /* filter2D, :141 */
                    float s = 0.0f;
                    s += wk[0] * r0[x]; s += wk[1] * r0[x + 1]; s += wk[2] * r0[x + 2];
                    s += wk[3] * r1[x]; s += wk[4] * r1[x + 1]; s += wk[5] * r1[x + 2];
                    s += wk[6] * r2[x]; s += wk[7] * r2[x + 1]; s += wk[8] * r2[x + 2];
                    tmp[x] = s;
                }
                for (int x = 0; x < w; x++) arow[x] = arow[x] + tmp[x];  /* cv::add, :144 */
            }
        }
        const float bias = (float)L->b[op];                         /* :147 */
        const float slope = 0.1f;                                   /* :152 */
        for (long i = 0; i < plane; i++) {
            float v = acc[i] + bias;
            float pos = v > 0.0f ? v : 0.0f;                        /* cv::max :150 */
            float neg = v < 0.0f ? v : 0.0f;                        /* cv::min :151 */
            acc[i] = neg * slope + pos;                             /* cv::scaleAdd :152 */
        }
    }
    free(tmp);
    return NULL;
}

/* ---- Model::filter -------------------------------------------------------
 * in:  n_in planes h x w, dense planar.  out: n_out planes h x w, dense planar.
 * Output planes are partitioned over n_job threads exactly like
 * src/modelHandler.cpp:42-65 (the last thread takes the remainder).
 * Returns 0 on success, -1 on bad arguments (the reference returns false). */
int w2xo_filter(const w2xo_layer *L, const float *in, float *out, int w, int h, int n_job)
{
    if (!L || !in || !out || w <= 0 || h <= 0 || n_job < 1 || L->k != 3) return -1;
    const int pw = w + 2, ph = h + 2;
    float *in_pad = (float *)malloc(sizeof(float) * (size_t)pw * ph * L->n_in);
    if (!in_pad) return -1;
    for (int ip = 0; ip < L->n_in; ip++)
        w2xo_pad_replicate(in + (long)w * h * ip, w, h, w, 1, in_pad + (long)pw * ph * ip);

    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_job);
    worker_arg *args = (worker_arg *)malloc(sizeof(worker_arg) * (size_t)n_job);
    int wpt = L->n_out / n_job;
    for (int idx = 0; idx < n_job; idx++) {
        args[idx].L = L; args[idx].in_pad = in_pad; args[idx].out = out;
        args[idx].w = w; args[idx].h = h;
        args[idx].begin = wpt * idx;
        if (!(idx == n_job - 1 && wpt * n_job != L->n_out)) args[idx].count = wpt;
        else args[idx].count = L->n_out - wpt * idx;
        pthread_create(&th[idx], NULL, filter_worker, &args[idx]);
    }
    for (int idx = 0; idx < n_job; idx++) pthread_join(th[idx], NULL);
    free(th); free(args); free(in_pad);
    return 0;
}

/* ---- convertWithModelsBasic ---------------------------------------------
 * in: one plane h x w with row stride in_stride (floats) -- the reference
 * hands a non-contiguous ROI here in the block-split path.  out: dense h x w. */
int w2xo_convert_basic(const w2xo_layer *layers, int n_layers, const float *in, int w, int h,
                       long in_stride, float *out, int n_job)
{
    if (n_layers < 1 || layers[0].n_in != 1 || layers[n_layers - 1].n_out != 1) return -1;
    const long plane = (long)w * h;
    float *cur = (float *)malloc(sizeof(float) * (size_t)plane);
    if (!cur) return -1;
    for (int y = 0; y < h; y++) memcpy(cur + (long)y * w, in + (long)y * in_stride, sizeof(float) * (size_t)w);
    for (int li = 0; li < n_layers; li++) {
        if (li > 0 && layers[li].n_in != layers[li - 1].n_out) { free(cur); return -1; }
        float *nxt = (float *)malloc(sizeof(float) * (size_t)plane * layers[li].n_out);
        if (!nxt) { free(cur); return -1; }
        if (w2xo_filter(&layers[li], cur, nxt, w, h, n_job) != 0) { free(cur); free(nxt); return -1; }
        free(cur);
        cur = nxt;
    }
    memcpy(out, cur, sizeof(float) * (size_t)plane);
    free(cur);
    return 0;
}

/* ---- block geometry of convertWithModelsBlockSplit -----------------------
 * Fills rows of 8 ints per block, row-major (r outer, c inner):
 *   { r, c, in_y0, in_y1, in_x0, in_x1, out_y0, out_x0 }
 * in_* index the pad-n_model plane ((h+2n) x (w+2n)); out_* index the h x w
 * output; the written region is (in_y1-in_y0-2n) x (in_x1-in_x0-2n).
 * Returns the number of blocks (splitRows*splitColumns); table may be NULL.
 * NOTE: the reference computes the output COLUMN offset with blockSize.height
 * (src/convertRoutine.cpp:153-154); reproduced here verbatim. */
int w2xo_block_table(int w, int h, int bw, int bh, int n_model, int *table, int *split_cols,
                     int *split_rows)
{
    unsigned n = (unsigned)n_model;
    unsigned sc = (unsigned)ceilf((float)w / (float)(bw - 2 * (int)n));   /* :100-102 */
    unsigned sr = (unsigned)ceilf((float)h / (float)(bh - 2 * (int)n));   /* :103-105 */
    if (split_cols) *split_cols = (int)sc;
    if (split_rows) *split_rows = (int)sr;
    int pw = w + 2 * n_model, ph = h + 2 * n_model;
    int idx = 0;
    for (unsigned r = 0; r < sr; r++) {
        int y0 = (int)(r * (bh - 2 * n));
        int y1 = (r == sr - 1) ? ph : y0 + bh;                            /* :115-121 */
        for (unsigned c = 0; c < sc; c++) {
            int x0 = (int)(c * (bw - 2 * n));
            int x1 = (c == sc - 1) ? pw : x0 + bw;                        /* :123-131 */
            if (table) {
                int *t = table + 8 * idx;
                t[0] = (int)r; t[1] = (int)c; t[2] = y0; t[3] = y1; t[4] = x0; t[5] = x1;
                t[6] = (int)(r * (bh - 2 * n));                           /* :150 */
                t[7] = (int)(c * (bh - 2 * n));                           /* :153 (height!) */
            }
            idx++;
        }
    }
    return idx;
}

/* ---- convertWithModels ----------------------------------------------------
 * block_splitting: the reference's 4th argument.  bw,bh: modelUtility block
 * size (512x512 default, src/modelHandler.hpp:99).  out: dense h x w. */
int w2xo_convert(const w2xo_layer *layers, int n_layers, const float *in, int w, int h,
                 long in_stride, float *out, int block_splitting, int bw, int bh, int n_job)
{
    if (!layers || n_layers < 1 || !in || !out || w <= 0 || h <= 0) return -1;
    const int n = n_layers;
    const int pw = w + 2 * n, ph = h + 2 * n;
    int require = (w * h) > bw * bh * 3 / 2;                              /* :25-26 int math */
    float *pad = (float *)malloc(sizeof(float) * (size_t)pw * ph);
    if (!pad) return -1;
    w2xo_pad_replicate(in, w, h, in_stride, n, pad);                      /* :35 / :96 */
    int rc = 0;
    if (block_splitting && require) {
        int nb = w2xo_block_table(w, h, bw, bh, n, NULL, NULL, NULL);
        int *tab = (int *)malloc(sizeof(int) * 8 * (size_t)nb);
        w2xo_block_table(w, h, bw, bh, n, tab, NULL, NULL);
        memset(out, 0, sizeof(float) * (size_t)w * h);                    /* :113 */
        for (int i = 0; i < nb && rc == 0; i++) {
            const int *t = tab + 8 * i;
            int bh_i = t[3] - t[2], bw_i = t[5] - t[4];
            float *bo = (float *)malloc(sizeof(float) * (size_t)bw_i * bh_i);
            rc = w2xo_convert_basic(layers, n_layers, pad + (long)t[2] * pw + t[4], bw_i, bh_i,
                                    pw, bo, n_job);
            if (rc == 0) {
                int oh = bh_i - 2 * n, ow = bw_i - 2 * n;
                if (t[6] + oh > h || t[7] + ow > w) rc = -2;              /* assert :157-160 */
                for (int y = 0; y < oh && rc == 0; y++)
                    memcpy(out + (long)(t[6] + y) * w + t[7], bo + (long)(y + n) * bw_i + n,
                           sizeof(float) * (size_t)ow);
            }
            free(bo);
        }
        free(tab);
    } else {
        float *full = (float *)malloc(sizeof(float) * (size_t)pw * ph);
        rc = w2xo_convert_basic(layers, n_layers, pad, pw, ph, pw, full, n_job);
        if (rc == 0)
            for (int y = 0; y < h; y++)                                   /* crop :40-46 */
                memcpy(out + (long)y * w, full + (long)(y + n) * pw + n, sizeof(float) * (size_t)w);
        free(full);
    }
    free(pad);
    return rc;
}
