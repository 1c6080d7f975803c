// tc_strip_kernel.cuh -- tc_conv3x3_strip_kernel: the narrow layers (Cin, Cout <= 64) as row strips with the three ky taps
// stacked along N
// Part of the tcgen05 engine's single translation unit: included by kernels_tc.cu inside namespace w2x::tc, in this order:
//   tc_ptx.cuh, tc_config.cuh, tc_issue.cuh, tc_epilogue.cuh, tc_kernel.cuh, tc_pair_kernel.cuh, tc_strip_kernel.cuh, tc_edge_kernels.cuh
//
// Why: an M128 x N x K16 tcgen05.mma fetches (128 + N) operand rows of 32 B from shared memory at 128 B/clk whatever N is
// (profiles/r01_umma_microbench.txt), while its math takes N/2 clocks.  With N = Cout = 32 / 64 the fetch (40 / 48 clk)
// hides the math (16 / 32 clk): the 16x16-tile kernel ran L1..L3 at 32 / 59 / 64 % tensor-pipe activity with the
// shared-memory pipe at 80..95 %.  Here one A fetch feeds THREE taps:
//
//   * an M-tile is 128 consecutive pixels of ONE frame row r (a "strip"); tap kx is the descriptor start offset
//     kx * ROWB into the staged row (130 pixels: the strip plus one pixel either side);
//   * the B operand of tap column kx is [W(ky=0,kx) ; W(ky=1,kx) ; W(ky=2,kx)], N = 3 * Cout rows: input row r
//     contributes W(ky) to output row r + 1 - ky, so the three N-blocks of one MMA belong to three different OUTPUT rows;
//   * output rows own TMEM column blocks of Cout columns laid out in DESCENDING row order, block(n) = NB-1 - (n mod NB):
//     the blocks of rows r+1, r, r-1 are then adjacent ascending columns and ONE N = 3*Cout MMA accumulates into all
//     three.  The tensor core sums the ky taps; nothing is added in the epilogue.  Where the ring wraps (2 of NB strips)
//     the MMA is issued as two (N = Cout + 2*Cout).
//   * every accumulate flag is 1: the epilogue re-zeroes a block (tcgen05.st) right after draining it.
//
// Per MMA: N = 96 -> fetch 56 clk, math 48;  N = 192 -> fetch 80, math 96 (math-bound).  Each input row is staged once per
// 128-pixel column (130/128 L2->SMEM amplification instead of 1.27x), all weights stay resident in shared memory
// (36..144 KB), the frame is walked in units of `seg_rows` rows x 128 columns, static round-robin over the CTAs.
//
// Warps: 0 = activation producer (TMA), 1 = MMA issuer (software-pipelined: the next strip's plan, tc_strip_plan.h, is computed
// between the current strip's tap groups), 2 = weights + TMEM owner, 4.. = one or two epilogue sets (alternate output rows).
// The 64 -> 64 shape, whose bound was the shared-memory pipe, stores its records straight from registers (StripCfg::DIRECT).
//
// Every output element still sees the same operations in the same order whatever the unit/strip geometry:
//   for ky (= strips r-1, r, r+1): for 32-channel chunk c: for kx: [xh*wh k0, xh*wh k1, corrections]
// so block-split, whole-plane, banded and multi-GPU runs stay bit-identical to each other.

#include "tc_strip_plan.h"              // the per-strip schedule (plain integer arithmetic, also compiled by a CPU test)

constexpr int STRIP_W = 128;            // pixels per strip = GEMM M
constexpr int STRIP_BOXW = STRIP_W + 2;  // staged pixels per row

// Input frames are RECORD frames (tc_epilogue.cuh): one TMA box {128 B, 1 block, 130 px, 1 row} per (row, 32-channel block)
// lands as 130 rows of 128 B in the SWIZZLE_128B pattern; the fp16 K steps, the xh8 and the xl8 slices of a pixel are the
// 32-byte quarters of its row.  The output frame is a RECORD frame as well.
template <int CIN, int COUT, bool F8>
struct StripCfg {
    static_assert((CIN == 32 || CIN == 64) && (COUT == 32 || COUT == 64), "strip kernel: narrow layers only");
    static constexpr int NCH = CIN / 32;                                     // 32-channel chunks
    static constexpr int ROWB = 128;                                         // bytes per pixel per 32-channel block (one record)
    static constexpr int A_TX = STRIP_BOXW * ROWB;                           // bytes one TMA box delivers
    static constexpr int A_SLOT = (A_TX + 1023) / 1024 * 1024;               // SWIZZLE_128B repeats every 1024 B
    static constexpr int NROWS = 3 * COUT;                                   // B rows of one stage: ky-major
    static constexpr int W_STAGE = NROWS * 128;                              // per (chunk, kx): [wh 64 B rows | wh8 | wl8 32 B rows] or [wh | wl]
    static constexpr int W_BYTES = NCH * 3 * W_STAGE;
    static constexpr int NB = 512 / COUT;                                    // accumulator blocks (output rows in flight) in TMEM
    static constexpr int SMEM_MAX = 227 * 1024;
    static constexpr int BAR_BYTES = 1024;
    static constexpr int STG_WARP = 4096;
    // 64 -> 64 (144 KB of weights, bound by the shared-memory pipe): the epilogue stores its records straight from registers
    // (32-byte global stores) -- no staging tiles, so two epilogue sets and a fourth activation slot fit
#ifndef W2X_STRIP_DIRECT
#define W2X_STRIP_DIRECT 1              // -DW2X_STRIP_DIRECT=0: staging tile + TMA store for every shape (A/B timing builds)
#endif
    static constexpr bool DIRECT = W2X_STRIP_DIRECT && CIN == 64 && COUT == 64;
    // two epilogue warp sets (alternate output rows) when the resident weights leave room for their staging tiles
    static constexpr int EPI_SETS = DIRECT ? 2 : (SMEM_MAX - 1024 - BAR_BYTES - W_BYTES - 8 * STG_WARP) / A_SLOT >= 3 ? 2 : 1;
    static constexpr int STG_BYTES = DIRECT ? 0 : EPI_SETS * 4 * STG_WARP;
    static constexpr int A_FIT = (SMEM_MAX - 1024 - BAR_BYTES - W_BYTES - STG_BYTES) / A_SLOT;
    static constexpr int A_SLOTS = A_FIT > 6 ? 6 : A_FIT;
    static constexpr int SMEM_BYTES = 1024 + W_BYTES + A_SLOTS * A_SLOT + STG_BYTES + BAR_BYTES;
    static constexpr int THREADS = (4 + 4 * EPI_SETS) * 32;                  // warps: 0 A producer | 1 MMA issuer | 2 weights + TMEM | 3 idle | 4.. epilogue
    static_assert(A_SLOTS >= 3, "need at least three staged rows");
    static_assert((1 + 2 * A_SLOTS + 2 * NB) * 8 + 4 <= BAR_BYTES, "barrier area overflow");
    static_assert(W_STAGE % 1024 == 0 && A_SLOT % 1024 == 0 && (NROWS * 64) % 256 == 0, "swizzle pattern alignment");
    static_assert(NB % 2 == 0, "block ownership alternates between the epilogue sets");
};

struct StripParams {
    const uint8_t *wpack;       // [chunk][kx] stages, exact shared-memory images (model.cpp: pack_tc_layer_strip)
    float bias[64];             // (float)bias * ACT_SCALE
    int Wp, Hp;
    int out_y0, out_rows;       // only frame rows [out_y0, out_y0 + out_rows) are stored
    int ncols, n_units, seg_rows;   // units = 128-pixel columns x segments of seg_rows rows, unit u = seg * ncols + col
    float out_scale;
    unsigned long long *prof;
    uint8_t *out_win;           // first stored row of the output frame (direct stores: StripCfg::DIRECT, or experiments 128 / 256 = 16- / 32-byte stores, results correct)
    int dbg;                    // always 0 in product builds; -DW2X_EPI_EXPERIMENTS + W2X_DEBUG_STRIP (timing only, results WRONG):
                                // 1 = no TMA stores, 2 = no staging either, 4 = no activation loads, 8 = no MMAs,
                                // 16 = the issuer neither waits for nor probes a barrier, 32 = the epilogue does not touch TMEM,
                                // 64 = the epilogue does not zero the blocks
};

template <int CIN, int COUT, bool F8>
__global__ void __launch_bounds__(StripCfg<CIN, COUT, F8>::THREADS, 1)
tc_conv3x3_strip_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_out, const StripParams p) {
    using C = StripCfg<CIN, COUT, F8>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_base = smem_base;
    const uint32_t a_base = w_base + C::W_BYTES;
    const uint32_t stg_base = a_base + C::A_SLOTS * C::A_SLOT;
    const uint32_t bar_base = stg_base + C::STG_BYTES;
    const uint32_t w_full = bar_base;
    auto a_full = [&](uint32_t i) { return bar_base + 8u * (1u + i); };
    auto a_empty = [&](uint32_t i) { return bar_base + 8u * (1u + C::A_SLOTS + i); };
    auto blk_full = [&](uint32_t i) { return bar_base + 8u * (1u + 2u * C::A_SLOTS + i); };
    auto blk_empty = [&](uint32_t i) { return bar_base + 8u * (1u + 2u * C::A_SLOTS + C::NB + i); };
    const uint32_t tmem_slot = bar_base + 8u * (1u + 2u * C::A_SLOTS + 2u * C::NB);
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const bool prof_on = p.prof != nullptr;
    unsigned long long *prof = prof_on ? p.prof + (size_t)blockIdx.x * PROF_N : nullptr;

    if (threadIdx.x == 0) {
        mbar_init(w_full, 1);
        for (uint32_t i = 0; i < (uint32_t)C::A_SLOTS; i++) {
            mbar_init(a_full(i), 1);
            mbar_init(a_empty(i), 1);
        }
        for (uint32_t i = 0; i < (uint32_t)C::NB; i++) {
            mbar_init(blk_full(i), 1);
            mbar_init(blk_empty(i), 4);   // the four lane-quarter warps of the owning epilogue set
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_in);
        prefetch_tmap(&tmap_out);
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);

    // unit u -> frame rows [y0, y1), column col
    auto unit_of = [&](int u, int &col, int &y0, int &y1) {
        const int seg = u / p.ncols;
        col = u - seg * p.ncols;
        y0 = seg * p.seg_rows;
        y1 = min(y0 + p.seg_rows, p.Hp);
    };
    constexpr uint32_t NB = (uint32_t)C::NB;
    auto blk_of = [](uint32_t n) { return NB - 1u - (n % NB); };   // descending: rows n, n+1 own adjacent blocks b, b-1

    if (warp == 0) {
        // ===================== A producer: one staged row (130 px x 32 ch, all planes) per (strip, chunk) =====================
        uint32_t it = 0;
        unsigned long long w_a = 0;
        for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
            int col, y0, y1;
            unit_of(u, col, y0, y1);
            const int r_first = max(y0 - 1, 0), r_last = min(y1, p.Hp - 1);
            const int x0 = col * STRIP_W - 1;
            for (int r = r_first; r <= r_last; r++) {
                for (int c = 0; c < C::NCH; c++, it++) {
                    const uint32_t slot = it % (uint32_t)C::A_SLOTS, round = it / (uint32_t)C::A_SLOTS;
                    mbar_wait_prof(a_empty(slot), (round & 1u) ^ 1u, prof_on, w_a);
                    if (p.dbg & 4) {
                        if (lane == 0) mbar_arrive(a_full(slot));
                        continue;
                    }
                    mbar_arrive_expect_tx(a_full(slot), (uint32_t)C::A_TX);
                    tma_load_4d(a_base + slot * C::A_SLOT, &tmap_in, a_full(slot), 0, c, x0, r);
                }
            }
        }
        if (prof_on && lane == 0) prof[PROF_APROD_WAIT] += w_a;
    } else if (warp == 2) {
        // ===================== weights: every stage once, resident for the whole launch =====================================
        mbar_arrive_expect_tx(w_full, (uint32_t)C::W_BYTES);
        for (int s = 0; s < C::NCH * 3; s++)
            bulk_load(w_base + (uint32_t)s * C::W_STAGE, p.wpack + (size_t)s * C::W_STAGE, C::W_STAGE, w_full);
    } else if (warp == 1) {
        // ===================== MMA issuer (one converged warp, an elected lane issues) =====================================
        // The tensor queue hides only about one MMA of issuer time (profiles/r02_strip_issuer_experiment.txt: the old loop's ~600 cycles of
        // index arithmetic between two strips were ~500 idle tensor cycles per strip), so the loop is software-pipelined: the
        // NEXT strip's plan (tc_strip_plan.h) and its barrier probes are computed in two slices BETWEEN the tap-column groups
        // of the current strip's last chunk; between the last MMA of a strip and the first of the next there are only the
        // commits and the (normally already satisfied) waits.
        constexpr uint32_t LO_FIXED = 1u << 16;
        constexpr uint32_t A_HI32 = (uint32_t)(make_desc_const(8 * C::ROWB, 2u) >> 32);     // SWIZZLE_128B, 8-pixel groups are contiguous: SBO = 1024 B
        constexpr uint32_t B_HI32 = (uint32_t)(make_desc_const(8 * 64, 4u) >> 32);
        constexpr uint32_t B8_HI32 = (uint32_t)(make_desc_const(8 * 32, 6u) >> 32);
        constexpr uint32_t A_STEP = (uint32_t)C::A_SLOT >> 4, W_STEP = (uint32_t)C::W_STAGE >> 4, KX_STEP = (uint32_t)C::ROWB >> 4;
        constexpr uint32_t B8H_OFF = (uint32_t)C::NROWS * 4u, B8L_OFF = (uint32_t)C::NROWS * 6u;   // e4m3 / lo rows behind the fp16 rows [16-byte units]
        constexpr uint32_t IDESC_0 = make_idesc(128, 0), IDESC_BLK = (uint32_t)(COUT >> 3) << 17;   // + one N-block
        constexpr uint32_t LOG_NB = NB == 16 ? 4u : 3u;
        static_assert(NB == 8 || NB == 16, "ring size");
        auto desc = [](uint32_t hi32, uint32_t lo32) { return ((uint64_t)hi32 << 32) | (uint64_t)lo32; };
        // shared-memory addresses are below 256 KB: descriptor start fields (addr >> 4, 14 bits) add without carries
        const uint32_t aa0 = ((a_base >> 4) & 0x3FFFu) | LO_FIXED, bb0 = ((w_base >> 4) & 0x3FFFu) | LO_FIXED;
        auto blk_bar_empty = [&](uint32_t n) { return blk_empty(NB - 1u - (n & (NB - 1u))); };
        auto blk_par = [&](uint32_t n) { return (n >> LOG_NB) & 1u; };
        const bool waits = !(p.dbg & 16), mmas = !(p.dbg & 8);

        uint32_t n_strips = 0;
        unsigned long long w_acc = 0, w_af = 0, w_bf = 0;
        const long long t_begin = clock64();
        mbar_wait_prof(w_full, 0u, prof_on, w_bf);
        tc_fence_after();

        int u = blockIdx.x;
        if (u < p.n_units) {
            // ---- walk state: unit (rows, nbase), strip j; activation slot ring ----
            int col, y0, y1;
            unit_of(u, col, y0, y1);
            int rows = y1 - y0, j = strip_j_first(y0), j_last = strip_j_last(y1, rows, p.Hp);
            uint32_t nbase = 0;
            uint32_t slot = 0, apar = 0;                      // a_full parity of the slot ring's current round
            StripPlan P = strip_plan(j, true, j == j_last, rows, nbase, NB);
            uint32_t a_ready = 0, new_ready = 0;
            // one strip's operands: D blocks, instruction descriptors, B row offsets of the two runs [16-byte units], rows to hand over
            struct Ops { uint32_t d0, id0, id1, bq0, bq1, run1, acq_n, acq_cnt, com_n, com_cnt; };
            auto ops_of = [&](const StripPlan &Q) {
                Ops o;
                o.d0 = tmem_base + Q.b0 * COUT;
                o.id0 = IDESC_0 + Q.cnt0 * IDESC_BLK;
                o.id1 = IDESC_0 + Q.cnt1 * IDESC_BLK;
                o.bq0 = Q.ky_lo * (uint32_t)(COUT * 4);
                o.bq1 = o.bq0 + Q.cnt0 * (uint32_t)(COUT * 4);
                o.run1 = Q.cnt1;
                o.acq_n = Q.acq_n, o.acq_cnt = Q.acq_cnt, o.com_n = Q.com_n, o.com_cnt = Q.com_cnt;
                return o;
            };
            Ops cur = ops_of(P), nxt = cur;
            uint32_t aa = aa0;                                // descriptor start field of the current activation slot
            for (;;) {
                // ---- blocks that receive their first tap (at most two): drained + zeroed?  (probed during the previous strip) ----
                if (waits) {
                    if (cur.acq_cnt > 0 && !new_ready) mbar_wait_prof(blk_bar_empty(cur.acq_n), blk_par(cur.acq_n), prof_on, w_acc);
                    if (cur.acq_cnt > 1) mbar_wait_prof(blk_bar_empty(cur.acq_n + 1u), blk_par(cur.acq_n + 1u), prof_on, w_acc);
                }
                bool more = true;
                n_strips++;
#pragma unroll
                for (int c = 0; c < C::NCH; c++) {
                    if (waits && !a_ready) mbar_wait_prof(a_full(slot), apar, prof_on, w_af);
                    tc_fence_after();
                    uint32_t slot_n = 0, apar_n = 0;
#pragma unroll
                    for (int kx = 0; kx < 3; kx++) {
                        if (mmas) {
                            const uint32_t ah = aa + (uint32_t)kx * KX_STEP, sb = bb0 + (uint32_t)(c * 3 + kx) * W_STEP;
#pragma unroll
                            for (int s = 0; s < 2; s++) {
                                if (s == 1 && !cur.run1) continue;
                                const uint32_t d = s ? tmem_base : cur.d0, idesc = s ? cur.id1 : cur.id0, bq = s ? cur.bq1 : cur.bq0;
                                const uint32_t bh = sb + bq;
                                umma_f16(d, desc(A_HI32, ah), desc(B_HI32, bh), idesc, 1u);
                                umma_f16(d, desc(A_HI32, ah + 2u), desc(B_HI32, bh + 2u), idesc, 1u);
                                // the record's quarters: +0 / +2 the fp16 K steps, +4 xh8 (or lo step 0), +6 xl8 (or lo step 1)   [16-byte units]
                                if constexpr (F8) {
                                    const uint32_t b8h = sb + B8H_OFF + (bq >> 1), b8l = sb + B8L_OFF + (bq >> 1);
                                    umma_f8(d, desc(A_HI32, ah + 6u), desc(B8_HI32, b8h), idesc, 1u);    // xl8 * wh8
                                    umma_f8(d, desc(A_HI32, ah + 4u), desc(B8_HI32, b8l), idesc, 1u);    // xh8 * wl8
                                } else {
                                    const uint32_t bl = sb + B8H_OFF + bq;
                                    umma_f16(d, desc(A_HI32, ah + 4u), desc(B_HI32, bh), idesc, 1u);     // xl * wh
                                    umma_f16(d, desc(A_HI32, ah + 6u), desc(B_HI32, bh + 2u), idesc, 1u);
                                    umma_f16(d, desc(A_HI32, ah), desc(B_HI32, bl), idesc, 1u);          // xh * wl
                                    umma_f16(d, desc(A_HI32, ah + 2u), desc(B_HI32, bl + 2u), idesc, 1u);
                                }
                            }
                        }
                        // ---- slices of the NEXT strip's bookkeeping, placed behind tap columns 0 and 1 ----
                        if (kx == 0) {
                            slot_n = slot + 1u == (uint32_t)C::A_SLOTS ? 0u : slot + 1u;
                            apar_n = slot_n == 0u ? apar ^ 1u : apar;
                            if (waits) a_ready = mbar_test(a_full(slot_n), apar_n);        // the next chunk's (or strip's) activations
                            if (c == C::NCH - 1) {
                                if (j < j_last) {
                                    j++;
                                    P = strip_plan(j, false, j == j_last, rows, nbase, NB);
                                } else {
                                    u += gridDim.x;
                                    more = u < p.n_units;
                                    if (more) {
                                        nbase += (uint32_t)rows;
                                        unit_of(u, col, y0, y1);
                                        rows = y1 - y0;
                                        j = strip_j_first(y0);
                                        j_last = strip_j_last(y1, rows, p.Hp);
                                        P = strip_plan(j, true, j == j_last, rows, nbase, NB);
                                    }
                                }
                                asm volatile("" : "+r"(P.b0), "+r"(P.cnt0), "+r"(P.cnt1), "+r"(P.ky_lo), "+r"(P.acq_n), "+r"(P.acq_cnt));   // keep the slice here
                            }
                        } else if (kx == 1 && c == C::NCH - 1) {
                            nxt = ops_of(P);
                            new_ready = 0;
                            if (waits && more && nxt.acq_cnt) new_ready = mbar_test(blk_bar_empty(nxt.acq_n), blk_par(nxt.acq_n));
                            asm volatile("" : "+r"(nxt.d0), "+r"(nxt.id0), "+r"(nxt.id1), "+r"(nxt.bq0), "+r"(nxt.bq1), "+r"(nxt.run1));
                        }
                    }
                    umma_commit_one(a_empty(slot));
                    slot = slot_n;
                    apar = apar_n;
                    aa = aa0 + slot * A_STEP;
                }
                if (cur.com_cnt > 0) umma_commit_one(blk_full(NB - 1u - (cur.com_n & (NB - 1u))));
                if (cur.com_cnt > 1) umma_commit_one(blk_full(NB - 1u - ((cur.com_n + 1u) & (NB - 1u))));
                if (!more) break;
                cur = nxt;
            }
        }
        if (prof_on && lane == 0) {
            prof[PROF_TOTAL] += (unsigned long long)(clock64() - t_begin);
            prof[PROF_MMA_WAIT_ACC] += w_acc;
            prof[PROF_MMA_WAIT_A] += w_af;
            prof[PROF_MMA_WAIT_B] += w_bf;
            prof[PROF_TILESETS] += n_strips;
        }
    } else if (warp >= 4) {
        // ===================== epilogue: set e drains output rows n with n % EPI_SETS == e ===================================
        const uint32_t q = (uint32_t)warp & 3u;                  // TMEM lane quarter = 32 pixels of the strip
        const uint32_t e = ((uint32_t)warp - 4u) >> 2;
        const uint32_t stg = stg_base + (e * 4u + q) * (uint32_t)C::STG_WARP;
        const uint32_t lane_base = tmem_base + ((q * 32u) << 16);
        unsigned long long w_e = 0, work_e = 0;
        // all accumulate flags are 1: hand every block over zeroed
        for (uint32_t n = e; n < NB; n += (uint32_t)C::EPI_SETS) {
#pragma unroll
            for (int cb = 0; cb < COUT / 32; cb++) tmem_st32_zero(lane_base + blk_of(n) * COUT + (uint32_t)cb * 32u);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0)
            for (uint32_t n = e; n < NB; n += (uint32_t)C::EPI_SETS) mbar_arrive(blk_empty(blk_of(n)));
        uint32_t nrow = 0;
        for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
            int col, y0, y1;
            unit_of(u, col, y0, y1);
            const int rows = y1 - y0;
            const int gx0 = col * STRIP_W + (int)q * 32;
            for (int i = 0; i < rows; i++) {
                const uint32_t n = nrow + (uint32_t)i;
                if (n % (uint32_t)C::EPI_SETS != e) continue;
                const uint32_t blk = blk_of(n);
                mbar_wait_prof(blk_full(blk), (n / NB) & 1u, prof_on, w_e);
                const long long t_work = prof_on ? clock64() : 0;
                tc_fence_after();
                const uint32_t tcol = lane_base + blk * COUT;
                uint32_t r[32];
                if (p.dbg & 32) {
#pragma unroll
                    for (int k = 0; k < 32; k++) r[k] = (uint32_t)(lane + k);
                } else tmem_ld32(tcol, r);
#pragma unroll
                for (int cb = 0; cb < COUT / 32; cb++) {
                    float act[32];
                    if (!(p.dbg & 32)) tmem_ld_wait_dep(r);
#pragma unroll
                    for (int k = 0; k < 32; k++) act[k] = __uint_as_float(r[k]);
                    if (cb + 1 < COUT / 32) {
                        if (!(p.dbg & 32)) tmem_ld32(tcol + (uint32_t)(cb + 1) * 32u, r);
                    } else {   // the row is in registers: zero its block and hand it back before the last conversion
                        if (!(p.dbg & (32 | 64))) {
#pragma unroll
                            for (int z = 0; z < COUT / 32; z++) tmem_st32_zero(tcol + (uint32_t)z * 32u);
                            tmem_st_wait();
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(blk_empty(blk));
                    }
#pragma unroll
                    for (int k = 0; k < 32; k++) {
                        const float v = fmaf(act[k], p.out_scale, p.bias[cb * 32 + k]);     // = ACT_SCALE * (conv + bias)
                        act[k] = fmaxf(v, 0.1f * v);                                         // leaky 0.1
                    }
                    const int gy = y0 + i - p.out_y0;
#ifdef W2X_EPI_EXPERIMENTS
                    const bool direct = C::DIRECT || (p.dbg & (128 | 256)) != 0;
#else
                    constexpr bool direct = C::DIRECT;
#endif
                    if (direct) {
                        if (gx0 + lane < p.Wp && gy >= 0 && gy < p.out_rows)
                            epilogue_store32_direct<F8>(act, p.out_win + (((size_t)gy * p.Wp + gx0 + lane) * (COUT / 32) + cb) * 128u, !(p.dbg & 128));
                        continue;
                    }
                    if (gx0 < p.Wp && gy >= 0 && gy < p.out_rows) epilogue_store32_rec<F8>(act, &tmap_out, p.dbg, stg, lane, gx0, gy, cb);
                }
                if (prof_on) work_e += (unsigned long long)(clock64() - t_work);
            }
            nrow += (uint32_t)rows;
        }
        bulk_wait_all();
        if (prof_on && warp == 4 && lane == 0) {
            prof[PROF_EPI_WAIT] += w_e;
            prof[PROF_EPI_WORK] += work_e;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}
