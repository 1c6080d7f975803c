// tc_pair_kernel.cuh -- tc_conv3x3_pair_kernel: the cta_group::2 variant for the 128-wide layers
// Part of the tcgen05 engine's single translation unit: included by kernels_tc.cu inside namespace w2x::tc, in this order:
//   tc_ptx.cuh, tc_config.cuh, tc_issue.cuh, tc_epilogue.cuh, tc_kernel.cuh, tc_pair_kernel.cuh, tc_strip_kernel.cuh, tc_edge_kernels.cuh
// (pure code organisation: the generated SASS is the same as with one file).

// ================================================================================================
// The CTA-pair variant (cta_group::2) for Cout = 128
// ================================================================================================
// Two CTAs of a cluster (the two SMs of a TPC) each stage THEIR 16x16 region like the single-CTA kernel, but every
// tcgen05.mma is M = 256: rows 0-127 come from CTA 0's shared memory, rows 128-255 from CTA 1's, and the N = 128
// weight rows are split -- each CTA loads and holds only 64 of them.  One thread pair in the leader CTA drives both
// SMs' tensor cores.  Per CTA this halves the weight bytes pulled from L2 and the B-operand bytes read from shared
// memory per MMA (the single-CTA N = 128 MMAs sit at the 128 B/clk shared-memory operand limit).
//   * all TMA loads of both CTAs signal the LEADER's mbarriers (cp.async.bulk.tensor ... .cta_group::2),
//   * tcgen05.commit ... .multicast::cluster releases stages / slots / accumulators in both CTAs,
//   * both CTAs' epilogue warps arrive (remotely) on the leader's accumulator-empty barriers.
template <int CIN, int COUT, bool FUSE, bool F8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
tc_conv3x3_pair_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_w,
                       const __grid_constant__ CUtensorMap tmap_out, const TcParams p) {
    using C = PairCfg<CIN, COUT, FUSE, F8>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base;
    constexpr int NAS = C::A_SLOTS;
    const uint32_t b_base = a_base + NAS * C::A_SLOT;
    const uint32_t bar_base = b_base + C::NBP * C::B_HALF;
    auto a_full = [&](int i) { return bar_base + 8u * (uint32_t)i; };
    auto a_empty = [&](int i) { return bar_base + 8u * (uint32_t)(NAS + i); };
    auto acc_full = [&](int i) { return bar_base + 8u * (uint32_t)(2 * NAS + i); };
    auto acc_empty = [&](int i) { return bar_base + 8u * (uint32_t)(2 * NAS + 2 + i); };
    auto b_full = [&](int i) { return bar_base + 8u * (uint32_t)(2 * NAS + 4 + i); };
    auto b_empty = [&](int i) { return bar_base + 8u * (uint32_t)(2 * NAS + 4 + C::NBP + i); };
    const uint32_t tmem_slot = bar_base + 8u * (uint32_t)(2 * NAS + 4 + 2 * C::NBP);
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    // warp index through a shuffle: ptxas then knows it is warp-uniform, and with it the role branch, the M-tile index and
    // every descriptor derived from them (uniform registers feed tcgen05.mma directly, no per-MMA R2UR waterfall)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const bool prof_on = p.prof != nullptr;
    unsigned long long *prof = prof_on ? p.prof + (size_t)blockIdx.x * PROF_N : nullptr;
    const uint32_t rank = cluster_ctarank();
    const bool is_leader = rank == 0;
    const int n_pairs_cl = (int)(gridDim.x >> 1), pair_id = (int)(blockIdx.x >> 1);
    const int n_pair_sets = (p.n_tilesets + 1) / 2;          // tile-sets are taken two at a time: (2q, 2q+1) -> (CTA 0, CTA 1)
    const int tiles_y = (p.out_rows + REGION - 1) / REGION;   // tile-sets tile the store window [out_y0, out_y0 + out_rows)

    if (threadIdx.x == 0) {
        for (int i = 0; i < NAS; i++) {
            mbar_init(a_full(i), 1);        // leader's: its A producer's arrive.expect_tx covers the bytes of BOTH CTAs
            mbar_init(a_empty(i), 2);       // one multicast tcgen05.commit per issuer
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(acc_full(i), 2);
            mbar_init(acc_empty(i), 16);    // leader's: 8 local + 8 remote epilogue warps
        }
        for (int i = 0; i < C::NBP; i++) {
            mbar_init(b_full(i), 1);
            mbar_init(b_empty(i), 2);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_in);
        prefetch_tmap(&tmap_w);
        if constexpr (!FUSE) prefetch_tmap(&tmap_out);
    }
    cluster_sync_all();                     // both CTAs' barriers exist before anything can signal them
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);   // uniform for the compiler as well
    // this CTA's tile-set of pair-set q (a phantom region below the frame when the count is odd: loads zero-fill, stores are masked)
    auto region_of = [&](int q, int &tx, int &ty) {
        const int ts = 2 * q + (int)rank;
        if (ts < p.n_tilesets) { ty = ts / p.tiles_x; tx = ts - ty * p.tiles_x; }
        else { ty = tiles_y; tx = 0; }
    };

    if (warp == 0) {
        // ===================== A producer (both CTAs): boxes land locally, completion is counted on the LEADER's barrier ====
        {
            uint32_t it = 0;
            unsigned long long w_a = 0;
            for (int q = pair_id; q < n_pair_sets; q += n_pairs_cl) {
                int tx, ty;
                region_of(q, tx, ty);
                const int x0 = tx * REGION - 1, y0 = p.out_y0 + ty * REGION - 1;
                for (int c = 0; c < C::NCHUNK; c++, it++) {
                    const uint32_t slot = it % (uint32_t)NAS, round = it / (uint32_t)NAS;
                    mbar_wait_prof(a_empty(slot), (round & 1u) ^ 1u, prof_on, w_a);
                    if (is_leader) mbar_arrive_expect_tx(a_full(slot), 2u * (uint32_t)C::A_TX);
                    tma_load_4d_2cta(a_base + slot * C::A_SLOT, &tmap_in, mapa_rank(a_full(slot), 0), 0, c, x0, y0);   // 324 records of this 32-channel block
                }
            }
            if (prof_on && lane == 0) prof[PROF_APROD_WAIT] += w_a;
        }
    } else if (warp == 2) {
        // ===================== B producer (both CTAs): this CTA's 64 rows of every weight stage ========================
        // tmap_w views the packed stream as rows of 1 KB, box = 2 rows (2 KB).  A 128-row fp16 block is 8 KB
        // (this CTA's half: 4 KB at +rank*4 KB); an e4m3 stage is [wh8 4 KB | wl8 4 KB] (halves: 2 KB at +rank*2 KB each).
        {
            uint32_t stage = 0, phase = 0;
            unsigned long long w_b = 0;
            constexpr int N_STEPS = C::NCHUNK * 9;                             // one stage per (32-channel block, tap)
            for (int q = pair_id; q < n_pair_sets; q += n_pairs_cl) {
                for (int blk = 0; blk < N_STEPS; blk++) {
                    mbar_wait_prof(b_empty(stage), phase ^ 1u, prof_on, w_b);
                    if (is_leader) mbar_arrive_expect_tx(b_full(stage), 2u * (uint32_t)C::B_HALF);
                    const uint32_t bar = mapa_rank(b_full(stage), 0);
                    const uint32_t dst = b_base + stage * C::B_HALF;
                    const int row0 = blk * (2 * C::B_BLOCK / 1024);           // first 1 KB row of this step in the stream (16 rows per step)
                    // first block (128 rows x 64 B = 8 KB, fp16): this CTA's operand rows 64*rank .. +64 = 4 KB at +rank*4 KB
                    tma_load_2d_2cta(dst, &tmap_w, bar, 0, row0 + (int)rank * 4);
                    tma_load_2d_2cta(dst + 2048u, &tmap_w, bar, 0, row0 + (int)rank * 4 + 2);
                    if constexpr (F8) {   // [wh8 | wl8]: 128 rows x 32 B = 4 KB each; this CTA's half = 2 KB
                        tma_load_2d_2cta(dst + 4096u, &tmap_w, bar, 0, row0 + 8 + (int)rank * 2);
                        tma_load_2d_2cta(dst + 6144u, &tmap_w, bar, 0, row0 + 12 + (int)rank * 2);
                    } else {              // lo block (fp16)
                        tma_load_2d_2cta(dst + 4096u, &tmap_w, bar, 0, row0 + 8 + (int)rank * 4);
                        tma_load_2d_2cta(dst + 6144u, &tmap_w, bar, 0, row0 + 8 + (int)rank * 4 + 2);
                    }
                    if (++stage == (uint32_t)C::NBP) { stage = 0; phase ^= 1u; }
                }
            }
            if (prof_on && lane == 0) prof[PROF_BPROD_WAIT] += w_b;
        }
    } else if (warp == 1 || warp == 7) {
        // ===================== MMA issuers: LEADER CTA only, M = 256 across the pair ====================================
        if (is_leader) {
            const uint32_t leader = lane == 0 ? 1u : 0u;
            const uint32_t jt = warp == 1 ? 0u : 1u;
            constexpr uint32_t idesc_c = make_idesc(256, COUT);
            constexpr uint32_t A_SBO = HALO * C::ROWB;
            constexpr uint32_t B_SBO = 8 * C::B_ROWB;
            constexpr uint32_t A_HI32 = (uint32_t)(make_desc_const(A_SBO, C::A_LAYOUT) >> 32);
            constexpr uint32_t B_HI32 = (uint32_t)(make_desc_const(B_SBO, C::B_LAYOUT) >> 32);
            constexpr uint32_t LO_FIXED = 1u << 16;
            constexpr uint32_t B8_HI32 = (uint32_t)(make_desc_const(8 * 32, 6u) >> 32);
            auto desc = [](uint32_t hi32, uint32_t lo32) { return ((uint64_t)hi32 << 32) | (uint64_t)lo32; };
            uint32_t a_it = 0, stage = 0, phase = 0, n = 0, b_ready = 0;
            unsigned long long w_acc = 0, w_af = 0, w_bf = 0;
            const long long t_begin = clock64();
            auto acquire_b = [&](uint32_t &b0_out) {
                if (!b_ready) mbar_wait_prof(b_full(stage), phase, prof_on, w_bf);
                tc_fence_after();
                b0_out = (((b_base + stage * C::B_HALF) >> 4) & 0x3FFFu) | LO_FIXED;
                uint32_t ns = stage + 1, np = phase;
                if (ns == (uint32_t)C::NBP) { ns = 0; np ^= 1u; }
                b_ready = mbar_test(b_full(ns), np);
            };
            auto release_b = [&]() {
                umma2_commit_one(b_empty(stage));
                if (++stage == (uint32_t)C::NBP) { stage = 0; phase ^= 1u; }
            };
            for (int q = pair_id; q < n_pair_sets; q += n_pairs_cl, n++) {
                const uint32_t set = n & 1u;
                mbar_wait_prof(acc_empty(set), ((n >> 1) & 1u) ^ 1u, prof_on, w_acc);
                tc_fence_after();
                const uint32_t dj = tmem_base + (set * 2u + jt) * C::TILE_COLS;
                for (int c = 0; c < C::NCHUNK; c++, a_it++) {
                    const uint32_t slot = a_it % (uint32_t)NAS;
                    mbar_wait_prof(a_full(slot), (a_it / (uint32_t)NAS) & 1u, prof_on, w_af);
                    tc_fence_after();
                    // this issuer's window into the staged records; a record's quarters: +0 / +2 the fp16 K steps, +4 xh8 (f16x3: lo
                    // step 0), +6 xl8 (f16x3: lo step 1)   [16-byte units]
                    const uint32_t ah0 = ((((a_base + slot * C::A_SLOT) >> 4) & 0x3FFFu) | LO_FIXED) + jt * (8u * C::ROWB >> 4);
                    uint32_t tap_off = 0;
                    for (int t = 0; t < 9; t++) {
                        const uint32_t first = (c | t) != 0 ? 1u : 0u;
                        {
                            const uint32_t ah = ah0 + tap_off, al = ah + 4u;
                            const uint32_t acc0 = first;
                            uint32_t b0;
                            acquire_b(b0);                  // one stage per 32-channel step: this CTA's rows of both blocks
                            if constexpr (F8) {
                                issue_tap_f8<true>(dj, ah, b0, COUT / 2, A_HI32, B_HI32, B8_HI32, idesc_c, acc0);   // this CTA holds half of the B rows
                            } else {
                                umma2_f16(dj, desc(A_HI32, ah), desc(B_HI32, b0), idesc_c, acc0);
                                umma2_f16(dj, desc(A_HI32, ah + 2u), desc(B_HI32, b0 + 2u), idesc_c, 1u);
                                umma2_f16(dj, desc(A_HI32, al), desc(B_HI32, b0), idesc_c, 1u);
                                umma2_f16(dj, desc(A_HI32, al + 2u), desc(B_HI32, b0 + 2u), idesc_c, 1u);
                                umma2_f16(dj, desc(A_HI32, ah), desc(B_HI32, b0 + (4096u >> 4)), idesc_c, 1u);
                                umma2_f16(dj, desc(A_HI32, ah + 2u), desc(B_HI32, b0 + (4096u >> 4) + 2u), idesc_c, 1u);
                            }
                            release_b();
                        }
                        tap_off += tap_step<C::ROWB>(t);
                    }
                    umma2_commit_one(a_empty(slot));
                }
                umma2_commit_one(acc_full(set));
            }
            if (prof_on && leader && jt == 0) {
                prof[PROF_TOTAL] += (unsigned long long)(clock64() - t_begin);
                prof[PROF_MMA_WAIT_ACC] += w_acc;
                prof[PROF_MMA_WAIT_A] += w_af;
                prof[PROF_MMA_WAIT_B] += w_bf;
                prof[PROF_TILESETS] += n;
            }
        }
    } else {
        // ===================== epilogue (both CTAs), same math as the single-CTA kernel ===================================
        const uint32_t q4 = (uint32_t)warp & 3u;
        const int j = warp >= 8 ? 1 : 0;
        const uint32_t row = q4 * 32u + (uint32_t)lane;
        const int oy = (int)(row >> 3), ox = (int)(row & 7u);
        const uint32_t stg = bar_base + C::BAR_BYTES + C::W6_BYTES + (uint32_t)(j * 4 + (int)q4) * (uint32_t)C::STG_WARP;
        uint32_t n = 0;
        unsigned long long w_e = 0, work_e = 0;
        for (int q = pair_id; q < n_pair_sets; q += n_pairs_cl, n++) {
            const uint32_t set = n & 1u;
            int tx, ty;
            region_of(q, tx, ty);
            mbar_wait_prof(acc_full(set), (n >> 1) & 1u, prof_on, w_e);
            const long long t_work = prof_on ? clock64() : 0;
            tc_fence_after();
            const uint32_t tcol = tmem_base + ((q4 * 32u) << 16) + (set * 2u + (uint32_t)j) * C::TILE_COLS;
            tile_epilogue<COUT, FUSE, F8, false>(p, &tmap_out, tcol, stg, lane, tx * REGION + 8 * j, ty * REGION + 4 * (int)q4,
                                                 tx * REGION + 8 * j + ox, p.out_y0 + ty * REGION + oy,
                                                 [&] { mbar_arrive_cluster(mapa_rank(acc_empty(set), 0)); });   // both CTAs' warps arrive on the LEADER's barrier
            if (prof_on) work_e += (unsigned long long)(clock64() - t_work);
        }
        if constexpr (!FUSE) bulk_wait_all();    // this warp's TMA stores are complete before the CTA may exit
        if (prof_on && warp == 3 && lane == 0) {
            prof[PROF_EPI_WAIT] += w_e;
            prof[PROF_EPI_WORK] += work_e;
        }
    }

    tc_fence_before();
    cluster_sync_all();                     // nobody may free TMEM / exit while the peer can still signal or read
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
}
