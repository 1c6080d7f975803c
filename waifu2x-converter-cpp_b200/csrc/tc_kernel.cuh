// tc_kernel.cuh -- tc_conv3x3_kernel: the single-CTA layer kernel (warp roles, pipelines)
// Part of the tcgen05 engine's single translation unit: included by kernels_tc.cu inside namespace w2x::tc, in this order:
//   tc_ptx.cuh, tc_config.cuh, tc_issue.cuh, tc_epilogue.cuh, tc_kernel.cuh, tc_pair_kernel.cuh, tc_strip_kernel.cuh, tc_edge_kernels.cuh
// (pure code organisation: the generated SASS is the same as with one file).

// ================================================================================================
// The layer kernel
// ================================================================================================
template <int CIN, int COUT, bool FUSE, bool F8>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_conv3x3_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_out, const TcParams p) {
    using C = Cfg<CIN, COUT, FUSE, F8>;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment: the SWIZZLE_128B pattern repeats every 1024 B
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base;
    const uint32_t b_base = a_base + C::A_SLOTS * C::A_SLOT;
    const uint32_t bar_base = b_base + C::NB * C::B_STAGE;
    // barrier map (8 bytes each)
    auto a_full = [&](int i) { return bar_base + 8u * (uint32_t)i; };
    auto a_empty = [&](int i) { return bar_base + 8u * (uint32_t)(2 + i); };
    auto acc_full = [&](int i) { return bar_base + 8u * (uint32_t)(4 + i); };
    auto acc_empty = [&](int i) { return bar_base + 8u * (uint32_t)(6 + i); };
    auto b_full = [&](int i) { return bar_base + 8u * (uint32_t)(8 + i); };
    auto b_empty = [&](int i) { return bar_base + 8u * (uint32_t)(8 + C::NB + i); };
    const uint32_t tmem_slot = bar_base + 8u * (uint32_t)(8 + 2 * C::NB);   // 4 bytes: TMEM base address
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    // warp index through a shuffle: ptxas then knows it is warp-uniform, and with it the role branch, the M-tile index and
    // every descriptor derived from them (uniform registers feed tcgen05.mma directly, no per-MMA R2UR waterfall)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const bool prof_on = p.prof != nullptr;
    unsigned long long *prof = prof_on ? p.prof + (size_t)blockIdx.x * PROF_N : nullptr;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(a_full(i), 1);
            mbar_init(a_empty(i), 2);     // one tcgen05.commit per MMA issuer
            mbar_init(acc_full(i), 2);
            mbar_init(acc_empty(i), 8);   // one arrive per epilogue warp
        }
        for (int i = 0; i < C::NB; i++) {
            mbar_init(b_full(i), 1);
            mbar_init(b_empty(i), 2);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_in);
        if constexpr (!FUSE) prefetch_tmap(&tmap_out);
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, C::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);   // uniform for the compiler as well

    if (warp == 0) {
        // ===================== A producer: one halo'd box of records per (tile-set, 32-channel block) ==============
        // (whole warp walks the loop; the arrive and the TMA instructions elect one lane)
        {
            uint32_t it = 0;
            unsigned long long w_a = 0;
            for (int ts = blockIdx.x; ts < p.n_tilesets; ts += gridDim.x) {
                const int ty = ts / p.tiles_x, tx = ts - ty * p.tiles_x;
                const int x0 = tx * REGION - 1, y0 = p.out_y0 + ty * REGION - 1;   // box origin incl. ring (may be -1); tile-sets tile the store window
                for (int c = 0; c < C::NCHUNK; c++, it++) {
                    const uint32_t slot = it & 1u, round = it >> 1;
                    mbar_wait_prof(a_empty(slot), (round & 1u) ^ 1u, prof_on, w_a);
                    mbar_arrive_expect_tx(a_full(slot), (uint32_t)C::A_TX);
                    tma_load_4d(a_base + slot * C::A_SLOT, &tmap_in, a_full(slot), 0, c, x0, y0);
                }
            }
            if (prof_on && lane == 0) prof[PROF_APROD_WAIT] += w_a;
        }
    } else if (warp == 2) {
        // ===================== B producer: stream the packed weights in consumption order ============
        {
            uint32_t stage = 0, phase = 0;
            unsigned long long w_b = 0;
            for (int ts = blockIdx.x; ts < p.n_tilesets; ts += gridDim.x) {
                const uint8_t *src = reinterpret_cast<const uint8_t *>(p.wpack);
                if (C::RESIDENT && ts != (int)blockIdx.x) break;           // resident weights: one pass fills every stage for good
                for (int blk = 0; blk < C::STAGES_PER_TILESET; blk++) {
                    if constexpr (!C::RESIDENT) mbar_wait_prof(b_empty(stage), phase ^ 1u, prof_on, w_b);
                    mbar_arrive_expect_tx(b_full(stage), C::B_STAGE);
                    bulk_load(b_base + stage * C::B_STAGE, src + (size_t)blk * C::B_STAGE, C::B_STAGE, b_full(stage));
                    if (++stage == (uint32_t)C::NB) { stage = 0; phase ^= 1u; }
                }
            }
            if (prof_on && lane == 0) prof[PROF_BPROD_WAIT] += w_b;
        }
    } else if (warp == 1 || warp == 7) {
        // ===================== MMA issuers (warp 1: M-tile 0, warp 7: M-tile 1) ========================
        // The whole warp walks the loop converged; each MMA / commit elects one lane inside its asm block.
        const uint32_t leader = lane == 0 ? 1u : 0u;
        const uint32_t jt = warp == 1 ? 0u : 1u;
        constexpr uint32_t idesc_c = make_idesc(128, COUT);          // N = Cout
        constexpr uint32_t idesc_2c = make_idesc(128, 2 * COUT);     // N = 2*Cout (stacked [wh;wl]); only used when STACK
        constexpr uint32_t A_SBO = HALO * C::ROWB;                   // next output row = next halo row
        constexpr uint32_t B_SBO = 8 * C::B_ROWB;                    // dense rows
        constexpr uint32_t A_HI32 = (uint32_t)(make_desc_const(A_SBO, C::A_LAYOUT) >> 32);
        constexpr uint32_t B_HI32 = (uint32_t)(make_desc_const(B_SBO, C::B_LAYOUT) >> 32);
        constexpr uint32_t LO_FIXED = 1u << 16;                      // LBO field = 1
        // e4m3 weight blocks: 32-byte rows (SWIZZLE_32B); the e4m3 activation slices are quarters of the same 128-byte records
        constexpr uint32_t B8_HI32 = (uint32_t)(make_desc_const(8 * 32, 6u) >> 32);
        auto desc = [](uint32_t hi32, uint32_t lo32) { return ((uint64_t)hi32 << 32) | (uint64_t)lo32; };
        uint32_t a_it = 0, stage = 0, phase = 0, n = 0;
        uint32_t b_ready = 0;                         // result of the early probe of b_full(stage)
        unsigned long long w_acc = 0, w_af = 0, w_bf = 0;
        const long long t_begin = clock64();
        // wait for the current weight stage (usually already known to be full), then probe the NEXT one
        auto acquire_b = [&](uint32_t &b0_out) {
            if constexpr (C::RESIDENT) {
                if (n == 0) {                             // the stages arrive once, during the first tile-set
                    mbar_wait_prof(b_full(stage), 0u, prof_on, w_bf);
                    tc_fence_after();
                }
                b0_out = (((b_base + stage * C::B_STAGE) >> 4) & 0x3FFFu) | LO_FIXED;
            } else {
                if (!b_ready) mbar_wait_prof(b_full(stage), phase, prof_on, w_bf);
                tc_fence_after();
                b0_out = (((b_base + stage * C::B_STAGE) >> 4) & 0x3FFFu) | LO_FIXED;
                uint32_t ns = stage + 1, np = phase;
                if (ns == (uint32_t)C::NB) { ns = 0; np ^= 1u; }
                b_ready = mbar_test(b_full(ns), np);      // consumed at the next acquire_b
            }
        };
        auto release_b = [&]() {
            if constexpr (!C::RESIDENT) umma_commit_one(b_empty(stage));
            if (++stage == (uint32_t)C::NB) { stage = 0; phase ^= 1u; }
        };
        for (int ts = blockIdx.x; ts < p.n_tilesets; ts += gridDim.x, n++) {
            const uint32_t set = n & 1u;
            mbar_wait_prof(acc_empty(set), ((n >> 1) & 1u) ^ 1u, prof_on, w_acc);
            tc_fence_after();
            const uint32_t dj = tmem_base + (set * 2u + jt) * C::TILE_COLS;   // this issuer's accumulator columns
            for (int c = 0; c < C::NCHUNK; c++, a_it++) {
                const uint32_t slot = a_it & 1u;
                mbar_wait_prof(a_full(slot), (a_it >> 1) & 1u, prof_on, w_af);
                tc_fence_after();
                // descriptor low word (address >> 4) of this issuer's window into the staged records; a record's quarters:
                // +0 / +2 the fp16 K steps, +4 xh8 (f16x3: lo step 0), +6 xl8 (f16x3: lo step 1)   [16-byte units]
                const uint32_t ah0 = ((((a_base + slot * C::A_SLOT) >> 4) & 0x3FFFu) | LO_FIXED) + jt * (8u * C::ROWB >> 4);
                uint32_t tap_off = 0;                 // ((ky*HALO + kx) * ROWB) >> 4
                for (int t = 0; t < 9; t++) {
                    const uint32_t first = (c | t) != 0 ? 1u : 0u;
                    {
                        const uint32_t ah = ah0 + tap_off, al = ah + 4u;
                        const uint32_t acc0 = first;
                        uint32_t b0;
                        if constexpr (C::MERGE) {
                            // one stage = [wh fp16 | wh8 | wl8]: main product (two K=16 steps) + both e4m3 corrections (K=32 each)
                            acquire_b(b0);
                            issue_tap_f8<false>(dj, ah, b0, COUT, A_HI32, B_HI32, B8_HI32, idesc_c, acc0);
                            release_b();
                        } else if constexpr (C::STACK) {
                            // one stage = [wh ; wl]: xh*[wh;wl] (N = 2*Cout, D1|D2) then xl*wh (N = Cout, D1)
                            acquire_b(b0);
                            umma_f16(dj, desc(A_HI32, ah), desc(B_HI32, b0), idesc_2c, acc0);
                            umma_f16(dj, desc(A_HI32, ah + 2u), desc(B_HI32, b0 + 2u), idesc_2c, 1u);
                            umma_f16(dj, desc(A_HI32, al), desc(B_HI32, b0), idesc_c, 1u);
                            umma_f16(dj, desc(A_HI32, al + 2u), desc(B_HI32, b0 + 2u), idesc_c, 1u);
                            release_b();
                        } else {
                            // ---- hi weights: xh*wh and xl*wh ----
                            acquire_b(b0);
                            umma_f16(dj, desc(A_HI32, ah), desc(B_HI32, b0), idesc_c, acc0);
                            umma_f16(dj, desc(A_HI32, ah + 2u), desc(B_HI32, b0 + 2u), idesc_c, 1u);
                            umma_f16(dj, desc(A_HI32, al), desc(B_HI32, b0), idesc_c, 1u);
                            umma_f16(dj, desc(A_HI32, al + 2u), desc(B_HI32, b0 + 2u), idesc_c, 1u);
                            release_b();
                            // ---- lo weights: xh*wl ----
                            acquire_b(b0);
                            umma_f16(dj, desc(A_HI32, ah), desc(B_HI32, b0), idesc_c, 1u);
                            umma_f16(dj, desc(A_HI32, ah + 2u), desc(B_HI32, b0 + 2u), idesc_c, 1u);
                            release_b();
                        }
                    }
                    // next tap: kx+1, or the next halo row
                    tap_off += tap_step<C::ROWB>(t);
                }
                umma_commit_one(a_empty(slot));   // the staged boxes may be overwritten once these MMAs retire
            }
            umma_commit_one(acc_full(set));       // this issuer's accumulators of the tile-set are final
        }
        if (prof_on && leader && jt == 0) {
            prof[PROF_TOTAL] += (unsigned long long)(clock64() - t_begin);
            prof[PROF_MMA_WAIT_ACC] += w_acc;
            prof[PROF_MMA_WAIT_A] += w_af;
            prof[PROF_MMA_WAIT_B] += w_bf;
            prof[PROF_TILESETS] += n;
        }
    } else {
        // ===================== epilogue: warps 3..6 drain M-tile 0, warps 8..11 drain M-tile 1 ==========
        const uint32_t q = (uint32_t)warp & 3u;          // TMEM lane quarter this warp may access
        const int j = warp >= 8 ? 1 : 0;                 // M-tile
        const uint32_t row = q * 32u + (uint32_t)lane;   // GEMM row = pixel inside the 8x16 M-tile
        const int oy = (int)(row >> 3), ox = (int)(row & 7u);
        const uint32_t stg = bar_base + C::BAR_BYTES + C::W6_BYTES + (uint32_t)(j * 4 + (int)q) * (uint32_t)C::STG_WARP;   // this warp's staging tile
        uint32_t n = 0;
        unsigned long long w_e = 0, work_e = 0;
        for (int ts = blockIdx.x; ts < p.n_tilesets; ts += gridDim.x, n++) {
            const uint32_t set = n & 1u;
            const int ty = ts / p.tiles_x, tx = ts - ty * p.tiles_x;
            mbar_wait_prof(acc_full(set), (n >> 1) & 1u, prof_on, w_e);
            const long long t_work = prof_on ? clock64() : 0;
            tc_fence_after();
            const uint32_t tcol = tmem_base + ((q * 32u) << 16) + (set * 2u + (uint32_t)j) * C::TILE_COLS;
            tile_epilogue<COUT, FUSE, F8, C::STACK>(p, &tmap_out, tcol, stg, lane, tx * REGION + 8 * j, ty * REGION + 4 * (int)q,
                                                    tx * REGION + 8 * j + ox, p.out_y0 + ty * REGION + oy, [&] { mbar_arrive(acc_empty(set)); });
            if (prof_on) work_e += (unsigned long long)(clock64() - t_work);
        }
        if constexpr (!FUSE) bulk_wait_all();    // this warp's TMA stores are complete before the CTA may exit
        if (prof_on && warp == 3 && lane == 0) {
            prof[PROF_EPI_WAIT] += w_e;
            prof[PROF_EPI_WORK] += work_e;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}
