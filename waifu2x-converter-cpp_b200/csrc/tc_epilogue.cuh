// tc_epilogue.cuh -- the epilogue store: accumulator block -> frame planes -> swizzled staging tile -> TMA store
// Part of the tcgen05 engine's single translation unit: included by kernels_tc.cu inside namespace w2x::tc, in this order:
//   tc_ptx.cuh, tc_config.cuh, tc_issue.cuh, tc_epilogue.cuh, tc_kernel.cuh, tc_pair_kernel.cuh, tc_strip_kernel.cuh, tc_edge_kernels.cuh
// (pure code organisation: the generated SASS is the same as with one file).

// Epilogue store of 32 activated output channels of ONE pixel per thread (lane = pixel inside this warp's 32-pixel block:
// 8x4 pixels of an M-tile in the tile kernels, 32 consecutive pixels of a strip in the strip kernel).  The warp converts
// to the record's slices, writes them into its 4 KB staging tile in the TMA swizzle pattern (conflict-free 16-byte stores)
// and one lane issues the TMA store of the box: the bytes leave asynchronously while the warp converts the next 32
// channels, the frame edge is clipped by the TMA unit, and the shared-memory pipe (which the tensor core's operand fetches
// saturate) sees one pass instead of a store + load round trip.
// RECORD frames ([Hp][Wp][C/32][128 B], one 128-byte record per pixel per 32-channel block:
// {xh fp16 x32 | xh8 x32 | xl8 x32} or {hi fp16 x32 | lo fp16 x32}): the warp's staging tile is [32 px][128 B] in the
// SWIZZLE_128B pattern and leaves as ONE box of 32 rows of 128 B -- a third of the TMA row requests of the planar
// frame (the TMA unit serves ~1 row per 1.45 clk whatever its length: profiles/r02_strip_experiments.txt).
template <bool F8>
__device__ __forceinline__ void epilogue_store32_rec(const float (&act)[32], const CUtensorMap *tmo, int dbg, uint32_t stg, int lane, int gx0, int gy0,
                                                     int cb) {
    uint32_t g0[16], g1[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const float v0 = act[2 * i], v1 = act[2 * i + 1];
        __half2 h = __floats2half2_rn(v0, v1);
        float2 hf = __half22float2(h);
        g0[i] = *reinterpret_cast<uint32_t *>(&h);
        if constexpr (F8) {
            constexpr float kDown = 1.0f / (float)(1 << F8_C), kUp = (float)(1 << F8_A);
            const __half2 hd = __hmul2(h, __float2half2_rn(kDown));
            const uint32_t h8 = __nv_cvt_halfraw2_to_fp8x2(static_cast<__half2_raw>(hd), __NV_SATFINITE, __NV_E4M3);
            const uint32_t l8 = __nv_cvt_float2_to_fp8x2(make_float2((v0 - hf.x) * kUp, (v1 - hf.y) * kUp), __NV_SATFINITE, __NV_E4M3);
            if (i & 1) { g1[i >> 1] |= h8 << 16; g1[8 + (i >> 1)] |= l8 << 16; }
            else { g1[i >> 1] = h8; g1[8 + (i >> 1)] = l8; }
        } else {
            __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
            g1[i] = *reinterpret_cast<uint32_t *>(&l);
        }
    }
    if (dbg & 2) {
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) x ^= g0[i] ^ g1[i];
        if (x == 0x7fc12345u) sts128(stg, make_uint4(x, x, x, x));
        return;
    }
    bulk_wait_read();
    __syncwarp();
    const uint32_t row = stg + (uint32_t)lane * 128u, sw = (uint32_t)lane & 7u;
#pragma unroll
    for (int u = 0; u < 4; u++) {   // units 0..3: the fp16 half; units 4..7: [xh8 | xl8] or lo
        sts128(row + (((uint32_t)u ^ sw) << 4), make_uint4(g0[4 * u], g0[4 * u + 1], g0[4 * u + 2], g0[4 * u + 3]));
        sts128(row + (((uint32_t)(4 + u) ^ sw) << 4), make_uint4(g1[4 * u], g1[4 * u + 1], g1[4 * u + 2], g1[4 * u + 3]));
    }
    fence_proxy_async();
    __syncwarp();
    if (!(dbg & 1)) {
        tma_store_4d(tmo, stg, 0, cb, gx0, gy0);
        bulk_commit();
    }
}

// The same record straight from registers: 4 x 32-byte (STG.256) or 8 x 16-byte global stores per lane, no staging tile and no
// TMA store -- nothing of the epilogue crosses the shared-memory pipe (which bounds the 64 -> 64 strip layer), but a
// warp-wide store touches 32 different 128-byte lines: 16-byte stores are LSU-bound everywhere, 32-byte stores pay off where
// the shared-memory pipe is the bound and lose where it is not (profiles/r02_strip_direct_store_experiment.txt).
template <bool F8>
__device__ __forceinline__ void epilogue_store32_direct(const float (&act)[32], uint8_t *rec, bool wide) {
    uint32_t g[32];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const float v0 = act[2 * i], v1 = act[2 * i + 1];
        __half2 h = __floats2half2_rn(v0, v1);
        float2 hf = __half22float2(h);
        g[i] = *reinterpret_cast<uint32_t *>(&h);
        if constexpr (F8) {
            constexpr float kDown = 1.0f / (float)(1 << F8_C), kUp = (float)(1 << F8_A);
            const __half2 hd = __hmul2(h, __float2half2_rn(kDown));
            const uint32_t h8 = __nv_cvt_halfraw2_to_fp8x2(static_cast<__half2_raw>(hd), __NV_SATFINITE, __NV_E4M3);
            const uint32_t l8 = __nv_cvt_float2_to_fp8x2(make_float2((v0 - hf.x) * kUp, (v1 - hf.y) * kUp), __NV_SATFINITE, __NV_E4M3);
            if (i & 1) { g[16 + (i >> 1)] |= h8 << 16; g[24 + (i >> 1)] |= l8 << 16; }
            else { g[16 + (i >> 1)] = h8; g[24 + (i >> 1)] = l8; }
        } else {
            __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
            g[16 + i] = *reinterpret_cast<uint32_t *>(&l);
        }
    }
    if (wide) {
#pragma unroll
        for (int u = 0; u < 4; u++)
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(rec + 32 * u), "r"(g[8 * u]), "r"(g[8 * u + 1]), "r"(g[8 * u + 2]),
                         "r"(g[8 * u + 3]), "r"(g[8 * u + 4]), "r"(g[8 * u + 5]), "r"(g[8 * u + 6]), "r"(g[8 * u + 7])
                         : "memory");
    } else {
#pragma unroll
        for (int u = 0; u < 8; u++) *reinterpret_cast<uint4 *>(rec + 16 * u) = make_uint4(g[4 * u], g[4 * u + 1], g[4 * u + 2], g[4 * u + 3]);
    }
}

// One M-tile's share of a tile-set in the 16x16-tile kernels (single-CTA and CTA-pair): this warp's 32 pixels x COUT
// accumulator columns -> scale, bias, leaky-ReLU -> either records (TMA store) or, with the last layer folded in (FUSE),
// the nine per-tap dot products of the pixel.  `release()` hands the accumulator columns back to the MMA issuer (a local
// mbarrier arrive in the single-CTA kernel, a remote one on the leader's barrier in the pair kernel) as soon as the last
// columns are in registers.  STACK (f16x3, Cout <= 64): the accumulator is D1 | D2 (xh*[wh;wl]), summed here.
template <int COUT, bool FUSE, bool F8, bool STACK, typename Release>
__device__ __forceinline__ void tile_epilogue(const TcParams &p, const CUtensorMap *tmap_out, uint32_t tcol, uint32_t stg, int lane, int gx0, int gy0,
                                              int fx, int fy, Release &&release) {
    const bool inside = fy < p.Hp && fx < p.Wp && fy >= p.out_y0 && fy < p.out_y0 + p.out_rows;
    float pt[9];                                   // FUSE: nine per-tap dot products of this pixel
#pragma unroll
    for (int t = 0; t < 9; t++) pt[t] = 0.f;
    uint32_t r[32];
    if constexpr (!STACK) tmem_ld32(tcol, r);
#pragma unroll
    for (int cb = 0; cb < COUT / 32; cb++) {
        // ---- 32 output channels of this pixel: accumulator -> scale, bias, leaky-ReLU ----
        float act[32];
        if constexpr (STACK) {
            uint32_t r2[32];
            tmem_ld32(tcol + (uint32_t)cb * 32u, r);
            tmem_ld32(tcol + (uint32_t)(COUT + cb * 32), r2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i++) act[i] = __uint_as_float(r[i]) + __uint_as_float(r2[i]);
        } else {
            tmem_ld_wait_dep(r);
#pragma unroll
            for (int i = 0; i < 32; i++) act[i] = __uint_as_float(r[i]);
            // the next 32 columns travel from TMEM while this block is converted and stored
            if (cb + 1 < COUT / 32) tmem_ld32(tcol + (uint32_t)(cb + 1) * 32u, r);
            else {   // the accumulators are in registers: hand the TMEM columns back before the last block's conversion
                tc_fence_before();
                __syncwarp();
                if (lane == 0) release();
            }
        }
#pragma unroll
        for (int i = 0; i < 32; i++) {
            const float v = fmaf(act[i], p.out_scale, p.bias[cb * 32 + i]);     // = ACT_SCALE * (conv + bias)
            act[i] = fmaxf(v, 0.1f * v);                                         // leaky 0.1: min(v,0)*0.1 + max(v,0)
        }
        if constexpr (!FUSE) {
            epilogue_store32_rec<F8>(act, tmap_out, p.dbg, stg, lane, gx0, gy0, cb);   // (row coordinates of the store map are window-relative, never negative)
        } else {
            // last layer folded in: accumulate the nine tap dot products over these 32 channels
#pragma unroll
            for (int g = 0; g < 8; g++) {
#pragma unroll
                for (int t = 0; t < 9; t++) {
                    const float *w = p.last_w + t * COUT + cb * 32 + 4 * g;   // compile-time offsets into the parameter bank
                    pt[t] = fmaf(act[4 * g + 0], w[0], pt[t]);
                    pt[t] = fmaf(act[4 * g + 1], w[1], pt[t]);
                    pt[t] = fmaf(act[4 * g + 2], w[2], pt[t]);
                    pt[t] = fmaf(act[4 * g + 3], w[3], pt[t]);
                }
            }
        }
    }
    if constexpr (FUSE) {
        if (inside) {
            float4 *dst = reinterpret_cast<float4 *>(p.partial + ((size_t)fy * p.Wp + fx) * 12);
            dst[0] = make_float4(pt[0], pt[1], pt[2], pt[3]);
            dst[1] = make_float4(pt[4], pt[5], pt[6], pt[7]);
            dst[2] = make_float4(pt[8], 0.f, 0.f, 0.f);
        }
    }
    if constexpr (STACK) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release();
    }
}
