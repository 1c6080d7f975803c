// tc_epilogue.cuh -- the epilogue store: accumulator block -> frame planes -> swizzled staging tile -> TMA store
// Part of the tcgen05 engine's single translation unit: included by kernels_tc.cu inside namespace w2x::tc, in this order:
//   tc_ptx.cuh, tc_config.cuh, tc_epilogue.cuh, tc_kernel.cuh, tc_pair_kernel.cuh, tc_edge_kernels.cuh
// (pure code organisation: the generated SASS is the same as with one file).

// Epilogue store of 32 activated output channels of ONE pixel per thread (lane = pixel inside this warp's 4x8 pixel
// block of an M-tile).  The warp converts to the frame's planes, writes them into its 4 KB staging tile in the TMA
// swizzle pattern (conflict-free 16-byte stores) and one lane issues TMA stores of the 8x4-pixel boxes: the bytes leave
// asynchronously while the warp converts the next 32 channels, the frame edge is clipped by the TMA unit, and the
// shared-memory pipe (which the tensor core's operand fetches saturate) sees one pass instead of a store + load round trip.
//   tile + 0    : fp16 plane (hi | xh), 32 px x 64 B, SWIZZLE_64B
//   tile + 2048 : f16x3: lo plane, same shape (one store of a {32, 8, 4, 2} box covers both planes)
//                 F8   : xh8 (32 px x 32 B) then xl8 at +1024, SWIZZLE_32B, one {32, 8, 4, 2} box of the e4m3 tensor
template <int COUT, bool F8>
__device__ __forceinline__ void epilogue_store32(const float (&act)[32], const CUtensorMap *tmo, const CUtensorMap *tmo8, int dbg, uint32_t stg,
                                                 int lane, int gx0, int gy0, int cb) {
    uint32_t g0[16], g1[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const float v0 = act[2 * i], v1 = act[2 * i + 1];      // already x ACT_SCALE (folded into out_scale / bias)
        __half2 h = __floats2half2_rn(v0, v1);
        float2 hf = __half22float2(h);
        g0[i] = *reinterpret_cast<uint32_t *>(&h);
        if constexpr (F8) {
            // xh8 = e4m3(xh * 2^-F8_C) in g1[0..7], xl8 = e4m3((x16 - xh) * 2^F8_A) in g1[8..15]
            constexpr float kDown = 1.0f / (float)(1 << F8_C), kUp = (float)(1 << F8_A);
            const __half2 hd = __hmul2(h, __float2half2_rn(kDown));        // exact (power of two), one op for both channels
            const uint32_t h8 = __nv_cvt_halfraw2_to_fp8x2(static_cast<__half2_raw>(hd), __NV_SATFINITE, __NV_E4M3);
            const uint32_t l8 = __nv_cvt_float2_to_fp8x2(make_float2((v0 - hf.x) * kUp, (v1 - hf.y) * kUp), __NV_SATFINITE, __NV_E4M3);
            if (i & 1) { g1[i >> 1] |= h8 << 16; g1[8 + (i >> 1)] |= l8 << 16; }
            else { g1[i >> 1] = h8; g1[8 + (i >> 1)] = l8; }
        } else {
            __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
            g1[i] = *reinterpret_cast<uint32_t *>(&l);
        }
    }
    if (dbg & 2) {   // timing experiment: conversion only
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) x ^= g0[i] ^ g1[i];
        if (x == 0x7fc12345u) sts128(stg, make_uint4(x, x, x, x));
        return;
    }
    bulk_wait_read();          // the previous boxes of this tile are on their way
    __syncwarp();
    const uint32_t sw64 = (uint32_t)((lane >> 1) & 3), sw32 = (uint32_t)((lane >> 2) & 1);
#pragma unroll
    for (int v = 0; v < 4; v++)
        sts128(stg + (uint32_t)lane * 64u + (((uint32_t)v ^ sw64) << 4), make_uint4(g0[4 * v], g0[4 * v + 1], g0[4 * v + 2], g0[4 * v + 3]));
    if constexpr (F8) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
            sts128(stg + 2048u + (uint32_t)lane * 32u + (((uint32_t)c ^ sw32) << 4), make_uint4(g1[4 * c], g1[4 * c + 1], g1[4 * c + 2], g1[4 * c + 3]));
            sts128(stg + 3072u + (uint32_t)lane * 32u + (((uint32_t)c ^ sw32) << 4), make_uint4(g1[8 + 4 * c], g1[9 + 4 * c], g1[10 + 4 * c], g1[11 + 4 * c]));
        }
    } else {
#pragma unroll
        for (int v = 0; v < 4; v++)
            sts128(stg + 2048u + (uint32_t)lane * 64u + (((uint32_t)v ^ sw64) << 4), make_uint4(g1[4 * v], g1[4 * v + 1], g1[4 * v + 2], g1[4 * v + 3]));
    }
    fence_proxy_async();       // generic-proxy writes -> visible to the TMA unit
    __syncwarp();
    if (!(dbg & 1)) {
        tma_store_4d(tmo, stg, cb * 32, gx0, gy0, 0);
        if constexpr (F8) tma_store_4d(tmo8, stg + 2048u, cb * 32, gx0, gy0, 0);
        bulk_commit();
    }
}

// The same for RECORD frames ([Hp][Wp][C/32][128 B], one 128-byte record per pixel per 32-channel block:
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
