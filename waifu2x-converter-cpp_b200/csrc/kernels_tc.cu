// kernels_tc.cu -- the tcgen05 engine (W2X_ENGINE_TC): sm_100a only.
//
// What it computes (reference src/modelHandler.cpp:134-152 for all output planes of a layer at
// once): out[o](y,x) = leaky( sum_i sum_{ky,kx} W[o][i][ky][kx] * in[i](y+ky-1, x+kx-1) + bias[o] ).
//
// How: implicit GEMM, D[pixel][o] += A[pixel][(tap,i)] * B[(tap,i)][o], on the 5th-generation
// tensor cores (tcgen05.mma, fp32 accumulators in TMEM).  fp32 fidelity comes from a 2-term split of
// both operands (x = xh + xl, w = wh + wl, h = the fp16 rounding) and three accumulated products
// xh*wh + xl*wh + xh*wl (the dropped xl*wl term is ~2^-22 relative); SURVEY.md section 7 shows a
// single fp16/tf32 pass misses the 1e-4 gate by 10x.  Two arithmetic modes (template flag F8):
//   f16x3      all three products as kind::f16 MMAs on fp16 hi/lo planes
//   f16+f8x2   xh*wh as kind::f16; the two correction products as kind::f8f6f4 MMAs on e4m3 copies of the
//              operands (K = 32 per instruction, twice the rate) -- the default, 2.0 instead of 3.0 passes
//
// Data layout in HBM: every activation is an NHWC "frame" of 4 bytes per element holding value*ACT_SCALE,
// [hi fp16][lo fp16] or [xh fp16][xh8 e4m3][xl8 e4m3] planes of [Hp][Wp][C]; all layers of one pass share
// the frame size (the padded plane), reads outside the frame are zero-filled by TMA, so each layer is a
// same-size convolution whose polluted ring grows by one pixel per layer and is cropped at the end -- the
// same argument that makes the reference's per-layer BORDER_REPLICATE harmless.
//
// Per CTA (persistent, 1 per SM, 12 warps):
//   warp 0      A producer   one TMA box {KC ch, 18, 18} per (tile-set, channel chunk, plane): the
//                            16x16 output region plus a 1-pixel ring, staged ONCE and addressed nine
//                            times (the 3x3 taps are UMMA-descriptor start-address offsets into it)
//   warps 1, 7  MMA issuers  one per M-tile (8 wide x 16 tall pixels): tcgen05.mma, M=128 (M=256 across a CTA
//                            pair for the 128-wide layers), N=Cout, K=16 (fp16) / 32 (e4m3) per instruction;
//                            operands provably warp-uniform, so the MMAs issue back to back from uniform registers
//   warp 2      B producer   pre-swizzled 32-channel weight stages: a cp.async.bulk ring, or resident for the
//                            narrow layers; owns TMEM
//   warps 3-6, 8-11 epilogue one set per M-tile: tcgen05.ld -> scale, +bias, leaky-ReLU -> either the frame's
//                            planes, staged in the TMA swizzle pattern and TMA-stored, or (FUSE) the last
//                            layer's nine tap partials; overlaps the next tile-set (TMEM double buffer)
//
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstdio>

#include "kernels.h"

namespace w2x {
namespace tc {

// ================================================================================================
// PTX wrappers
// ================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// Producer-side helpers are called by a whole converged warp; ONE elected lane executes the instruction (elect.sync inside
// the asm block).  For the TMA / tcgen05 instructions this is what lets ptxas keep their operands in uniform registers.
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Spin on try_wait; a protocol bug must not hang the GPU, so give up (trap -> launch error) after ~4 s.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t spins = 0;; spins++) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if ((spins & 1023u) == 1023u) {
            long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 8000000000LL) __trap();
        }
    }
}
// Non-blocking probe of a phase: issued EARLY (before the MMAs of the current stage) so that the ~100-cycle
// mbarrier round trip of the next stage's wait overlaps with issue work instead of draining the tensor queue.
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMA: 4-D tiled load global -> shared, completion on an mbarrier
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// bulk (1-D) copy global -> shared, completion on an mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
                 : "memory");
}
// TMA store of one 4-D box shared -> global (bulk async-group completion); whole warp calls, one elected lane issues
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n\t}"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q cp.async.bulk.commit_group;\n\t}" ::: "memory");
}
// Every lane waits for ITS OWN bulk groups (lanes that issued none return at once), so whichever lane the elect.sync of
// tma_store_4d / bulk_commit picked is covered; callers follow with __syncwarp().
// ... have finished READING shared memory (the staging tile may be rewritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... have completed (before the CTA exits)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// The issuing WARP walks the MMA loop converged; each tcgen05.mma / commit is predicated by an elect.sync inside its asm
// block.  Together with a shuffle-derived (provably uniform) warp index and TMEM base this lets ptxas keep every
// descriptor in uniform registers and emit back-to-back UTC*MMA -- a lane predicate or a thread-derived operand
// costs an ELECT / R2UR / BRA.U.ANY waterfall of ~20 dependent instructions per MMA (measured ~140 cycles per MMA per
// issuer: that, not the tensor pipe, was what bounded the narrow layers).
#define W2X_UMMA_VARIANT(NAME, OPCODE)                                                                        \
    __device__ __forceinline__ void NAME(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,      \
                                         uint32_t accum) {                                                     \
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t@q " OPCODE \
                     " [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),                                                \
                     "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)                                            \
                     : "memory");                                                                              \
    }
W2X_UMMA_VARIANT(umma_f16, "tcgen05.mma.cta_group::1.kind::f16")
W2X_UMMA_VARIANT(umma_f8, "tcgen05.mma.cta_group::1.kind::f8f6f4")   // e4m3 x e4m3 -> f32, K = 32 per instruction, twice the f16 rate
#undef W2X_UMMA_VARIANT
__device__ __forceinline__ void umma_commit_one(uint32_t bar) {   // whole (converged) warp calls, one elected lane commits
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}

// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp receives TMEM lane (base_lane + i), 32 consecutive columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait::ld that also names the destination registers of the load it waits for, so the compiler cannot move their first
// use above the wait when another tcgen05.ld has already been issued in between (software-pipelined epilogue)
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

// ================================================================================================
// Descriptors
// ================================================================================================
// Shared-memory matrix descriptor (K-major, swizzled).  Field layout as in CUTLASS
// cute/arch/mma_sm100_desc.hpp (UMMA::SmemDescriptor): start address >>4 in [0,14), leading byte
// offset >>4 in [16,30), stride byte offset >>4 in [32,46), version=1 in [46,48), base_offset in
// [49,52), layout type in [61,64) (2 = SWIZZLE_128B, 4 = SWIZZLE_64B).
// Canonical K-major layout, 16-byte units: ((8, n), 2) : ((ROWB/16, SBO), 1) -- eight rows ROWB
// bytes apart form a group, groups are SBO bytes apart, the swizzle XOR is a function of the
// shared-memory ADDRESS bits (Swizzle<B,4,3> o smem_ptr), which is what lets a descriptor start
// anywhere inside a TMA-written box.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)1u << 16;                               // LBO: unused for swizzled K-major; CUTLASS writes 1
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1u << 46;                               // descriptor version (Blackwell)
    d |= (uint64_t)(base_off & 7u) << 49;
    d |= (uint64_t)(layout_type & 7u) << 61;
    return d;
}

// Everything of a descriptor except the start address (compile-time part).
__host__ __device__ constexpr uint64_t make_desc_const(uint32_t sbo_bytes, uint32_t layout_type) {
    return ((uint64_t)1u << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1u << 46) |
           ((uint64_t)(layout_type & 7u) << 61);
}

// Instruction descriptor (UMMA::InstrDescriptor): c_format F32 (1) at [4,6), a/b format F16 (0) at
// [7,10)/[10,13), a/b major K (0) at 15/16, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ================================================================================================
// Per-layer configuration
// ================================================================================================
// Channels per staged activation box: 32 for layers up to 64 inputs (two small boxes per tile-set instead of one 83 KB
// one leave room for the store staging and a deep weight ring), 64 for the 128-input layers.
__host__ __device__ constexpr int act_kc(int cin) { return cin <= 64 ? 32 : 64; }

// F8 = false: three kind::f16 products xh*wh + xl*wh + xh*wl ("f16x3").
// F8 = true : xh*wh in kind::f16, the two correction products in kind::f8f6f4 on e4m3 copies
//             xl8*wh8 + xh8*wl8 (K = 32 per MMA at twice the rate: 2.0 instead of 3.0 pass-equivalents).
//             Activation frames then hold [xh fp16][xh8][xl8] planes (same 4 bytes per element).
constexpr int F8_A = 10, F8_C = 1;   // xl8 = e4m3((x16 - xh) * 2^F8_A), xh8 = e4m3(xh * 2^-F8_C); must match w2x_internal.h

template <int CIN, int COUT, bool FUSE = false, bool F8 = false>
struct Cfg {
    // ---- A operand (activations): one TMA box per (tile-set, 64-channel chunk, hi|lo) ----
    static constexpr int KC = act_kc(CIN);              // channels per activation chunk
    static constexpr int NCHUNK = CIN / KC;
    static constexpr int ROWB = KC * 2;                 // bytes per pixel per chunk (= swizzle span)
    static constexpr uint32_t A_LAYOUT = ROWB == 128 ? 2u : 4u;                // SWIZZLE_128B : SWIZZLE_64B
    static constexpr int A_PLANE = HALO * HALO * ROWB;                       // bytes one TMA box delivers
    static constexpr int A_PLANE_PAD = (A_PLANE + 1023) / 1024 * 1024;
    static constexpr int ROWB8 = KC;                                         // e4m3 planes: one byte per channel
    static constexpr uint32_t A8_LAYOUT = ROWB8 == 64 ? 4u : 6u;               // SWIZZLE_64B : SWIZZLE_32B
    static constexpr int A8_PLANE = HALO * HALO * ROWB8;
    static constexpr int A8_PLANE_PAD = (A8_PLANE + 1023) / 1024 * 1024;
    static constexpr int A_SLOT = F8 ? A_PLANE_PAD + 2 * A8_PLANE_PAD : 2 * A_PLANE_PAD;   // xh + (xh8, xl8)  |  hi + lo
    static constexpr int A_TX = F8 ? A_PLANE + 2 * A8_PLANE : 2 * A_PLANE;   // bytes the TMA loads of one slot deliver
    static constexpr int A_SLOTS = 2;
    // ---- B operand (weights): stages of 32 input channels (two K=16 steps), SWIZZLE_64B rows of 64 B ----
    static constexpr int KB = 32;
    static constexpr int KBLOCKS = KC / KB;             // weight stages per (chunk, tap, part)
    static constexpr int B_ROWB = KB * 2;
    static constexpr uint32_t B_LAYOUT = 4u;
    // Cout <= 64: hi and lo weights form ONE stage of 2*Cout rows, so xh*[wh;wl] is a single N = 2*Cout MMA
    // (accumulators D1 | D2 side by side, summed in the epilogue) -- two MMAs per K step instead of three.
    static constexpr bool STACK = COUT <= 64 && !F8;
    static constexpr int B_BLOCK = COUT * B_ROWB;                            // one (chunk, tap, kblock, hi|lo) block
    // F8: per 32-channel block ONE stage [wh fp16 (Cout x 64 B) | wh8 | wl8 (e4m3, Cout x 32 B each)]: four MMAs per
    // issuer per barrier round trip.
    static constexpr bool MERGE = F8 && (COUT <= 64 || FUSE);   // (a stored 128-in/128-out layer has no room for 16 KB stages beside its two 83 KB activation slots)
    static constexpr int B_STAGE = (STACK || MERGE) ? 2 * B_BLOCK : B_BLOCK;
    static constexpr int STAGES_PER_TILESET = NCHUNK * 9 * KBLOCKS * ((STACK || MERGE) ? 1 : 2);
    // ---- accumulators ----
    static constexpr int TILE_COLS = STACK ? 2 * COUT : COUT;                // TMEM columns per M-tile
    static constexpr int ACC_COLS = 4 * TILE_COLS;                           // 2 sets x 2 M-tiles
    static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
    // ---- shared memory map: [A slots][B stages][barriers + bias (1 KB)][last-layer weights][store staging] ----
    static constexpr int BAR_BYTES = 1024;
    static constexpr int W6_BYTES = 0;                                       // (the fused last layer's weights travel as kernel parameters)
    static constexpr int STG_WARP = 4096;                                    // [fp16 plane 2 KB | lo plane 2 KB, or xh8 1 KB | xl8 1 KB] of 32 px x 32 ch
    static constexpr int STG_BYTES = FUSE ? 0 : 8 * STG_WARP;               // epilogue store staging per epilogue warp
    static constexpr int SMEM_MAX = 227 * 1024;
    static constexpr int NB_FIT = (SMEM_MAX - 1024 - BAR_BYTES - W6_BYTES - STG_BYTES - A_SLOTS * A_SLOT) / B_STAGE;
    // Narrow layers: ALL weight stages of a tile-set fit -> loaded once per CTA and kept (no ring traffic, no stage barriers
    // after the first tile-set; the TMA unit is left to the activation boxes and the epilogue's stores).
    static constexpr bool RESIDENT = STAGES_PER_TILESET <= NB_FIT && STAGES_PER_TILESET <= 24;
    static constexpr int NB = RESIDENT ? STAGES_PER_TILESET : (NB_FIT > 8 ? 8 : NB_FIT);
    static constexpr int SMEM_BYTES = 1024 + A_SLOTS * A_SLOT + NB * B_STAGE + BAR_BYTES + W6_BYTES + STG_BYTES;
    static_assert(NB >= 3, "need at least three weight stages");
    static_assert((8 + 2 * NB) * 8 + 4 <= 512 && COUT * 4 <= 512, "barrier/bias area overflow");
    static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
    static_assert(B_STAGE % 512 == 0, "weight stage must keep the 512-byte SWIZZLE_64B pattern alignment");
    static_assert(CIN % KC == 0 && KC % KB == 0 && COUT % 16 == 0 && COUT <= 128, "shape");
};

// warps: 0 A producer | 1, 7 MMA issuers (M-tile 0, 1) | 2 B producer + TMEM owner | 3-6 epilogue of M-tile 0 | 8-11 epilogue of M-tile 1
constexpr int NUM_THREADS = 12 * 32;

struct TcParams {
    const uint16_t *wpack;   // [chunk][tap][kblock][hi|lo][COUT rows x 64 B], pre-swizzled (see model.cpp)
    float bias[128];         // [COUT] (float)bias, by value (constant bank, see last_w)
    __half *out;             // [2][Hp][Wp][COUT]
    int Wp, Hp;
    int tiles_x, n_tilesets;
    float out_scale;         // 1 / wscale  (accumulator -> ACT_SCALE * conv)
    unsigned long long *prof;   // optional [gridDim.x][16] cycle counters (see PROF_* below), nullptr = off
    int dbg;                    // always 0 in product builds; -DW2X_EPI_EXPERIMENTS + W2X_DEBUG_EPI: 1 = no global stores, 2 = no staging either (timing only, results WRONG)
    // fused last layer (FUSE kernels only): this layer's activations never reach HBM; instead each pixel's
    // nine tap partials P[t] = sum_c act[c] * w_last[c][t] are written ([Hp][Wp][12] fp32, 3 pad words).
    float *partial;             // nullptr = not fused
    float last_w[9 * 128];      // [9][COUT] tap-major, by value: the epilogue's FFMAs read them straight from the constant
                                // bank (kernel parameters), which keeps 288 broadcast LDS.128 per pixel off the shared-memory
                                // pipe the tensor core's operand fetches saturate
};

// per-CTA profile record (cycles, accumulated over launches)
enum { PROF_TOTAL = 0, PROF_MMA_WAIT_ACC, PROF_MMA_WAIT_A, PROF_MMA_WAIT_B, PROF_APROD_WAIT, PROF_BPROD_WAIT,
       PROF_EPI_WAIT, PROF_EPI_WORK, PROF_TILESETS, PROF_N = 16 };

__device__ __forceinline__ void mbar_wait_prof(uint32_t bar, uint32_t parity, bool on, unsigned long long &acc) {
    if (on) {
        long long t0 = clock64();
        mbar_wait(bar, parity);
        acc += (unsigned long long)(clock64() - t0);
    } else {
        mbar_wait(bar, parity);
    }
}

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

// ================================================================================================
// The layer kernel
// ================================================================================================
template <int CIN, int COUT, bool FUSE, bool F8>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_conv3x3_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_in8,
                  const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_out8, const TcParams p) {
    using C = Cfg<CIN, COUT, FUSE, F8>;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment: swizzle patterns repeat every 1024 B (SWIZZLE_128B) / 512 B (SWIZZLE_64B)
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
        if constexpr (F8) prefetch_tmap(&tmap_in8);
        if constexpr (!FUSE) {
            prefetch_tmap(&tmap_out);
            if constexpr (F8) prefetch_tmap(&tmap_out8);
        }
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
        // ===================== A producer: one halo'd box per (tile-set, chunk, hi|lo) ==============
        // (whole warp walks the loop; the arrive and the TMA instructions elect one lane)
        {
            uint32_t it = 0;
            unsigned long long w_a = 0;
            for (int ts = blockIdx.x; ts < p.n_tilesets; ts += gridDim.x) {
                const int ty = ts / p.tiles_x, tx = ts - ty * p.tiles_x;
                const int x0 = tx * REGION - 1, y0 = ty * REGION - 1;   // box origin incl. ring (may be -1)
                for (int c = 0; c < C::NCHUNK; c++, it++) {
                    const uint32_t slot = it & 1u, round = it >> 1;
                    mbar_wait_prof(a_empty(slot), (round & 1u) ^ 1u, prof_on, w_a);
                    mbar_arrive_expect_tx(a_full(slot), (uint32_t)C::A_TX);
                    const uint32_t dst = a_base + slot * C::A_SLOT;
                    tma_load_4d(dst, &tmap_in, a_full(slot), c * C::KC, x0, y0, 0);
                    if constexpr (F8) {
                        tma_load_4d(dst + C::A_PLANE_PAD, &tmap_in8, a_full(slot), c * C::KC, x0, y0, 0);                     // xh8
                        tma_load_4d(dst + C::A_PLANE_PAD + C::A8_PLANE_PAD, &tmap_in8, a_full(slot), c * C::KC, x0, y0, 1);   // xl8
                    } else {
                        tma_load_4d(dst + C::A_PLANE_PAD, &tmap_in, a_full(slot), c * C::KC, x0, y0, 1);
                    }
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
        // e4m3 operands (F8): activation planes with ROWB8-byte rows, weight blocks with 32-byte rows (SWIZZLE_32B)
        constexpr uint32_t A8_HI32 = (uint32_t)(make_desc_const(HALO * C::ROWB8, C::A8_LAYOUT) >> 32);
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
                // descriptor low words (address >> 4) of this issuer's window into the hi / lo activation planes
                const uint32_t ah0 = ((((a_base + slot * C::A_SLOT) >> 4) & 0x3FFFu) | LO_FIXED) + jt * (8u * C::ROWB >> 4);
                const uint32_t al0 = ah0 + (C::A_PLANE_PAD >> 4);
                // F8: windows into the xh8 / xl8 planes of this slot
                const uint32_t a8h0 = ((((a_base + slot * C::A_SLOT + C::A_PLANE_PAD) >> 4) & 0x3FFFu) | LO_FIXED) + jt * (8u * C::ROWB8 >> 4);
                const uint32_t a8l0 = a8h0 + (C::A8_PLANE_PAD >> 4);
                uint32_t tap_off = 0;                 // ((ky*HALO + kx) * ROWB) >> 4
                uint32_t tap_off8 = 0;                // ((ky*HALO + kx) * ROWB8) >> 4
                for (int t = 0; t < 9; t++) {
                    const uint32_t first = (c | t) != 0 ? 1u : 0u;
#pragma unroll
                    for (int kb = 0; kb < C::KBLOCKS; kb++) {
                        const uint32_t ah = ah0 + tap_off + 4u * kb, al = al0 + tap_off + 4u * kb;   // 32 channels = 64 B = 4 units
                        const uint32_t acc0 = kb ? 1u : first;
                        uint32_t b0;
                        if constexpr (C::MERGE) {
                            // one stage = [wh fp16 | wh8 | wl8]: main product (two K=16 steps) + both e4m3 corrections (K=32 each)
                            acquire_b(b0);
                            umma_f16(dj, desc(A_HI32, ah), desc(B_HI32, b0), idesc_c, acc0);
                            umma_f16(dj, desc(A_HI32, ah + 2u), desc(B_HI32, b0 + 2u), idesc_c, 1u);
                            umma_f8(dj, desc(A8_HI32, a8l0 + tap_off8 + 2u * kb), desc(B8_HI32, b0 + (COUT * 64u >> 4)), idesc_c, 1u);
                            umma_f8(dj, desc(A8_HI32, a8h0 + tap_off8 + 2u * kb), desc(B8_HI32, b0 + (COUT * 96u >> 4)), idesc_c, 1u);
                            release_b();
                        } else if constexpr (F8) {
                            // ---- stage 1: wh (fp16): the main product xh*wh, two K=16 steps ----
                            acquire_b(b0);
                            umma_f16(dj, desc(A_HI32, ah), desc(B_HI32, b0), idesc_c, acc0);
                            umma_f16(dj, desc(A_HI32, ah + 2u), desc(B_HI32, b0 + 2u), idesc_c, 1u);
                            release_b();
                            // ---- stage 2: [wh8 | wl8] (e4m3): corrections xl8*wh8 and xh8*wl8, one K=32 step each ----
                            acquire_b(b0);
                            umma_f8(dj, desc(A8_HI32, a8l0 + tap_off8 + 2u * kb), desc(B8_HI32, b0), idesc_c, 1u);
                            umma_f8(dj, desc(A8_HI32, a8h0 + tap_off8 + 2u * kb), desc(B8_HI32, b0 + (COUT * 32u >> 4)), idesc_c, 1u);
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
                    tap_off += (t % 3 == 2) ? ((HALO - 2) * C::ROWB >> 4) : (C::ROWB >> 4);
                    tap_off8 += (t % 3 == 2) ? ((HALO - 2) * C::ROWB8 >> 4) : (C::ROWB8 >> 4);
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
            const int fy = ty * REGION + oy, fx = tx * REGION + 8 * j + ox;
            const bool inside = fy < p.Hp && fx < p.Wp;
            float pt[9];                                   // FUSE: nine per-tap dot products of this pixel
#pragma unroll
            for (int t = 0; t < 9; t++) pt[t] = 0.f;
            uint32_t r[32];
            if constexpr (!C::STACK) tmem_ld32(tcol, r);
#pragma unroll
            for (int cb = 0; cb < COUT / 32; cb++) {
                // ---- 32 output channels of this pixel: accumulator -> scale, bias, leaky-ReLU ----
                float act[32];
                if constexpr (C::STACK) {
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
                        if (lane == 0) mbar_arrive(acc_empty(set));
                    }
                }
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const float v = fmaf(act[i], p.out_scale, p.bias[cb * 32 + i]);     // = ACT_SCALE * (conv + bias)
                    act[i] = fmaxf(v, 0.1f * v);                                         // leaky 0.1: min(v,0)*0.1 + max(v,0)
                }
                if constexpr (!FUSE) {
                    epilogue_store32<COUT, F8>(act, &tmap_out, &tmap_out8, p.dbg, stg, lane, tx * REGION + 8 * j, ty * REGION + 4 * (int)q, cb);
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
            if constexpr (C::STACK) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty(set));
            }
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
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma_load_4d_2cta(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
#define W2X_UMMA2_VARIANT(NAME, OPCODE)                                                                       \
    __device__ __forceinline__ void NAME(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,      \
                                         uint32_t accum) {                                                     \
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t@q " OPCODE \
                     " [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),                                                \
                     "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)                                            \
                     : "memory");                                                                              \
    }
W2X_UMMA2_VARIANT(umma2_f16, "tcgen05.mma.cta_group::2.kind::f16")
W2X_UMMA2_VARIANT(umma2_f8, "tcgen05.mma.cta_group::2.kind::f8f6f4")
#undef W2X_UMMA2_VARIANT
__device__ __forceinline__ void umma2_commit_one(uint32_t bar) {   // arrives on `bar` in BOTH CTAs of the pair; one elected lane commits
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(bar),
        "h"((uint16_t)3)
        : "memory");
}

template <int CIN, int COUT, bool FUSE, bool F8>
struct PairCfg : Cfg<CIN, COUT, FUSE, F8> {
    using Base = Cfg<CIN, COUT, FUSE, F8>;
    static_assert(COUT == 128, "the CTA-pair kernel is built for the 128-wide layers");
    static constexpr int B_HALF = Base::B_BLOCK;                         // bytes of one weight stage held by ONE CTA: its 64 rows of BOTH blocks
                                                                         // of a 32-channel step ([hi | lo] or [wh | wh8 | wl8])
    static constexpr int NBP_FIT = (Base::SMEM_MAX - 1024 - Base::BAR_BYTES - Base::W6_BYTES - Base::STG_BYTES - Base::A_SLOTS * Base::A_SLOT) / B_HALF;
    static constexpr int NBP = NBP_FIT > 12 ? 12 : NBP_FIT;
    static constexpr int SMEM_BYTES = 1024 + Base::A_SLOTS * Base::A_SLOT + NBP * B_HALF + Base::BAR_BYTES + Base::W6_BYTES + Base::STG_BYTES;
    static_assert((8 + 2 * NBP) * 8 + 4 <= 512, "barrier area overflow");
    static_assert(B_HALF % 2048 == 0, "weight halves are moved as 2 KB TMA boxes");
};

template <int CIN, int COUT, bool FUSE, bool F8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
tc_conv3x3_pair_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_in8,
                       const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_out,
                       const __grid_constant__ CUtensorMap tmap_out8, const TcParams p) {
    using C = PairCfg<CIN, COUT, FUSE, F8>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base;
    const uint32_t b_base = a_base + C::A_SLOTS * C::A_SLOT;
    const uint32_t bar_base = b_base + C::NBP * C::B_HALF;
    auto a_full = [&](int i) { return bar_base + 8u * (uint32_t)i; };
    auto a_empty = [&](int i) { return bar_base + 8u * (uint32_t)(2 + i); };
    auto acc_full = [&](int i) { return bar_base + 8u * (uint32_t)(4 + i); };
    auto acc_empty = [&](int i) { return bar_base + 8u * (uint32_t)(6 + i); };
    auto b_full = [&](int i) { return bar_base + 8u * (uint32_t)(8 + i); };
    auto b_empty = [&](int i) { return bar_base + 8u * (uint32_t)(8 + C::NBP + i); };
    const uint32_t tmem_slot = bar_base + 8u * (uint32_t)(8 + 2 * C::NBP);
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
    const int tiles_y = (p.Hp + REGION - 1) / REGION;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(a_full(i), 1);        // leader's: its A producer's arrive.expect_tx covers the bytes of BOTH CTAs
            mbar_init(a_empty(i), 2);       // one multicast tcgen05.commit per issuer
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
        if constexpr (F8) prefetch_tmap(&tmap_in8);
        if constexpr (!FUSE) {
            prefetch_tmap(&tmap_out);
            if constexpr (F8) prefetch_tmap(&tmap_out8);
        }
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
                const int x0 = tx * REGION - 1, y0 = ty * REGION - 1;
                for (int c = 0; c < C::NCHUNK; c++, it++) {
                    const uint32_t slot = it & 1u, round = it >> 1;
                    mbar_wait_prof(a_empty(slot), (round & 1u) ^ 1u, prof_on, w_a);
                    if (is_leader) mbar_arrive_expect_tx(a_full(slot), 2u * (uint32_t)C::A_TX);
                    const uint32_t bar = mapa_rank(a_full(slot), 0);
                    const uint32_t dst = a_base + slot * C::A_SLOT;
                    tma_load_4d_2cta(dst, &tmap_in, bar, c * C::KC, x0, y0, 0);
                    if constexpr (F8) {
                        tma_load_4d_2cta(dst + C::A_PLANE_PAD, &tmap_in8, bar, c * C::KC, x0, y0, 0);
                        tma_load_4d_2cta(dst + C::A_PLANE_PAD + C::A8_PLANE_PAD, &tmap_in8, bar, c * C::KC, x0, y0, 1);
                    } else {
                        tma_load_4d_2cta(dst + C::A_PLANE_PAD, &tmap_in, bar, c * C::KC, x0, y0, 1);
                    }
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
            constexpr int N_STEPS = C::NCHUNK * 9 * C::KBLOCKS;                // one stage per (chunk, tap, 32-channel block)
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
            constexpr uint32_t A8_HI32 = (uint32_t)(make_desc_const(HALO * C::ROWB8, C::A8_LAYOUT) >> 32);
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
                    const uint32_t slot = a_it & 1u;
                    mbar_wait_prof(a_full(slot), (a_it >> 1) & 1u, prof_on, w_af);
                    tc_fence_after();
                    const uint32_t ah0 = ((((a_base + slot * C::A_SLOT) >> 4) & 0x3FFFu) | LO_FIXED) + jt * (8u * C::ROWB >> 4);
                    const uint32_t al0 = ah0 + (C::A_PLANE_PAD >> 4);
                    const uint32_t a8h0 = ((((a_base + slot * C::A_SLOT + C::A_PLANE_PAD) >> 4) & 0x3FFFu) | LO_FIXED) + jt * (8u * C::ROWB8 >> 4);
                    const uint32_t a8l0 = a8h0 + (C::A8_PLANE_PAD >> 4);
                    uint32_t tap_off = 0, tap_off8 = 0;
                    for (int t = 0; t < 9; t++) {
                        const uint32_t first = (c | t) != 0 ? 1u : 0u;
#pragma unroll
                        for (int kb = 0; kb < C::KBLOCKS; kb++) {
                            const uint32_t ah = ah0 + tap_off + 4u * kb, al = al0 + tap_off + 4u * kb;
                            const uint32_t acc0 = kb ? 1u : first;
                            uint32_t b0;
                            acquire_b(b0);                  // one stage per 32-channel step: this CTA's rows of both blocks
                            if constexpr (F8) {
                                umma2_f16(dj, desc(A_HI32, ah), desc(B_HI32, b0), idesc_c, acc0);
                                umma2_f16(dj, desc(A_HI32, ah + 2u), desc(B_HI32, b0 + 2u), idesc_c, 1u);
                                umma2_f8(dj, desc(A8_HI32, a8l0 + tap_off8 + 2u * kb), desc(B8_HI32, b0 + (4096u >> 4)), idesc_c, 1u);
                                umma2_f8(dj, desc(A8_HI32, a8h0 + tap_off8 + 2u * kb), desc(B8_HI32, b0 + (6144u >> 4)), idesc_c, 1u);
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
                        tap_off += (t % 3 == 2) ? ((HALO - 2) * C::ROWB >> 4) : (C::ROWB >> 4);
                        tap_off8 += (t % 3 == 2) ? ((HALO - 2) * C::ROWB8 >> 4) : (C::ROWB8 >> 4);
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
            const int fy = ty * REGION + oy, fx = tx * REGION + 8 * j + ox;
            const bool inside = fy < p.Hp && fx < p.Wp;
            float pt[9];
#pragma unroll
            for (int t = 0; t < 9; t++) pt[t] = 0.f;
            uint32_t r[32];
            tmem_ld32(tcol, r);
#pragma unroll
            for (int cb = 0; cb < COUT / 32; cb++) {
                float act[32];
                tmem_ld_wait_dep(r);
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const float v = fmaf(__uint_as_float(r[i]), p.out_scale, p.bias[cb * 32 + i]);   // = ACT_SCALE * (conv + bias)
                    act[i] = fmaxf(v, 0.1f * v);                                                       // leaky 0.1
                }
                // the next 32 columns travel from TMEM while this block is converted and stored
                if (cb + 1 < COUT / 32) tmem_ld32(tcol + (uint32_t)(cb + 1) * 32u, r);
                else {   // the accumulators are in registers: hand the TMEM columns back before the last block's conversion
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(mapa_rank(acc_empty(set), 0));
                }
                if constexpr (FUSE) {
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
                } else {
                    epilogue_store32<COUT, F8>(act, &tmap_out, &tmap_out8, p.dbg, stg, lane, tx * REGION + 8 * j, ty * REGION + 4 * (int)q4, cb);
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

// ================================================================================================
// First layer (Cin = 1), last layer (Cout = 1), layout converters -- CUDA-core, HBM-bound
// ================================================================================================
// First layer: Model::filterWorker with nInputPlanes = 1 on the (already replicate-padded) plane;
// writes the NHWC frame the tcgen05 layers consume.  One thread per pixel, 32 x 8 pixels per block.
//   * weights and biases travel as kernel parameters: the 9*COUT FFMAs per pixel take them straight from the constant
//     bank (as shared-memory broadcasts they were one LDS per FFMA -- the LSU, not the FP32 pipe, bounded the kernel);
//   * 32 channels at a time are converted into a swizzled shared-memory image of the block's 8 x 32 pixels and leave
//     through TMA stores (the same path as the tcgen05 epilogue): a thread's own 16-byte stores sat at a 64-byte stride.
template <int COUT>
struct FirstParams {
    float w[COUT * 9];    // [COUT][3][3]
    float b[COUT];        // (float)bias
};
constexpr int FIRST_TILE_BYTES = 32 * 1024;   // [fp16 plane 16 KB | lo plane 16 KB]  or  [xh 16 KB | xh8 8 KB | xl8 8 KB]

template <int COUT, bool F8>
__global__ void __launch_bounds__(256, 4)
first_layer_kernel(const float *__restrict__ in, long in_stride, int pw, int ph, const __grid_constant__ CUtensorMap tmap_out,
                   const __grid_constant__ CUtensorMap tmap_out8, const __grid_constant__ FirstParams<COUT> prm) {
    extern __shared__ uint8_t first_smem[];
    const uint32_t tile = (smem_u32(first_smem) + 1023u) & ~1023u;
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    const int x = blockIdx.x * 32 + lane, y = blockIdx.y * 8 + wy;     // threads past the frame edge compute clamped copies; TMA clips them
    float v[9];
#pragma unroll
    for (int ky = 0; ky < 3; ky++)
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
            int gy = min(max(y + ky - 1, 0), ph - 1), gx = min(max(x + kx - 1, 0), pw - 1);
            v[ky * 3 + kx] = __ldg(in + (long)gy * in_stride + gx);
        }
    const uint32_t r = (uint32_t)threadIdx.x;                            // pixel index inside the block = row of the staged image
    const uint32_t sw64 = (r >> 1) & 3u, sw32 = (r >> 2) & 1u;
#pragma unroll 1
    for (int cb = 0; cb < COUT / 32; cb++) {
        if (cb) {   // the previous 32 channels' boxes must have left shared memory
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncthreads();
        }
#pragma unroll
        for (int c8 = 0; c8 < 4; c8++) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float a[2];
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const float *w = prm.w + (cb * 32 + c8 * 8 + 2 * i + e) * 9;
                    float t = w[0] * v[0];
#pragma unroll
                    for (int k = 1; k < 9; k++) t = fmaf(w[k], v[k], t);
                    float rr = (0.f + t) + prm.b[cb * 32 + c8 * 8 + 2 * i + e];
                    a[e] = (fminf(rr, 0.f) * 0.1f + fmaxf(rr, 0.f)) * ACT_SCALE;
                }
                __half2 h = __floats2half2_rn(a[0], a[1]);
                float2 hf = __half22float2(h);
                hi[i] = *reinterpret_cast<uint32_t *>(&h);
                if constexpr (F8) {
                    constexpr float kDown = 1.0f / (float)(1 << F8_C), kUp = (float)(1 << F8_A);
                    const uint32_t h8 = __nv_cvt_float2_to_fp8x2(make_float2(hf.x * kDown, hf.y * kDown), __NV_SATFINITE, __NV_E4M3);
                    const uint32_t l8 = __nv_cvt_float2_to_fp8x2(make_float2((a[0] - hf.x) * kUp, (a[1] - hf.y) * kUp), __NV_SATFINITE, __NV_E4M3);
                    if (i & 1) { lo[i >> 1] |= h8 << 16; lo[2 + (i >> 1)] |= l8 << 16; }     // lo[0..1] = xh8 (8 bytes), lo[2..3] = xl8
                    else { lo[i >> 1] = h8; lo[2 + (i >> 1)] = l8; }
                } else {
                    __half2 l = __floats2half2_rn(a[0] - hf.x, a[1] - hf.y);
                    lo[i] = *reinterpret_cast<uint32_t *>(&l);
                }
            }
            // 16-byte unit c8 of this pixel's 64-byte fp16 row (SWIZZLE_64B image)
            sts128(tile + r * 64u + (((uint32_t)c8 ^ sw64) << 4), make_uint4(hi[0], hi[1], hi[2], hi[3]));
            if constexpr (F8) {
                // 8 bytes of the pixel's 32-byte e4m3 rows (SWIZZLE_32B images): unit c8/2, half c8%2
                const uint32_t off = r * 32u + ((((uint32_t)c8 >> 1) ^ sw32) << 4) + ((uint32_t)c8 & 1u) * 8u;
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(tile + 16384u + off), "r"(lo[0]), "r"(lo[1]) : "memory");
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(tile + 24576u + off), "r"(lo[2]), "r"(lo[3]) : "memory");
            } else {
                sts128(tile + 16384u + r * 64u + (((uint32_t)c8 ^ sw64) << 4), make_uint4(lo[0], lo[1], lo[2], lo[3]));
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&tmap_out)), "r"(tile), "r"(cb * 32), "r"(x0), "r"(y0), "r"(0) : "memory");
            if constexpr (F8)
                asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                             ::"l"(reinterpret_cast<uint64_t>(&tmap_out8)), "r"(tile + 16384u), "r"(cb * 32), "r"(x0), "r"(y0), "r"(0) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");    // shared memory stays valid until the boxes are out
}

// Last layer: nOutputPlanes = 1.  fp32 arithmetic in the reference's association: per input plane a
// 9-tap sum, planes added in ascending order, then bias and leaky-ReLU.  One thread per pixel.
template <int CIN, bool F8>
__global__ void __launch_bounds__(256)
last_layer_kernel(const __half *__restrict__ in, int pw, int ph, const float *__restrict__ wgt, float bias, int crop,
                  float *__restrict__ dst, long dst_stride) {
    __shared__ float s_w[CIN * 9];
    for (int i = threadIdx.x; i < CIN * 9; i += blockDim.x) s_w[i] = wgt[i];
    __syncthreads();
    const int x = crop + blockIdx.x * 32 + (threadIdx.x & 31), y = crop + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= pw - crop || y >= ph - crop) return;
    const size_t plane_elems = (size_t)ph * pw * CIN;
    const float inv = 1.0f / ACT_SCALE;
    float acc = 0.f;
    for (int c8 = 0; c8 < CIN / 8; c8++) {
        float t[8];
#pragma unroll
        for (int e = 0; e < 8; e++) t[e] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ky++)
#pragma unroll
            for (int kx = 0; kx < 3; kx++) {
                // frame reads outside [0,pw)x[0,ph) cannot happen: crop >= 1 keeps the 3x3 window inside
                const size_t pixo = ((size_t)(y + ky - 1) * pw + (x + kx - 1)) * CIN + c8 * 8;
                const __half *ph_ = in + pixo;
                uint4 uh = __ldg(reinterpret_cast<const uint4 *>(ph_));
                const __half2 *h2 = reinterpret_cast<const __half2 *>(&uh);
                uint4 ul = make_uint4(0, 0, 0, 0);
                uint2 ul8 = make_uint2(0, 0);
                if constexpr (F8) ul8 = __ldg(reinterpret_cast<const uint2 *>(reinterpret_cast<const uint8_t *>(in) + 3 * plane_elems + pixo));
                else ul = __ldg(reinterpret_cast<const uint4 *>(ph_ + plane_elems));
                const __half2 *l2 = reinterpret_cast<const __half2 *>(&ul);
                const __nv_fp8x2_storage_t *l8 = reinterpret_cast<const __nv_fp8x2_storage_t *>(&ul8);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    float2 hf = __half22float2(h2[i]), lf;
                    if constexpr (F8) {
                        __half2_raw r = __nv_cvt_fp8x2_to_halfraw2(l8[i], __NV_E4M3);
                        lf = __half22float2(*reinterpret_cast<__half2 *>(&r));
                        lf.x *= 1.0f / (float)(1 << F8_A);
                        lf.y *= 1.0f / (float)(1 << F8_A);
                    } else lf = __half22float2(l2[i]);
                    float a0 = (hf.x + lf.x) * inv, a1 = (hf.y + lf.y) * inv;
                    const int tap = ky * 3 + kx;
                    t[2 * i] = fmaf(s_w[(c8 * 8 + 2 * i) * 9 + tap], a0, t[2 * i]);
                    t[2 * i + 1] = fmaf(s_w[(c8 * 8 + 2 * i + 1) * 9 + tap], a1, t[2 * i + 1]);
                }
            }
#pragma unroll
        for (int e = 0; e < 8; e++) acc += t[e];
    }
    float r = acc + bias;
    dst[(long)(y - crop) * dst_stride + (x - crop)] = fminf(r, 0.f) * 0.1f + fmaxf(r, 0.f);
}

// Second half of the fused last layer: out(y,x) = leaky(bias + sum_t P[(y+ky-1, x+kx-1)][t]), taps in
// row-major order, for the interior [crop, ph-crop) x [crop, pw-crop).
__global__ void __launch_bounds__(256)
last_gather_kernel(const float *__restrict__ partial, int pw, int ph, float bias, int crop_x, int crop_top,
                   int crop_bottom, float *__restrict__ dst, long dst_stride) {
    const int x = crop_x + blockIdx.x * 32 + (threadIdx.x & 31), y = crop_top + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= pw - crop_x || y >= ph - crop_bottom) return;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ky++)
#pragma unroll
        for (int kx = 0; kx < 3; kx++)
            acc += __ldg(partial + ((size_t)(y + ky - 1) * pw + (x + kx - 1)) * 12 + ky * 3 + kx);
    const float r = acc + bias;
    dst[(long)(y - crop_top) * dst_stride + (x - crop_x)] = fminf(r, 0.f) * 0.1f + fmaxf(r, 0.f);
}

__global__ void planar_to_nhwc_kernel(const float *__restrict__ in, int C, int w, int h, __half *__restrict__ out, int f8) {
    const int pw = w + 2, ph = h + 2;
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = (long)pw * ph * C;
    if (idx >= total) return;
    int c = (int)(idx % C);
    long pix = idx / C;
    int x = (int)(pix % pw), y = (int)(pix / pw);
    int sx = min(max(x - 1, 0), w - 1), sy = min(max(y - 1, 0), h - 1);
    float a = in[((long)c * h + sy) * w + sx] * ACT_SCALE;
    __half hh = __float2half_rn(a);
    out[idx] = hh;
    if (f8) {
        uint8_t *b = reinterpret_cast<uint8_t *>(out);
        const float hf = __half2float(hh);
        b[2 * total + idx] = (uint8_t)__nv_cvt_float_to_fp8(hf * (1.0f / (float)(1 << F8_C)), __NV_SATFINITE, __NV_E4M3);
        b[3 * total + idx] = (uint8_t)__nv_cvt_float_to_fp8((a - hf) * (float)(1 << F8_A), __NV_SATFINITE, __NV_E4M3);
    } else {
        out[idx + total] = __float2half_rn(a - __half2float(hh));
    }
}

__global__ void nhwc_to_planar_kernel(const __half *__restrict__ in, int C, int w, int h, float *__restrict__ out, int f8) {
    const int pw = w + 2, ph = h + 2;
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = (long)w * h * C;
    if (idx >= total) return;
    int x = (int)(idx % w);
    long r = idx / w;
    int y = (int)(r % h), c = (int)(r / h);
    long src = ((long)(y + 1) * pw + (x + 1)) * C + c;
    long plane = (long)pw * ph * C;
    float lo;
    if (f8) {
        __half_raw r = __nv_cvt_fp8_to_halfraw(reinterpret_cast<const uint8_t *>(in)[3 * plane + src], __NV_E4M3);
        lo = __half2float(*reinterpret_cast<__half *>(&r)) * (1.0f / (float)(1 << F8_A));
    } else lo = __half2float(in[src + plane]);
    out[idx] = (__half2float(in[src]) + lo) * (1.0f / ACT_SCALE);
}

// ================================================================================================
// Host side
// ================================================================================================
bool layer_supported(int cin, int cout) {
    auto ok = [](int c) { return c == 32 || c == 64 || c == 128; };
    return ok(cin) && ok(cout);
}

template <int CIN, int COUT, bool FUSE, bool F8>
static cudaError_t set_attr1() {
    return cudaFuncSetAttribute(tc_conv3x3_kernel<CIN, COUT, FUSE, F8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                Cfg<CIN, COUT, FUSE, F8>::SMEM_BYTES);
}
template <int CIN, int COUT>
static cudaError_t set_attr() {
    cudaError_t e;
    if ((e = set_attr1<CIN, COUT, false, false>()) != cudaSuccess) return e;
    if ((e = set_attr1<CIN, COUT, true, false>()) != cudaSuccess) return e;
    if ((e = set_attr1<CIN, COUT, false, true>()) != cudaSuccess) return e;
    return set_attr1<CIN, COUT, true, true>();
}

#define W2X_TC_SHAPES(X) \
    X(32, 32) X(32, 64) X(32, 128) X(64, 32) X(64, 64) X(64, 128) X(128, 32) X(128, 64) X(128, 128)

size_t layer_smem_bytes(int cin, int cout) {
#define X(ci, co) \
    if (cin == ci && cout == co) return Cfg<ci, co, true, false>::SMEM_BYTES;
    W2X_TC_SHAPES(X)
#undef X
    return 0;
}

template <int CIN, bool FUSE, bool F8>
static cudaError_t set_attr_pair() {
    return cudaFuncSetAttribute(tc_conv3x3_pair_kernel<CIN, 128, FUSE, F8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                PairCfg<CIN, 128, FUSE, F8>::SMEM_BYTES);
}
#define W2X_PAIR_CINS(X) X(32) X(64) X(128)

cudaError_t init_kernels() {
    cudaError_t e;
#define X(ci)                                                            \
    if ((e = set_attr_pair<ci, false, false>()) != cudaSuccess) return e; \
    if ((e = set_attr_pair<ci, true, false>()) != cudaSuccess) return e;  \
    if ((e = set_attr_pair<ci, false, true>()) != cudaSuccess) return e;  \
    if ((e = set_attr_pair<ci, true, true>()) != cudaSuccess) return e;
    W2X_PAIR_CINS(X)
#undef X
#define X(ci, co) \
    if ((e = set_attr<ci, co>()) != cudaSuccess) return e;
    W2X_TC_SHAPES(X)
#undef X
    return cudaSuccess;
}

template <int CIN, int COUT, bool FUSE, bool F8>
static cudaError_t launch_k(const CUtensorMap *tmap, const CUtensorMap *tmap8, const CUtensorMap *p_out_maps, const TcParams &p, int grid, cudaStream_t s) {
    tc_conv3x3_kernel<CIN, COUT, FUSE, F8><<<grid, NUM_THREADS, Cfg<CIN, COUT, FUSE, F8>::SMEM_BYTES, s>>>(*tmap, *tmap8, p_out_maps[0], p_out_maps[1], p);
    return cudaGetLastError();
}

template <int CIN, int COUT>
static cudaError_t launch_one(const CUtensorMap *tmap, const CUtensorMap *tmap8, const CUtensorMap *omaps, const TcParams &p, int num_sms, bool f8, cudaStream_t s) {
    int grid = p.n_tilesets < num_sms ? p.n_tilesets : num_sms;
    if (f8) return p.partial ? launch_k<CIN, COUT, true, true>(tmap, tmap8, omaps, p, grid, s) : launch_k<CIN, COUT, false, true>(tmap, tmap8, omaps, p, grid, s);
    return p.partial ? launch_k<CIN, COUT, true, false>(tmap, tmap8, omaps, p, grid, s) : launch_k<CIN, COUT, false, false>(tmap, tmap8, omaps, p, grid, s);
}

static int make_weight_stream_map(CUtensorMap *map, const void *base, size_t bytes);
static int make_out_tensor_maps(CUtensorMap *map16, CUtensorMap *map8, void *base, int C, int Wp, int Hp, bool f8, int box_w = 8, int box_h = 4);

template <int CIN, bool FUSE, bool F8>
static cudaError_t launch_pair_k(const CUtensorMap *tmap, const CUtensorMap *tmap8, const CUtensorMap *tmapw, const CUtensorMap *p_out_maps,
                                 const TcParams &p, int grid, cudaStream_t s) {
    tc_conv3x3_pair_kernel<CIN, 128, FUSE, F8><<<grid, NUM_THREADS, PairCfg<CIN, 128, FUSE, F8>::SMEM_BYTES, s>>>(*tmap, *tmap8, *tmapw, p_out_maps[0], p_out_maps[1], p);
    return cudaGetLastError();
}

template <int CIN>
static cudaError_t launch_pair(const CUtensorMap *tmap, const CUtensorMap *tmap8, const CUtensorMap *omaps, const TcParams &p, int num_sms, bool f8, cudaStream_t s) {
    using C0 = Cfg<CIN, 128, false, false>;
    const size_t bytes = (size_t)C0::NCHUNK * 9 * C0::KBLOCKS * 2 * C0::B_BLOCK;   // both flavours stream the same number of bytes per tile-set
    CUtensorMap tmapw;
    if (make_weight_stream_map(&tmapw, p.wpack, bytes)) return cudaErrorInvalidValue;
    const int n_pair_sets = (p.n_tilesets + 1) / 2;
    int grid = 2 * (n_pair_sets < num_sms / 2 ? n_pair_sets : num_sms / 2);
    if (f8) return p.partial ? launch_pair_k<CIN, true, true>(tmap, tmap8, &tmapw, omaps, p, grid, s) : launch_pair_k<CIN, false, true>(tmap, tmap8, &tmapw, omaps, p, grid, s);
    return p.partial ? launch_pair_k<CIN, true, false>(tmap, tmap8, &tmapw, omaps, p, grid, s) : launch_pair_k<CIN, false, false>(tmap, tmap8, &tmapw, omaps, p, grid, s);
}

cudaError_t launch_tc_layer(const CUtensorMap *tmap_in, const void *wpack, const float *bias, __half *out, int cin,
                            int cout, int pw, int ph, float out_scale, int f8, int num_sms, cudaStream_t s,
                            unsigned long long *prof, const float *last_w, float *partial, const CUtensorMap *tmap_in8, int pair) {
    TcParams p;
    p.wpack = reinterpret_cast<const uint16_t *>(wpack);
    // ACT_SCALE (a power of two) is folded into the epilogue's affine step: leaky(16 v) = 16 leaky(v) exactly, so the
    // kernel produces the frame's x16 values without a separate multiply; the fused last layer gets weights / 16.
    for (int i = 0; i < cout; i++) p.bias[i] = bias[i] * ACT_SCALE;      // HOST pointer
    p.out = out;
    p.Wp = pw;
    p.Hp = ph;
    p.tiles_x = (pw + REGION - 1) / REGION;
    p.n_tilesets = p.tiles_x * ((ph + REGION - 1) / REGION);
    p.out_scale = out_scale * ACT_SCALE;
    p.prof = prof;
#ifdef W2X_EPI_EXPERIMENTS   // timing experiments only (results are wrong): build with -DW2X_EPI_EXPERIMENTS, then W2X_DEBUG_EPI=1|2
    static const int dbg_epi = std::getenv("W2X_DEBUG_EPI") ? std::atoi(std::getenv("W2X_DEBUG_EPI")) : 0;
    p.dbg = dbg_epi;
#else
    p.dbg = 0;
#endif
    p.partial = partial;
    if (partial) {
        if (!last_w) return cudaErrorInvalidValue;
        for (int i = 0; i < 9 * cout; i++) p.last_w[i] = last_w[i] * (1.0f / ACT_SCALE);      // HOST pointer: [9][cout]
    }
    if (f8 && !tmap_in8) return cudaErrorInvalidValue;
    const CUtensorMap *t8 = tmap_in8 ? tmap_in8 : tmap_in;
    // the epilogue's TMA stores: 8x4-pixel x 32-channel boxes of this layer's output frame (fused layers store no frame)
    CUtensorMap omaps[2];
    if (partial) { omaps[0] = *tmap_in; omaps[1] = *t8; }
    else if (make_out_tensor_maps(&omaps[0], &omaps[1], out, cout, pw, ph, f8 != 0)) return cudaErrorInvalidValue;
    if (pair && cout == 128 && num_sms >= 2) {
#define X(ci) \
    if (cin == ci) return launch_pair<ci>(tmap_in, t8, omaps, p, num_sms, f8 != 0, s);
        W2X_PAIR_CINS(X)
#undef X
    }
#define X(ci, co) \
    if (cin == ci && cout == co) return launch_one<ci, co>(tmap_in, t8, omaps, p, num_sms, f8 != 0, s);
    W2X_TC_SHAPES(X)
#undef X
    return cudaErrorInvalidValue;
}

template <int COUT>
static cudaError_t launch_first_c(const float *in, long in_stride_floats, int pw, int ph, const float *wgt, const float *bias, __half *out,
                                  cudaStream_t s, int f8) {
    static_assert(FIRST_TILE_BYTES + 1024 <= 48 * 1024, "the first layer's staging tile stays under the default dynamic shared memory limit");
    FirstParams<COUT> prm;
    for (int i = 0; i < COUT * 9; i++) prm.w[i] = wgt[i];     // HOST pointers
    for (int i = 0; i < COUT; i++) prm.b[i] = bias[i];
    CUtensorMap omaps[2];
    if (make_out_tensor_maps(&omaps[0], &omaps[1], out, COUT, pw, ph, f8 != 0, 32, 8)) return cudaErrorInvalidValue;
    dim3 grid((pw + 31) / 32, (ph + 7) / 8);
    if (grid.y > 65535) return cudaErrorInvalidConfiguration;
    if (f8) first_layer_kernel<COUT, true><<<grid, 256, FIRST_TILE_BYTES + 1024, s>>>(in, in_stride_floats, pw, ph, omaps[0], omaps[1], prm);
    else first_layer_kernel<COUT, false><<<grid, 256, FIRST_TILE_BYTES + 1024, s>>>(in, in_stride_floats, pw, ph, omaps[0], omaps[1], prm);
    return cudaGetLastError();
}

cudaError_t launch_first(const float *in, long in_stride_floats, int pw, int ph, const float *wgt, const float *bias,
                         int cout, __half *out, cudaStream_t s, int f8) {
    switch (cout) {
        case 32: return launch_first_c<32>(in, in_stride_floats, pw, ph, wgt, bias, out, s, f8);
        case 64: return launch_first_c<64>(in, in_stride_floats, pw, ph, wgt, bias, out, s, f8);
        case 128: return launch_first_c<128>(in, in_stride_floats, pw, ph, wgt, bias, out, s, f8);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_last(const __half *in, int cin, int pw, int ph, const float *wgt, float bias, int crop, float *dst,
                        long dst_stride_floats, cudaStream_t s, int f8) {
    const int ow = pw - 2 * crop, oh = ph - 2 * crop;
    if (ow < 1 || oh < 1 || crop < 1) return cudaErrorInvalidValue;
    dim3 grid((ow + 31) / 32, (oh + 7) / 8);
    if (grid.y > 65535) return cudaErrorInvalidConfiguration;
#define W2X_LAST(C)                                                                                              \
    case C:                                                                                                          \
        if (f8) last_layer_kernel<C, true><<<grid, 256, 0, s>>>(in, pw, ph, wgt, bias, crop, dst, dst_stride_floats);    \
        else last_layer_kernel<C, false><<<grid, 256, 0, s>>>(in, pw, ph, wgt, bias, crop, dst, dst_stride_floats);      \
        break;
    switch (cin) {
        W2X_LAST(32) W2X_LAST(64) W2X_LAST(128)
        default: return cudaErrorInvalidValue;
    }
#undef W2X_LAST
    return cudaGetLastError();
}

cudaError_t launch_last_gather(const float *partial, int pw, int ph, float bias, int crop, float *dst,
                               long dst_stride_floats, cudaStream_t s) {
    return launch_last_gather_xy(partial, pw, ph, bias, crop, crop, crop, dst, dst_stride_floats, s);
}

cudaError_t launch_last_gather_xy(const float *partial, int pw, int ph, float bias, int crop_x, int crop_top,
                                  int crop_bottom, float *dst, long dst_stride_floats, cudaStream_t s) {
    const int ow = pw - 2 * crop_x, oh = ph - crop_top - crop_bottom;
    if (ow < 1 || oh < 1 || crop_x < 1 || crop_top < 1 || crop_bottom < 1) return cudaErrorInvalidValue;
    dim3 grid((ow + 31) / 32, (oh + 7) / 8);
    if (grid.y > 65535) return cudaErrorInvalidConfiguration;
    last_gather_kernel<<<grid, 256, 0, s>>>(partial, pw, ph, bias, crop_x, crop_top, crop_bottom, dst, dst_stride_floats);
    return cudaGetLastError();
}

cudaError_t launch_planar_to_nhwc(const float *in, int C, int w, int h, __half *out, cudaStream_t s, int f8) {
    long total = (long)(w + 2) * (h + 2) * C;
    planar_to_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, C, w, h, out, f8);
    return cudaGetLastError();
}

cudaError_t launch_nhwc_to_planar(const __half *in, int C, int w, int h, float *out, cudaStream_t s, int f8) {
    long total = (long)w * h * C;
    nhwc_to_planar_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, C, w, h, out, f8);
    return cudaGetLastError();
}

// ---- TMA descriptor ----------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// The packed weight stream as rows of 1 KB (256 x 32-bit words), box = 2 rows (2 KB), no swizzle: the CTA-pair kernel's
// B loads.  (Long rows matter: the TMA unit issues one request per box row, and 64-byte rows made the weight ring
// request-bound.)
static int make_weight_stream_map(CUtensorMap *map, const void *base, size_t bytes) {
    PFN_encodeTiled enc = get_encode();
    if (!enc || bytes % 2048) return -1;
    cuuint64_t dims[2] = {256, (cuuint64_t)(bytes / 1024)};
    cuuint64_t strides[1] = {1024};
    cuuint32_t box[2] = {256, 2};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

// Output frame as the epilogue stores it: boxes of 32 channels x 8 px x 4 rows (one epilogue warp's pixels; the first
// layer's blocks store 32 x 8 pixels).
//   f16x3: ONE map over [2][Hp][Wp][C] fp16, box {32, 8, 4, 2} (hi and lo planes in one store), SWIZZLE_64B; map8 = copy.
//   F8:    map16 over the xh plane, box {32, 8, 4, 1}, SWIZZLE_64B; map8 over the two e4m3 planes, box {32, 8, 4, 2}, SWIZZLE_32B.
static int make_out_tensor_maps(CUtensorMap *map16, CUtensorMap *map8, void *base, int C, int Wp, int Hp, bool f8, int box_w, int box_h) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return -1;
    cuuint32_t estr[4] = {1, 1, 1, 1};
    {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)(f8 ? 1 : 2)};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)Wp * C * 2, (cuuint64_t)Hp * Wp * C * 2};
        cuuint32_t box[4] = {32, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)(f8 ? 1 : 2)};
        CUresult r = enc(map16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return (int)r;
    }
    if (!f8) { *map8 = *map16; return 0; }
    char *b8 = reinterpret_cast<char *>(base) + (size_t)2 * Hp * Wp * C;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wp, (cuuint64_t)Hp, 2};
    cuuint64_t strides[3] = {(cuuint64_t)C, (cuuint64_t)Wp * C, (cuuint64_t)Hp * Wp * C};
    cuuint32_t box[4] = {32, (cuuint32_t)box_w, (cuuint32_t)box_h, 2};
    CUresult r = enc(map8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, b8, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

int make_act_tensor_map(CUtensorMap *map, const void *base, int C, int Wp, int Hp) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return -1;
    const int kc = act_kc(C);
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wp, (cuuint64_t)Hp, 2};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)Wp * C * 2, (cuuint64_t)Hp * Wp * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)kc, HALO, HALO, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

// F8 frames: [xh fp16 [Hp][Wp][C]] [xh8 [Hp][Wp][C]] [xl8 [Hp][Wp][C]]  (bytes 2 + 1 + 1 per element)
int make_act_tensor_maps_f8(CUtensorMap *map16, CUtensorMap *map8, const void *base, int C, int Wp, int Hp) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return -1;
    const int kc = act_kc(C);
    cuuint32_t estr[4] = {1, 1, 1, 1};
    {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wp, (cuuint64_t)Hp, 1};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)Wp * C * 2, (cuuint64_t)Hp * Wp * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)kc, HALO, HALO, 1};
        CUresult r = enc(map16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return (int)r;
    }
    {
        const char *b8 = reinterpret_cast<const char *>(base) + (size_t)2 * Hp * Wp * C;
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wp, (cuuint64_t)Hp, 2};
        cuuint64_t strides[3] = {(cuuint64_t)C, (cuuint64_t)Wp * C, (cuuint64_t)Hp * Wp * C};
        cuuint32_t box[4] = {(cuuint32_t)kc, HALO, HALO, 1};
        CUresult r = enc(map8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<char *>(b8), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, kc == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return (int)r;
    }
    return 0;
}

}  // namespace tc
}  // namespace w2x
