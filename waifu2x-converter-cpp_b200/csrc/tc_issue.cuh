// tc_issue.cuh -- what the MMA issuers of the 16x16-tile kernels share (single-CTA tc_kernel.cuh: cta_group::1, CTA-pair
// tc_pair_kernel.cuh: cta_group::2): the walk over the nine taps of a staged 18x18 box of records and the MMAs of one tap in the
// default precision.  Part of the tcgen05 engine's single translation unit (kernels_tc.cu), after tc_config.cuh.

// UMMA<PAIR>: the cta_group::1 / cta_group::2 spellings of the same three instructions
template <bool PAIR>
struct UMMA {
    static __device__ __forceinline__ void f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
        if constexpr (PAIR) umma2_f16(d, a, b, idesc, acc); else umma_f16(d, a, b, idesc, acc);
    }
    static __device__ __forceinline__ void f8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
        if constexpr (PAIR) umma2_f8(d, a, b, idesc, acc); else umma_f8(d, a, b, idesc, acc);
    }
    static __device__ __forceinline__ void commit_one(uint32_t bar) {
        if constexpr (PAIR) umma2_commit_one(bar); else umma_commit_one(bar);
    }
};

// descriptor start-field step to the next tap of an 18-wide box: kx + 1, or the first column of the next halo row   [16-byte units]
template <int ROWB>
__device__ __forceinline__ uint32_t tap_step(int t) { return (t % 3 == 2) ? ((HALO - 2) * ROWB >> 4) : (ROWB >> 4); }

// One tap of one 32-channel step in W2X_PRECISION_F16_F8X2: xh*wh (two K = 16 steps) + xl8*wh8 + xh8*wl8 (K = 32 each) from ONE
// weight stage [wh fp16 rows of 64 B | wh8 | wl8 rows of 32 B]; `rows` = B rows this CTA holds (Cout, or Cout/2 in a pair).
// A record's quarters: +0 / +2 the fp16 K steps, +4 xh8, +6 xl8   [16-byte units].
template <bool PAIR>
__device__ __forceinline__ void issue_tap_f8(uint32_t d, uint32_t ah, uint32_t b0, uint32_t rows, uint32_t a_hi32, uint32_t b_hi32, uint32_t b8_hi32,
                                             uint32_t idesc, uint32_t acc0) {
    auto desc = [](uint32_t hi32, uint32_t lo32) { return ((uint64_t)hi32 << 32) | (uint64_t)lo32; };
    UMMA<PAIR>::f16(d, desc(a_hi32, ah), desc(b_hi32, b0), idesc, acc0);
    UMMA<PAIR>::f16(d, desc(a_hi32, ah + 2u), desc(b_hi32, b0 + 2u), idesc, 1u);
    UMMA<PAIR>::f8(d, desc(a_hi32, ah + 6u), desc(b8_hi32, b0 + (rows * 64u >> 4)), idesc, 1u);    // xl8 * wh8
    UMMA<PAIR>::f8(d, desc(a_hi32, ah + 4u), desc(b8_hi32, b0 + (rows * 96u >> 4)), idesc, 1u);    // xh8 * wl8
}
