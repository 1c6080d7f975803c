// umma_microbench.cu -- how fast does B200 retire tcgen05.mma when nothing else is in the way?
// Operands sit in shared memory (no loads), 1 or 2 issuing threads loop over back-to-back MMAs into private
// accumulators, one commit at the end.  Reports cycles per MMA for
//   cta_group::1  M=128  N in {64,128,256}   (K=16 fp16, K=32 e4m3)
//   cta_group::2  M=256  N in {128,256}      (each CTA holds half the B rows)
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_microbench tools/umma_microbench.cu
// Run under gpurun:  tools/umma_microbench > gpurun_out/umma_microbench.txt
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = clock64();
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__host__ __device__ constexpr uint64_t desc_const(uint32_t sbo, uint32_t layout) {
    return ((uint64_t)1u << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1u << 46) | ((uint64_t)(layout & 7u) << 61);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

template <int CG, bool F8>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id) {
    if constexpr (CG == 1 && !F8) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
    if constexpr (CG == 1 && F8) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
    if constexpr (CG == 2 && !F8) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
    if constexpr (CG == 2 && F8) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
}

// ISSUERS threads (one per warp 0..ISSUERS-1) each issue `iters` MMAs; rows of A: 128 x 128 B (SWIZZLE_128B, 64 fp16),
// B: N rows x 64 B (SWIZZLE_64B).  The K=16/32 slice is rotated through the row so consecutive MMAs read different bytes.
// Halo mode (rowb > 0) replays the conv kernel's operand addressing: A is one 18x18-pixel box with rowb bytes per pixel
// (swizzle span = rowb), an M=128 tile is 16 image rows of 8 pixels (SBO = 18*rowb), the nine taps are start-address
// offsets (ky*18 + kx)*rowb, and a row holds `nslice` 32-byte K slices.  rowb = 0: dense 128-byte rows, aligned atoms.
template <int CG, int N, int ISSUERS, bool F8>
__global__ void __launch_bounds__(128, 1) bench_kernel(int iters, unsigned long long *out, uint32_t rowb, uint32_t nslice, uint32_t a_lay, uint32_t b_lay,
                                                       uint32_t b_rowb) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = base, b_base = base + 4 * 16384, bar = b_base + 65536, slot = bar + 64;
    uint32_t *slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (slot - smem_u32(smem_raw)));
    for (uint32_t i = threadIdx.x; i < (4 * 16384 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank = 0;
    if constexpr (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        mbar_init(bar, ISSUERS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if constexpr (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 3) {
        if constexpr (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot_ptr;
    long long t0 = 0, t1 = 0;
    if (warp < ISSUERS && lane == 0 && rank == 0) {
        const uint32_t A_HI = (uint32_t)(desc_const(rowb ? 18u * rowb : (a_lay == 2 ? 1024u : a_lay == 4 ? 512u : 256u), a_lay) >> 32);
        const uint32_t B_HI = (uint32_t)(desc_const(8u * b_rowb, b_lay) >> 32);
        const uint32_t a0 = (((a_base + (rowb ? (uint32_t)warp * 8u * rowb : (uint32_t)warp * 16384u)) >> 4) & 0x3FFF) | (1u << 16);
        const uint32_t b0 = ((b_base >> 4) & 0x3FFF) | (1u << 16);
        const uint32_t d = tmem + (uint32_t)warp * 256u;                    // private accumulator columns
        const uint32_t id = idesc(CG == 2 ? 256 : 128, N) | (F8 ? 0u : 0u);
        uint32_t tap = 0, sl = 0;
        t0 = clock64();
        for (int i = 0; i < iters; i++) {
            const uint32_t off = rowb ? ((tap / 3u) * 18u + (tap % 3u)) * rowb : 0u;
            mma<CG, F8>(d, ((uint64_t)A_HI << 32) | (a0 + ((off + sl * 32u) >> 4)), ((uint64_t)B_HI << 32) | (b0 + ((sl * 32u) % b_rowb >> 4)), id);
            if (++sl == nslice) {
                sl = 0;
                if (++tap == 9u) tap = 0;
            }
        }
        if constexpr (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
    }
    mbar_wait(bar, 0);
    t1 = clock64();
    if (warp == 0 && lane == 0 && rank == 0 && blockIdx.x < 2) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if constexpr (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    else __syncthreads();
    if (warp == 3) {
        if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int CG, int N, int ISSUERS, bool F8>
static void run(const char *name, int grid, uint32_t rowb = 0, uint32_t nslice = 2, uint32_t a_lay = 2, uint32_t b_lay = 4, uint32_t b_rowb = 64) {
    const int smem = 1024 + 4 * 16384 + 65536 + 256, iters = 4096;
    cudaFuncSetAttribute(bench_kernel<CG, N, ISSUERS, F8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    unsigned long long *d, h[2] = {0, 0};
    cudaMalloc(&d, 16);
    cudaMemset(d, 0, 16);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; rep++) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, bench_kernel<CG, N, ISSUERS, F8>, iters, d, rowb, nslice, a_lay, b_lay, b_rowb);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            std::printf("%-44s  ERROR %s\n", name, cudaGetErrorString(e));
            cudaGetLastError();
            cudaFree(d);
            return;
        }
    }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const double per = (double)h[0] / (iters * ISSUERS);
    const double macs = (double)(CG == 2 ? 256 : 128) * N * (F8 ? 32 : 16);
    std::printf("%-44s  grid %3d  %7.1f cycles per MMA  (%5.0f MAC/clk per SM; fp16 dense peak 4096, e4m3 8192)\n", name, grid, per, macs / per / CG);
    cudaFree(d);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    std::printf("tcgen05.mma back-to-back issue, operands resident in shared memory, %d SMs\n", sms);
    for (int grid : {2, sms}) {
        run<1, 64, 1, false>("cta_group::1 M128 N64  f16  1 issuer", grid);
        run<1, 128, 1, false>("cta_group::1 M128 N128 f16  1 issuer", grid);
        run<1, 128, 2, false>("cta_group::1 M128 N128 f16  2 issuers", grid);
        run<1, 256, 1, false>("cta_group::1 M128 N256 f16  1 issuer", grid);
        run<1, 256, 2, false>("cta_group::1 M128 N256 f16  2 issuers", grid);
        run<1, 128, 2, true>("cta_group::1 M128 N128 e4m3 2 issuers", grid);
        run<2, 128, 1, false>("cta_group::2 M256 N128 f16  1 issuer", grid);
        run<2, 128, 2, false>("cta_group::2 M256 N128 f16  2 issuers", grid);
        run<2, 256, 1, false>("cta_group::2 M256 N256 f16  1 issuer", grid);
        run<2, 256, 2, false>("cta_group::2 M256 N256 f16  2 issuers", grid);
        run<2, 128, 2, true>("cta_group::2 M256 N128 e4m3 2 issuers", grid);
    }
    // the conv kernel's operand addressing (one halo box, taps = start offsets), 2 issuers, full grid
    std::printf("halo-box operand addressing (A = 18x18 px box, SBO = 18*rowb, tap offsets), 2 issuers\n"
                "NOTE: these rows are bounded by THIS loop's own per-MMA address arithmetic in the single issuing thread (~150 cycles per\n"
                "MMA per issuer, i.e. ~75 with two) -- every shape reads the same.  That observation is what exposed the issue-side\n"
                "bottleneck of the conv kernel (descriptor operands in vector registers -> ELECT/R2UR waterfall per MMA, DESIGN.md 3.2);\n"
                "the pipe rates are the dense rows above.\n");
    run<1, 32, 2, false>("halo f16  rowb 64  SW64  N32  (L1 main)", sms, 64, 2, 4, 4, 64);
    run<1, 64, 2, false>("halo f16  rowb 64  SW64  N64  (L2 main)", sms, 64, 2, 4, 4, 64);
    run<1, 64, 2, false>("halo f16  rowb 128 SW128 N64  (L3 main)", sms, 128, 4, 2, 4, 64);
    run<1, 128, 2, false>("halo f16  rowb 128 SW128 N128 (L4 main)", sms, 128, 4, 2, 4, 64);
    run<1, 32, 2, true>("halo e4m3 rowb 32  SW32  N32  (L1 corr)", sms, 32, 1, 6, 6, 32);
    run<1, 64, 2, true>("halo e4m3 rowb 32  SW32  N64  (L2 corr)", sms, 32, 1, 6, 6, 32);
    run<1, 64, 2, true>("halo e4m3 rowb 64  SW64  N64  (L3 corr)", sms, 64, 2, 4, 6, 32);
    run<1, 128, 2, true>("halo e4m3 rowb 64  SW64  N128 (L4 corr)", sms, 64, 2, 4, 6, 32);
    run<1, 128, 2, true>("halo e4m3 rowb 128 SW128 N128 (L5 corr)", sms, 128, 4, 2, 6, 32);
    run<1, 32, 2, false>("dense f16 SW128 N32", sms);
    run<1, 64, 2, false>("dense f16 SW128 N64", sms);
    run<1, 32, 1, false>("dense f16 SW128 N32 1 issuer", sms);
    run<1, 64, 2, false>("dense f16 rowb-64-like: SW64 aligned atoms N64", sms, 0, 2, 4, 4, 64);
    return 0;
}
