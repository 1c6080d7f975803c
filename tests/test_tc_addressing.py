"""Index arithmetic of the tcgen05 engine, checked on the CPU against a byte-level model of what the
hardware units do with shared memory:

  * TMA (SWIZZLE_64B / SWIZZLE_128B) writes box element e at  swz(box_base + logical_offset(e));
  * a UMMA K-major descriptor (start S, stride-byte-offset SBO) reads GEMM row r, K element k at
    swz(S + (r // 8) * SBO + (r % 8) * ROWB + 2k),

with swz(a) = a ^ (((a >> 7) & (ROWB/16 - 1)) << 4) applied to the shared-memory ADDRESS
(CUTLASS: Swizzle<B,4,3> o smem_ptr).  The test proves that the descriptor offsets used in
csrc/tc_kernel.cuh ((ky*18 + 8j + kx)*ROWB + 32*s, SBO = 18*ROWB) address exactly the 3x3-shifted
windows of the ONE staged 18x18 box -- i.e. that no per-tap reload is needed."""
import numpy as np
import pytest

HALO = 18


def swz(a, rowb):
    return a ^ (((a >> 7) & (rowb // 16 - 1)) << 4)


@pytest.mark.parametrize("kc", [32, 64])
def test_tap_windows_come_from_one_staged_box(kc):
    rowb = kc * 2
    rng = np.random.default_rng(kc)
    box = rng.integers(1, 60000, size=(HALO, HALO, kc), dtype=np.uint16)   # [hy][hx][c] as TMA delivers it
    base = 7 * 1024                                                         # any 1024-aligned shared address
    smem = np.zeros(64 * 1024, np.uint16)                                   # indexed in 2-byte units
    for hy in range(HALO):
        for hx in range(HALO):
            for c in range(kc):
                smem[swz(base + (hy * HALO + hx) * rowb + 2 * c, rowb) // 2] = box[hy, hx, c]
    sbo = HALO * rowb
    for ky in range(3):
        for kx in range(3):
            for j in range(2):
                for s in range(kc // 16):
                    start = base + (ky * HALO + 8 * j + kx) * rowb + 32 * s
                    for r in range(128):
                        oy, ox = r // 8, r % 8
                        for k in (0, 5, 15):
                            a = swz(start + (r // 8) * sbo + (r % 8) * rowb + 2 * k, rowb)
                            assert smem[a // 2] == box[oy + ky, 8 * j + ox + kx, 16 * s + k]


def test_descriptor_fields_fit():
    # 14-bit fields in 16-byte units: addresses < 256 KiB, SBO 18*128 = 2304
    assert (227 * 1024) >> 4 < (1 << 14) and (HALO * 128) >> 4 < (1 << 14)


def test_epilogue_staging_tile_is_the_tma_store_image():
    """The epilogue (csrc/tc_epilogue.cuh::epilogue_store32) writes each lane's 32 channels of ONE pixel into the warp's
    staging tile at 16-byte units XOR-ed with ((lane >> 1) & 3) (fp16 rows of 64 B) / ((lane >> 2) & 1) (e4m3 rows of
    32 B); the TMA store (SWIZZLE_64B / SWIZZLE_32B box {32 ch, 8 px, 4 rows}) reads box element e from
    swz(tile + logical_offset(e)).  Both must describe the same bytes: pixel (h, w) of the box = lane h*8 + w."""
    rng = np.random.default_rng(5)
    tile = 9 * 1024                                         # any 512-byte aligned shared address
    smem = np.zeros(32 * 1024, np.uint8)
    px16 = rng.integers(0, 256, size=(32, 64), dtype=np.uint8)     # [lane][64 B]: 32 channels fp16
    px8 = rng.integers(0, 256, size=(32, 32), dtype=np.uint8)      # [lane][32 B]: 32 channels e4m3
    for lane in range(32):
        for v in range(4):                                  # the kernel's sts128 addresses
            a = tile + lane * 64 + ((v ^ ((lane >> 1) & 3)) << 4)
            smem[a:a + 16] = px16[lane, 16 * v:16 * v + 16]
        for c in range(2):
            a = tile + 2048 + lane * 32 + ((c ^ ((lane >> 2) & 1)) << 4)
            smem[a:a + 16] = px8[lane, 16 * c:16 * c + 16]
    for h in range(4):
        for w in range(8):
            lane = h * 8 + w
            for byte in range(64):                          # what the TMA unit reads for box element (c, w, h)
                assert smem[swz(tile + (h * 8 + w) * 64 + byte, 64)] == px16[lane, byte]
            for byte in range(32):
                assert smem[swz(tile + 2048 + (h * 8 + w) * 32 + byte, 32)] == px8[lane, byte]


def test_first_layer_staging_tile_is_the_tma_store_image():
    """first_layer_kernel: 256 threads = 8 rows x 32 pixels, row index r = threadIdx.x; fp16 units at (c8 ^ ((r>>1)&3)),
    e4m3 8-byte halves at unit ((c8>>1) ^ ((r>>2)&1)), half c8 & 1; TMA box {32 ch, 32 px, 8 rows}."""
    rng = np.random.default_rng(6)
    tile = 3 * 1024
    smem = np.zeros(40 * 1024, np.uint8)
    px16 = rng.integers(0, 256, size=(256, 64), dtype=np.uint8)
    px8 = rng.integers(0, 256, size=(256, 32), dtype=np.uint8)
    for r in range(256):
        for c8 in range(4):
            a = tile + r * 64 + ((c8 ^ ((r >> 1) & 3)) << 4)
            smem[a:a + 16] = px16[r, 16 * c8:16 * c8 + 16]
            a8 = tile + 16384 + r * 32 + (((c8 >> 1) ^ ((r >> 2) & 1)) << 4) + (c8 & 1) * 8
            smem[a8:a8 + 8] = px8[r, 8 * c8:8 * c8 + 8]
    for r in range(256):
        for byte in range(64):
            assert smem[swz(tile + r * 64 + byte, 64)] == px16[r, byte]
        for byte in range(32):
            assert smem[swz(tile + 16384 + r * 32 + byte, 32)] == px8[r, byte]
