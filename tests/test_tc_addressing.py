"""Index arithmetic of the tcgen05 engine, checked on the CPU against a byte-level model of what the
hardware units do with shared memory:

  * TMA (SWIZZLE_64B / SWIZZLE_128B) writes box element e at  swz(box_base + logical_offset(e));
  * a UMMA K-major descriptor (start S, stride-byte-offset SBO) reads GEMM row r, K element k at
    swz(S + (r // 8) * SBO + (r % 8) * ROWB + 2k),

with swz(a) = a ^ (((a >> 7) & (ROWB/16 - 1)) << 4) applied to the shared-memory ADDRESS
(CUTLASS: Swizzle<B,4,3> o smem_ptr).  The test proves that the descriptor offsets used in
csrc/kernels_tc.cu ((ky*18 + 8j + kx)*ROWB + 32*s, SBO = 18*ROWB) address exactly the 3x3-shifted
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
