"""The CPU oracle (oracle/w2x_oracle.c) pinned against the golden vectors that the reference's own
arithmetic backend produced (oracle/gen_golden.py -> tests/golden/, OpenCV through cv2)."""
import json

import numpy as np
import pytest

from conftest import golden_path

# OpenCV's SIMD filter2D and the scalar C restatement differ only by fp32 re-association / FMA.
ORACLE_VS_CV2_TOL = 3e-6


@pytest.mark.parametrize("name,kind", [("scale2.0x", "uniform"), ("scale2.0x", "smooth"),
                                       ("noise1", "uniform"), ("noise2", "uniform")])
def test_cfg1_256_matches_cv2_golden(oracle_mod, oracle_models, ncpu, name, kind):
    x = oracle_mod.seeded_plane(256, 256, 0, kind)
    y = oracle_models[name].convert(x, n_job=ncpu)
    g = np.load(golden_path(f"cfg1_{name}_{kind}.npy"))
    assert y.shape == g.shape == (256, 256)
    assert np.abs(y - g).max() <= ORACLE_VS_CV2_TOL


def test_odd_sizes_match_cv2_golden(oracle_mod, oracle_models, ncpu):
    z = np.load(golden_path("odd_sizes.npz"))
    for (w, h) in ((1, 1), (15, 13), (37, 61)):
        x = oracle_mod.seeded_plane(w, h, 10 + w, "uniform")
        y = oracle_models["scale2.0x"].convert(x, n_job=ncpu)
        assert y.shape == (h, w)
        assert np.abs(y - z[f"out_{w}x{h}"]).max() <= ORACLE_VS_CV2_TOL


def test_per_layer_filter_matches_cv2_golden(oracle_models, ncpu):
    z = np.load(golden_path("layers_32x24.npz"))
    om = oracle_models["scale2.0x"]
    for li in range(len(om)):
        out = om.filter(li, z[f"in{li}"], n_job=ncpu)
        assert out.shape == z[f"out{li}"].shape
        assert np.abs(out - z[f"out{li}"]).max() <= ORACLE_VS_CV2_TOL, li


def test_block_tables_match_reference_arithmetic(oracle_mod):
    tabs = json.load(open(golden_path("block_tables.json")))
    for key, t in tabs.items():
        w, h = map(int, key.split("x"))
        tab, sc, sr = oracle_mod.block_table(w, h)
        assert (sc, sr) == (t["split_cols"], t["split_rows"]), key
        assert tab.tolist() == [list(r) for r in t["rows"]], key
        # the blocks tile the output exactly (square default block)
        cover = np.zeros((h, w), np.int32)
        for (r, c, y0, y1, x0, x1, oy, ox) in tab:
            cover[oy:oy + (y1 - y0 - 14), ox:ox + (x1 - x0 - 14)] += 1
        assert cover.min() == 1 and cover.max() == 1, key


def test_known_block_counts(oracle_mod):
    # 1920x1080 -> 4x3, 3840x2160 -> 8x5, 4096^2 -> 9x9, 8192^2 -> 17x17
    for (w, h, sc, sr) in ((1920, 1080, 4, 3), (3840, 2160, 8, 5), (4096, 4096, 9, 9), (8192, 8192, 17, 17)):
        _, c, r = oracle_mod.block_table(w, h)
        assert (c, r) == (sc, sr)


def test_split_equals_nosplit_small_blocks(oracle_mod, oracle_models, ncpu):
    """convertWithModelsBlockSplit and the unsplit path agree (same operands per pixel)."""
    x = oracle_mod.seeded_plane(150, 131, 3, "uniform")
    om = oracle_models["noise1"]
    a = om.convert(x, block_splitting=True, block=(64, 64), n_job=ncpu)     # 150*131 > 64*64*1.5 -> split
    b = om.convert(x, block_splitting=False, n_job=ncpu)
    assert np.array_equal(a, b)


@pytest.mark.slow
def test_split_513x768_matches_cv2_golden(oracle_mod, oracle_models, ncpu):
    z = np.load(golden_path("split_513x768.npz"))
    x = oracle_mod.seeded_plane(513, 768, 5, "uniform")
    y = oracle_models["scale2.0x"].convert(x, n_job=ncpu)
    assert np.abs(y[::16, ::16] - z["lattice"]).max() <= ORACLE_VS_CV2_TOL
    assert np.abs(y[494:502, :] - z["rows_494_502"]).max() <= ORACLE_VS_CV2_TOL
    assert np.abs(y[:, 494:502] - z["cols_494_502"]).max() <= ORACLE_VS_CV2_TOL


def test_strided_input_roi(oracle_mod, oracle_models, ncpu):
    big = oracle_mod.seeded_plane(80, 40, 9, "uniform")
    roi = big[5:30, 7:50]
    a = oracle_models["noise2"].convert(roi, n_job=ncpu)
    b = oracle_models["noise2"].convert(np.ascontiguousarray(roi), n_job=ncpu)
    assert np.array_equal(a, b)
