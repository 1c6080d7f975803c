"""GPU parity at the sizes BASELINE.json names beyond config 3, in the DEFAULT precision (fp16 + two e4m3 correction
products -- what bench.py, the CLI and every multi-GPU number run), through the C ABI:

  * config 4: 8192x8192 scale pass -- random windows and corners against the CPU oracle on the window + its 7-pixel
    context, and the whole plane against the fp32 CUDA-core engine (itself pinned to the oracle at 5e-6);
  * config 2: 1920x1080 noise1 pass -> 2x nearest upscale -> 3840x2160 scale2.0x pass, the chained error against the
    oracle running the same chain on windows;
  * the adversarial planes of tests/test_numerics_model.py (binary noise, checkerboards, stripes, impulses, saturated and
    out-of-range values: the e4m3 `lo` planes saturate silently) on the GPU instead of the CPU emulation.
Tolerances: GOLD 1e-4 (the stated gate), F8_TOL 6e-5 (what this precision is expected to reach on white noise).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD_TOL = 1e-4
F8_TOL = 6e-5
FP32_TOL = 5e-6


@pytest.fixture(scope="module")
def ctx8(w2x):
    c = w2x.Context(0, engine=w2x.ENGINE_TC)
    c.set_precision(w2x.PRECISION_F16_F8X2)
    f = w2x.Context(0, engine=w2x.ENGINE_FP32)
    yield c, f
    c.close()
    f.close()


@pytest.fixture(scope="module")
def models(w2x, oracle_models):
    return {n: w2x.Model.from_arrays(om.weights, om.biases) for n, om in oracle_models.items()}


def _windows(y, x, om, ncpu, rng, n_win, size=48):
    H, W = x.shape
    worst = 0.0
    for _ in range(n_win):
        x0, y0 = int(rng.integers(7, W - size - 7)), int(rng.integers(7, H - size - 7))
        ref = om.convert(x[y0 - 7:y0 + size + 7, x0 - 7:x0 + size + 7], n_job=ncpu)[7:-7, 7:-7]
        worst = max(worst, float(np.abs(y[y0:y0 + size, x0:x0 + size] - ref).max()))
    ref = om.convert(x[:size + 7, :size + 7], n_job=ncpu)[:size, :size]                 # corners: replicate padding
    worst = max(worst, float(np.abs(y[:size, :size] - ref).max()))
    ref = om.convert(x[-size - 7:, -size - 7:], n_job=ncpu)[-size:, -size:]
    return max(worst, float(np.abs(y[-size:, -size:] - ref).max()))


@pytest.mark.parametrize("name", ["scale2.0x", "noise1", "noise2"])
def test_cfg4_8192_default_precision(ctx8, models, oracle_mod, oracle_models, ncpu, name):
    ctx, fp32 = ctx8
    x = oracle_mod.seeded_plane(8192, 8192, 2, "uniform")
    y = ctx.convert_plane(models[name], x)
    assert np.isfinite(y).all()
    err_w = _windows(y, x, oracle_models[name], ncpu, np.random.default_rng(8), 3)
    assert err_w <= F8_TOL, err_w
    full = float(np.abs(y - fp32.convert_plane(models[name], x)).max())
    print(f"8192x8192 {name}: windows vs oracle {err_w:.2e}, whole plane vs fp32 engine {full:.2e}")
    assert full <= F8_TOL + FP32_TOL, full
    assert full <= GOLD_TOL


def test_cfg2_chain_noise1_then_scale2x_3840x2160(ctx8, models, oracle_mod, oracle_models, ncpu):
    """What src/main.cpp:96,140-148 feeds the path for `-m noise_scale` on a 1920x1080 image: the noise1 pass on the Y
    plane, cv::resize INTER_NEAREST x2, the scale pass on 3840x2160."""
    ctx, _ = ctx8
    x = oracle_mod.seeded_plane(1920, 1080, 6, "uniform")
    d = ctx.convert_plane(models["noise1"], x)
    up = np.ascontiguousarray(np.repeat(np.repeat(d, 2, axis=0), 2, axis=1))
    y = ctx.convert_plane(models["scale2.0x"], up)
    assert y.shape == (2160, 3840)
    rng = np.random.default_rng(3)
    worst = 0.0
    for _ in range(4):
        # a 24x24 window of the input with 11 pixels of context: 7 for the noise pass, 4 (= ceil(7/2)) for the scale pass
        x0, y0 = int(rng.integers(11, 1920 - 35)), int(rng.integers(11, 1080 - 35))
        win = x[y0 - 11:y0 + 24 + 11, x0 - 11:x0 + 24 + 11]
        dn = oracle_models["noise1"].convert(win, n_job=ncpu)[7:-7, 7:-7]            # valid: 4 px of context left
        upw = np.ascontiguousarray(np.repeat(np.repeat(dn, 2, axis=0), 2, axis=1))   # 8 px of context
        ref = oracle_models["scale2.0x"].convert(upw, n_job=ncpu)[8:-8, 8:-8]
        got = y[2 * y0:2 * y0 + 48, 2 * x0:2 * x0 + 48]
        worst = max(worst, float(np.abs(got - ref).max()))
    print(f"cfg2 chain (1080p noise1 -> x2 -> scale2.0x): windows vs oracle chain {worst:.2e}")
    assert worst <= GOLD_TOL
    assert worst <= 1.5 * F8_TOL, worst          # two passes


@pytest.mark.parametrize("name", ["scale2.0x", "noise1", "noise2"])
def test_adversarial_and_out_of_range_planes_on_the_gpu(ctx8, models, oracle_models, ncpu, name):
    ctx, _ = ctx8
    rng = np.random.default_rng(12)
    h, w = 48, 56
    yy, xx = np.mgrid[0:h, 0:w]
    planes = {
        "binary noise": (rng.random((h, w)) > 0.5).astype(np.float32),
        "checkerboard": ((yy + xx) % 2).astype(np.float32),
        "2px stripes": ((xx // 2) % 2).astype(np.float32),
        "impulses": (rng.random((h, w)) > 0.97).astype(np.float32),
        "all ones": np.ones((h, w), np.float32),
        "ramp": (xx / (w - 1)).astype(np.float32),
        "out of range": (rng.random((h, w)) * 3.0 - 1.0).astype(np.float32),
        "far out of range": (rng.random((h, w)) * 40.0 - 20.0).astype(np.float32),
    }
    for label, x in planes.items():
        ref = oracle_models[name].convert(x, n_job=ncpu)
        got = ctx.convert_plane(models[name], x)
        scale = max(1.0, float(np.abs(ref).max()))
        assert np.abs(got - ref).max() <= F8_TOL * scale, (name, label, float(np.abs(got - ref).max()), scale)
