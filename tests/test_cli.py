"""the drop-in CLI (flag surface, messages, exit codes, output naming) and the image
plumbing around the hot path, pinned against OpenCV (cv2) -- the library the reference uses for exactly these calls
(src/main.cpp:74-76, 132-146, 158-167, 171-190)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

cv2 = pytest.importorskip("cv2")
PKG = os.path.join(ROOT, "waifu2x-converter-cpp_b200")
CLI = os.path.join(PKG, "w2x-converter")


@pytest.fixture(scope="module")
def cli(w2x):
    w2x.build()
    assert os.path.exists(CLI)
    return CLI


@pytest.fixture(scope="module")
def imgtool(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "test_imgproc")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(PKG, "host"),
                           os.path.join(ROOT, "tests", "cpp", "test_imgproc.cpp"), "-o", out, "-lz", "-pthread"])
    return out


def _test_image(w=37, h=29, seed=3):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([128 + 100 * np.sin(xx / 5.0 + c) * np.cos(yy / 4.0 - c) for c in range(3)], -1) + rng.normal(0, 12, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)        # B,G,R like cv2


def test_png_read_convert_and_rgb2yuv_match_cv2(imgtool, tmp_path):
    bgr = _test_image()
    cv2.imwrite(str(tmp_path / "in.png"), bgr)
    subprocess.check_call([imgtool, "yuv", str(tmp_path / "in.png"), str(tmp_path / "yuv.f32")])
    got = np.fromfile(tmp_path / "yuv.f32", np.float32).reshape(bgr.shape)
    ref = cv2.imread(str(tmp_path / "in.png"), cv2.IMREAD_COLOR)
    ref = ref.astype(np.float32) * np.float32(1.0 / 255.0)      # convertTo(CV_32F, 1/255)
    ref = cv2.cvtColor(ref, cv2.COLOR_RGB2YUV)                  # on B,G,R data, like src/main.cpp:76
    assert np.abs(got - ref).max() <= 2e-7
    # other PNG flavours imread accepts: grey, RGBA, palette-free 16-bit
    for name, arr in (("grey.png", bgr[..., 0]), ("rgba.png", np.dstack([bgr, np.full(bgr.shape[:2], 77, np.uint8)])),
                      ("deep.png", (bgr.astype(np.uint16) << 8) | 3)):
        cv2.imwrite(str(tmp_path / name), arr)
        subprocess.check_call([imgtool, "yuv", str(tmp_path / name), str(tmp_path / "o.f32")])
        got = np.fromfile(tmp_path / "o.f32", np.float32).reshape(bgr.shape)
        ref = cv2.cvtColor(cv2.imread(str(tmp_path / name), cv2.IMREAD_COLOR).astype(np.float32) * np.float32(1 / 255.0), cv2.COLOR_RGB2YUV)
        assert np.abs(got - ref).max() <= 2e-7, name


@pytest.mark.parametrize("interp,flag,dw,dh", [("nearest", cv2.INTER_NEAREST, 74, 58), ("cubic", cv2.INTER_CUBIC, 74, 58),
                                               ("linear", cv2.INTER_LINEAR, 59, 46), ("linear", cv2.INTER_LINEAR, 27, 21)])
def test_resize_matches_cv2(imgtool, tmp_path, interp, flag, dw, dh):
    src = cv2.cvtColor(_test_image().astype(np.float32) / 255, cv2.COLOR_RGB2YUV)
    src.tofile(tmp_path / "s.f32")
    subprocess.check_call([imgtool, "resize", str(tmp_path / "s.f32"), "37", "29", str(dw), str(dh), interp, str(tmp_path / "d.f32")])
    got = np.fromfile(tmp_path / "d.f32", np.float32).reshape(dh, dw, 3)
    ref = cv2.resize(src, (dw, dh), interpolation=flag)
    assert np.abs(got - ref).max() <= (0 if interp == "nearest" else 2e-6)


def test_yuv2rgb_to_u8_and_png_write_match_cv2(imgtool, tmp_path):
    rng = np.random.default_rng(1)
    yuv = cv2.cvtColor(_test_image(41, 23, 9).astype(np.float32) / 255, cv2.COLOR_RGB2YUV)
    yuv += rng.normal(0, 0.3, yuv.shape).astype(np.float32)         # push values out of range: saturation paths
    yuv.tofile(tmp_path / "y.f32")
    subprocess.check_call([imgtool, "rgb8", str(tmp_path / "y.f32"), "41", "23", str(tmp_path / "o.png")])
    got = cv2.imread(str(tmp_path / "o.png"), cv2.IMREAD_COLOR)
    f = cv2.cvtColor(yuv, cv2.COLOR_YUV2RGB)
    ref = np.clip(np.rint(f * np.float32(255.0)), 0, 255).astype(np.uint8)   # convertTo(CV_8U, 255): cvRound + saturate
    diff = np.abs(got.astype(int) - ref.astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 0.002              # only float-rounding ties may differ


def test_row_parallel_paths_match_cv2_on_a_large_image(imgtool, tmp_path):
    """640x400 pushes every imgproc routine past its single-thread threshold: the row-parallel sweeps must give the same
    numbers (per-pixel arithmetic does not depend on the split)."""
    w, h = 640, 400
    bgr = _test_image(w, h, 21)
    cv2.imwrite(str(tmp_path / "in.png"), bgr)
    subprocess.check_call([imgtool, "yuv", str(tmp_path / "in.png"), str(tmp_path / "yuv.f32")])
    got = np.fromfile(tmp_path / "yuv.f32", np.float32).reshape(bgr.shape)
    yuv = cv2.cvtColor(bgr.astype(np.float32) * np.float32(1.0 / 255.0), cv2.COLOR_RGB2YUV)
    assert np.abs(got - yuv).max() <= 2e-7
    yuv.tofile(tmp_path / "s.f32")
    for interp, flag, dw, dh, tol in (("nearest", cv2.INTER_NEAREST, 2 * w, 2 * h, 0), ("cubic", cv2.INTER_CUBIC, 2 * w, 2 * h, 2e-6),
                                      ("linear", cv2.INTER_LINEAR, 480, 300, 2e-6)):
        subprocess.check_call([imgtool, "resize", str(tmp_path / "s.f32"), str(w), str(h), str(dw), str(dh), interp, str(tmp_path / "d.f32")])
        got = np.fromfile(tmp_path / "d.f32", np.float32).reshape(dh, dw, 3)
        assert np.abs(got - cv2.resize(yuv, (dw, dh), interpolation=flag)).max() <= tol, interp
    subprocess.check_call([imgtool, "rgb8", str(tmp_path / "s.f32"), str(w), str(h), str(tmp_path / "o.png")])
    got8 = cv2.imread(str(tmp_path / "o.png"), cv2.IMREAD_COLOR)
    ref8 = np.clip(np.rint(cv2.cvtColor(yuv, cv2.COLOR_YUV2RGB) * np.float32(255.0)), 0, 255).astype(np.uint8)
    diff = np.abs(got8.astype(int) - ref8.astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 0.002


def test_cli_flag_surface_and_exit_codes(cli, tmp_path, json_models):
    def run(*a):
        return subprocess.run([cli, *a], capture_output=True, text=True, cwd=tmp_path)
    r = run("--version")
    assert r.returncode == 0 and "version: 1.0.0" in r.stdout                           # src/main.cpp:26
    r = run("--help")
    assert r.returncode == 0
    for flag in ("--input_file", "--output_file", "--mode", "--noise_level", "--scale_ratio", "--model_dir", "--jobs", "-i", "-o", "-m", "-j"):
        assert flag in r.stdout, flag
    assert "waifu2x reimplementation using OpenCV" in r.stdout
    r = run()
    assert r.returncode == 1 and "PARSE ERROR" in r.stderr and "input_file" in r.stderr  # TCLAP: required arg, exit(1)
    assert run("-i", "a.png", "-m", "bogus").returncode == 1
    assert run("-i", "a.png", "--noise_level", "3").returncode == 1
    assert run("-i", "a.png", "--jobs", "x").returncode == 1
    assert run("-i", "a.png", "--nope", "1").returncode == 1
    assert run("-i", "a.png", "-i", "b.png").returncode == 1
    # model file missing -> message + exit(-1) (src/main.cpp:88-89), after a successful image read
    cv2.imwrite(str(tmp_path / "in.png"), _test_image(16, 16))
    r = run("-i", "in.png", "--model_dir", "no_such_dir")
    assert r.returncode == 255 and "couldn't open no_such_dir/noise1_model.json" in r.stderr
    r = run("-i", "missing.png")
    assert r.returncode == 255


@pytest.mark.parametrize("mode", ["noise", "scale", "noise_scale"])
def test_cli_without_a_device_fails_instead_of_writing_an_unprocessed_image(cli, tmp_path, json_models, mode):
    """No CUDA device: the conversion cannot happen and there is no CPU fallback.  The reference cannot fail silently here
    (a failing layer ends the process: src/convertRoutine.cpp:69 std::exit(-1)); neither may the drop-in -- in particular
    `-m noise` must not write the un-denoised input and report success."""
    cv2.imwrite(str(tmp_path / "in.png"), _test_image(24, 20))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([cli, "-i", "in.png", "-m", mode, "--model_dir", os.path.dirname(json_models["scale2.0x"])], capture_output=True, text=True,
                       cwd=tmp_path, env=env)
    assert r.returncode != 0, r.stdout
    assert "process successfully done!" not in r.stdout
    assert [f for f in os.listdir(tmp_path) if f != "in.png"] == []
    assert "no CUDA device" in r.stderr or "CPU fallback" in r.stderr or "Error" in r.stderr


def _reference_pipeline(bgr, mode, level, ratio, oracle_models, ncpu):
    """src/main.cpp restated with cv2 for the plumbing and the CPU oracle for convertWithModels."""
    image = cv2.cvtColor(bgr.astype(np.float32) * np.float32(1 / 255.0), cv2.COLOR_RGB2YUV)
    if mode in ("noise", "noise_scale"):
        ch = cv2.split(image)
        ch = [oracle_models[f"noise{level}"].convert(ch[0].copy(), n_job=ncpu), ch[1], ch[2]]
        image = cv2.merge(ch)
    if mode in ("scale", "noise_scale"):
        it = int(np.ceil(np.log2(ratio)))
        shrink = 0.0 if int(ratio) == 2 ** it else ratio / 2.0 ** it
        for _ in range(it):
            size = (image.shape[1] * 2, image.shape[0] * 2)
            y = cv2.split(cv2.resize(image, size, interpolation=cv2.INTER_NEAREST))[0].copy()
            ch = list(cv2.split(cv2.resize(image, size, interpolation=cv2.INTER_CUBIC)))
            ch[0] = oracle_models["scale2.0x"].convert(y, n_job=ncpu)
            image = cv2.merge(ch)
        if shrink != 0.0:
            size = (int(image.shape[1] * shrink), int(image.shape[0] * shrink))
            image = cv2.resize(image, size, interpolation=cv2.INTER_LINEAR)
    out = cv2.cvtColor(image, cv2.COLOR_YUV2RGB)
    return np.clip(np.rint(out * np.float32(255.0)), 0, 255).astype(np.uint8)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,level,ratio,name", [("noise_scale", 1, 2.0, "in(noise_scale)(Level1)(x2.000000).png"),
                                                  ("scale", 1, 3.0, "in(scale)(x3.000000).png"),
                                                  ("noise", 2, 2.0, "in(noise)(Level2).png")])
def test_cli_end_to_end_matches_reference_pipeline(cli, tmp_path, json_models, oracle_models, ncpu, mode, level, ratio, name):
    bgr = _test_image(45, 33, 11)
    cv2.imwrite(str(tmp_path / "in.png"), bgr)
    mdir = os.path.dirname(json_models["scale2.0x"])
    r = subprocess.run([cli, "-i", str(tmp_path / "in.png"), "-m", mode, "--noise_level", str(level), "--scale_ratio", str(ratio),
                        "--model_dir", mdir, "-j", "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("process successfully done!")                       # src/main.cpp:192
    if "scale" in mode:
        assert "start scaling" in r.stdout and "#1 2x scaling..." in r.stdout            # :123,129-130
    assert "Iteration #7..." in r.stdout                                                  # src/convertRoutine.cpp:67
    out = cv2.imread(str(tmp_path / name), cv2.IMREAD_COLOR)                              # auto name rule, :173-189
    assert out is not None, os.listdir(tmp_path)
    ref = _reference_pipeline(bgr, mode, level, ratio, oracle_models, ncpu)
    assert out.shape == ref.shape
    diff = np.abs(out.astype(int) - ref.astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 0.01                                   # 8-bit rounding ties only
