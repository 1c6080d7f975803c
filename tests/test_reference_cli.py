"""The reference's OWN CLI -- src/main.cpp + src/modelHandler.cpp + src/convertRoutine.cpp compiled unmodified against
the OpenCV API shim (oracle/cvshim; recipe oracle/Makefile -> oracle/_ref/waifu2x-reference-cli) -- pins:

  * the restated pipeline tests/test_cli.py::_reference_pipeline (cv2 plumbing + CPU oracle), which is what the GPU test
    of the product CLI is compared with: same pixels (<= 1 LSB), same auto output name, same progress lines;
  * the product CLI's flag surface: for every malformed / failing invocation both binaries give the same exit code and
    the same messages (no GPU needed: these paths end before any conversion).

The binary is built only where /root/reference exists and travels prebuilt; nothing here reads /root/reference."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

cv2 = pytest.importorskip("cv2")
from test_cli import CLI, _reference_pipeline, _test_image, cli  # noqa: E402,F401  (the product CLI fixture and the restated pipeline)

REF_CLI = os.path.join(ROOT, "oracle", "_ref", "waifu2x-reference-cli")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_CLI), reason="oracle/_ref/waifu2x-reference-cli not built (needs /root/reference at build time)")


@pytest.mark.parametrize("mode,level,ratio,name", [("noise_scale", 1, 2.0, "in(noise_scale)(Level1)(x2.000000).png"),
                                                  ("scale", 1, 3.0, "in(scale)(x3.000000).png"),
                                                  ("scale", 1, 1.5, "in(scale)(x1.500000).png"),
                                                  ("noise", 2, 2.0, "in(noise)(Level2).png")])
def test_reference_cli_equals_the_restated_pipeline(tmp_path, json_models, oracle_models, ncpu, mode, level, ratio, name):
    bgr = _test_image(21, 17, 11)
    cv2.imwrite(str(tmp_path / "in.png"), bgr)
    mdir = os.path.dirname(json_models["scale2.0x"])
    r = subprocess.run([REF_CLI, "-i", str(tmp_path / "in.png"), "-m", mode, "--noise_level", str(level), "--scale_ratio", str(ratio),
                        "--model_dir", mdir, "-j", "4"], capture_output=True, text=True)
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


def _norm(text, exe):
    """messages modulo the program name / path TCLAP prints"""
    text = text.replace(exe, "PROG").replace(os.path.basename(exe), "PROG")
    return re.sub(r"[ \t]+", " ", text).strip()


CASES = [
    [],                                                     # required -i missing -> TCLAP parse error, exit 1
    ["-i", "a.png", "-m", "bogus"],                         # value not in the allowed set
    ["-i", "a.png", "--noise_level", "3"],
    ["-i", "a.png", "--jobs", "x"],                         # not an integer
    ["-i", "a.png", "--nope", "1"],                         # unknown flag
    ["-i", "a.png", "-i", "b.png"],                         # flag given twice
    ["--version"],
    ["-i", "in.png", "--model_dir", "no_such_dir"],         # model file missing after a successful image read: exit(-1)
    ["-i", "in.png", "-m", "scale", "--model_dir", "no_such_dir"],
]


@pytest.mark.parametrize("args", CASES, ids=[" ".join(a) or "(none)" for a in CASES])
def test_product_cli_fails_exactly_like_the_reference_cli(cli, tmp_path, args):  # noqa: F811
    cv2.imwrite(str(tmp_path / "in.png"), _test_image(16, 16))
    ours = subprocess.run([cli, *args], capture_output=True, text=True, cwd=tmp_path)
    ref = subprocess.run([REF_CLI, *args], capture_output=True, text=True, cwd=tmp_path)
    assert ours.returncode == ref.returncode, (ours.returncode, ref.returncode, ours.stderr, ref.stderr)
    if args == ["--version"]:
        assert _norm(ours.stdout, cli).split("version:")[1] == _norm(ref.stdout, REF_CLI).split("version:")[1]
        return
    # first line of the diagnostic: "PARSE ERROR: ..." + the offending argument, or the model loader's message
    def key_lines(r, exe):
        lines = [l for l in _norm(r.stderr, exe).splitlines() if l.strip()]
        return [l for l in lines if "PARSE ERROR" in l or "Argument" in l or "couldn't open" in l or "Required" in l or "Value" in l or "Couldn't" in l]
    assert key_lines(ours, cli) == key_lines(ref, REF_CLI), (ours.stderr, ref.stderr)


def test_help_lists_the_same_flags(cli):  # noqa: F811
    ours = subprocess.run([cli, "--help"], capture_output=True, text=True)
    ref = subprocess.run([REF_CLI, "--help"], capture_output=True, text=True)
    assert ours.returncode == ref.returncode == 0
    flags = lambda t: sorted(set(re.findall(r"(?<![\w-])(--?[a-z_]+)", t)))
    assert flags(ours.stdout) == flags(ref.stdout)
    for line in ("number of threads launching at the same time", "path to custom model directory (don't append last / )",
                 "custom scale ratio", "noise reduction level", "image processing mode", "waifu2x reimplementation using OpenCV"):
        assert line in ours.stdout and line in ref.stdout, line
