"""INTEGRATION.md option A, proven: oracle/_ref/waifu2x-optionA is the reference's UNMODIFIED src/main.cpp compiled with
host/dropin/{modelHandler,convertRoutine}.hpp in front of its own headers -- i.e. w2xc::Model, w2xc::modelUtility and
w2xc::convertWithModels(cv::Mat&, cv::Mat&, ...) come from the product's host/w2xc.hpp (-DW2X_WITH_OPENCV) and
libw2x_b200.so -- with oracle/cvshim standing in for OpenCV (recipe: oracle/Makefile).  On a GPU it must behave exactly like
the reference's own CPU build of the same main.cpp (oracle/_ref/waifu2x-reference-cli): same stdout, same file name, same
pixels within 1 LSB.  Built only where /root/reference exists; travels prebuilt."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

cv2 = pytest.importorskip("cv2")
from test_cli import _test_image  # noqa: E402

OPTION_A = os.path.join(ROOT, "oracle", "_ref", "waifu2x-optionA")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "waifu2x-reference-cli")
pytestmark = pytest.mark.skipif(not (os.path.exists(OPTION_A) and os.path.exists(REF_CLI)), reason="oracle/_ref binaries not built (need /root/reference at build time)")


def test_option_a_binary_links_the_product_library_and_keeps_the_reference_flag_surface():
    ldd = subprocess.check_output(["ldd", OPTION_A], text=True)
    assert "libw2x_b200.so" in ldd and "not found" not in ldd
    ours = subprocess.run([OPTION_A, "--version"], capture_output=True, text=True)
    ref = subprocess.run([REF_CLI, "--version"], capture_output=True, text=True)
    assert ours.returncode == ref.returncode == 0 and "1.0.0" in ours.stdout
    assert subprocess.run([OPTION_A], capture_output=True, text=True).returncode == 1      # TCLAP: required -i missing


@pytest.mark.gpu
@pytest.mark.parametrize("mode,level,ratio,w,h", [("noise_scale", 1, 2.0, 37, 29), ("scale", 1, 2.0, 700, 600), ("noise", 2, 2.0, 64, 48)])
def test_option_a_equals_the_reference_build_of_the_same_main_cpp(tmp_path, json_models, mode, level, ratio, w, h):
    bgr = _test_image(w, h, 5)
    mdir = os.path.dirname(json_models["scale2.0x"])
    outs = {}
    for name, exe in (("gpu", OPTION_A), ("cpu", REF_CLI)):
        d = tmp_path / name
        d.mkdir()
        cv2.imwrite(str(d / "in.png"), bgr)
        r = subprocess.run([exe, "-i", "in.png", "-m", mode, "--noise_level", str(level), "--scale_ratio", str(ratio), "--model_dir", mdir,
                            "-j", "16"], capture_output=True, text=True, cwd=d)
        assert r.returncode == 0, r.stdout + r.stderr
        files = sorted(f for f in os.listdir(d) if f != "in.png")
        outs[name] = (r.stdout, files, cv2.imread(str(d / files[0]), cv2.IMREAD_COLOR))
    assert outs["gpu"][1] == outs["cpu"][1]                      # auto output name (src/main.cpp:173-189)
    assert outs["gpu"][0] == outs["cpu"][0]                      # progress lines: per-block "start process block (c,r) ..." + "Iteration #k..." (1400x1200 is split)
    diff = np.abs(outs["gpu"][2].astype(int) - outs["cpu"][2].astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 0.01
