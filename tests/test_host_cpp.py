"""The C++ host mirror (waifu2x-converter-cpp_b200/host/w2xc.hpp): same names and behaviour as the
reference's w2xc::Model / modelUtility / convertWithModels, driven like the reference's main.cpp."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

PKG = os.path.join(ROOT, "waifu2x-converter-cpp_b200")


@pytest.fixture(scope="module")
def exe(w2x, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "test_host_api")
    cmd = ["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "host"),
           os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp"), "-o", out, "-L", PKG, "-lw2x_b200", f"-Wl,-rpath,{PKG}"]
    subprocess.check_call(cmd)
    return out


def test_cpp_host_api_without_gpu(exe, json_models):
    p = subprocess.run([exe, json_models["scale2.0x"]], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "ALL OK" in p.stdout
    assert "couldn't open /nonexistent/model.json" in p.stderr          # src/modelHandler.cpp:176-179


@pytest.mark.gpu
def test_cpp_convert_with_models_on_gpu(exe, json_models, oracle_mod, oracle_models, ncpu, tmp_path):
    x = oracle_mod.seeded_plane(70, 45, 6, "uniform")
    fin, fout = tmp_path / "in.f32", tmp_path / "out.f32"
    x.tofile(fin)
    p = subprocess.run([exe, json_models["noise1"], "--gpu", str(fin), "70", "45", str(fout)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert [l for l in p.stdout.splitlines() if l.startswith("Iteration")][:7] == [f"Iteration #{k}..." for k in range(1, 8)]
    assert "number of input planes mismatch" in p.stderr
    y = np.fromfile(fout, np.float32).reshape(45, 70)
    ref = oracle_models["noise1"].convert(x, n_job=ncpu)
    assert np.abs(y - ref).max() <= 6e-5   # library default precision: fp16 + 2 x e4m3 corrections
