"""bench.py's output contract (the driver parses ONE JSON line): the reference arm is run for real on the CPU (with the
fast C restatement as the timed code, W2X_BENCH_CPU=oracle); the product arm needs a GPU, so here only its key set is
checked against the source."""
import json
import os
import re
import subprocess
import sys

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def test_reference_arm_prints_one_valid_json_line():
    env = dict(os.environ, W2X_BENCH_CPU="oracle")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "Mpix/s full scale2.0x model pass" and d["unit"] == "Mpix/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_is_rank0_only_under_torchrun_env():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", W2X_BENCH_CPU="oracle")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_line_carries_the_contract_keys():
    src = open(os.path.join(ROOT, "bench.py")).read()
    body = src[src.index("def run_ours"):src.index("def main")]
    line = body[body.index("line = {"):]
    for k in BASE_KEYS | {"gpu_launches", "clocks", "roofline", "cpu_baseline"}:
        assert f'"{k}"' in line, k
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert f'"{k}"' in body, k                       # the roofline object
    assert '"model"' not in line                          # domain vocabulary only
    assert re.search(r"from oracle import", body) is None   # the product arm never imports oracle/
