"""gen_golden.py -- TEST INFRASTRUCTURE.  Regenerates tests/golden/ from the reference.

Run in the authoring container only (needs /root/reference and cv2):

    python oracle/gen_golden.py

The reference ships no golden vectors, so the vectors committed under
tests/golden/ are outputs of the reference's own arithmetic backend: oracle/ref_cv2.py drives
the same cv::filter2D / add / max / min / scaleAdd / copyMakeBorder calls the reference makes
(src/modelHandler.cpp:117-159, src/convertRoutine.cpp:21-169) on seeded inputs, using the
reference's model files loaded with the reference's double->float rule.

Written files
  models/<name>_model.npz      fp32 weights [o][i][3][3] + fp64 biases of models/<name>_model.json
  model_kat.json               sha256 of each JSON file, first/last weight + first bias per layer
  cfg1_<name>_<kind>.npy       convertWithModels output on the 256x256 seeded plane (config 1)
  odd_sizes.npz                convertWithModels outputs for 1x1, 15x13, 37x61 (scale2.0x)
  split_513x768.npz            lattice + block-boundary strips of the 513x768 (block-split) output
  layers_32x24.npz             Model::filter output of every layer on a seeded 32x24 input
  block_tables.json            block geometry for the sizes SURVEY.md section 8(c) lists
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle, ref_cv2  # noqa: E402

REF_MODELS = "/root/reference/models"
OUT = os.path.join(ROOT, "tests", "golden")


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def main():
    os.makedirs(os.path.join(OUT, "models"), exist_ok=True)
    nj = os.cpu_count() or 4
    kat = {}
    cv_models = {}
    for name in oracle.MODEL_NAMES:
        path = os.path.join(REF_MODELS, f"{name}_model.json")
        om = oracle.OracleModel.from_json(path)
        om.save_npz(os.path.join(OUT, "models", f"{name}_model.npz"))
        cv_models[name] = ref_cv2.load_models_json(path)
        kat[name] = {
            "sha256": sha256(path),
            "bytes": os.path.getsize(path),
            "dims": om.dims,
            "layers": [{"w_first": float(w.reshape(-1)[0]), "w_last": float(w.reshape(-1)[-1]),
                        "w_sum64": float(w.astype(np.float64).sum()),
                        "b_first": float(b[0]), "b_last": float(b[-1])}
                       for w, b in zip(om.weights, om.biases)],
        }
    with open(os.path.join(OUT, "model_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)

    # config 1: 256x256 Y plane, all three models on noise, scale2.0x also on a smooth image
    for name in oracle.MODEL_NAMES:
        for kind in (("uniform", "smooth") if name == "scale2.0x" else ("uniform",)):
            x = oracle.seeded_plane(256, 256, 0, kind)
            y = ref_cv2.convert_with_models(x, cv_models[name], n_job=nj)
            np.save(os.path.join(OUT, f"cfg1_{name}_{kind}.npy"), y)
            print("cfg1", name, kind, hashlib.sha256(x.tobytes()).hexdigest()[:16], float(y.mean()))

    # odd sizes (non-split)
    odd = {}
    for (w, h) in ((1, 1), (15, 13), (37, 61)):
        x = oracle.seeded_plane(w, h, 10 + w, "uniform")
        odd[f"out_{w}x{h}"] = ref_cv2.convert_with_models(x, cv_models["scale2.0x"], n_job=nj)
    np.savez_compressed(os.path.join(OUT, "odd_sizes.npz"), **odd)

    # 513x768: first size past the 512x768 no-split edge -> block-split path (2x2 blocks)
    x = oracle.seeded_plane(513, 768, 5, "uniform")
    y = ref_cv2.convert_with_models(x, cv_models["scale2.0x"], n_job=nj)
    y_ns = ref_cv2.convert_with_models(x, cv_models["scale2.0x"], block_splitting=False, n_job=nj)
    print("513x768 split vs non-split max-abs", float(np.abs(y - y_ns).max()))
    np.savez_compressed(os.path.join(OUT, "split_513x768.npz"), lattice=y[::16, ::16],
                        rows_494_502=y[494:502, :], cols_494_502=y[:, 494:502],
                        split_vs_nosplit_maxabs=np.float64(np.abs(y - y_ns).max()))

    # per-layer Model::filter (same-size, BORDER_REPLICATE) on a 32x24 plane
    lay = {}
    rng = np.random.default_rng(77)
    for li, m in enumerate(cv_models["scale2.0x"]):
        xin = (rng.random((m.n_in, 24, 32), dtype=np.float32) - 0.25).astype(np.float32)
        out = m.filter([xin[i] for i in range(m.n_in)], n_job=nj)
        lay[f"in{li}"] = xin
        lay[f"out{li}"] = np.stack(out)
    np.savez_compressed(os.path.join(OUT, "layers_32x24.npz"), **lay)

    # block geometry tables
    sizes = [(1920, 1080), (3840, 2160), (4096, 4096), (8192, 8192), (512, 768), (513, 768),
             (499, 1), (1, 1), (498, 1000), (997, 790)]
    tabs = {}
    for (w, h) in sizes:
        rows, sc, sr = ref_cv2.block_table(w, h, 512, 512, 7)
        tabs[f"{w}x{h}"] = {"split_cols": sc, "split_rows": sr,
                            "require_split": bool(w * h > 512 * 512 * 3 // 2), "rows": rows}
    with open(os.path.join(OUT, "block_tables.json"), "w") as f:
        json.dump(tabs, f)
    print("done")


if __name__ == "__main__":
    main()
