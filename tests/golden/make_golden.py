"""Generates the committed fixtures under tests/golden/ from the reference checkout.

Run once in the build container (`python tests/golden/make_golden.py`); `/root/reference` does
not exist on the GPU box, so tests only ever read the generated files.

* reference_kats.json : the known-answer vectors of the reference's own test-suite
  (test/runtests.jl line numbers are recorded per entry) and the byte-exact `save` fixtures
  test/data/save1..4.
* iris.npz : the tutorial data docs/src/tutorial/data/iris.{simmat,classes,features} (150x150 similarity,
  150x3 one-hot labels) as arrays + names.  No expected outputs are stored in the reference for
  it; it is used as an extra CUDA-vs-oracle input.
"""
import json
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def read_named(path):
    with open(path) as f:
        lines = [l.rstrip("\n") for l in f if l.strip()]
    cols = lines[0].split()
    rows, vals = [], []
    for l in lines[1:]:
        p = l.split()
        rows.append(p[0])
        vals.append([float(x) for x in p[1:]])
    return np.array(vals, dtype=np.float64), rows, cols


def main():
    kats = {
        "k": {"src": "test/runtests.jl:20-26",
              "M": [[0, 0, 0], [0, 0, 1], [0, 1, 1], [1, 1, 1]], "expect": [0, 1, 2, 3]},
        "cutoff": {"src": "test/runtests.jl:36-71",
                   "x": 0.8, "y": [round(0.1 * i, 10) for i in range(11)],
                   "y_julia_range": "0.0:0.1:1.0",
                   "z": [[0.1, 0.5], [0.5, 1.0]],
                   "cases": [
                       {"alpha": 0.5, "x_bin": 1.0, "x_w": 0.8,
                        "y_bin": [0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1],
                        "y_w": [0, 0, 0, 0, 0, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0],
                        "z_bin": [[0, 1], [1, 1]], "z_w": [[0, 0.5], [0.5, 1.0]]},
                       {"alpha": -0.01, "x_bin": 1.0, "x_w": 0.8,
                        "y_bin": [1] * 11, "y_w": "y",
                        "z_bin": [[1, 1], [1, 1]], "z_w": "z"},
                       {"alpha": 1.01, "x_bin": 0.0, "x_w": 0.0,
                        "y_bin": [0] * 11, "y_w": [0] * 11,
                        "z_bin": [[0, 0], [0, 0]], "z_w": [[0, 0], [0, 0]]}]},
        "featurize": {"src": "test/runtests.jl:73-81",
                      "M0": [[0.1, 0.5], [0.5, 1.0]], "names": ["s1", "s2"], "alpha": 0.5,
                      "bin": [[0, 1], [1, 1]], "w": [[0, 0.5], [0.5, 1.0]],
                      "colnames": ["fs1", "fs2"]},
        "construct": {"src": "test/runtests.jl:83-110",
                      "X": [[1, 0, 1], [1, 1, 0], [0, 1, 1]], "y": [[0, 1], [1, 1], [1, 0]],
                      "xrows": ["s1", "s2", "s3"], "xcols": ["fs1", "fs2", "fs3"],
                      "ycols": ["t1", "t2"], "queries": ["s1"],
                      "names": ["s1", "s2", "s3", "fs2", "fs3", "t1", "t2"],
                      "err_same_names": "Source and Features nodes have the same names!",
                      "err_rows": "Labels and features have different number of source nodes"},
        "spread": {"src": "test/runtests.jl:113-118",
                   "M": [[1, 0, 0], [1, 1, 0], [1, 1, 1]],
                   "W": [[1, 0, 0], [0.5, 0.5, 0], [0.33333, 0.33333, 0.33333]], "rtol": 1e-5},
        "predict": {"src": "test/runtests.jl:120-158",
                    "names": ["q1", "s1", "s2", "s3", "f1", "f2", "f3", "t1", "t2"],
                    "A": [[0, 0, 0, 0, 0, 0, 1, 0, 0], [0, 0, 0, 0, 1, 1, 0, 1, 0],
                          [0, 0, 0, 0, 1, 1, 0, 1, 0], [0, 0, 0, 0, 0, 0, 1, 0, 1],
                          [0, 1, 1, 0, 0, 0, 0, 0, 0], [0, 1, 1, 0, 0, 0, 0, 0, 0],
                          [1, 0, 0, 1, 0, 0, 0, 0, 0], [0, 1, 1, 0, 0, 0, 0, 0, 0],
                          [0, 0, 0, 1, 0, 0, 0, 0, 0]],
                    "B": [[0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 1, 1, 0, 1, 0],
                          [0, 0, 0, 0, 1, 1, 0, 1, 0], [0, 0, 0, 0, 0, 0, 1, 0, 1],
                          [0, 1, 1, 0, 0, 0, 0, 0, 0], [0, 1, 1, 0, 0, 0, 0, 0, 0],
                          [0, 0, 0, 1, 0, 0, 0, 0, 0], [0, 1, 1, 0, 0, 0, 0, 0, 0],
                          [0, 0, 0, 1, 0, 0, 0, 0, 0]],
                    "rows": ["q1"], "cols": ["t1", "t2"], "yhat": [[0, 0.5]], "exact": True},
        "clean": {"src": "test/runtests.jl:160-183",
                  "names": ["q1", "s1", "s2", "s3", "f1", "f2", "f3", "t1", "t2"],
                  "A": [[0, 0, 0, 0, 0, 0, 1, 0, 0], [0, 0, 0, 0, 1, 1, 0, 1, 0],
                        [0, 0, 0, 0, 1, 1, 0, 1, 0], [0, 0, 0, 0, 0, 0, 1, 0, 0],
                        [0, 1, 1, 0, 0, 0, 0, 0, 0], [0, 1, 1, 0, 0, 0, 0, 0, 0],
                        [1, 0, 0, 1, 0, 0, 0, 0, 0], [0, 1, 1, 0, 0, 0, 0, 0, 0],
                        [0, 0, 0, 0, 0, 0, 0, 0, 0]],
                  "yhat": [[0, 0.5]], "targets": ["t1", "t2"], "expect": [[0, -99]]},
        "save": {"src": "test/runtests.jl:185-203",
                 "y": [[1, 0, 1], [0, 1, 0]], "rows": ["s1", "s2"], "cols": ["t1", "t2", "t3"],
                 "files": {}},
        "atL": {"src": "test/runtests.jl:226-243",
                "yhat": list(range(1, 11)), "y": [0, 0, 0, 0, 0, 0, 0, 1, 1, 1],
                "grouping": [1.0] * 10,
                "recall": {"5": 1.0, "1": 1.0 / 3.0}, "precision": {"5": 0.6, "1": 1.0}},
        "confusion": {"src": "test/runtests.jl:246-266", "tn_fp_fn_tp": [3, 2, 2, 3],
                      "f1": 0.6, "mcc": 0.2, "acc": 0.6, "bacc": 0.6, "recall": 0.6,
                      "precision": 0.6,
                      "y": [1, 1, 0, 1, 0, 0, 0, 1, 1, 0], "yhat": [1, 1, 1, 1, 1, 0, 0, 0, 0, 0]},
        "mcc_limits": {"src": "test/runtests.jl:280-287",
                       "yhat": [1, 1, 0, 1, 0, 0, 0, 1, 1, 0], "y": [1, 1, 1, 0, 0, 0, 0, 0, 0, 0],
                       "tol": 1e-5},
    }
    for i in (1, 2, 3, 4):
        with open(f"{REF}/test/data/save{i}") as f:
            kats["save"]["files"][f"save{i}"] = f.read()
    kats["read_namedmatrix"] = {"src": "test/runtests.jl:8-18, test/data/mat1..4", "files": {},
                                "expect": {"mat1": {"rows": True, "cols": True, "rownames": ["s1", "s2"], "colnames": ["t1", "t2", "t3"]},
                                           "mat2": {"rows": True, "cols": False, "rownames": ["s1", "s2"], "colnames": ["C#1", "C#2", "C#3"]},
                                           "mat3": {"rows": False, "cols": True, "rownames": ["R#1", "R#2"], "colnames": ["t1", "t2", "t3"]},
                                           "mat4": {"rows": False, "cols": False, "rownames": ["R#1", "R#2"], "colnames": ["C#1", "C#2", "C#3"]}},
                                "values": [[0.0, 0.0, 0.0], [0.0, 0.0, 0.0]]}
    for i in (1, 2, 3, 4):
        with open(f"{REF}/test/data/mat{i}") as f:
            kats["read_namedmatrix"]["files"][f"mat{i}"] = f.read()
    with open(f"{REF}/docs/src/tutorial/data/iris.classes") as f:
        kats["read_namedmatrix"]["files"]["iris.classes"] = f.read()
    with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
        json.dump(kats, f, indent=1)

    S, srows, scols = read_named(f"{REF}/docs/src/tutorial/data/iris.simmat")
    C, crows, ccols = read_named(f"{REF}/docs/src/tutorial/data/iris.classes")
    F, frows, fcols = read_named(f"{REF}/docs/src/tutorial/data/iris.features")
    assert srows == crows == frows and S.shape == (150, 150) and C.shape == (150, 3) and F.shape == (150, 4)
    np.savez_compressed(os.path.join(HERE, "iris.npz"), S=S, C=C, F=F, names=np.array(srows),
                        classes=np.array(ccols), descriptors=np.array(fcols))
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    sys.exit(main())
