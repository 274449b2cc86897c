"""Generate tests/golden/volume_prep_golden.json by running the REAL `npz_to_tensor` of the reference
(scripts/data.py:49-111) in the build container:

    python -m oracle.make_golden_volume

scripts/data.py cannot be imported here (nibabel, data_inference, ... are absent), so the function's own source is
located with `ast`, compiled from where it lies under /root/reference and executed with numpy / torch in scope -
nothing is copied into this repository.  Outputs are 221 MB each, so the fixture stores their SHA-256 and a few
probe values instead; tests/test_volume_prep_cpu.py recomputes them through oracle/volume_prep_oracle.py.
"""
from __future__ import annotations

import ast
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/scripts/data.py"

from oracle.volume_prep_oracle import digest, synthetic_volume  # noqa: E402

# (D, H, W) of the stored array, dtype, seed: larger / smaller than the 240x480x480 target on every axis, odd sizes
CASES = [((240, 480, 480), "float32", 0), ((301, 512, 512), "float16", 1), ((200, 400, 500), "float32", 2),
         ((241, 479, 481), "float16", 3), ((96, 600, 333), "float32", 4)]
PROBES = [(0, 0, 0, 0), (0, 120, 240, 240), (0, 239, 479, 479), (0, 17, 333, 41), (0, 200, 20, 470)]


def reference_fn():
    tree = ast.parse(open(REF).read(), REF)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "npz_to_tensor")
    mod = ast.Module(body=[node], type_ignores=[])
    ns = {"np": np, "torch": torch}
    exec(compile(mod, REF, "exec"), ns)
    return ns["npz_to_tensor"], (node.lineno, node.end_lineno)


def main():
    fn, lines = reference_fn()
    out = {"reference": f"scripts/data.py:{lines[0]}-{lines[1]} npz_to_tensor", "numpy": np.__version__, "cases": []}
    with tempfile.TemporaryDirectory() as tmp:
        for shape, dtype, seed in CASES:
            arr = synthetic_volume(shape, dtype, seed)
            path = os.path.join(tmp, "v.npz")
            np.savez(path, arr)
            t = fn(path)
            assert tuple(t.shape) == (1, 240, 480, 480) and t.dtype == torch.float32
            a = t.numpy()
            out["cases"].append({"shape": list(shape), "dtype": dtype, "seed": seed, "sha256": digest(a),
                                 "probes": [[list(p), float(a[p])] for p in PROBES],
                                 "n_pad": int((a == -1.0).sum()), "sum": float(a.astype(np.float64).sum())})
            print(shape, dtype, out["cases"][-1]["sha256"][:16], out["cases"][-1]["n_pad"])
    dst = os.path.join(ROOT, "tests", "golden", "volume_prep_golden.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    main()
