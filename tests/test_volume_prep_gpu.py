"""ctk_volume_prep (vit_exp_b200.data.npz_to_tensor) against the oracle pinned to scripts/data.py:49-111: BIT-EXACT.

Validated on a B200 in round 2 (6 passed).
"""
import json
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "volume_prep_golden.json")))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: "x".join(map(str, c["shape"])) + "_" + c["dtype"])
def test_kernel_matches_reference_digest(cuda_dev, case):
    from oracle import volume_prep_oracle as V
    from vit_exp_b200 import data
    arr = V.synthetic_volume(tuple(case["shape"]), case["dtype"], case["seed"])
    out = data.npz_to_tensor(arr, cuda_dev)
    assert out.shape == (1, 240, 480, 480) and out.dtype == torch.float32
    assert V.digest(out.cpu().numpy()) == case["sha256"]


def test_small_targets_nan_and_batch(cuda_dev, tmp_path):
    from oracle import volume_prep_oracle as V
    from vit_exp_b200 import data, ops
    arr = V.synthetic_volume((7, 9, 13), "float16", 5)
    arr[3, 4, 5] = np.nan
    out = torch.empty(1, 4, 12, 16, device=cuda_dev)
    ops.volume_prep(torch.from_numpy(arr).to(cuda_dev), out)
    ref = V.npz_array_to_tensor(arr, target_hwd=(12, 16, 4))
    assert np.array_equal(out.cpu().numpy(), ref, equal_nan=True)
    p = tmp_path / "v.npz"
    np.savez(p, V.synthetic_volume((250, 470, 490), "float32", 6))
    b = data.batch_to_tensor([str(p), str(p)], cuda_dev)
    assert b.shape == (2, 1, 240, 480, 480) and torch.equal(b[0], b[1])
    assert V.digest(b[0].cpu().numpy()) == V.digest(V.npz_array_to_tensor(np.load(p)["arr_0"]))
