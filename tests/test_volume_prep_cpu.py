"""Loader-side volume preparation (scripts/data.py:49-111 `npz_to_tensor`): the numpy oracle against golden digests
produced by the reference's own function (oracle/make_golden_volume.py), plus the index plan the CUDA kernel shares."""
import json
import os

import numpy as np
import pytest

from oracle import volume_prep_oracle as V

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "volume_prep_golden.json")))


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: "x".join(map(str, c["shape"])) + "_" + c["dtype"])
def test_oracle_matches_reference_digest(case):
    arr = V.synthetic_volume(tuple(case["shape"]), case["dtype"], case["seed"])
    out = V.npz_array_to_tensor(arr)
    assert out.shape == (1, 240, 480, 480) and out.dtype == np.float32
    for idx, want in case["probes"]:
        assert float(out[tuple(idx)]) == want
    assert int((out == -1.0).sum()) == case["n_pad"]
    assert V.digest(out) == case["sha256"]              # bit-exact with the reference function


def test_float16_arithmetic_is_done_in_float16():
    """data.py:59-61 keeps the stored dtype: (x + 1) / 2 rounds in fp16 for fp16 volumes (differs from fp32 math)."""
    arr = np.array([[[0.3337, -0.7771, 0.0001, 0.9995]]], dtype=np.float16)
    out = V.npz_array_to_tensor(arr, target_hwd=(1, 4, 1))[0, 0, 0]
    want16 = ((arr[0, 0] + np.float16(1)) / np.float16(2)).astype(np.float32)
    assert np.array_equal(out, want16)
    assert not np.array_equal(out, (arr[0, 0].astype(np.float32) + 1) / 2)


@pytest.mark.parametrize("n,t", [(480, 480), (512, 480), (400, 480), (479, 480), (481, 480), (1, 240), (1000, 240)])
def test_axis_plan_covers_target(n, t):
    start, length, pad = V.axis_plan(n, t)
    assert 0 <= start and start + length <= n and 0 < length <= t and 0 <= pad and pad + length <= t
    if n >= t:
        assert length == t and pad == 0 and start == (n - t) // 2
    else:
        assert start == 0 and length == n and pad == (t - n) // 2


def test_data_module_host_side(tmp_path):
    """vit_exp_b200.data: staging keeps the stored dtype, unsupported dtypes are refused, and there is no CPU path."""
    import torch

    from vit_exp_b200 import data
    arr16 = V.synthetic_volume((5, 6, 7), "float16", 1)
    p = tmp_path / "v.npz"
    np.savez(p, arr16)
    t = data.stage(str(p), pin=False)
    assert t.dtype == torch.float16 and tuple(t.shape) == (5, 6, 7) and np.array_equal(t.numpy(), arr16)
    assert data.stage(arr16.astype(np.float32), pin=False).dtype == torch.float32
    with pytest.raises(TypeError):
        data.stage(arr16.astype(np.float64), pin=False)
    with pytest.raises(AssertionError):
        data.npz_to_tensor(arr16, device="cpu")
    assert data.TARGET_DHW == (V.TARGET_HWD[2], V.TARGET_HWD[0], V.TARGET_HWD[1])
