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
