"""CPU oracle for the loader-side volume preparation.  TEST INFRASTRUCTURE ONLY
(same rules as oracle/ctclip_oracle.py: imported by tests/, smoke() and bench.py's CPU legs only).

Restates scripts/data.py:49-111 `npz_to_tensor` on an in-memory array, in numpy:
  * the stored array `arr_0` is (D, H, W); the reference transposes it to (H, W, D) (data.py:53), works there
    and permutes back at the end (data.py:104), so the net layout change is none;
  * clip to [-1, 1], then (x + 1) / 2 IN THE STORED DTYPE (data.py:59-61: numpy keeps float16 arithmetic for a
    float16 array, python scalars are weak), then cast to float32 (data.py:62);
  * centre crop to at most (480, 480, 240) per axis (data.py:77-85), centre pad with the constant -1 up to exactly
    that shape (data.py:87-98) - the pad value is -1 even though valid data now lives in [0, 1];
  * add the channel axis: (1, 240, 480, 480) (data.py:104-106).
Pinned by tests/golden/volume_prep_golden.json, produced by oracle/make_golden_volume.py, which executes the
reference's own function body (extracted from /root/reference/scripts/data.py with `ast`, not copied) on
synthetic .npz files.
"""
from __future__ import annotations

import hashlib

import numpy as np

TARGET_HWD = (480, 480, 240)      # data.py:73


def synthetic_volume(shape, dtype, seed: int = 0) -> np.ndarray:
    """Deterministic (formula-based, no RNG) stand-in for a preprocessed CT-RATE array: values in [-2.2, 2.2] so that
    both clip bounds are exercised; shape is (D, H, W)."""
    d, h, w = shape
    z = np.arange(d, dtype=np.int64)[:, None, None]
    y = np.arange(h, dtype=np.int64)[None, :, None]
    x = np.arange(w, dtype=np.int64)[None, None, :]
    v = (z * 7919 + y * 104729 + x * 1299709 + seed * 15485863) % 4401
    return (v.astype(np.float64) / 1000.0 - 2.2).astype(dtype)


def axis_plan(n: int, target: int):
    """(start, length, pad_before) of one axis: data.py:77-98."""
    start = max((n - target) // 2, 0)
    end = min(start + target, n)
    length = end - start
    pad_before = (target - length) // 2
    return start, length, pad_before


def npz_array_to_tensor(arr: np.ndarray, target_hwd=TARGET_HWD) -> np.ndarray:
    """float32 (1, D_t, H_t, W_t) exactly as data.py:49-111 returns it for np.load(path)['arr_0'] == arr."""
    assert arr.ndim == 3
    lo, hi = -1, 1
    x = np.clip(arr, lo, hi)
    x = (x - lo) / (hi - lo)                       # stays in arr.dtype for floating arrays (weak python scalars)
    x = x.astype(np.float32)
    th, tw, td = target_hwd
    out = np.full((td, th, tw), -1.0, dtype=np.float32)
    (z0, zl, zp), (y0, yl, yp), (x0, xl, xp) = (axis_plan(n, t) for n, t in zip(arr.shape, (td, th, tw)))
    out[zp:zp + zl, yp:yp + yl, xp:xp + xl] = x[z0:z0 + zl, y0:y0 + yl, x0:x0 + xl]
    return out[None]


def digest(t: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(t, dtype=np.float32).tobytes()).hexdigest()
