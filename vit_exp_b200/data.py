"""Loader-side volume preparation on the GPU.

Mirror of the reference's `npz_to_tensor(path)` (scripts/data.py:49-111): same name, same result - a float32
`(1, 240, 480, 480)` tensor holding `(clip(x, -1, 1) + 1) / 2`, centre-cropped / padded with -1 - but the stored
array travels to the device in its stored dtype (float16 datasets: 2 bytes per voxel, half the PCIe bytes of the
float32 result the reference's loader ships) and the arithmetic runs in libctk (`ctk_volume_prep`), bit-exact with
the reference (tests/test_volume_prep_cpu.py pins the oracle to the reference function; tests/test_volume_prep_gpu.py
compares the kernel with the oracle).  There is no CPU path: without a CUDA device this raises.
"""
from __future__ import annotations

from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import ops

TARGET_DHW = (240, 480, 480)          # data.py:73 target_shape (480, 480, 240) in (h, w, d) order


def _as_stored_array(src) -> np.ndarray:
    if isinstance(src, (str, bytes)) or hasattr(src, "__fspath__"):
        src = np.load(src)["arr_0"]                       # data.py:51
    arr = np.ascontiguousarray(src)
    if arr.dtype not in (np.float32, np.float16):
        raise TypeError(f"npz_to_tensor: stored dtype {arr.dtype} is not one the reference datasets use "
                        "(float32: data_preprocess/preprocess_ctrate_train.py:103; float16: *_fp16 directories)")
    assert arr.ndim == 3, "stored volumes are (D, H, W)"
    return arr


def stage(src, pin: bool = True) -> torch.Tensor:
    """Host side of the transfer: the stored array as a (pinned) torch tensor in its stored dtype."""
    t = torch.from_numpy(_as_stored_array(src))
    return t.pin_memory() if pin and torch.cuda.is_available() else t


def npz_to_tensor(src: Union[str, np.ndarray, torch.Tensor], device: Optional[torch.device] = None,
                  out: Optional[torch.Tensor] = None, non_blocking: bool = True) -> torch.Tensor:
    """`src`: path of a .npz with `arr_0`, a numpy array, or a staged / device tensor (D, H, W).
    Returns float32 `(1, 240, 480, 480)` on `device` (default: current CUDA device); `out` reuses a buffer."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    assert device.type == "cuda", "vit_exp_b200.data has no CPU path (the reference's loader is the CPU path)"
    t = src if isinstance(src, torch.Tensor) else stage(src)
    if not t.is_cuda:
        t = t.to(device, non_blocking=non_blocking)       # H2D in the stored dtype
    if out is None:
        out = torch.empty((1,) + TARGET_DHW, dtype=torch.float32, device=device)
    ops.volume_prep(t.contiguous(), out)
    return out


def batch_to_tensor(srcs: Sequence, device: Optional[torch.device] = None) -> torch.Tensor:
    """(B, 1, 240, 480, 480): what the reference's DataLoader collates from `npz_to_tensor` results
    (CTCLIPTrainer.py:596 `batch["image"]`)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.empty((len(srcs), 1) + TARGET_DHW, dtype=torch.float32, device=device)
    for i, s in enumerate(srcs):
        npz_to_tensor(s, device, out=out[i])
    return out
