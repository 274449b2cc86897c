"""The kernels' per-element code, run on the CPU.

`vit_exp_b200/csrc/volume_prep_math.cuh` holds the __host__ __device__ arithmetic and index plan that
`ctk_volume_prep`'s kernel executes per output vector.  tests/host_checks.cu loops over the same functions on the
host (a test-only shared library built here with nvcc - no GPU needed to build or run it), so the arithmetic (fp16
double rounding like numpy, NaN propagation) and the crop / pad indexing are compared with the oracle, and with the
golden digests of the reference function, without a GPU.  Only the kernel launch itself is left to the GPU suite.
"""
import ctypes
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import volume_prep_oracle as V

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_checks.cu")
OUT = os.path.join(ROOT, "tests", "_build", "libctk_hostcheck.so")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "volume_prep_golden.json")))


@pytest.fixture(scope="module")
def lib():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC] + [os.path.join(ROOT, "vit_exp_b200", "csrc", h) for h in ("volume_prep_math.cuh", "clip_epilogue_math.cuh", "mha_dropout.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        r = subprocess.run([nvcc, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets",
                            "-x", "cu", SRC, "-o", OUT], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    lib = ctypes.CDLL(OUT)
    lib.hostcheck_volume_prep.restype = None
    lib.hostcheck_volume_prep.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_int] * 3 + [ctypes.c_void_p] + \
                                         [ctypes.c_int] * 3
    return lib


def _run(lib, arr, target_dhw):
    arr = np.ascontiguousarray(arr)
    out = np.empty(target_dhw, dtype=np.float32)
    lib.hostcheck_volume_prep(arr.ctypes.data, int(arr.dtype == np.float16), *arr.shape, out.ctypes.data, *target_dhw)
    return out


def test_every_float16_value(lib):
    """all 65536 bit patterns: the kernel's fp16 path is bit-identical to numpy's (x + 1) / 2 in float16, NaNs included"""
    bits = np.arange(65536, dtype=np.uint16)
    arr = bits.view(np.float16).reshape(1, 1, 65536)
    got = _run(lib, arr, (1, 1, 65536))
    want = V.npz_array_to_tensor(arr, target_hwd=(1, 65536, 1))[0]
    assert np.array_equal(got.view(np.uint32)[~np.isnan(want)], want.view(np.uint32)[~np.isnan(want)])
    assert np.array_equal(np.isnan(got), np.isnan(want))


def test_float32_values(lib):
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.standard_normal(200000).astype(np.float32) * 1.5,
                           np.array([0.0, -0.0, 1.0, -1.0, np.nextafter(np.float32(1), np.float32(2)), -1.0000001, 1e-45,
                                     -1e-45, 3e38, -3e38, np.inf, -np.inf, np.nan], dtype=np.float32)])
    vals = np.resize(vals, (vals.size + 3) // 4 * 4)
    arr = vals.reshape(1, 1, -1)
    got = _run(lib, arr, (1, 1, arr.shape[2]))
    want = V.npz_array_to_tensor(arr, target_hwd=(1, arr.shape[2], 1))[0]
    ok = ~np.isnan(want)
    assert np.array_equal(got.view(np.uint32)[ok], want.view(np.uint32)[ok]) and np.array_equal(np.isnan(got), ~ok)


@pytest.mark.parametrize("shape,target", [((7, 9, 13), (4, 12, 16)), ((3, 20, 8), (6, 8, 8)), ((10, 10, 10), (10, 12, 8)),
                                          ((1, 1, 1), (2, 2, 4)), ((9, 5, 21), (9, 5, 20))])
@pytest.mark.parametrize("dtype", ["float16", "float32"])
def test_crop_pad_indexing_small(lib, shape, target, dtype):
    arr = V.synthetic_volume(shape, dtype, seed=7)
    got = _run(lib, arr, target)
    want = V.npz_array_to_tensor(arr, target_hwd=(target[1], target[2], target[0]))[0]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("case", GOLD["cases"][1:4], ids=lambda c: "x".join(map(str, c["shape"])) + "_" + c["dtype"])
def test_full_size_digest_of_reference_function(lib, case):
    """full 240x480x480 targets: the SHA-256 the reference's own npz_to_tensor produced"""
    arr = V.synthetic_volume(tuple(case["shape"]), case["dtype"], case["seed"])
    got = _run(lib, arr, (240, 480, 480))
    assert V.digest(got) == case["sha256"]


# ------------------------------------------------------------------------------------------------
# contrastive-loss GEMM epilogues: the device functions of csrc/clip_epilogue_math.cuh over host accumulators
# ------------------------------------------------------------------------------------------------
def _bind_clip(lib):
    f32p, i, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.hostcheck_lse_part.restype = None
    lib.hostcheck_lse_part.argtypes = [f32p, i, i, f, i, f32p, f32p]
    lib.hostcheck_clip_grad.restype = None
    lib.hostcheck_clip_grad.argtypes = [f32p, i, i, f, f, f32p, f32p, i, i, f32p, f32p]


def _merge(part):
    """clip_reduce_kernel: (lse, sum(x e^x) / sum(e^x)) per row from the [blocks, M, 3] partials"""
    m = part[:, :, 0].max(axis=0)
    sc = np.exp(part[:, :, 0] - m)
    l = (part[:, :, 1] * sc).sum(axis=0)
    w = (part[:, :, 2] * sc).sum(axis=0)
    return m + np.log(l), w / l


@pytest.mark.parametrize("N,W,rank,lt", [(256, 1, 0, 1.0), (384, 4, 2, 2.0), (200, 1, 0, 0.5)])
def test_clip_epilogue_device_functions_vs_oracle(lib, N, W, rank, lt):
    """the loss pipeline of head.cu `clip_loss_tc` with the REAL epilogue functions (run on the host) and numpy for
    the matrix products: loss, d(log temperature) and both latent gradients against the fp64 oracle.  N = 200 has a
    ragged last block (LSE_PART only; the gradient pass needs b_local % 32 == 0 and is skipped there)."""
    import torch
    from oracle import ctclip_oracle as orc
    _bind_clip(lib)
    d, B, row0 = 64, N // W, rank * (N // W)
    g = torch.Generator().manual_seed(0)
    T = torch.nn.functional.normalize(torch.randn(N, d, generator=g), dim=-1)
    I = torch.nn.functional.normalize(torch.randn(N, d, generator=g) + 0.5 * T, dim=-1)
    ref = orc.clip_loss_and_local_grads(T.double(), I.double(), torch.tensor(lt).double(), B, rank)

    def split(x):
        hi = x.bfloat16().float()
        return hi.numpy(), (x - hi).bfloat16().float().numpy()
    (Th, Tl), (Ih, Il) = split(T), split(I)
    catT, catI = np.concatenate([Th, Th, Tl], 1), np.concatenate([Ih, Il, Ih], 1)
    nblk = (N + 127) // 128

    def lse_pass(A, Bm, want_diag):
        acc = np.ascontiguousarray(A @ Bm.T, dtype=np.float32)
        part = np.full((nblk, N, 3), np.nan, dtype=np.float32)
        diag = np.full(N, np.nan, dtype=np.float32)
        lib.hostcheck_lse_part(acc.ctypes.data, N, N, lt, 0, part.ctypes.data, diag.ctypes.data if want_diag else None)
        return part, diag
    rowpart, diag = lse_pass(catT, catI, True)
    colpart, _ = lse_pass(catI, catT, False)
    assert np.isfinite(rowpart).all() and np.isfinite(colpart).all() and np.isfinite(diag).all()
    row_lse, row_mean = _merge(rowpart)
    col_lse, col_mean = _merge(colpart)
    inv = 1.0 / (2.0 * N * B)
    loss = ((row_lse - diag).sum() + (col_lse - diag).sum()) * inv
    dtemp = ((row_mean - diag).sum() + (col_mean - diag).sum()) * inv
    assert abs(loss - float(ref["loss"])) <= 2e-6 * abs(float(ref["loss"]))
    assert abs(dtemp - float(ref["dlog_temp"])) <= 1e-4 * abs(float(ref["dlog_temp"])) + 1e-8
    if B % 32:
        return

    def grad_pass(A, Bm, vec0, bias):
        acc = np.ascontiguousarray(A @ Bm.T, dtype=np.float32)                      # [N, b_local]
        hi, lo = np.empty_like(acc), np.empty_like(acc)
        v0, bb = np.ascontiguousarray(vec0, dtype=np.float32), np.ascontiguousarray(bias, dtype=np.float32)
        lib.hostcheck_clip_grad(acc.ctypes.data, N, B, lt, inv, v0.ctypes.data, bb.ctypes.data, 0, row0,
                                hi.ctypes.data, lo.ctypes.data)
        assert np.array_equal(hi, torch.from_numpy(hi).bfloat16().float().numpy())     # representable in bf16
        return hi, lo
    g0h, g0l = grad_pass(catI, catT[row0:row0 + B], col_lse, row_lse[row0:row0 + B])
    g1h, g1l = grad_pass(catT, catI[row0:row0 + B], row_lse, col_lse[row0:row0 + B])
    dT = g0h.T @ Ih + g0h.T @ Il + g0l.T @ Ih
    dI = g1h.T @ Th + g1h.T @ Tl + g1l.T @ Th
    for got, want in ((dT, ref["dT_local"].numpy()), (dI, ref["dI_local"].numpy())):
        assert np.abs(got - want).max() <= 2e-4 * np.abs(want).max()


def test_mha_dropout_mask_matches_torch_restatement(lib):
    """csrc/mha_dropout.cuh (what ctk_mha_fwd / ctk_mha_bwd evaluate per probability) == tests/emulated_ops.mha_keep_mask,
    the torch restatement the GPU parity tests build their reference attention from; and the mask is a fair coin."""
    import ctypes as C
    import sys
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emulated_ops as E
    lib.hostcheck_mha_keep.restype = None
    lib.hostcheck_mha_keep.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_int, C.c_int, C.c_float, C.c_void_p]
    for base, off, nbh, L, p in [(1234567890123456789, 0, 3, 70, 0.1), (2 ** 63 + 17, 11, 2, 130, 0.5), (0, 0, 1, 64, 0.9)]:
        got = np.empty((nbh, L, L), dtype=np.uint8)
        lib.hostcheck_mha_keep(base, off, nbh, L, p, got.ctypes.data)
        want = E.mha_keep_mask(E.mha_seed(base, off), nbh, L, p).numpy().astype(np.uint8)
        assert np.array_equal(got, want)
        frac = got.mean()
        assert abs(frac - (1 - p)) < 4 * np.sqrt(p * (1 - p) / got.size) + 1e-3, (p, frac)
    # rows / columns are not correlated: per-row keep rates spread like a binomial
    got = np.empty((4, 256, 256), dtype=np.uint8)
    lib.hostcheck_mha_keep(99, 3, 4, 256, 0.1, got.ctypes.data)
    assert abs(got.mean(axis=2).std() - np.sqrt(0.09 / 256)) < 0.006 and abs(got.mean(axis=1).std() - np.sqrt(0.09 / 256)) < 0.006
