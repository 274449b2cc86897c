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
    deps = [SRC, os.path.join(ROOT, "vit_exp_b200", "csrc", "volume_prep_math.cuh")]
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
