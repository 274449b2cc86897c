"""Thin tensor-level wrappers over the C ABI (one Python function per ctk_* entry point).

Every function takes CUDA tensors, validates dtype/contiguity, and launches on the current
torch stream. Nothing here computes on the host and nothing falls back to torch ops.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import (EPI_ARGMAX, EPI_ATOMIC_F32, EPI_BF16, EPI_F32, EPI_GEGLU, EPI_GEGLU_BWD,
                   EPI_QKV, EPI_RESID_F32, GemmEpilogue, check)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    assert t.is_cuda, "libctk has no CPU path: tensor must live on a CUDA device"
    return t.data_ptr()


def _chk(t: torch.Tensor, dtype, name: str):
    assert t.is_cuda and t.dtype == dtype and t.is_contiguous(), f"{name}: need contiguous CUDA {dtype}, got {t.dtype} {t.device}"


# --------------------------------------------------------------------------- GEMM
def gemm(a: torch.Tensor, b: torch.Tensor, epilogue: int, c: torch.Tensor, *, M: int, N: int, K: int,
         mn_major: bool = False, lda: Optional[int] = None, ldb: Optional[int] = None,
         ldc: Optional[int] = None, bias=None, resid=None, ldr: Optional[int] = None, aux0=None,
         ld_aux0: int = 0, vec0=None, vec1=None, row_map=None, alpha: float = 1.0, split_k: int = 0, i0: int = 0, i1: int = 0):
    """D = A[M,K] B[N,K]^T with a fused epilogue (ctk_gemm_bf16). a, b are bf16; K-major:
    a [M, lda>=K], b [N, ldb>=K]; MN-major: a [K, lda>=M], b [K, ldb>=N]."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    lib = _lib.load()
    e = GemmEpilogue()
    e.C = _p(c)
    e.ldc = ldc if ldc is not None else (c.stride(0) if c.dim() >= 2 else 0)
    e.bias = _p(bias)
    e.resid = _p(resid)
    e.ldr = ldr if ldr is not None else (resid.stride(0) if resid is not None else 0)
    e.aux0 = _p(aux0)
    e.ld_aux0 = ld_aux0
    e.vec0 = _p(vec0)
    e.vec1 = _p(vec1)
    e.row_map = _p(row_map)
    e.alpha = alpha
    e.i0 = i0
    e.i1 = i1
    lda = lda if lda is not None else a.stride(0)
    ldb = ldb if ldb is not None else b.stride(0)
    check(lib.ctk_gemm_bf16(_p(a), lda, int(mn_major), _p(b), ldb, int(mn_major), M, N, K, epilogue,
                            C.byref(e), split_k, _stream()), "ctk_gemm_bf16")
    return c


# --------------------------------------------------------------------------- weight prep
def cast_bf16(src: torch.Tensor, ld: Optional[int] = None, col_scale=None, out=None) -> torch.Tensor:
    _chk(src, torch.float32, "src")
    rows, cols = src.shape
    ld = ld or cols
    out = out if out is not None else torch.empty(rows, ld, dtype=torch.bfloat16, device=src.device)
    check(_lib.load().ctk_cast_bf16(_p(src), _p(out), rows, cols, ld, _p(col_scale), _stream()), "ctk_cast_bf16")
    return out


def transpose_cast_bf16(src: torch.Tensor, ld: Optional[int] = None, out=None) -> torch.Tensor:
    _chk(src, torch.float32, "src")
    rows, cols = src.shape
    ld = ld or rows
    out = out if out is not None else torch.empty(cols, ld, dtype=torch.bfloat16, device=src.device)
    check(_lib.load().ctk_transpose_cast_bf16(_p(src), _p(out), rows, cols, ld, _stream()), "ctk_transpose_cast_bf16")
    return out


def pack_ff_w1(w1: torch.Tensor, inner: int, inner_pad: int, want_t: bool = True):
    _chk(w1, torch.float32, "w1")
    dim = w1.shape[1]
    dev = w1.device
    dst = torch.empty(2 * inner_pad, dim, dtype=torch.bfloat16, device=dev)
    dst_t = torch.empty(dim, 2 * inner_pad, dtype=torch.bfloat16, device=dev) if want_t else None
    row_map = torch.empty(2 * inner_pad, dtype=torch.int32, device=dev)
    check(_lib.load().ctk_pack_ff_w1(_p(w1), _p(dst), _p(dst_t), _p(row_map), inner, inner_pad, dim, _stream()),
          "ctk_pack_ff_w1")
    return dst, dst_t, row_map


# --------------------------------------------------------------------------- contrastive head
def mean_pool(x: torch.Tensor) -> torch.Tensor:
    _chk(x, torch.float32, "x")
    B, n, dim = x.shape
    out = torch.empty(B, dim, dtype=torch.float32, device=x.device)
    check(_lib.load().ctk_mean_pool_fwd(_p(x), _p(out), B, n, dim, _stream()), "ctk_mean_pool_fwd")
    return out


def latent_fwd(x: torch.Tensor, W: torch.Tensor):
    """x fp32 [B, din] (rows may be strided views, last dim contiguous); W fp32 [dl, din]."""
    assert x.dtype == torch.float32 and x.stride(-1) == 1 and x.dim() == 2
    _chk(W, torch.float32, "W")
    B, din = x.shape
    dl = W.shape[0]
    lat = torch.empty(B, dl, dtype=torch.float32, device=x.device)
    rn = torch.empty(B, dtype=torch.float32, device=x.device)
    check(_lib.load().ctk_latent_fwd(_p(x), x.stride(0), _p(W), _p(lat), _p(rn), B, din, dl, _stream()), "ctk_latent_fwd")
    return lat, rn


def latent_bwd(dlat, lat, rn, x, W, need_dx: bool = True):
    B, din = x.shape
    dl = W.shape[0]
    _chk(dlat, torch.float32, "dlat")
    dW = torch.empty_like(W)
    dx = torch.empty(B, din, dtype=torch.float32, device=x.device) if need_dx else None
    check(_lib.load().ctk_latent_bwd(_p(dlat), _p(lat), _p(rn), _p(x), x.stride(0), _p(W), _p(dW), _p(dx),
                                     din, B, din, dl, _stream()), "ctk_latent_bwd")
    return dW, dx


def clip_loss_fwd_bwd(T: torch.Tensor, I: torch.Tensor, log_temp: torch.Tensor, b_local: int, row0: int,
                      need_grad: bool = True):
    """Returns (out[2] = {loss, dlog_temp}, d_local[2, b_local, d] or None)."""
    _chk(T, torch.float32, "T")
    _chk(I, torch.float32, "I")
    assert log_temp.dtype == torch.float32 and log_temp.numel() == 1
    N, d = T.shape
    lib = _lib.load()
    nbytes = lib.ctk_clip_loss_ws_bytes(N, b_local)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=T.device)
    out = torch.empty(2, dtype=torch.float32, device=T.device)
    d_local = torch.empty(2, b_local, d, dtype=torch.float32, device=T.device) if need_grad else None
    check(lib.ctk_clip_loss_fwd_bwd(_p(T), _p(I), _p(log_temp), _p(out), _p(d_local), _p(ws), nbytes, N, d,
                                    b_local, row0, _stream()), "ctk_clip_loss_fwd_bwd")
    return out, d_local


def pair_logits(text_lat: torch.Tensor, image_lat: torch.Tensor, log_temp: torch.Tensor) -> torch.Tensor:
    _chk(text_lat, torch.float32, "text_lat")
    _chk(image_lat, torch.float32, "image_lat")
    P, d = text_lat.shape
    out = torch.empty(P, dtype=torch.float32, device=text_lat.device)
    check(_lib.load().ctk_pair_logits(_p(text_lat), _p(image_lat), _p(log_temp), _p(out), P, d, _stream()), "ctk_pair_logits")
    return out


def fill_(t: torch.Tensor, v: float):
    _chk(t, torch.float32, "t")
    check(_lib.load().ctk_fill_f32(_p(t), v, t.numel(), _stream()), "ctk_fill_f32")
    return t
