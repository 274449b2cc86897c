"""Test doubles of the `vit_exp_b200.ops` entry points (plain torch, run anywhere).

TEST INFRASTRUCTURE ONLY: these let the `-m "not gpu"` suite check the *host-side orchestration* of a module
(which kernel gets which operand, the hand-derived backward chain, gradient bookkeeping) on a machine without
a GPU by monkeypatching `<module>.ops`.  Nothing under `vit_exp_b200/` imports this file; the product path has
no fallback and asserts on non-CUDA tensors.  Each double follows the contract documented in include/ctk.h.
"""
from __future__ import annotations

import math

import torch

from vit_exp_b200._lib import (EPI_ATOMIC_F32, EPI_BF16, EPI_F32, EPI_GELU, EPI_GELU_BWD, EPI_RESID_F32)  # noqa: F401

OPERAND = torch.bfloat16        # set to torch.float32 for exact checks of the host math
CALLS = []                      # (name, epilogue or None) per call, for launch-plan assertions


def _gelu(u):
    return 0.5 * u * (1.0 + torch.erf(u / math.sqrt(2.0)))


def _gelu_grad(u):
    cdf = 0.5 * (1.0 + torch.erf(u / math.sqrt(2.0)))
    pdf = torch.exp(-0.5 * u * u) / math.sqrt(2.0 * math.pi)
    return cdf + u * pdf


def gemm(a, b, epilogue, c, *, M, N, K, mn_major=False, lda=None, ldb=None, ldc=None, bias=None, resid=None,
         ldr=None, aux0=None, ld_aux0=0, vec0=None, vec1=None, row_map=None, alpha=1.0, split_k=0, i0=0, i1=0):
    """ctk_gemm_bf16: D = A[M,K] B[N,K]^T (K-major) or A[K,M]^T B[K,N] (MN-major), fp32 accumulate."""
    assert lda is None and ldb is None and row_map is None, "emulation covers the calls the text tower makes"
    assert a.dtype == OPERAND and b.dtype == OPERAND
    CALLS.append(("gemm", epilogue))
    if mn_major:
        assert a.shape[0] == K and b.shape[0] == K and a.shape[1] >= M and b.shape[1] >= N
        acc = a[:, :M].float().t() @ b[:, :N].float()
    else:
        assert a.shape == (M, K) and b.shape == (N, K), (a.shape, b.shape, M, N, K)
        acc = a.float() @ b.float().t()
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.shape == (N,)
        assert epilogue in (EPI_BF16, EPI_F32, EPI_RESID_F32, EPI_GELU), "only these epilogues add a bias"
        acc = acc + bias
    if epilogue == EPI_BF16:
        assert c.dtype == OPERAND and c.shape == (M, N)
        c.copy_((acc * alpha).to(c.dtype))
    elif epilogue == EPI_F32:
        assert c.dtype == torch.float32
        c.copy_(acc)
    elif epilogue == EPI_RESID_F32:
        assert c.dtype == torch.float32 and resid.dtype == torch.float32 and resid.shape == (M, N)
        c.copy_(acc + resid)
    elif epilogue == EPI_GELU:
        assert aux0.shape == (M, N) and ld_aux0 == N and c.dtype == OPERAND and aux0.dtype == OPERAND
        c.copy_(acc.to(c.dtype))
        aux0.copy_(_gelu(acc).to(aux0.dtype))
    elif epilogue == EPI_GELU_BWD:
        assert aux0.shape == (M, N) and ld_aux0 == N and c.dtype == OPERAND
        c.copy_((acc * _gelu_grad(aux0.float())).to(c.dtype))
    elif epilogue == EPI_ATOMIC_F32:
        assert c.dtype == torch.float32 and c.shape == (M, N) and (ldc is None or ldc == N)
        c.add_(alpha * acc)
    else:
        raise AssertionError(f"epilogue {epilogue} not emulated")
    return c


def cast_bf16(src, ld=None, col_scale=None, out=None):
    assert src.dtype == torch.float32 and src.is_contiguous() and ld is None and col_scale is None and out is None
    CALLS.append(("cast_bf16", None))
    return src.to(OPERAND)


def transpose_cast_bf16(src, ld=None, out=None):
    assert src.dtype == torch.float32 and src.is_contiguous() and ld is None and out is None
    CALLS.append(("transpose_cast_bf16", None))
    return src.t().contiguous().to(OPERAND)


def layernorm_fwd(x, gamma, beta=None, *, want_bf16=True, want_f32=False, want_raw=False, eps=1e-5,
                  perm_outer=0, perm_inner=0, save_stats=True):
    assert x.dtype == torch.float32 and x.is_contiguous() and perm_inner == 0
    CALLS.append(("layernorm_fwd", None))
    mean = x.mean(1)
    var = x.var(1, unbiased=False)
    rstd = torch.rsqrt(var + eps)
    y = (x - mean[:, None]) * rstd[:, None] * gamma
    if beta is not None:
        y = y + beta
    return (y.to(OPERAND) if want_bf16 else None, y if want_f32 else None, x.to(OPERAND) if want_raw else None,
            mean if save_stats else None, rstd if save_stats else None)


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta=None, *, dx=None, accum=False, perm_outer=0,
                  perm_inner=0, bcast_rows=0, dy_scale=1.0, dx_bf16=None):
    assert perm_inner == 0 and bcast_rows == 0 and dy.is_contiguous()
    CALLS.append(("layernorm_bwd", None))
    dyf = dy.float() * dy_scale
    xhat = (x - mean[:, None]) * rstd[:, None]
    dgamma.add_((dyf * xhat).sum(0))
    if dbeta is not None:
        dbeta.add_(dyf.sum(0))
    wdy = dyf * gamma
    res = rstd[:, None] * (wdy - wdy.mean(1, keepdim=True) - xhat * (wdy * xhat).mean(1, keepdim=True))
    if dx is None:
        assert not accum
        dx = res
    elif accum:
        dx.add_(res)
    else:
        dx.copy_(res)
    if dx_bf16 is not None:
        dx_bf16.copy_(dx.to(dx_bf16.dtype))
    return dx


def colsum_(dy, out):
    assert out.dtype == torch.float32 and out.shape == (dy.shape[1],)
    CALLS.append(("colsum", None))
    out.add_(dy.float().sum(0))
    return out
