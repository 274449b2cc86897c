"""Numerical design check of the tensor-core contrastive loss (head.cu `clip_loss_tc`, GEMM epilogues LSE_PART /
CLIP_GRAD), on CPU: a line-by-line torch restatement of its four passes - bf16 hi/lo split laid along K, per-128-column
online statistics merged as `clip_reduce_kernel` does, the two [N, b_local] gradient stripes with the kernel's
(vec0, bias, diagonal) conventions, hi/lo split of the stripes, three MN-major products per latent gradient - against
the fp64 oracle.  It pins the precision the split-bf16 scheme can reach (the tolerances of
tests/test_clip_loss_tc_gpu.py are set an order of magnitude above it) and the index conventions of the passes; the
CUDA code itself is checked on the GPU.
"""
import math

import pytest
import torch

from oracle import ctclip_oracle as orc


def _split(x):
    hi = x.bfloat16().float()
    return hi, (x - hi).bfloat16().float()


def _online_stats(S, block=128, chunk=32):
    """LSE_PART: per row and 128-column block, (m, l, w) accumulated over 32-column chunks with rescaling"""
    N = S.shape[1]
    parts = []
    for c0 in range(0, N, block):
        m = torch.full((S.shape[0],), -math.inf)
        l = torch.zeros(S.shape[0])
        w = torch.zeros(S.shape[0])
        for cc in range(c0, min(c0 + block, N), chunk):
            v = S[:, cc:cc + chunk]
            mx = torch.maximum(m, v.max(dim=1).values)
            r = torch.exp(m - mx)
            e = torch.exp(v - mx[:, None])
            l = l * r + e.sum(1)
            w = w * r + (e * v).sum(1)
            m = mx
        parts.append((m, l, w))
    return parts


def _merge(parts):
    """clip_reduce_kernel: lse and sum(x e^x)/sum(e^x) per index from the block partials"""
    m = torch.stack([p[0] for p in parts]).max(dim=0).values
    l = sum(p[1] * torch.exp(p[0] - m) for p in parts)
    w = sum(p[2] * torch.exp(p[0] - m) for p in parts)
    return m + torch.log(l), w / l


@pytest.mark.parametrize("N,W,rank,lt", [(1024, 1, 0, 1.0), (1024, 8, 5, 1.0), (2048, 4, 1, 2.5), (1152, 4, 3, 0.0)])
def test_split_bf16_pipeline_reaches_fp32_level_accuracy(N, W, rank, lt):
    d, B, row0 = 512, N // W, rank * (N // W)
    g = torch.Generator().manual_seed(0)
    T = torch.nn.functional.normalize(torch.randn(N, d, generator=g), dim=-1)
    I = torch.nn.functional.normalize(torch.randn(N, d, generator=g) + 0.5 * T, dim=-1)
    ref = orc.clip_loss_and_local_grads(T.double(), I.double(), torch.tensor(lt).double(), B, rank)
    s = math.exp(lt)
    (Th, Tl), (Ih, Il) = _split(T), _split(I)
    catT = torch.cat([Th, Th, Tl], dim=1)                    # text rows  [hi | hi | lo]
    catI = torch.cat([Ih, Il, Ih], dim=1)                    # image rows [hi | lo | hi]
    # pass 1 + 2
    S_rows = (catT @ catI.T) * s                             # rows = texts
    S_cols = (catI @ catT.T) * s                             # rows = images
    row_lse, row_mean = _merge(_online_stats(S_rows))
    col_lse, col_mean = _merge(_online_stats(S_cols))
    diag = S_rows.diagonal()
    inv = 1.0 / (2.0 * N * B)
    loss = ((row_lse - diag).sum() + (col_lse - diag).sum()) * inv
    dtemp = ((row_mean - diag).sum() + (col_mean - diag).sum()) * inv
    # pass 3: gemm rows = all N, gemm cols = the rank's b_local; diagonal where gemm_row == gemm_col + row0
    eye = (torch.arange(N)[:, None] == torch.arange(B)[None, :] + row0).float()
    x0 = (catI @ catT[row0:row0 + B].T) * s                  # [images, local texts]
    g0 = (torch.exp(x0 - col_lse[:, None]) + torch.exp(x0 - row_lse[row0:row0 + B][None, :]) - 2 * eye) * inv * s
    x1 = (catT @ catI[row0:row0 + B].T) * s                  # [texts, local images]
    g1 = (torch.exp(x1 - row_lse[:, None]) + torch.exp(x1 - col_lse[row0:row0 + B][None, :]) - 2 * eye) * inv * s
    # pass 4: hi.hi + hi.lo + lo.hi, contraction over the N gathered samples
    (g0h, g0l), (g1h, g1l) = _split(g0), _split(g1)
    dT = g0h.T @ Ih + g0h.T @ Il + g0l.T @ Ih
    dI = g1h.T @ Th + g1h.T @ Tl + g1l.T @ Th
    assert abs(loss.double() - ref["loss"]) <= 1e-6 * abs(ref["loss"])
    assert abs(dtemp.double() - ref["dlog_temp"]) <= 1e-4 * abs(ref["dlog_temp"]) + 1e-8
    for got, want in ((dT, ref["dT_local"]), (dI, ref["dI_local"])):
        assert (got.double() - want).abs().max() <= 1e-4 * want.abs().max()
