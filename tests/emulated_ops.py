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

from vit_exp_b200.ops import ZeroArena  # noqa: E402,F401  (plain torch: a pre-zeroed buffer handed out in slices)

OPERAND = torch.bfloat16        # set to torch.float32 for exact checks of the host math
CALLS = []                      # (name, epilogue or None) per call, for launch-plan assertions


def _gelu(u):
    return 0.5 * u * (1.0 + torch.erf(u / math.sqrt(2.0)))


def _gelu_grad(u):
    cdf = 0.5 * (1.0 + torch.erf(u / math.sqrt(2.0)))
    pdf = torch.exp(-0.5 * u * u) / math.sqrt(2.0 * math.pi)
    return cdf + u * pdf


def gemm(a, b, epilogue, c, *, M, N, K, mn_major=False, lda=None, ldb=None, ldc=None, bias=None, resid=None,
         ldr=None, aux0=None, ld_aux0=0, vec0=None, vec1=None, row_map=None, alpha=1.0, split_k=0, i0=0, i1=0,
         b_mn_major=None):
    """ctk_gemm_bf16: D = A[M,K] B[N,K]^T (K-major) or A[K,M]^T B[K,N] (MN-major), fp32 accumulate."""
    assert lda is None and ldb is None and row_map is None, "emulation covers the calls the text tower makes"
    assert a.dtype == OPERAND and b.dtype == OPERAND
    CALLS.append(("gemm", epilogue))
    if b_mn_major and not mn_major:                 # K-major A, MN-major B: dX = dY W with W [K, N] as stored
        assert a.shape == (M, K) and b.shape == (K, N), (a.shape, b.shape, M, N, K)
        acc = a.float() @ b.float()
    elif mn_major:
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
    assert src.dtype == torch.float32 and src.is_contiguous() and ld is None and col_scale is None
    CALLS.append(("cast_bf16", None))
    if out is not None:
        assert out.shape == src.shape and out.dtype == OPERAND
        out.copy_(src.to(OPERAND))
        return out
    return src.to(OPERAND)


def transpose_cast_bf16_slice(src, dst):
    assert src.dtype == torch.float32 and dst.shape == (src.shape[1], src.shape[0]) and dst.dtype == OPERAND
    CALLS.append(("transpose_cast_bf16", None))
    dst.copy_(src.t().to(OPERAND))
    return dst


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


# ---- text tower: BertEmbeddings and the attention core (include/ctk.h: ctk_bert_embed_*, ctk_mha_*)
def bert_embed_fwd(ids, token_type, word, pos, typ):
    CALLS.append(("bert_embed_fwd", None))
    B, L = ids.shape
    e = word[ids] + (typ[0] if token_type is None else typ[token_type]) + pos[:L].unsqueeze(0)
    return e.reshape(B * L, -1).float().contiguous()


def bert_embed_bwd(de, ids, token_type, word_shape, pos_shape, typ_shape, pad_idx):
    CALLS.append(("bert_embed_bwd", None))
    B, L = ids.shape
    dword = torch.zeros(word_shape).index_add_(0, ids.reshape(-1), de)
    if pad_idx is not None:
        dword[pad_idx].zero_()
    dpos = torch.zeros(pos_shape)
    dpos[:L] = de.view(B, L, -1).sum(0)
    dtyp = torch.zeros(typ_shape)
    if token_type is None:
        dtyp[0] = de.sum(0)
    else:
        dtyp.index_add_(0, token_type.reshape(-1), de)
    return dword, dpos, dtyp


_M32 = 0xFFFFFFFF


def _mix32(x):
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & _M32
    return x ^ (x >> 16)


def mha_hash(seed: int, bh, i, j):
    """torch restatement of csrc/mha_dropout.cuh: mha_hash (uint32 arithmetic carried in int64 tensors)"""
    bh, i, j = (torch.as_tensor(v, dtype=torch.int64) for v in (bh, i, j))
    x = ((i * 0x9E3779B1) & _M32) ^ ((j * 0x85EBCA77) & _M32) ^ ((bh * 0xC2B2AE3D) & _M32) ^ (seed & _M32)
    x = _mix32(x)
    x = (x + ((seed >> 32) & _M32) + ((j * 0x27D4EB2F) & _M32)) & _M32
    return _mix32(x)


def mha_seed(base: int, offset: int) -> int:
    return (base + offset * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF


def mha_keep_mask(seed: int, nbh: int, L: int, p: float):
    """bool [nbh, L, L]: element (bh, i, j) survives attention-probability dropout"""
    thresh = min(int(p * 4294967296.0), 0xFFFFFFFF)
    bh = torch.arange(nbh).view(-1, 1, 1)
    i = torch.arange(L).view(1, -1, 1)
    j = torch.arange(L).view(1, 1, -1)
    return mha_hash(seed, bh, i, j) >= thresh


def _mha_math(qkv, key_mask, B, L, heads, scale, p_drop, seed, seed_off, null_k=None, null_v=None):
    """fp32 softmax attention on the packed projections; null pairs [heads, n_null, dh] are extra keys every query sees
    (appended here: the order of keys does not change a softmax; the kernels hash them at key index ceil(L/64)*64 + j)."""
    H = qkv.shape[1] // 3
    q, k, v = (qkv.float().view(B, L, 3, heads, H // heads)[:, :, t].transpose(1, 2) for t in range(3))
    valid = torch.ones(B, 1, 1, L, dtype=torch.bool) if key_mask is None else (key_mask.view(B, 1, 1, L) != 0)
    n_null = 0
    if null_k is not None:
        n_null = null_k.shape[1]
        k = torch.cat([k, null_k.float()[None].expand(B, -1, -1, -1)], dim=2)
        v = torch.cat([v, null_v.float()[None].expand(B, -1, -1, -1)], dim=2)
        valid = torch.cat([valid, torch.ones(B, 1, 1, n_null, dtype=torch.bool)], dim=-1)
    sc = (q @ k.transpose(-1, -2)) * scale
    sc = sc.masked_fill(~valid, float("-inf"))
    lse = torch.logsumexp(sc, dim=-1)
    pr = torch.softmax(sc, dim=-1)
    if p_drop > 0:
        assert n_null == 0, "the doubles restate the dropout mask for plain keys only"
        keep = mha_keep_mask(mha_seed(int(seed.item()) & 0xFFFFFFFFFFFFFFFF, seed_off), B * heads, L, p_drop)
        pr = pr * keep.view(B, heads, L, L).to(pr.dtype) / (1.0 - p_drop)
    return (pr @ v).transpose(1, 2).reshape(B * L, H), lse


def mha_fwd(qkv, key_mask, B, L, heads, scale, p_drop=0.0, seed=None, seed_off=0, null_k=None, null_v=None):
    CALLS.append(("mha_fwd", p_drop))
    assert qkv.dtype == OPERAND
    ctx, lse = _mha_math(qkv, key_mask, B, L, heads, scale, p_drop, seed, seed_off, null_k, null_v)
    return ctx.to(OPERAND), lse


def mha_bwd(qkv, key_mask, out, dout, lse, B, L, heads, scale, p_drop=0.0, seed=None, seed_off=0, null_k=None, null_v=None):
    CALLS.append(("mha_bwd", p_drop))
    with torch.enable_grad():
        x = qkv.detach().float().requires_grad_(True)
        leaves = [x]
        nk = nv = None
        if null_k is not None:
            nk = null_k.detach().float().requires_grad_(True)
            nv = null_v.detach().float().requires_grad_(True)
            leaves += [nk, nv]
        ctx, _ = _mha_math(x, key_mask, B, L, heads, scale, p_drop, seed, seed_off, nk, nv)
        grads = torch.autograd.grad(ctx, leaves, dout.float())
    if null_k is None:
        return grads[0].to(OPERAND)
    # per-sequence slabs [B, heads, n_null, dh] as the kernel writes them: put the whole gradient in slab 0
    dnk = torch.zeros(B, *null_k.shape)
    dnv = torch.zeros(B, *null_v.shape)
    dnk[0], dnv[0] = grads[1], grads[2]
    return grads[0].to(OPERAND), dnk, dnv


def colsum_(dy, out):
    assert out.dtype == torch.float32 and out.shape == (dy.shape[1],)
    CALLS.append(("colsum", None))
    out.add_(dy.float().sum(0))
    return out


# ================================================================================================
# image encoder + contrastive head (contracts: include/ctk.h; callers: vit_exp_b200/transformer_maskgit.py,
# vit_exp_b200/ct_clip.py).  `gemm_full` extends `gemm` with the epilogues / operand views those callers use.
# ================================================================================================
import torch.nn.functional as F  # noqa: E402

from vit_exp_b200._lib import EPI_ARGMAX, EPI_GEGLU, EPI_GEGLU_BWD, EPI_QKV  # noqa: E402

GRAPH_LAUNCHES = 0
GEMM_PROFILE = None


def gemm_full(a, b, epilogue, c, *, M, N, K, mn_major=False, lda=None, ldb=None, ldc=None, bias=None, resid=None,
              ldr=None, aux0=None, ld_aux0=0, vec0=None, vec1=None, row_map=None, alpha=1.0, split_k=0, i0=0, i1=0,
              b_mn_major=None):
    """all epilogues; operands may be column-sliced views (their strides play the role of lda / ldb)."""
    CALLS.append(("gemm", epilogue))
    if b_mn_major and not mn_major:
        acc = a[:M, :K].float() @ b[:K, :N].float()
    elif mn_major:
        acc = a[:K, :M].float().t() @ b[:K, :N].float()
    else:
        acc = a[:M, :K].float() @ b[:N, :K].float().t()
    if bias is not None:
        acc = acc + bias
    if epilogue == EPI_BF16:
        c[:M, :N] = (acc * alpha).to(c.dtype)
    elif epilogue == EPI_F32:
        c[:M, :N] = acc
    elif epilogue == EPI_RESID_F32:
        c[:M, :N] = acc + resid[:M, :N]
    elif epilogue == EPI_GEGLU:                     # U interleaved per 128 hidden units: 128 value | 128 gate
        u = acc.view(M, N // 256, 2, 128)
        c[:M, :N] = acc.to(c.dtype)
        aux0[:M, : N // 2] = (_gelu(u[:, :, 1]) * u[:, :, 0]).reshape(M, N // 2).to(aux0.dtype)
    elif epilogue == EPI_GEGLU_BWD:                 # acc = dH [M, N]; aux0 = U [M, 2N]; c = dU [M, 2N]
        u = aux0[:M, : 2 * N].float().view(M, N // 128, 2, 128)
        dh = acc.view(M, N // 128, 128)
        val, gate = u[:, :, 0], u[:, :, 1]
        du = torch.stack([dh * _gelu(gate), dh * val * _gelu_grad(gate)], dim=2)
        c[:M, : 2 * N] = du.reshape(M, 2 * N).to(c.dtype)
    elif epilogue == EPI_QKV:
        out = acc.clone()
        nh = i0 // 32
        if nh:
            hd = acc[:, :i0].view(M, nh, 32)
            rn = 1.0 / hd.norm(dim=-1).clamp_min(1e-12)
            out[:, :i0] = (hd * rn[..., None] * vec0 * alpha).reshape(M, i0)
            aux0[:M, i1 // 32: i1 // 32 + nh] = rn
        c[:M, i1: i1 + N] = out.to(c.dtype)
    elif epilogue == EPI_ATOMIC_F32:
        upd = alpha * acc
        if row_map is None:
            c[:M, :N] += upd
        else:
            keep = row_map[:M] >= 0
            c.index_add_(0, row_map[:M][keep].long(), upd[keep])
    elif epilogue == EPI_ARGMAX:                    # the double stores the winning column itself (first maximum)
        c[:M] = acc.argmax(dim=1)
    else:
        raise AssertionError(f"epilogue {epilogue} not emulated")
    return c


def cast_bf16_full(src, ld=None, col_scale=None, out=None):
    rows, cols = src.shape
    ld = ld or cols
    v = src if col_scale is None else src * col_scale
    res = torch.zeros(rows, ld, dtype=OPERAND) if out is None else out
    res[:, :cols] = v.to(OPERAND)
    return res


def transpose_cast_bf16_full(src, ld=None, out=None):
    rows, cols = src.shape
    ld = ld or rows
    res = torch.zeros(cols, ld, dtype=OPERAND) if out is None else out
    res[:, :rows] = src.t().to(OPERAND)
    return res


def pack_ff_w1(w1, inner, inner_pad, want_t=True):
    dim = w1.shape[1]
    dst = torch.zeros(2 * inner_pad, dim, dtype=OPERAND)
    row_map = torch.full((2 * inner_pad,), -1, dtype=torch.int32)
    for blk in range(inner_pad // 128):
        j0 = blk * 128
        n = max(0, min(128, inner - j0))
        if n:
            dst[blk * 256: blk * 256 + n] = w1[j0: j0 + n].to(OPERAND)                          # value rows
            dst[blk * 256 + 128: blk * 256 + 128 + n] = w1[inner + j0: inner + j0 + n].to(OPERAND)  # gate rows
            row_map[blk * 256: blk * 256 + n] = torch.arange(j0, j0 + n, dtype=torch.int32)
            row_map[blk * 256 + 128: blk * 256 + 128 + n] = torch.arange(inner + j0, inner + j0 + n, dtype=torch.int32)
    return dst, (dst.t().contiguous() if want_t else None), row_map


def patch_norm_fwd(video, pt, p1, p2, eps=1e-5, out=None):
    B, c, D, H, W = video.shape
    t, h, w = D // pt, H // p1, W // p2
    x = video.reshape(B, c, t, pt, h, p1, w, p2).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(B * t * h * w, -1)
    K = x.shape[1]
    ld = (K + 7) // 8 * 8
    mean = x.mean(1)
    rstd = torch.rsqrt(x.var(1, unbiased=False) + eps)
    xhat = torch.zeros(x.shape[0], ld, dtype=OPERAND)
    xhat[:, :K] = ((x - mean[:, None]) * rstd[:, None]).to(OPERAND)
    if out is not None:
        for dst, src in zip(out, (xhat, mean, rstd)):
            dst.copy_(src)
        return out
    return xhat, mean, rstd


def _perm_index(rows, perm_outer, perm_inner):
    """output row of input row r: (g*outer + o)*inner + i -> (g*inner + i)*outer + o"""
    r = torch.arange(rows)
    if perm_inner == 0:
        return r
    g, rem = r // (perm_outer * perm_inner), r % (perm_outer * perm_inner)
    o, i = rem // perm_inner, rem % perm_inner
    return (g * perm_inner + i) * perm_outer + o


def layernorm_fwd_full(x, gamma, beta=None, *, want_bf16=True, want_f32=False, want_raw=False, eps=1e-5,
                       perm_outer=0, perm_inner=0, save_stats=True):
    CALLS.append(("layernorm_fwd", None))
    mean = x.mean(1)
    rstd = torch.rsqrt(x.var(1, unbiased=False) + eps)
    y = (x - mean[:, None]) * rstd[:, None] * gamma
    if beta is not None:
        y = y + beta
    dst = _perm_index(x.shape[0], perm_outer, perm_inner)
    yp = torch.empty_like(y)
    yp[dst] = y
    return (yp.to(OPERAND) if want_bf16 else None, yp if want_f32 else None, x.to(OPERAND) if want_raw else None,
            mean if save_stats else None, rstd if save_stats else None)


def layernorm_bwd_full(dy, x, gamma, mean, rstd, dgamma, dbeta=None, *, dx=None, accum=False, perm_outer=0,
                       perm_inner=0, bcast_rows=0, dy_scale=1.0, dx_bf16=None):
    CALLS.append(("layernorm_bwd", None))
    rows = x.shape[0]
    if bcast_rows:
        dyf = dy.float().repeat_interleave(bcast_rows, dim=0) * dy_scale
    else:
        dyf = dy.float()[_perm_index(rows, perm_outer, perm_inner)] * dy_scale
    xhat = (x - mean[:, None]) * rstd[:, None]
    dgamma.add_((dyf * xhat).sum(0))
    if dbeta is not None:
        dbeta.add_(dyf.sum(0))
    wdy = dyf * gamma
    res = rstd[:, None] * (wdy - wdy.mean(1, keepdim=True) - xhat * (wdy * xhat).mean(1, keepdim=True))
    if dx is None:
        dx = res
    elif accum:
        dx.add_(res)
    else:
        dx.copy_(res)
    if dx_bf16 is not None:
        dx_bf16.copy_(dx.to(dx_bf16.dtype))
    return dx


def _peg_conv(x, w, b, shape):
    B, n0, n1, n2 = shape
    dim = x.shape[-1]
    v = x.reshape(B, n0, n1, n2, dim).permute(0, 4, 1, 2, 3)
    v = F.conv3d(F.pad(v, (1, 1, 1, 1, 2, 0)), w.reshape(dim, 1, 3, 3, 3), b, groups=dim)
    return v.permute(0, 2, 3, 4, 1).reshape(x.shape)


def peg_fwd(x, w, b, shape):
    return _peg_conv(x, w, b, shape) + x


def peg_bwd(dy, x, w, shape, dw, db, dx_bf16=None):
    with torch.enable_grad():
        xx, ww, bb = x.detach().clone().requires_grad_(), w.detach().clone().requires_grad_(), torch.zeros_like(db).requires_grad_()
        y = _peg_conv(xx, ww, bb, shape) + xx
        gx, gw, gb = torch.autograd.grad(y, (xx, ww, bb), dy)
    dw.add_(gw)
    db.add_(gb)
    if dx_bf16 is not None:
        dx_bf16.copy_(gx.to(dx_bf16.dtype))
    return gx


def _cpb_table(w0, b0, w1, b1, w2, b2, gh, gw):
    dy, dx = torch.meshgrid(torch.arange(-(gh - 1), gh), torch.arange(-(gw - 1), gw), indexing="ij")
    rel = torch.stack([dy, dx], dim=-1).reshape(-1, 2).float()
    rel = torch.sign(rel) * torch.log(rel.abs() + 1)
    h0 = F.leaky_relu(rel @ w0.t() + b0, 0.1)
    h1 = F.leaky_relu(h0 @ w1.t() + b1, 0.1)
    return (h1 @ w2.t() + b2).t().reshape(-1, 2 * gh - 1, 2 * gw - 1), h0, h1


def cpb_fwd(w0, b0, w1, b1, w2, b2, gh, gw):
    table, h0, h1 = _cpb_table(w0, b0, w1, b1, w2, b2, gh, gw)
    return table.contiguous(), h0, h1


def cpb_bwd(dtable, w0, w1, w2, h0, h1, gh, gw):
    """gradients from the saved (post-LeakyReLU) activations, as the kernel computes them"""
    heads = w2.shape[0]
    dy, dx = torch.meshgrid(torch.arange(-(gh - 1), gh), torch.arange(-(gw - 1), gw), indexing="ij")
    rel = torch.stack([dy, dx], dim=-1).reshape(-1, 2).float()
    rel = torch.sign(rel) * torch.log(rel.abs() + 1)
    slope = lambda h: torch.where(h > 0, torch.ones_like(h), torch.full_like(h, 0.1))
    dout = dtable.reshape(heads, -1).t()
    dz1 = (dout @ w2) * slope(h1)
    dz0 = (dz1 @ w1) * slope(h0)
    return [dz0.t() @ rel, dz0.sum(0), dz1.t() @ h0, dz1.sum(0), dout.t() @ h1, dout.sum(0)]


def _bias_from_table(table, gh, gw):
    ys, xs = torch.meshgrid(torch.arange(gh), torch.arange(gw), indexing="ij")
    ys, xs = ys.reshape(-1), xs.reshape(-1)
    iy = ys[:, None] - ys[None, :] + gh - 1
    ix = xs[:, None] - xs[None, :] + gw - 1
    return table[:, iy, ix]                                     # [heads, L, L]


def _attn_core(qkv, table, nseq, L, heads, gh, gw):
    inner = heads * 32
    q, k, v = (qkv[:, i * inner:(i + 1) * inner].float().reshape(nseq, L, heads, 32).permute(0, 2, 1, 3) for i in range(3))
    sim = q @ k.transpose(-1, -2)
    if table is not None:
        sim = sim + _bias_from_table(table, gh, gw)
    lse = torch.logsumexp(sim, dim=-1)
    out = (sim.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(nseq * L, inner)
    return out, lse


def attn_fwd(qkv, table, nseq, L, heads, gh=0, gw=0):
    out, lse = _attn_core(qkv, table, nseq, L, heads, gh, gw)
    return out.to(OPERAND), lse


def attn_bwd(qkv, table, out, dout, lse, dtable, nseq, L, heads, gh=0, gw=0):
    with torch.enable_grad():
        qq = qkv.detach().float().clone().requires_grad_()
        tt = table.detach().clone().requires_grad_() if table is not None else None
        o, _ = _attn_core(qq, tt, nseq, L, heads, gh, gw)
        gs = torch.autograd.grad(o, (qq, tt) if tt is not None else (qq,), dout.float())
    if tt is not None:
        dtable.add_(gs[1])
    return gs[0].to(OPERAND)


def qknorm_bwd_(dqkv, qkv, rnorm, q_scale, k_scale, alpha, dq_scale, dk_scale, heads):
    inner = heads * 32
    rows = qkv.shape[0]
    for off, scale, a, dscale, hcol in ((0, q_scale, alpha, dq_scale, 0), (inner, k_scale, 1.0, dk_scale, heads)):
        g = dqkv[:, off: off + inner].float().reshape(rows, heads, 32)
        y = qkv[:, off: off + inner].float().reshape(rows, heads, 32)           # = raw * rn * scale * a
        xhat = y / (scale * a)
        dscale.add_((g * xhat * a).sum((0, 1)))
        gx = g * scale * a
        rn = rnorm[:, hcol: hcol + heads]
        draw = rn[..., None] * (gx - xhat * (gx * xhat).sum(-1, keepdim=True))
        dqkv[:, off: off + inner] = draw.reshape(rows, inner).to(dqkv.dtype)
    return dqkv


def l2norm_rows(x, want_f32=False):
    xn = x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    return xn.to(OPERAND), (xn if want_f32 else None)


def vq_search(x, embed, want_xn_f32=False):
    """double of ops.vq_search: the fp32 arg-max of the cosine similarities (what pass 1 + pass 2 deliver)"""
    CALLS.append(("vq_search",))
    xn = x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    en = embed / embed.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    ind = (xn @ en.t()).argmax(dim=1)
    return ind, embed[ind].float(), (xn if want_xn_f32 else None)


def vq_gather(best, embed):
    ind = best.long()
    return ind, embed[ind].float()


def vq_ema_update_(xn_f32, ind, cluster_size, embed, decay=0.8):
    C = embed.shape[0]
    bins = torch.bincount(ind, minlength=C).float()
    cluster_size.mul_(decay).add_(bins * (1 - decay))
    esum = torch.zeros_like(embed).index_add_(0, ind, xn_f32)
    hit = bins > 0
    mean = esum[hit] / bins[hit][:, None]
    en = embed / embed.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    en[hit] = mean / mean.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    embed.mul_(decay).add_(en * (1 - decay))


def patch_affine_bwd(P, W, gamma, beta, db):
    return gamma * P + beta * db[:, None], (W * P).sum(0), (W * db[:, None]).sum(0)


def mean_pool(x):
    return x.mean(1)


def latent_fwd(x, W):
    y = x @ W.t()
    rn = 1.0 / y.norm(dim=-1).clamp_min(1e-12)
    return y * rn[:, None], rn


def latent_bwd(dlat, lat, rn, x, W, need_dx=True):
    dy = rn[:, None] * (dlat - lat * (dlat * lat).sum(-1, keepdim=True))
    return dy.t() @ x, (dy @ W if need_dx else None)


def clip_loss_fwd_bwd(T, I, log_temp, b_local, row0, need_grad=True):
    with torch.enable_grad():
        t, i, lt = (v.detach().clone().requires_grad_() for v in (T, I, log_temp))
        S = t @ i.t() * lt.exp()
        n = S.shape[0]
        lab = torch.arange(n)
        loss = (F.cross_entropy(S, lab) + F.cross_entropy(S.t(), lab)) / 2 / b_local
        gt, gi, glt = torch.autograd.grad(loss, (t, i, lt))
    sl = slice(row0, row0 + b_local)
    return torch.stack([loss.detach(), glt.reshape(())]), (torch.stack([gt[sl], gi[sl]]) if need_grad else None)


def pair_logits(text_lat, image_lat, log_temp):
    return (text_lat * image_lat).sum(-1) * log_temp.exp()


def fill_(t, v):
    return t.fill_(v)
