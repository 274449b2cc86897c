"""Thin tensor-level wrappers over the C ABI (one Python function per ctk_* entry point).

Every function takes CUDA tensors, validates dtype/contiguity, and launches on the current
torch stream. Nothing here computes on the host and nothing falls back to torch ops.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import (EPI_ARGMAX, EPI_ARGMAX_PART, EPI_ATOMIC_F32, EPI_BF16, EPI_CLIP_GRAD, EPI_F32, EPI_GEGLU, EPI_GEGLU_BWD,  # noqa: F401
                   EPI_GELU, EPI_GELU_BWD, EPI_LSE_PART, EPI_QKV, EPI_RESID_F32, GemmEpilogue, check)


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
GEMM_PROFILE = None     # bench.py sets this to a list: (start_event, end_event, tag) per GEMM launch
# libctk kernels launched by CUDA-graph replays (the C-side counter only sees direct launches); bench.py adds this
# to ctk_launch_count() for its `gpu_launches` line
GRAPH_LAUNCHES = 0


def gemm(a: torch.Tensor, b: torch.Tensor, epilogue: int, c: torch.Tensor, *, M: int, N: int, K: int,
         mn_major: bool = False, lda: Optional[int] = None, ldb: Optional[int] = None,
         ldc: Optional[int] = None, bias=None, resid=None, ldr: Optional[int] = None, aux0=None,
         ld_aux0: int = 0, vec0=None, vec1=None, row_map=None, alpha: float = 1.0, split_k: int = 0, i0: int = 0, i1: int = 0,
         b_mn_major: Optional[bool] = None):
    """D = A[M,K] B[N,K]^T with a fused epilogue (ctk_gemm_bf16). a, b are bf16; K-major:
    a [M, lda>=K], b [N, ldb>=K]; MN-major: a [K, lda>=M], b [K, ldb>=N].  `b_mn_major=True` with a K-major A is the
    input-gradient product dX = dY W with the weight W [K = out, N = in] used as stored."""
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
    prof = GEMM_PROFILE
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib.ctk_gemm_bf16(_p(a), lda, int(mn_major), _p(b), ldb, int(mn_major if b_mn_major is None else b_mn_major),
                            M, N, K, epilogue,
                            C.byref(e), split_k, _stream()), "ctk_gemm_bf16")
    if prof is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        prof.append((e0, e1, f"epi{epilogue}{'_wgrad' if mn_major else ''}"))
    return c


class ZeroArena:
    """One zero-filled fp32 buffer handed out in 256-byte aligned slices: the backward passes accumulate ~250 gradient
    tensors with atomics (split-K weight gradients, LayerNorm / PEG parameter gradients), and zeroing them with one
    fill kernel each cost 0.4 ms of tiny launches per step; this is a single fill.  `take` falls back to torch.zeros
    if the arena was sized too small."""

    def __init__(self, n_elems: int, device):
        self.buf = torch.zeros(int(n_elems), dtype=torch.float32, device=device)
        self.off = 0

    @staticmethod
    def room(*numels) -> int:
        return sum((int(n) + 63) // 64 * 64 for n in numels)

    def take(self, *shape) -> torch.Tensor:
        shape = tuple(shape[0]) if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)) else tuple(shape)
        n = 1
        for d in shape:
            n *= int(d)
        if self.off + n > self.buf.numel():
            return torch.zeros(shape, dtype=torch.float32, device=self.buf.device)
        v = self.buf[self.off: self.off + n].view(shape)
        self.off += (n + 63) // 64 * 64
        return v


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


def transpose_cast_bf16_slice(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """dst[:, :rows] (a bf16 column slice of a wider row-major matrix, `dst.stride(0)` elements per row) = src^T;
    nothing outside the slice is written."""
    _chk(src, torch.float32, "src")
    rows, cols = src.shape
    assert dst.dtype == torch.bfloat16 and dst.shape == (cols, rows) and dst.stride(1) == 1
    check(_lib.load().ctk_transpose_cast_bf16_slice(_p(src), _p(dst), rows, cols, dst.stride(0), _stream()),
          "ctk_transpose_cast_bf16_slice")
    return dst


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


# --------------------------------------------------------------------------- encoder ops
def patch_norm_fwd(video: torch.Tensor, pt: int, p1: int, p2: int, eps: float = 1e-5, out=None):
    """video fp32 [B,1,D,H,W] -> (xhat bf16 [B*T*Hp*Wp, ld], mean, rstd); ld = K rounded up to 8.
    `out` = (xhat, mean, rstd) writes into existing buffers (CUDA-graph replays keep their addresses)."""
    _chk(video, torch.float32, "video")
    B, Cc, D, H, W = video.shape
    assert Cc == 1, "CT volumes are single channel"
    K = pt * p1 * p2
    ld = (K + 7) // 8 * 8
    rows = B * (D // pt) * (H // p1) * (W // p2)
    if out is not None:
        xhat, mean, rstd = out
        assert xhat.shape == (rows, ld) and xhat.dtype == torch.bfloat16
    else:
        xhat = torch.empty(rows, ld, dtype=torch.bfloat16, device=video.device)
        mean = torch.empty(rows, dtype=torch.float32, device=video.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=video.device)
    check(_lib.load().ctk_patch_norm_fwd(_p(video), _p(xhat), ld, _p(mean), _p(rstd), B, D, H, W, pt, p1, p2, eps,
                                         _stream()), "ctk_patch_norm_fwd")
    return xhat, mean, rstd


def layernorm_fwd(x, gamma, beta=None, *, want_bf16=True, want_f32=False, want_raw=False, eps=1e-5,
                  perm_outer=0, perm_inner=0, save_stats=True):
    _chk(x, torch.float32, "x")
    rows, dim = x.shape
    dev = x.device
    ob = torch.empty(rows, dim, dtype=torch.bfloat16, device=dev) if want_bf16 else None
    of = torch.empty(rows, dim, dtype=torch.float32, device=dev) if want_f32 else None
    raw = torch.empty(rows, dim, dtype=torch.bfloat16, device=dev) if want_raw else None
    mean = torch.empty(rows, dtype=torch.float32, device=dev) if save_stats else None
    rstd = torch.empty(rows, dtype=torch.float32, device=dev) if save_stats else None
    check(_lib.load().ctk_layernorm_fwd(_p(x), _p(gamma), _p(beta), _p(ob), _p(of), _p(raw), _p(mean), _p(rstd),
                                        rows, dim, eps, perm_outer, perm_inner, _stream()), "ctk_layernorm_fwd")
    return ob, of, raw, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta=None, *, dx=None, accum=False, perm_outer=0,
                  perm_inner=0, bcast_rows=0, dy_scale=1.0, dx_bf16=None):
    """dy bf16 or fp32; returns dx (fp32). dgamma/dbeta accumulate (must be zero-initialised)."""
    rows, dim = x.shape
    if dx is None:
        assert not accum
        dx = torch.empty_like(x)
    dyb = dy if dy.dtype == torch.bfloat16 else None
    dyf = dy if dy.dtype == torch.float32 else None
    assert dy.is_contiguous() and (dyb is not None or dyf is not None)
    check(_lib.load().ctk_layernorm_bwd(_p(dyb), _p(dyf), _p(x), _p(gamma), _p(mean), _p(rstd), _p(dx), _p(dx_bf16), int(accum),
                                        _p(dgamma), _p(dbeta), rows, dim, perm_outer, perm_inner, bcast_rows,
                                        dy_scale, _stream()), "ctk_layernorm_bwd")
    return dx


def peg_fwd(x, w, b, shape):
    """x fp32 [B*n0*n1*n2, dim] viewed as the grid `shape`=(B,n0,n1,n2); returns conv(x)+b+x."""
    _chk(x, torch.float32, "x")
    B, n0, n1, n2 = shape
    dim = x.shape[-1]
    y = torch.empty_like(x)
    check(_lib.load().ctk_peg_fwd(_p(x), _p(w), _p(b), _p(y), B, n0, n1, n2, dim, _stream()), "ctk_peg_fwd")
    return y


def peg_bwd(dy, x, w, shape, dw, db, dx_bf16=None):
    B, n0, n1, n2 = shape
    dim = x.shape[-1]
    dx = torch.empty_like(dy)
    check(_lib.load().ctk_peg_bwd(_p(dy), _p(x), _p(w), _p(dx), _p(dx_bf16), _p(dw), _p(db), B, n0, n1, n2, dim,
                                  _stream()), "ctk_peg_bwd")
    return dx


def cpb_fwd(w0, b0, w1, b1, w2, b2, gh: int, gw: int):
    dim, heads = w1.shape[0], w2.shape[0]
    n_off = (2 * gh - 1) * (2 * gw - 1)
    dev = w0.device
    h0 = torch.empty(n_off, dim, dtype=torch.float32, device=dev)
    h1 = torch.empty(n_off, dim, dtype=torch.float32, device=dev)
    table = torch.empty(heads, 2 * gh - 1, 2 * gw - 1, dtype=torch.float32, device=dev)
    check(_lib.load().ctk_cpb_fwd(_p(w0), _p(b0), _p(w1), _p(b1), _p(w2), _p(b2), _p(h0), _p(h1), _p(table), gh, gw,
                                  dim, heads, _stream()), "ctk_cpb_fwd")
    return table, h0, h1


def cpb_bwd(dtable, w0, w1, w2, h0, h1, gh: int, gw: int):
    dim, heads = w1.shape[0], w2.shape[0]
    dev = w0.device
    n_off = (2 * gh - 1) * (2 * gw - 1)
    g = [torch.empty_like(w0), torch.empty(dim, device=dev), torch.empty_like(w1), torch.empty(dim, device=dev),
         torch.empty_like(w2), torch.empty(heads, device=dev)]
    ws = torch.empty(2 * n_off * dim, dtype=torch.float32, device=dev)
    check(_lib.load().ctk_cpb_bwd(_p(dtable), _p(w0), _p(w1), _p(w2), _p(h0), _p(h1), *[_p(t) for t in g], _p(ws), gh,
                                  gw, dim, heads, _stream()), "ctk_cpb_bwd")
    return g


def attn_fwd(qkv, table, nseq: int, L: int, heads: int, gh: int = 0, gw: int = 0):
    _chk(qkv, torch.bfloat16, "qkv")
    dev = qkv.device
    out = torch.empty(nseq * L, heads * 32, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(nseq, heads, L, dtype=torch.float32, device=dev)
    check(_lib.load().ctk_attn_fwd(_p(qkv), _p(table), _p(out), _p(lse), nseq, L, heads, gh, gw, _stream()), "ctk_attn_fwd")
    return out, lse


def attn_bwd(qkv, table, out, dout, lse, dtable, nseq: int, L: int, heads: int, gh: int = 0, gw: int = 0):
    _chk(dout, torch.bfloat16, "dout")
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    check(_lib.load().ctk_attn_bwd(_p(qkv), _p(table), _p(out), _p(dout), _p(lse), _p(delta), _p(dqkv), _p(dtable), nseq,
                                   L, heads, gh, gw, _stream()), "ctk_attn_bwd")
    return dqkv


def qknorm_bwd_(dqkv, qkv, rnorm, q_scale, k_scale, alpha: float, dq_scale, dk_scale, heads: int):
    rows = qkv.shape[0]
    check(_lib.load().ctk_qknorm_bwd(_p(dqkv), _p(qkv), _p(rnorm), _p(q_scale), _p(k_scale), alpha, _p(dq_scale),
                                     _p(dk_scale), rows, heads, _stream()), "ctk_qknorm_bwd")
    return dqkv


def l2norm_rows(x, want_f32: bool = False):
    _chk(x, torch.float32, "x")
    rows, dim = x.shape
    xb = torch.empty(rows, dim, dtype=torch.bfloat16, device=x.device)
    xf = torch.empty(rows, dim, dtype=torch.float32, device=x.device) if want_f32 else None
    check(_lib.load().ctk_l2norm_rows(_p(x), _p(xb), _p(xf), rows, dim, _stream()), "ctk_l2norm_rows")
    return xb, xf


def vq_gather(best, embed):
    rows = best.shape[0]
    C, dim = embed.shape
    ind = torch.empty(rows, dtype=torch.int64, device=embed.device)
    quant = torch.empty(rows, dim, dtype=torch.float32, device=embed.device)
    check(_lib.load().ctk_vq_gather(_p(best), _p(embed), _p(ind), _p(quant), rows, dim, C, _stream()), "ctk_vq_gather")
    return ind, quant


# two bf16-rounded unit vectors: |sim_bf16 - sim_fp32| <= 2 * 2^-9 + O(2^-18), so two candidates can swap order only if
# their bf16 similarities are closer than 2^-7 (plus fp32 accumulation noise)
VQ_RESCORE_MARGIN = 2.0 ** -7 + 1e-5


def vq_search(x: torch.Tensor, embed: torch.Tensor, want_xn_f32: bool = False):
    """Cosine-similarity code search of VectorQuantize(use_cosine_sim=True) (ctvit.py:188,403) with the fp32 arg-max:
    pass 1 = tcgen05 GEMM of the bf16-rounded unit vectors keeping, per row and 128-code block, the best code and the
    second-best value (the 13824 x 8192 similarities are never stored); pass 2 (ctk_vq_select) re-scores in fp32 every
    code inside the bf16 rounding bound of the row maximum.  x fp32 [rows, dim] (raw), embed fp32 [C, dim] (raw).
    Returns (ind int64 [rows], quant fp32 [rows, dim] = embed[ind], xn fp32 or None)."""
    _chk(x, torch.float32, "x")
    _chk(embed, torch.float32, "embed")
    rows, dim = x.shape
    C = embed.shape[0]
    dev = x.device
    xb, xf = l2norm_rows(x, want_f32=want_xn_f32)
    eb, en = l2norm_rows(embed, want_f32=True)
    nblk = (C + 127) // 128
    key = torch.empty(nblk, rows, dtype=torch.int64, device=dev)
    sec = torch.empty(nblk, rows, dtype=torch.float32, device=dev)
    gemm(xb, eb, EPI_ARGMAX_PART, key, M=rows, N=C, K=dim, ldc=0, aux0=sec)
    ind = torch.empty(rows, dtype=torch.int64, device=dev)
    quant = torch.empty(rows, dim, dtype=torch.float32, device=dev)
    check(_lib.load().ctk_vq_select(_p(key), _p(sec), nblk, _p(x), _p(en), _p(embed), _p(ind), _p(quant), rows, dim, C,
                                    VQ_RESCORE_MARGIN, _stream()), "ctk_vq_select")
    return ind, quant, xf


def vq_ema_update_(xn_f32, ind, cluster_size, embed, decay: float = 0.8):
    C, dim = embed.shape
    ws = torch.empty(C * dim + C, dtype=torch.float32, device=embed.device)
    check(_lib.load().ctk_vq_ema_update(_p(xn_f32), _p(ind), _p(cluster_size), _p(embed), _p(ws), xn_f32.shape[0], dim, C,
                                        decay, _stream()), "ctk_vq_ema_update")


def patch_affine_bwd(P, W, gamma, beta, db):
    n, k = W.shape
    dW = torch.empty_like(W)
    dgamma = torch.empty_like(gamma)
    dbeta = torch.empty_like(beta)
    check(_lib.load().ctk_patch_affine_bwd(_p(P), _p(W), _p(gamma), _p(beta), _p(db), _p(dW), _p(dgamma), _p(dbeta), n, k,
                                           _stream()), "ctk_patch_affine_bwd")
    return dW, dgamma, dbeta


def colsum_(dy, out):
    """out[cols] += column sums of dy [rows, cols] (bf16 or fp32)."""
    rows, cols = dy.shape
    dyb = dy if dy.dtype == torch.bfloat16 else None
    dyf = dy if dy.dtype == torch.float32 else None
    check(_lib.load().ctk_colsum(_p(dyb), _p(dyf), _p(out), rows, cols, _stream()), "ctk_colsum")
    return out


# --------------------------------------------------------------------------- loader-side volume preparation
def volume_prep(src: torch.Tensor, out: torch.Tensor):
    """scripts/data.py:49-111 on the device: src = stored array (D, H, W), fp32 or fp16, CUDA or pinned host memory;
    out fp32 CUDA (..., Dt, Ht, Wt) contiguous (the last three dims are the target volume)."""
    assert src.dim() == 3 and src.is_contiguous() and src.dtype in (torch.float32, torch.float16)
    assert src.is_cuda or src.is_pinned(), "volume_prep reads device or pinned host memory"
    _chk(out, torch.float32, "out")
    D, H, W = src.shape
    Dt, Ht, Wt = out.shape[-3:]
    assert out.numel() == Dt * Ht * Wt
    check(_lib.load().ctk_volume_prep(src.data_ptr(), int(src.dtype == torch.float16), D, H, W, _p(out), Dt, Ht, Wt,
                                      _stream()), "ctk_volume_prep")
    return out


# --------------------------------------------------------------------------- text tower (HF BertModel, ct_clip.py:1271)
def bert_embed_fwd(ids, token_type, word, pos, typ):
    """fp32 [B*L, H] = word[ids] + typ[token_type or 0] + pos[0..L-1] (BertEmbeddings before its LayerNorm)."""
    assert ids.dtype == torch.int64 and ids.is_contiguous() and ids.dim() == 2
    _chk(word, torch.float32, "word"); _chk(pos, torch.float32, "pos"); _chk(typ, torch.float32, "typ")
    B, L = ids.shape
    H = word.shape[1]
    out = torch.empty(B * L, H, dtype=torch.float32, device=ids.device)
    tt = None if token_type is None else token_type.contiguous()
    check(_lib.load().ctk_bert_embed_fwd(_p(ids), _p(tt), _p(word), _p(pos), _p(typ), _p(out), B * L, L, H, _stream()),
          "ctk_bert_embed_fwd")
    return out


def bert_embed_bwd(de, ids, token_type, word_shape, pos_shape, typ_shape, pad_idx):
    """gradients of the three embedding tables from de fp32 [B*L, H]"""
    _chk(de, torch.float32, "de")
    B, L = ids.shape
    H = de.shape[1]
    dev = de.device
    dword = torch.zeros(word_shape, dtype=torch.float32, device=dev)
    dpos = torch.zeros(pos_shape, dtype=torch.float32, device=dev)
    dtyp = torch.zeros(typ_shape, dtype=torch.float32, device=dev)
    tt = None if token_type is None else token_type.contiguous()
    check(_lib.load().ctk_bert_embed_bwd(_p(de), _p(ids), _p(tt), _p(dword), _p(dpos), _p(dtyp), B * L, L, H,
                                         -1 if pad_idx is None else int(pad_idx), _stream()), "ctk_bert_embed_bwd")
    if tt is None:
        colsum_(de, dtyp[0])
    return dword, dpos, dtyp


def mha_fwd(qkv, key_mask, B: int, L: int, heads: int, scale: float, p_drop: float = 0.0, seed=None, seed_off: int = 0,
            null_k=None, null_v=None):
    """Softmax attention core (head dim 64: BertSelfAttention; 32: CTViT3D's FlashAttention) on the packed bf16 projections
    qkv [B*L, 3H]; key_mask uint8 [B, L] or None; null_k / null_v bf16 [heads, n_null, dh]: learned null pairs every query
    also attends to; seed = int64 device tensor [1] (read by the kernel: graph replays see its current value).
    Returns (ctx bf16 [B*L, H], lse fp32 [B, heads, L])."""
    _chk(qkv, torch.bfloat16, "qkv")
    H = qkv.shape[1] // 3
    assert qkv.shape == (B * L, 3 * H) and H % heads == 0
    assert key_mask is None or (key_mask.dtype == torch.uint8 and key_mask.is_contiguous() and key_mask.shape == (B, L))
    assert p_drop == 0.0 or (seed is not None and seed.dtype == torch.int64 and seed.is_cuda)
    n_null = 0
    if null_k is not None:
        _chk(null_k, torch.bfloat16, "null_k"); _chk(null_v, torch.bfloat16, "null_v")
        n_null = null_k.shape[1]
        assert null_k.shape == (heads, n_null, H // heads) and null_v.shape == null_k.shape
    out = torch.empty(B * L, H, dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty(B, heads, L, dtype=torch.float32, device=qkv.device)
    check(_lib.load().ctk_mha_fwd(_p(qkv), _p(key_mask), _p(null_k), _p(null_v), n_null, _p(out), _p(lse), B, L, heads,
                                  H // heads, scale, p_drop, _p(seed), seed_off, _stream()), "ctk_mha_fwd")
    return out, lse


def mha_bwd(qkv, key_mask, out, dout, lse, B: int, L: int, heads: int, scale: float, p_drop: float = 0.0, seed=None,
            seed_off: int = 0, null_k=None, null_v=None):
    """dqkv bf16 [B*L, 3H] = (dq | dk | dv) of mha_fwd; with null pairs returns (dqkv, dnull_k, dnull_v) with the null
    gradients fp32 [B, heads, n_null, dh] (one slab per sequence: sum over dim 0)."""
    _chk(dout, torch.bfloat16, "dout")
    H = qkv.shape[1] // 3
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    n_null, dnk, dnv = 0, None, None
    if null_k is not None:
        n_null = null_k.shape[1]
        dnk = torch.empty(B, heads, n_null, H // heads, dtype=torch.float32, device=qkv.device)
        dnv = torch.empty_like(dnk)
    check(_lib.load().ctk_mha_bwd(_p(qkv), _p(key_mask), _p(null_k), _p(null_v), n_null, _p(out), _p(dout), _p(lse), _p(delta),
                                  _p(dqkv), _p(dnk), _p(dnv), B, L, heads, H // heads, scale, p_drop, _p(seed), seed_off,
                                  _stream()), "ctk_mha_bwd")
    return dqkv if null_k is None else (dqkv, dnk, dnv)
