"""Drop-in `CTViT` (reference: transformer_maskgit/transformer_maskgit/ctvit.py + attention.py).

Same constructor keywords, attributes, forward signature and state-dict keys as the reference
module, so a reference checkpoint loads unchanged and callers (`CTCLIP`, the trainer, zero-shot
inference) keep working.  Only the encoder branch (`return_encoded_tokens=True`, ctvit.py:353-412)
is implemented; the VQ-GAN decoder / discriminator branches raise NotImplementedError.

The nn.Module tree below only *holds parameters* under the reference's names.  All arithmetic runs
in libctk.so (sm_100a CUDA) through one autograd.Function that owns the forward and the hand
written backward of the whole encoder.  There is no CPU or torch fallback.
"""
from __future__ import annotations

import os
import weakref
from pathlib import Path
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import ops

ATTN_SCALE = 8.0          # attention.py:105 (scale = 8)


def pair(val):
    ret = (val, val) if not isinstance(val, tuple) else val
    assert len(ret) == 2
    return ret


# ------------------------------------------------------------------------------------------------
# parameter containers (names/shapes/init order identical to the reference modules)
# ------------------------------------------------------------------------------------------------
class LayerNorm(nn.Module):
    """attention.py:34-41: learnable gamma, zero `beta` buffer."""

    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))
        self.register_buffer("beta", torch.zeros(dim))


class PEG(nn.Module):
    """attention.py:62-90: parameter holder for the depthwise 3x3x3 conv."""

    def __init__(self, dim, causal=False):
        super().__init__()
        self.causal = causal
        self.dsconv = nn.Conv3d(dim, dim, 3, groups=dim)


class Attention(nn.Module):
    """attention.py:94-131 (self-attention, no null kv on this path: attention.py:422)."""

    def __init__(self, dim, dim_head=64, heads=8, num_null_kv=0, scale=8):
        super().__init__()
        self.heads, self.scale, self.dim_head = heads, scale, dim_head
        inner_dim = dim_head * heads
        self.norm = LayerNorm(dim)
        self.context_norm = LayerNorm(dim)
        self.num_null_kv = num_null_kv
        self.null_kv = nn.Parameter(torch.randn(heads, 2 * num_null_kv, dim_head))
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_kv = nn.Linear(dim, inner_dim * 2, bias=False)
        self.q_scale = nn.Parameter(torch.ones(dim_head))
        self.k_scale = nn.Parameter(torch.ones(dim_head))
        self.to_out = nn.Linear(inner_dim, dim, bias=False)


class _GEGLUSlot(nn.Module):
    """placeholder so FeedForward keeps the reference's Sequential indices (0,1,2,3,4)."""


def FeedForward(dim, mult=4, dropout=0.0):
    """attention.py:50-58."""
    inner_dim = int(mult * (2 / 3) * dim)
    return nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, inner_dim * 2, bias=False), _GEGLUSlot(),
                         nn.Dropout(dropout), nn.Linear(inner_dim, dim, bias=False))


class ContinuousPositionBias(nn.Module):
    """attention.py:335-361 (num_dims 2, layers 2)."""

    def __init__(self, *, dim, heads, num_dims=2, layers=2):
        super().__init__()
        self.net = nn.ModuleList([])
        self.net.append(nn.Sequential(nn.Linear(num_dims, dim), nn.LeakyReLU(0.1)))
        for _ in range(layers - 1):
            self.net.append(nn.Sequential(nn.Linear(dim, dim), nn.LeakyReLU(0.1)))
        self.net.append(nn.Linear(dim, heads))


class Transformer(nn.Module):
    """attention.py:386-439: layers[i] = [PEG, Attention, None (no cross-attn), FeedForward]."""

    def __init__(self, dim, *, depth, dim_head=64, heads=8, ff_mult=4, peg=False, peg_causal=False,
                 attn_dropout=0.0, ff_dropout=0.0, attn_num_null_kv=0):
        super().__init__()
        assert attn_dropout == 0.0 and ff_dropout == 0.0, "dropout is 0 on the reference path (run_train.py:56-66)"
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                PEG(dim=dim, causal=peg_causal) if peg else None,
                # attention.py:422 builds `Attention` without null kv; the FlashAttention of CTViT3D (attention.py:413)
                # gets attn_num_null_kv = 2 learned null key/value pairs
                Attention(dim=dim, dim_head=dim_head, heads=heads, num_null_kv=attn_num_null_kv),
                None,
                FeedForward(dim=dim, mult=ff_mult, dropout=ff_dropout),
            ]))
        self.norm_out = LayerNorm(dim)


class _CosineSimCodebook(nn.Module):
    def __init__(self, dim, codebook_size):
        super().__init__()
        embed = F.normalize(nn.init.kaiming_uniform_(torch.empty(1, codebook_size, dim)), dim=-1)
        self.register_buffer("initted", torch.Tensor([True]))
        self.register_buffer("cluster_size", torch.zeros(1, codebook_size))
        self.register_buffer("embed", embed)


class VectorQuantize(nn.Module):
    """State holder for VectorQuantize(dim, codebook_size, use_cosine_sim=True) (ctvit.py:188):
    buffers `_codebook.{initted, cluster_size, embed}` as in vector-quantize-pytorch 1.1.2.
    The search / gather / EMA kernels live in libctk (parity unpinned: see oracle header)."""

    def __init__(self, dim, codebook_size, use_cosine_sim=True, decay=0.8):
        super().__init__()
        assert use_cosine_sim
        self.decay = decay
        self.codebook_size = codebook_size
        self._codebook = _CosineSimCodebook(dim, codebook_size)

    @property
    def codebook(self):
        return self._codebook.embed[0]


# ------------------------------------------------------------------------------------------------
# flat parameter view used by the autograd.Function
# ------------------------------------------------------------------------------------------------
def _layer_params(tr: Transformer) -> List[torch.Tensor]:
    out = []
    for peg, attn, _, ff in tr.layers:
        out += [peg.dsconv.weight, peg.dsconv.bias, attn.norm.gamma, attn.to_q.weight, attn.to_kv.weight,
                attn.q_scale, attn.k_scale, attn.to_out.weight, ff[0].weight, ff[0].bias, ff[1].weight, ff[4].weight]
    out.append(tr.norm_out.gamma)
    return out


N_PER_LAYER = 12


class _Cfg:
    """static shape bundle"""

    def __init__(self, vit: "CTViT", video: torch.Tensor):
        B, C, D, H, W = video.shape
        self.B, self.D, self.H, self.W = B, D, H, W
        self.p1, self.p2 = vit.patch_size
        self.pt = vit.temporal_patch_size
        self.t, self.h, self.w = D // self.pt, H // self.p1, W // self.p2
        self.dim = vit.dim
        self.heads = vit.heads
        self.inner = vit.heads * vit.dim_head
        self.M = B * self.t * self.h * self.w
        self.ff_inner = vit.ff_inner
        self.ff_pad = (self.ff_inner + 127) // 128 * 128
        self.K = self.pt * self.p1 * self.p2
        self.Kp = (self.K + 7) // 8 * 8
        self.sd, self.td = vit.spatial_depth, vit.temporal_depth


def _prep_layer_weights(lp: List[torch.Tensor], cfg: _Cfg, need_bwd: bool) -> Dict[str, torch.Tensor]:
    (pw, pb, gamma, wq, wkv, qs, ks, wo, fg, fb, w1, w2) = lp
    d: Dict[str, torch.Tensor] = {}
    d["peg_w"] = pw.reshape(cfg.dim, 27)
    d["wq"] = ops.cast_bf16(wq)
    d["wkv"] = ops.cast_bf16(wkv)
    d["wo"] = ops.cast_bf16(wo)
    # the input-gradient products read these same [out, in] operands MN-major (ops.gemm(b_mn_major=True)): no
    # transposed weight copies
    d["w1p"], _, d["w1_map"] = ops.pack_ff_w1(w1, cfg.ff_inner, cfg.ff_pad, want_t=False)
    d["w2"] = ops.cast_bf16(w2, ld=cfg.ff_pad)                # pad columns (hidden units >= ff_inner) are zero
    return d


def _transformer_fwd(x, lps, norm_gamma, cfg: _Cfg, table, nseq, L, gh, gw, perm, save: bool, need_bwd: bool):
    """x fp32 [M, dim] -> norm_out(x) written through the row permutation `perm`=(outer, inner)."""
    dim, heads, inner, M = cfg.dim, cfg.heads, cfg.inner, cfg.M
    dev = x.device
    shape = (cfg.B, cfg.t, cfg.h, cfg.w)           # PEG always reshapes with video_shape (ctvit.py:289,303)
    saved = []
    for lp in lps:
        w = _prep_layer_weights(lp, cfg, need_bwd)
        (pw, pb, gamma, wq, wkv, qs, ks, wo, fg, fb, w1, w2) = lp
        x1 = ops.peg_fwd(x, w["peg_w"], pb, shape)
        xn, _, xraw, mu1, rs1 = ops.layernorm_fwd(x1, gamma, None, want_raw=True)
        qkv = torch.empty(M, 3 * inner, dtype=torch.bfloat16, device=dev)
        rn = torch.empty(M, 2 * heads, dtype=torch.float32, device=dev)
        ops.gemm(xn, w["wq"], ops.EPI_QKV, qkv, M=M, N=inner, K=dim, aux0=rn, ld_aux0=2 * heads, vec0=qs,
                 alpha=ATTN_SCALE, i0=inner, i1=0)
        ops.gemm(xraw, w["wkv"], ops.EPI_QKV, qkv, M=M, N=2 * inner, K=dim, aux0=rn, ld_aux0=2 * heads, vec0=ks,
                 alpha=1.0, i0=inner, i1=inner)
        o, lse = ops.attn_fwd(qkv, table, nseq, L, heads, gh, gw)
        x2 = torch.empty_like(x1)
        ops.gemm(o, w["wo"], ops.EPI_RESID_F32, x2, M=M, N=dim, K=inner, resid=x1)
        hn, _, _, mu2, rs2 = ops.layernorm_fwd(x2, fg, fb)
        U = torch.empty(M, 2 * cfg.ff_pad, dtype=torch.bfloat16, device=dev)
        Hh = torch.empty(M, cfg.ff_pad, dtype=torch.bfloat16, device=dev)
        ops.gemm(hn, w["w1p"], ops.EPI_GEGLU, U, M=M, N=2 * cfg.ff_pad, K=dim, aux0=Hh, ld_aux0=cfg.ff_pad)
        x3 = torch.empty_like(x2)
        ops.gemm(Hh, w["w2"], ops.EPI_RESID_F32, x3, M=M, N=dim, K=cfg.ff_pad, resid=x2)
        if save:
            saved.append(dict(w=w, x=x, x1=x1, xn=xn, xraw=xraw, mu1=mu1, rs1=rs1, qkv=qkv, rn=rn, o=o, lse=lse,
                              x2=x2, hn=hn, mu2=mu2, rs2=rs2, U=U, H=Hh))
        x = x3
    _, y, _, mu, rs = ops.layernorm_fwd(x, norm_gamma, None, want_bf16=False, want_f32=True,
                                        perm_outer=perm[0], perm_inner=perm[1])
    fin = dict(x=x, mu=mu, rs=rs) if save else None
    return y, saved, fin


def _transformer_bwd(dy, dy_bcast, lps, norm_gamma, cfg: _Cfg, table, dtable, nseq, L, gh, gw, perm, saved, fin,
                     grads_out: List[Optional[torch.Tensor]], want_bf16_out: bool, arena: "ops.ZeroArena"):
    """dy: gradient wrt the (permuted) norm_out output. Returns (dx fp32, dx bf16 or None); fills
    grads_out (same order as _layer_params)."""
    dim, heads, inner, M = cfg.dim, cfg.heads, cfg.inner, cfg.M
    dev = norm_gamma.device
    shape = (cfg.B, cfg.t, cfg.h, cfg.w)
    f32 = dict(dtype=torch.float32, device=dev)
    bf = dict(dtype=torch.bfloat16, device=dev)
    zeros = arena.take                     # zero-initialised accumulation targets: slices of one pre-filled buffer
    dgo = zeros(dim)
    g_bf = torch.empty(M, dim, **bf)
    if dy_bcast is not None:
        rows_per, scale = dy_bcast
        g = ops.layernorm_bwd(dy, fin["x"], norm_gamma, fin["mu"], fin["rs"], dgo, None, bcast_rows=rows_per,
                              dy_scale=scale, dx_bf16=g_bf)
    else:
        g = ops.layernorm_bwd(dy, fin["x"], norm_gamma, fin["mu"], fin["rs"], dgo, None, perm_outer=perm[0],
                              perm_inner=perm[1], dx_bf16=g_bf)
    grads_out[len(lps) * N_PER_LAYER] = dgo
    for li in range(len(lps) - 1, -1, -1):
        (pw, pb, gamma, wq, wkv, qs, ks, wo, fg, fb, w1, w2) = lps[li]
        s = saved[li]
        w = s["w"]
        # ---- feed-forward: x3 = x2 + W2 geglu(W1 LN(x2))
        dw2 = zeros(w2.shape)
        ops.gemm(g_bf, s["H"], ops.EPI_ATOMIC_F32, dw2, M=dim, N=cfg.ff_inner, K=M, mn_major=True, ldc=cfg.ff_inner)
        dU = torch.empty_like(s["U"])
        ops.gemm(g_bf, w["w2"], ops.EPI_GEGLU_BWD, dU, M=M, N=cfg.ff_pad, K=dim, aux0=s["U"], ld_aux0=2 * cfg.ff_pad,
                 b_mn_major=True)
        dw1 = zeros(w1.shape)
        ops.gemm(dU, s["hn"], ops.EPI_ATOMIC_F32, dw1, M=2 * cfg.ff_pad, N=dim, K=M, mn_major=True, ldc=dim,
                 row_map=w["w1_map"])
        dhn = torch.empty(M, dim, **bf)
        ops.gemm(dU, w["w1p"], ops.EPI_BF16, dhn, M=M, N=dim, K=2 * cfg.ff_pad, b_mn_major=True)
        dfg, dfb = zeros(dim), zeros(dim)
        ops.layernorm_bwd(dhn, s["x2"], fg, s["mu2"], s["rs2"], dfg, dfb, dx=g, accum=True, dx_bf16=g_bf)
        # ---- attention: x2 = x1 + Wo attn(q(LN(x1)), kv(x1))
        dwo = zeros(wo.shape)
        ops.gemm(g_bf, s["o"], ops.EPI_ATOMIC_F32, dwo, M=dim, N=inner, K=M, mn_major=True, ldc=inner)
        do = torch.empty(M, inner, **bf)
        ops.gemm(g_bf, w["wo"], ops.EPI_BF16, do, M=M, N=inner, K=dim, b_mn_major=True)
        dqkv = ops.attn_bwd(s["qkv"], table, s["o"], do, s["lse"], dtable, nseq, L, heads, gh, gw)
        dqs, dks = zeros(32), zeros(32)
        ops.qknorm_bwd_(dqkv, s["qkv"], s["rn"], qs, ks, ATTN_SCALE, dqs, dks, heads)
        dwq = zeros(wq.shape)
        ops.gemm(dqkv, s["xn"], ops.EPI_ATOMIC_F32, dwq, M=inner, N=dim, K=M, mn_major=True, lda=3 * inner, ldc=dim)
        dwkv = zeros(wkv.shape)
        dkv_view = dqkv[:, inner:]
        ops.gemm(dkv_view, s["xraw"], ops.EPI_ATOMIC_F32, dwkv, M=2 * inner, N=dim, K=M, mn_major=True,
                 lda=3 * inner, ldc=dim)
        # k/v read the raw stream: their input gradient joins the residual gradient directly
        ops.gemm(dkv_view, w["wkv"], ops.EPI_RESID_F32, g, M=M, N=dim, K=2 * inner, lda=3 * inner, resid=g,
                 b_mn_major=True)
        dxn = torch.empty(M, dim, **bf)
        ops.gemm(dqkv, w["wq"], ops.EPI_BF16, dxn, M=M, N=dim, K=inner, lda=3 * inner, b_mn_major=True)
        dgamma = zeros(dim)
        ops.layernorm_bwd(dxn, s["x1"], gamma, s["mu1"], s["rs1"], dgamma, None, dx=g, accum=True)
        # ---- PEG: x1 = conv(x) + b + x
        dpw, dpb = zeros(dim, 27), zeros(dim)
        need_bf = li > 0 or want_bf16_out
        g_new = ops.peg_bwd(g, s["x"], w["peg_w"], shape, dpw, dpb, dx_bf16=g_bf if need_bf else None)
        g = g_new
        base = li * N_PER_LAYER
        grads_out[base:base + N_PER_LAYER] = [dpw.reshape(pw.shape), dpb, dgamma, dwq, dwkv, dqs, dks, dwo, dfg, dfb,
                                              dw1, dw2]
        saved[li] = None        # free activations as we go
    return g, (g_bf if want_bf16_out else None)


def _encode_forward(vit: "CTViT", video: torch.Tensor, params: List[torch.Tensor], save: bool, training: bool,
                    xhat: Optional[torch.Tensor] = None):
    """`xhat` (normalised patches from ops.patch_norm_fwd) may be supplied by the caller: the CUDA-graph path runs
    that one kernel outside the graph because it is the only one that reads the (per-step) volume pointer."""
    cfg = _Cfg(vit, video)
    dim, M = cfg.dim, cfg.M
    it = iter(params)
    cpb = [next(it) for _ in range(6)]
    g1, b1, wp, bp, g3, b3 = (next(it) for _ in range(6))
    sp = [[next(it) for _ in range(N_PER_LAYER)] for _ in range(cfg.sd)]
    sp_norm = next(it)
    tp = [[next(it) for _ in range(N_PER_LAYER)] for _ in range(cfg.td)]
    tp_norm = next(it)

    # ---- patch embedding (ctvit.py:170-175); LayerNorm(K) affine folded into the projection
    if xhat is None:
        xhat, _, _ = ops.patch_norm_fwd(video, cfg.pt, cfg.p1, cfg.p2)
    wp_eff = ops.cast_bf16(wp, ld=cfg.Kp, col_scale=g1)
    bias_eff = torch.addmv(bp, wp, b1)                    # parameter folding: b + W beta (dim x K mat-vec)
    y0 = torch.empty(M, dim, dtype=torch.float32, device=video.device)
    ops.gemm(xhat, wp_eff, ops.EPI_F32, y0, M=M, N=dim, K=cfg.Kp, bias=bias_eff)
    _, x0, _, mu0, rs0 = ops.layernorm_fwd(y0, g3, b3, want_bf16=False, want_f32=True)
    # ---- continuous position bias table (attention.py:363-382, deduplicated offsets)
    table, h0, h1 = ops.cpb_fwd(*cpb, cfg.h, cfg.w)
    # ---- spatial then temporal stacks (ctvit.py:291-305)
    hw = cfg.h * cfg.w
    xs, sp_saved, sp_fin = _transformer_fwd(x0, sp, sp_norm, cfg, table, cfg.B * cfg.t, hw, cfg.h, cfg.w,
                                            (cfg.t, hw), save, save)
    xt, tp_saved, tp_fin = _transformer_fwd(xs, tp, tp_norm, cfg, None, cfg.B * hw, cfg.t, 0, 0, (hw, cfg.t), save, save)
    # ---- vector quantisation (ctvit.py:403): cosine-sim code search on the tensor cores, fp32-exact selection
    embed = vit.vq._codebook.embed[0]
    ind, quant, xf = ops.vq_search(xt, embed.contiguous(), want_xn_f32=training)
    if training:
        ops.vq_ema_update_(xf, ind, vit.vq._codebook.cluster_size[0], embed, vit.vq.decay)
    out = quant.view(cfg.B, cfg.t, cfg.h, cfg.w, dim)
    ctx = None
    if save:
        ctx = dict(cfg=cfg, xhat=xhat, y0=y0, mu0=mu0, rs0=rs0, table=table, h0=h0, h1=h1, sp_saved=sp_saved,
                   sp_fin=sp_fin, tp_saved=tp_saved, tp_fin=tp_fin)
    return out, ind.view(cfg.B, cfg.t, cfg.h, cfg.w), xt, ctx


def _encode_backward(vit: "CTViT", params: List[torch.Tensor], ctx, dtokens: torch.Tensor, dy_in=None):
    """`dy_in` = (dy fp32 [rows, dim], bcast) replaces the unpacking of `dtokens` (CUDA-graph path: static buffer)."""
    cfg: _Cfg = ctx["cfg"]
    dim, M = cfg.dim, cfg.M
    dev = dtokens.device if dy_in is None else dy_in[0].device
    it = iter(params)
    cpb = [next(it) for _ in range(6)]
    g1, b1, wp, bp, g3, b3 = (next(it) for _ in range(6))
    sp = [[next(it) for _ in range(N_PER_LAYER)] for _ in range(cfg.sd)]
    sp_norm = next(it)
    tp = [[next(it) for _ in range(N_PER_LAYER)] for _ in range(cfg.td)]
    tp_norm = next(it)
    hw = cfg.h * cfg.w
    n_tok = cfg.t * hw

    # straight-through VQ (quantize = x + (quantize - x).detach()): d enc = d tokens.
    # A mean-pool gradient arrives as a stride-0 expand of [B, dim]: keep it un-materialised.
    if dy_in is not None:
        dy, bcast = dy_in
    else:
        dy, bcast = _unpack_dtokens(dtokens, cfg)

    sp_g: List[Optional[torch.Tensor]] = [None] * (cfg.sd * N_PER_LAYER + 1)
    tp_g: List[Optional[torch.Tensor]] = [None] * (cfg.td * N_PER_LAYER + 1)
    # every zero-initialised accumulation target of this backward pass comes out of one pre-filled buffer
    arena = ops.ZeroArena(ops.ZeroArena.room(*(p.numel() for p in params), dim * cfg.K, ctx["table"].numel(),
                                             *([dim] * 8)), dev)
    g, _ = _transformer_bwd(dy, bcast, tp, tp_norm, cfg, None, None, cfg.B * hw, cfg.t, 0, 0, (hw, cfg.t),
                            ctx["tp_saved"], ctx["tp_fin"], tp_g, False, arena)
    dtable = arena.take(ctx["table"].shape)
    g, _ = _transformer_bwd(g, None, sp, sp_norm, cfg, ctx["table"], dtable, cfg.B * cfg.t, hw, cfg.h, cfg.w,
                            (cfg.t, hw), ctx["sp_saved"], ctx["sp_fin"], sp_g, False, arena)
    # ---- patch embedding backward (no input gradient: the volume is data)
    dg3, db3 = arena.take(dim), arena.take(dim)
    dy0_bf = torch.empty(M, dim, dtype=torch.bfloat16, device=dev)
    dy0 = ops.layernorm_bwd(g, ctx["y0"], g3, ctx["mu0"], ctx["rs0"], dg3, db3, dx_bf16=dy0_bf)
    dbp = arena.take(dim)
    ops.colsum_(dy0, dbp)
    P = arena.take(dim, cfg.K)
    ops.gemm(dy0_bf, ctx["xhat"], ops.EPI_ATOMIC_F32, P, M=dim, N=cfg.K, K=M, mn_major=True, ldc=cfg.K)
    dwp, dg1, db1 = ops.patch_affine_bwd(P, wp, g1, b1, dbp)
    cpb_g = ops.cpb_bwd(dtable.reshape(cfg.heads, -1), cpb[0], cpb[2], cpb[4], ctx["h0"], ctx["h1"], cfg.h, cfg.w)
    return cpb_g + [dg1, db1, dwp, dbp, dg3, db3] + sp_g + tp_g


def _unpack_dtokens(dtokens: torch.Tensor, cfg: "_Cfg"):
    """(dy fp32 [rows, dim], bcast): a mean-pool gradient arrives as a stride-0 expand of [B, dim]."""
    n_tok = cfg.t * cfg.h * cfg.w
    if dtokens.dim() == 5 and dtokens.stride()[1:4] == (0, 0, 0) and dtokens.stride(4) == 1:
        return dtokens[:, 0, 0, 0, :].contiguous().float(), (n_tok, 1.0)
    return dtokens.reshape(cfg.M, cfg.dim).contiguous().float(), None


class _EncoderGraph:
    """CUDA graphs of one training-shape encoder forward and backward.

    The encoder is ~260 (forward) + ~330 (backward) kernel launches with static shapes; enqueued from
    Python they cost ~17 ms of host time per step, more than the forward takes on the GPU, and they delay
    the text tower's launches.  After `WARMUP` eager calls with the same key (shapes, parameter
    addresses, mode) the bodies of `_encode_forward` / `_encode_backward` are captured once and replayed.
    Only two kernels stay outside: the patch gather (it reads the caller's volume, whose address changes
    from step to step) writes into a static buffer, and the incoming token gradient is copied into a
    static buffer.  Saved activations live in the graphs' private pool and are reused every step.
    """
    WARMUP = 2

    def __init__(self):
        self.calls = 0
        self.fwd = None
        self.bwd = {}            # keyed by the layout of the incoming gradient (broadcast or dense)
        self.pool = None
        self.failed = False
        self.owner = None        # weakref to the token of the forward whose activations sit in the static buffers


class _Token:
    """Lives on the autograd ctx of a graphed forward; `done` once its backward has consumed the static activations."""
    __slots__ = ("done", "__weakref__")

    def __init__(self):
        self.done = False


def _graph_key(video, training, params, embed):
    return (tuple(video.shape), bool(training), tuple(p.data_ptr() for p in params), embed.data_ptr())


def _capture(fn, pool, stream=None):
    """capture fn() into a CUDA graph (nothing executes during capture); returns (graph, outputs, launches).
    `stream`: capture stream - kernel nodes inherit ITS priority, not that of the stream the graph is later launched on."""
    from . import _lib
    lib = _lib.load()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    n0 = lib.ctk_launch_count()
    kw = {} if stream is None else {"stream": stream}
    # thread_local: NCCL's watchdog / other streams' threads keep issuing CUDA calls while we capture
    with torch.cuda.graph(g, pool=pool, capture_error_mode="thread_local", **kw):
        out = fn()
    return g, out, int(lib.ctk_launch_count() - n0)


_EAGER = object()       # "the launch decision was taken: this call runs on eager launches"


def _graph_launch(vit: "CTViT", video: torch.Tensor, training: bool, params: List[torch.Tensor]):
    """Decide how this forward runs and, on the graph path, LAUNCH it now (patch gather + replay).
    Returns `_EAGER` or a dict(eg, out, ind, pre_vq, saved, token) that `_CTViTEncode.forward` binds to an autograd
    node later.  Splitting launch from binding lets CTCLIP enqueue the encoder before the text tower's ~350 small
    launches (9 ms of host time during which the GPU would otherwise idle) while the encoder's autograd node is still
    created last, so that its backward is launched first."""
    if not (vit.cuda_graphs and ops.GEMM_PROFILE is None and not torch.cuda.is_current_stream_capturing()):
        return _EAGER
    key = _graph_key(video, training, params, vit.vq._codebook.embed)
    eg = vit._graphs.get(key)
    if eg is None:
        if len(vit._graphs) >= 4:                       # shapes keep changing: stay eager
            vit._graphs.clear()
        eg = vit._graphs[key] = _EncoderGraph()
    eg.calls += 1
    prev = eg.owner() if eg.owner is not None else None
    if eg.failed or eg.calls <= _EncoderGraph.WARMUP:
        return _EAGER
    if prev is not None and not prev.done:
        # a previous graphed forward is still waiting for its backward (e.g. two micro-batches before one
        # backward): replaying would overwrite its saved activations, so this call runs eagerly
        return _EAGER
    cfg = _Cfg(vit, video)
    if eg.fwd is None:
        try:
            eg.pool = torch.cuda.graph_pool_handle()
            st_x = ops.patch_norm_fwd(video, cfg.pt, cfg.p1, cfg.p2)            # static patch buffers
            g, outs, n = _capture(lambda: _encode_forward(vit, video, params, save=True, training=training,
                                                          xhat=st_x[0]), eg.pool)
            eg.fwd = dict(graph=g, outs=outs, launches=n, xbuf=st_x)
        except Exception as e:                                                  # pragma: no cover
            import warnings
            warnings.warn(f"CTViT: CUDA-graph capture failed ({e}); staying on eager launches")
            eg.failed = True
            torch.cuda.synchronize()
            return _EAGER
    f = eg.fwd
    ops.patch_norm_fwd(video, cfg.pt, cfg.p1, cfg.p2, out=f["xbuf"])
    f["graph"].replay()
    ops.GRAPH_LAUNCHES += f["launches"]
    out, ind, pre_vq, saved = f["outs"]
    token = _Token()
    eg.owner = weakref.ref(token)
    # the caller gets private copies (3 small stream-ordered copies): the static buffers are rewritten by the next replay
    return dict(eg=eg, out=out.clone(), ind=ind.clone(), pre_vq=pre_vq.clone(), saved=saved, token=token)


def _eval_graph_forward(vit: "CTViT", video: torch.Tensor, params: List[torch.Tensor]):
    """No-grad, eval-mode forward replayed from a CUDA graph (`CTViT.eval_graphs`; CTK_EVAL_GRAPHS=0 disables).  At one volume per call (zero-shot scoring, BASELINE configs 2 and 4) the ~260 launches of the
    forward cost about as much host time as the GPU needs to execute them; replaying them as one graph removes that.
    Same mechanics as the training graphs: the patch gather stays outside (it reads the caller's volume) and writes into
    static buffers, the caller receives private copies of the outputs.  Returns None when this call runs eagerly."""
    if not (vit.eval_graphs and ops.GEMM_PROFILE is None and not torch.cuda.is_current_stream_capturing()):
        return None
    key = ("eval",) + _graph_key(video, False, params, vit.vq._codebook.embed)
    eg = vit._graphs.get(key)
    if eg is None:
        if len(vit._graphs) >= 4:
            vit._graphs.clear()
        eg = vit._graphs[key] = _EncoderGraph()
    eg.calls += 1
    if eg.failed or eg.calls <= _EncoderGraph.WARMUP:
        return None
    cfg = _Cfg(vit, video)
    if eg.fwd is None:
        try:
            eg.pool = torch.cuda.graph_pool_handle()
            st_x = ops.patch_norm_fwd(video, cfg.pt, cfg.p1, cfg.p2)
            g, outs, n = _capture(lambda: _encode_forward(vit, video, params, save=False, training=False, xhat=st_x[0]),
                                  eg.pool)
            eg.fwd = dict(graph=g, outs=outs, launches=n, xbuf=st_x)
        except Exception as e:                                                  # pragma: no cover
            import warnings
            warnings.warn(f"CTViT: CUDA-graph capture of the eval forward failed ({e}); staying on eager launches")
            eg.failed = True
            torch.cuda.synchronize()
            return None
    f = eg.fwd
    ops.patch_norm_fwd(video, cfg.pt, cfg.p1, cfg.p2, out=f["xbuf"])
    f["graph"].replay()
    ops.GRAPH_LAUNCHES += f["launches"]
    out, ind, pre_vq, _ = f["outs"]
    return out.clone(), ind.clone(), pre_vq.clone()


class _CTViTEncode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vit, video, training, launched, *params):
        params = list(params)
        ctx.vit, ctx.params, ctx.graph = vit, params, None
        if launched is None:
            launched = _graph_launch(vit, video, training, params)
        if launched is _EAGER:
            out, ind, pre_vq, saved = _encode_forward(vit, video, params, save=True, training=training)
            ctx.saved = saved
            ctx.mark_non_differentiable(ind)
            return out, ind, pre_vq.detach()
        ctx.saved, ctx.graph, ctx.token = launched["saved"], launched["eg"], launched["token"]
        out, ind, pre_vq = launched["out"], launched["ind"], launched["pre_vq"]
        ctx.mark_non_differentiable(ind)
        return out, ind, pre_vq

    @staticmethod
    def backward(ctx, dtokens, _dind, dpre):
        assert dtokens is not None
        eg = ctx.graph
        if eg is None or ops.GEMM_PROFILE is not None:
            grads = _encode_backward(ctx.vit, ctx.params, ctx.saved, dtokens)
            ctx.saved = None
            return (None, None, None, None, *grads)
        cfg = ctx.saved["cfg"]
        dy, bcast = _unpack_dtokens(dtokens, cfg)
        bkey = (bcast, tuple(dy.shape))
        b = eg.bwd.get(bkey)
        if b is None:
            dy_static = dy.clone()
            # _transformer_bwd drops its references to the saved activations layer by layer; the graph's static
            # activations must survive for later captures (another gradient layout), so it gets shallow copies
            saved_view = dict(ctx.saved, sp_saved=list(ctx.saved["sp_saved"]), tp_saved=list(ctx.saved["tp_saved"]))
            g, grads, n = _capture(lambda: _encode_backward(ctx.vit, ctx.params, saved_view, None, dy_in=(dy_static, bcast)),
                                   eg.pool)
            b = eg.bwd[bkey] = dict(graph=g, grads=grads, launches=n, dy=dy_static)
        b["dy"].copy_(dy)
        b["graph"].replay()
        ops.GRAPH_LAUNCHES += b["launches"]
        ctx.token.done = True
        return (None, None, None, None, *b["grads"])


class CTViT(nn.Module):
    """Reference: ctvit.py:118-200 (constructor), :353-412 (forward, encoder branch)."""

    def __init__(self, *, dim, codebook_size, image_size, patch_size, temporal_patch_size, spatial_depth,
                 temporal_depth, discr_base_dim=16, dim_head=64, heads=8, channels=1, use_vgg_and_gan=True, vgg=None,
                 discr_attn_res_layers=(16,), use_hinge_loss=True, attn_dropout=0.0, ff_dropout=0.0):
        super().__init__()
        assert channels == 1, "CT volumes are single channel (run_train.py:56-66)"
        self.image_size = pair(image_size)
        self.patch_size = pair(patch_size)
        patch_height, patch_width = self.patch_size
        self.temporal_patch_size = temporal_patch_size
        self.dim, self.heads, self.dim_head = dim, heads, dim_head
        assert dim_head == 32, "libctk attention kernels are specialised for dim_head 32 (run_train.py:64)"
        self.spatial_depth, self.temporal_depth = spatial_depth, temporal_depth
        self.ff_inner = int(4 * (2 / 3) * dim)

        self.spatial_rel_pos_bias = ContinuousPositionBias(dim=dim, heads=heads)
        image_height, image_width = self.image_size
        assert (image_height % patch_height) == 0 and (image_width % patch_width) == 0
        pdim_ff = channels * patch_width * patch_height
        pdim = pdim_ff * temporal_patch_size
        # index 0 of both Sequentials is the einops Rearrange in the reference (no parameters)
        self.to_patch_emb_first_frame = nn.Sequential(nn.Identity(), nn.LayerNorm(pdim_ff), nn.Linear(pdim_ff, dim),
                                                      nn.LayerNorm(dim))
        self.to_patch_emb = nn.Sequential(nn.Identity(), nn.LayerNorm(pdim), nn.Linear(pdim, dim), nn.LayerNorm(dim))
        kw = dict(dim=dim, dim_head=dim_head, heads=heads, attn_dropout=attn_dropout, ff_dropout=ff_dropout, peg=True,
                  peg_causal=True)
        self.enc_spatial_transformer = Transformer(depth=spatial_depth, **kw)
        self.enc_temporal_transformer = Transformer(depth=temporal_depth, **kw)
        self.vq = VectorQuantize(dim=dim, codebook_size=codebook_size, use_cosine_sim=True)
        self.to_pixels_first_frame = nn.Sequential(nn.Linear(dim, pdim_ff), nn.Identity())
        self.to_pixels = nn.Sequential(nn.Linear(dim, pdim), nn.Identity())
        # training-shape forward / backward are replayed from CUDA graphs after two eager steps (see _EncoderGraph);
        # set to False to keep every step on eager launches
        self.cuda_graphs = os.environ.get("CTK_CUDA_GRAPHS", "1") != "0"
        # the no-grad eval forward (zero-shot scoring, one volume per call) is replayed from a CUDA graph as well
        # (CTK_EVAL_GRAPHS=0 keeps eager launches)
        self.eval_graphs = os.environ.get("CTK_EVAL_GRAPHS", "1") != "0"
        self._graphs = {}

    # -- reference helpers kept for callers --------------------------------------------------------
    @property
    def image_num_tokens(self):
        return int(self.image_size[0] / self.patch_size[0]) * int(self.image_size[1] / self.patch_size[1])

    @property
    def patch_height_width(self):
        return self.image_size[0] // self.patch_size[0], self.image_size[1] // self.patch_size[1]

    def load(self, path):
        path = Path(path)
        assert path.exists()
        self.load_state_dict(torch.load(str(path)))

    def _flat_params(self) -> List[torch.Tensor]:
        cpb = self.spatial_rel_pos_bias.net
        p = [cpb[0][0].weight, cpb[0][0].bias, cpb[1][0].weight, cpb[1][0].bias, cpb[2].weight, cpb[2].bias]
        pe = self.to_patch_emb
        p += [pe[1].weight, pe[1].bias, pe[2].weight, pe[2].bias, pe[3].weight, pe[3].bias]
        p += _layer_params(self.enc_spatial_transformer)
        p += _layer_params(self.enc_temporal_transformer)
        return p

    def encode_begin(self, video: torch.Tensor):
        """Optional early launch (used by CTCLIP): if this training-shape forward can be replayed from its CUDA graph,
        enqueue it NOW on the current stream and return a handle for `forward(..., _launched=handle)`, which only
        binds the autograd node.  Returns None when there is nothing to gain (no-grad call)."""
        assert video.is_cuda, "CTViT runs on sm_100a only: move the module and its input to a CUDA device"
        params = self._flat_params()
        if not (torch.is_grad_enabled() and any(p.requires_grad for p in params)):
            return None
        video = video.contiguous().float()
        with torch.no_grad():
            return (_graph_launch(self, video, self.training, params), video)

    def encode_with_aux(self, video: torch.Tensor, _launched=None):
        """returns (tokens after VQ, code indices, tokens before VQ)"""
        assert video.is_cuda, "CTViT runs on sm_100a only: move the module and its input to a CUDA device"
        params = self._flat_params()
        if _launched is not None:
            launched, video = _launched
            return _CTViTEncode.apply(self, video, self.training, launched, *params)
        video = video.contiguous().float()
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _CTViTEncode.apply(self, video, self.training, None, *params)
        if self.eval_graphs and not self.training:
            replayed = _eval_graph_forward(self, video, params)
            if replayed is not None:
                return replayed
        out, ind, pre, _ = _encode_forward(self, video, params, save=False, training=self.training)
        return out, ind, pre

    def forward(self, video, mask=None, return_recons=False, return_recons_only=False, return_discr_loss=False,
                apply_grad_penalty=True, return_only_codebook_ids=False, return_encoded_tokens=False, _launched=None):
        assert video.ndim in {4, 5}
        if video.ndim == 4:
            video = video[:, :, None]
            assert mask is None
        b, c, f, *image_dims = video.shape
        assert tuple(image_dims) == self.image_size
        assert mask is None, "frame masks are not used on the CT-CLIP path"
        tokens, indices, _ = self.encode_with_aux(video, _launched=_launched)
        if return_only_codebook_ids:
            return indices                      # (b, t, h, w), as the reference unpacks them (ctvit.py:399-400)
        if return_encoded_tokens:
            return tokens
        raise NotImplementedError("only the encoder branch (return_encoded_tokens=True / return_only_codebook_ids=True) "
                                  "of CTViT is on the CT-CLIP hot path (ctvit.py:411-412)")
