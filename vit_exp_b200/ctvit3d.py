"""Drop-in `CTViT3D` (reference: transformer_maskgit/transformer_maskgit/ctvit3d.py:175-520, attention.py:189-284
`FlashAttention`) - SURVEY.md 8f rank 3, the encoder most of the reference's training configurations use
(scripts/run_train.py:36-50: dim 768, 8 joint blocks, heads 8 x 32, patch 20x20x10 of 480x480x240).

Encoder branch only (`return_encoded_tokens=True`, ctvit3d.py:455-486): patch embedding (same kernels as CTViT) +
fixed 3-D sin/cos position table, then 8 x [x += attn(x); x += ff(x)] over ALL 13 824 tokens of a volume jointly
(no PEG, no factorisation), final LayerNorm, no vector quantisation.  Attention differs from CTViT's: 2 learned
null key/value pairs are prepended per head, l2norm and q/k scales apply to them too, the logit scale is
1/sqrt(dim_head) (SDPA default) instead of 8, and the position bias is ignored (attention.py:257).

What runs where: everything in libctk.  LayerNorm, the q / kv projections with the per-head l2norm*scale epilogue
(1/sqrt(dh) folded into q), the output projection + residual, the GEGLU feed-forward and every gradient product are
the kernels validated for CTViT, at dim 768; the attention core over the 2 null pairs + 13 824 tokens is
`ctk_mha_fwd / ctk_mha_bwd` (csrc/attention_mha.cu, head dim 32: flash-style mma.sync kernels streaming 64-key blocks,
the null pairs as one extra block, no 13 826-wide row ever stored).  libctk's tcgen05 attention keeps all keys of a
sequence in shared memory, which fits the 576-token slices of CTViT but not 13 826 keys; a streaming-KV tcgen05 version
is the follow-up.  Host logic checked on CPU against the oracle pinned to the reference module
(tests/test_ctvit3d_cpu.py), kernels on the GPU (tests/test_ctvit3d_gpu.py).
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .transformer_maskgit import ContinuousPositionBias, Transformer, _unpack_dtokens, pair

N_PER_LAYER = 11
OPERAND_DTYPE = torch.bfloat16      # tests set this to float32 to check the host math exactly on CPU


def sincos_pos_embed_3d(dim: int, grid) -> torch.Tensor:
    """Fixed table of ctvit3d.py:122-173, (n_t*n_h*n_w, dim).  The reference builds np.meshgrid(t, w, h) in numpy's
    default 'xy' indexing - arrays of shape (n_w, n_t, n_h) - and then *reshapes* them to (n_t, n_w, n_h)
    (ctvit3d.py:131-135); that memory-order reinterpretation is part of every trained checkpoint's table, so it is
    reproduced here.  Each coordinate fills a third of the channels as [sin | cos] with frequencies
    10000^(-i / (dim / 6))."""
    n_t, n_h, n_w = grid
    assert dim % 6 == 0, "CTViT3D needs dim % 6 == 0 (ctvit3d.py:140)"
    gt = torch.arange(n_t, dtype=torch.float32)[None, :, None].expand(n_w, n_t, n_h)
    gw = torch.arange(n_w, dtype=torch.float32)[:, None, None].expand(n_w, n_t, n_h)
    gh = torch.arange(n_h, dtype=torch.float32)[None, None, :].expand(n_w, n_t, n_h)
    d3 = dim // 3
    omega = 1.0 / 10000 ** (torch.arange(d3 // 2, dtype=torch.float32) / (d3 / 2.0))
    parts = []
    for pos in (gt, gw, gh):
        ang = pos.reshape(-1)[:, None] * omega[None, :]
        parts += [torch.sin(ang), torch.cos(ang)]
    return torch.cat(parts, dim=1)


class _Cfg3D:
    def __init__(self, vit: "CTViT3D", video: torch.Tensor):
        B, C, D, H, W = video.shape
        self.B = B
        self.p1, self.p2 = vit.patch_size
        self.pt = vit.temporal_patch_size
        self.t, self.h, self.w = D // self.pt, H // self.p1, W // self.p2
        self.n = self.t * self.h * self.w
        self.M = B * self.n
        self.dim, self.heads = vit.dim, vit.heads
        self.inner = vit.heads * 32
        self.ff_inner = vit.ff_inner
        self.ff_pad = (self.ff_inner + 127) // 128 * 128
        self.K = self.pt * self.p1 * self.p2
        self.Kp = (self.K + 7) // 8 * 8
        self.depth = vit.transformer_blocks
        self.q_alpha = 32 ** -0.5                    # SDPA default scale folded into q (attention.py:254)


def _layer_params(tr: Transformer) -> List[torch.Tensor]:
    out = []
    for _, attn, _, ff in tr.layers:
        out += [attn.norm.gamma, attn.to_q.weight, attn.to_kv.weight, attn.q_scale, attn.k_scale, attn.null_kv,
                attn.to_out.weight, ff[0].weight, ff[0].bias, ff[1].weight, ff[4].weight]
    out.append(tr.norm_out.gamma)
    return out


def _prep(lp, cfg: _Cfg3D, need_bwd: bool) -> Dict[str, torch.Tensor]:
    (gamma, wq, wkv, qs, ks, null_kv, wo, fg, fb, w1, w2) = lp
    d: Dict[str, torch.Tensor] = dict(wq=ops.cast_bf16(wq), wkv=ops.cast_bf16(wkv), wo=ops.cast_bf16(wo))
    d["w1p"], d["w1p_t"], d["w1_map"] = ops.pack_ff_w1(w1, cfg.ff_inner, cfg.ff_pad, want_t=need_bwd)
    d["w2"] = ops.cast_bf16(w2, ld=cfg.ff_pad)
    if need_bwd:
        d["wq_t"] = ops.transpose_cast_bf16(wq)
        d["wkv_t"] = ops.transpose_cast_bf16(wkv)
        d["wo_t"] = ops.transpose_cast_bf16(wo)
        w2t = torch.zeros(cfg.ff_pad, cfg.dim, dtype=OPERAND_DTYPE, device=w2.device)
        ops.transpose_cast_bf16(w2, out=w2t[: cfg.ff_inner])
        d["w2_t"] = w2t
    return d


def _null_kv(null_kv: torch.Tensor, k_scale: torch.Tensor, need_bwd: bool):
    """(nk, nv) [heads, n_null, 32]: the learned null pairs as the attention sees them - keys l2-normalised and
    scaled like every other key (attention.py:240-248).  A few hundred numbers: plain torch, differentiable."""
    heads, two_n, dh = null_kv.shape
    with torch.enable_grad() if need_bwd else torch.no_grad():
        leaf_n = null_kv.detach().float().requires_grad_(need_bwd)
        leaf_s = k_scale.detach().float().requires_grad_(need_bwd)
        nk, nv = leaf_n.reshape(heads, two_n // 2, 2, dh).unbind(dim=-2)
        nk = F.normalize(nk, dim=-1) * leaf_s
    return nk, nv, (leaf_n, leaf_s)


def _attention(qkv, nk, nv, cfg: _Cfg3D, need_bwd: bool):
    """packed qkv [M, 3*inner] (q pre-scaled by q_scale/sqrt(dh), k by k_scale) + null pairs -> context [M, inner].
    `FlashAttention` core (attention.py:250-260: SDPA over the 2 null pairs + 13 824 tokens, no mask, no bias) on
    libctk's flash-style kernel (ctk_mha_fwd, head dim 32): the null pairs are one extra key block, the 13 826-wide
    probability rows never leave registers."""
    nkb = nk.detach().to(qkv.dtype).contiguous()
    nvb = nv.detach().to(qkv.dtype).contiguous()
    ctx, lse = ops.mha_fwd(qkv, None, cfg.B, cfg.n, cfg.heads, 1.0, null_k=nkb, null_v=nvb)
    return ctx, ((qkv, nkb, nvb, ctx, lse) if need_bwd else None)


def _attention_bwd(saved, dctx, cfg: _Cfg3D, n_null: int):
    """-> (dqkv [M, 3*inner] in the packed layout, d nk, d nv [heads, n_null, 32] fp32 summed over the batch)"""
    qkv, nkb, nvb, ctx, lse = saved
    dqkv, dnk, dnv = ops.mha_bwd(qkv, None, ctx, dctx.contiguous(), lse, cfg.B, cfg.n, cfg.heads, 1.0, null_k=nkb, null_v=nvb)
    return dqkv, dnk.sum(0), dnv.sum(0)


def _forward(vit: "CTViT3D", video, params: List[torch.Tensor], save: bool):
    cfg = _Cfg3D(vit, video)
    dim, M, heads, inner = cfg.dim, cfg.M, cfg.heads, cfg.inner
    dev = video.device
    od = OPERAND_DTYPE
    it = iter(params)
    g1, b1, wp, bp, g3, b3 = (next(it) for _ in range(6))
    lps = [[next(it) for _ in range(N_PER_LAYER)] for _ in range(cfg.depth)]
    norm_gamma = next(it)
    # ---- patch embedding (ctvit3d.py:240-245) as in CTViT: LayerNorm(K) affine folded into the projection
    xhat, _, _ = ops.patch_norm_fwd(video, cfg.pt, cfg.p1, cfg.p2)
    wp_eff = ops.cast_bf16(wp, ld=cfg.Kp, col_scale=g1)
    bias_eff = torch.addmv(bp, wp, b1)
    y0 = torch.empty(M, dim, dtype=torch.float32, device=dev)
    ops.gemm(xhat, wp_eff, ops.EPI_F32, y0, M=M, N=dim, K=cfg.Kp, bias=bias_eff)
    _, x0, _, mu0, rs0 = ops.layernorm_fwd(y0, g3, b3, want_bf16=False, want_f32=True)
    x = (x0.view(cfg.B, cfg.n, dim) + vit.pos_embed.detach()).view(M, dim)          # ctvit3d.py:373 (fixed table)
    saved = []
    for lp in lps:
        (gamma, wq, wkv, qs, ks, null_kv, wo, fg, fb, w1, w2) = lp
        w = _prep(lp, cfg, save)
        xn, _, xraw, mu1, rs1 = ops.layernorm_fwd(x, gamma, None, want_raw=True)
        qkv = torch.empty(M, 3 * inner, dtype=od, device=dev)
        rn = torch.empty(M, 2 * heads, dtype=torch.float32, device=dev)
        ops.gemm(xn, w["wq"], ops.EPI_QKV, qkv, M=M, N=inner, K=dim, aux0=rn, ld_aux0=2 * heads, vec0=qs,
                 alpha=cfg.q_alpha, i0=inner, i1=0)
        ops.gemm(xraw, w["wkv"], ops.EPI_QKV, qkv, M=M, N=2 * inner, K=dim, aux0=rn, ld_aux0=2 * heads, vec0=ks,
                 alpha=1.0, i0=inner, i1=inner)
        nk, nv, null_leaves = _null_kv(null_kv, ks, save)
        o, attn_saved = _attention(qkv, nk, nv, cfg, save)
        x2 = torch.empty(M, dim, dtype=torch.float32, device=dev)
        ops.gemm(o, w["wo"], ops.EPI_RESID_F32, x2, M=M, N=dim, K=inner, resid=x)
        hn, _, _, mu2, rs2 = ops.layernorm_fwd(x2, fg, fb)
        U = torch.empty(M, 2 * cfg.ff_pad, dtype=od, device=dev)
        Hh = torch.empty(M, cfg.ff_pad, dtype=od, device=dev)
        ops.gemm(hn, w["w1p"], ops.EPI_GEGLU, U, M=M, N=2 * cfg.ff_pad, K=dim, aux0=Hh, ld_aux0=cfg.ff_pad)
        x3 = torch.empty(M, dim, dtype=torch.float32, device=dev)
        ops.gemm(Hh, w["w2"], ops.EPI_RESID_F32, x3, M=M, N=dim, K=cfg.ff_pad, resid=x2)
        if save:
            saved.append(dict(w=w, x=x, xn=xn, xraw=xraw, mu1=mu1, rs1=rs1, qkv=qkv, rn=rn, o=o, attn=attn_saved,
                              nk=nk, nv=nv, null_leaves=null_leaves, x2=x2, hn=hn, mu2=mu2, rs2=rs2, U=U, H=Hh))
        x = x3
    _, y, _, mu, rs = ops.layernorm_fwd(x, norm_gamma, None, want_bf16=False, want_f32=True)
    ctx = dict(cfg=cfg, xhat=xhat, y0=y0, mu0=mu0, rs0=rs0, saved=saved, fin=dict(x=x, mu=mu, rs=rs)) if save else None
    return y.view(cfg.B, cfg.t, cfg.h, cfg.w, dim), ctx


def _backward(vit: "CTViT3D", params: List[torch.Tensor], ctx, dtokens: torch.Tensor):
    cfg: _Cfg3D = ctx["cfg"]
    dim, M, heads, inner = cfg.dim, cfg.M, cfg.heads, cfg.inner
    od = OPERAND_DTYPE
    it = iter(params)
    g1, b1, wp, bp, g3, b3 = (next(it) for _ in range(6))
    lps = [[next(it) for _ in range(N_PER_LAYER)] for _ in range(cfg.depth)]
    norm_gamma = next(it)
    dev = norm_gamma.device
    f32 = dict(dtype=torch.float32, device=dev)
    dy, bcast = _unpack_dtokens(dtokens, cfg)            # a mean-pool gradient stays an un-materialised broadcast
    fin = ctx["fin"]
    dgo = torch.zeros(dim, **f32)
    g_bf = torch.empty(M, dim, dtype=od, device=dev)
    if bcast is not None:
        g = ops.layernorm_bwd(dy, fin["x"], norm_gamma, fin["mu"], fin["rs"], dgo, None, bcast_rows=bcast[0],
                              dy_scale=bcast[1], dx_bf16=g_bf)
    else:
        g = ops.layernorm_bwd(dy, fin["x"], norm_gamma, fin["mu"], fin["rs"], dgo, None, dx_bf16=g_bf)
    layer_grads: List[Optional[torch.Tensor]] = [None] * (cfg.depth * N_PER_LAYER)
    for li in range(cfg.depth - 1, -1, -1):
        (gamma, wq, wkv, qs, ks, null_kv, wo, fg, fb, w1, w2) = lps[li]
        s = ctx["saved"][li]
        w = s["w"]
        # ---- feed-forward: x3 = x2 + W2 geglu(W1 LN(x2))
        dw2 = torch.zeros_like(w2)
        ops.gemm(g_bf, s["H"], ops.EPI_ATOMIC_F32, dw2, M=dim, N=cfg.ff_inner, K=M, mn_major=True, ldc=cfg.ff_inner)
        dU = torch.empty_like(s["U"])
        ops.gemm(g_bf, w["w2_t"], ops.EPI_GEGLU_BWD, dU, M=M, N=cfg.ff_pad, K=dim, aux0=s["U"], ld_aux0=2 * cfg.ff_pad)
        dw1 = torch.zeros_like(w1)
        ops.gemm(dU, s["hn"], ops.EPI_ATOMIC_F32, dw1, M=2 * cfg.ff_pad, N=dim, K=M, mn_major=True, ldc=dim,
                 row_map=w["w1_map"])
        dhn = torch.empty(M, dim, dtype=od, device=dev)
        ops.gemm(dU, w["w1p_t"], ops.EPI_BF16, dhn, M=M, N=dim, K=2 * cfg.ff_pad)
        dfg, dfb = torch.zeros(dim, **f32), torch.zeros(dim, **f32)
        ops.layernorm_bwd(dhn, s["x2"], fg, s["mu2"], s["rs2"], dfg, dfb, dx=g, accum=True, dx_bf16=g_bf)
        # ---- attention: x2 = x + Wo attn(q(LN(x)), kv(x), null kv)
        dwo = torch.zeros_like(wo)
        ops.gemm(g_bf, s["o"], ops.EPI_ATOMIC_F32, dwo, M=dim, N=inner, K=M, mn_major=True, ldc=inner)
        do = torch.empty(M, inner, dtype=od, device=dev)
        ops.gemm(g_bf, w["wo_t"], ops.EPI_BF16, do, M=M, N=inner, K=dim)
        n_null = null_kv.shape[1] // 2
        dqkv, dnk, dnv = _attention_bwd(s["attn"], do, cfg, n_null)
        dqs, dks = torch.zeros(32, **f32), torch.zeros(32, **f32)
        ops.qknorm_bwd_(dqkv, s["qkv"], s["rn"], qs, ks, cfg.q_alpha, dqs, dks, heads)
        # null pairs: through the tiny normalise-and-scale graph of _null_kv (adds the null keys' share of dk_scale)
        leaf_n, leaf_s = s["null_leaves"]
        dnull, dks_null = torch.autograd.grad((s["nk"], s["nv"]), (leaf_n, leaf_s), (dnk, dnv))
        dks = dks + dks_null
        dwq = torch.zeros_like(wq)
        ops.gemm(dqkv, s["xn"], ops.EPI_ATOMIC_F32, dwq, M=inner, N=dim, K=M, mn_major=True, lda=3 * inner, ldc=dim)
        dwkv = torch.zeros_like(wkv)
        dkv_view = dqkv[:, inner:]
        ops.gemm(dkv_view, s["xraw"], ops.EPI_ATOMIC_F32, dwkv, M=2 * inner, N=dim, K=M, mn_major=True,
                 lda=3 * inner, ldc=dim)
        ops.gemm(dkv_view, w["wkv_t"], ops.EPI_RESID_F32, g, M=M, N=dim, K=2 * inner, lda=3 * inner, resid=g)
        dxn = torch.empty(M, dim, dtype=od, device=dev)
        ops.gemm(dqkv, w["wq_t"], ops.EPI_BF16, dxn, M=M, N=dim, K=inner, lda=3 * inner)
        dgamma = torch.zeros(dim, **f32)
        ops.layernorm_bwd(dxn, s["x"], gamma, s["mu1"], s["rs1"], dgamma, None, dx=g, accum=True, dx_bf16=g_bf)
        base = li * N_PER_LAYER
        layer_grads[base: base + N_PER_LAYER] = [dgamma, dwq, dwkv, dqs, dks, dnull.to(null_kv.dtype), dwo, dfg, dfb,
                                                 dw1, dw2]
        ctx["saved"][li] = None
    # ---- patch embedding backward (pos_embed is a fixed table: no gradient; the volume is data)
    dg3, db3 = torch.zeros(dim, **f32), torch.zeros(dim, **f32)
    dy0_bf = torch.empty(M, dim, dtype=od, device=dev)
    dy0 = ops.layernorm_bwd(g, ctx["y0"], g3, ctx["mu0"], ctx["rs0"], dg3, db3, dx_bf16=dy0_bf)
    dbp = torch.zeros(dim, **f32)
    ops.colsum_(dy0, dbp)
    P = torch.zeros(dim, cfg.K, **f32)
    ops.gemm(dy0_bf, ctx["xhat"], ops.EPI_ATOMIC_F32, P, M=dim, N=cfg.K, K=M, mn_major=True, ldc=cfg.K)
    dwp, dg1, db1 = ops.patch_affine_bwd(P, wp, g1, b1, dbp)
    return [dg1, db1, dwp, dbp, dg3, db3] + layer_grads + [dgo]


class _CTViT3DEncode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vit, video, *params):
        plist = [p.detach() for p in params]
        out, saved = _forward(vit, video, plist, save=True)
        ctx.vit, ctx.plist, ctx.saved = vit, plist, saved
        return out

    @staticmethod
    def backward(ctx, dtokens):
        grads = _backward(ctx.vit, ctx.plist, ctx.saved, dtokens)
        ctx.saved = None
        return (None, None, *grads)


class CTViT3D(nn.Module):
    """Reference constructor: ctvit3d.py:176-199 (keyword-only; `use_seg` heads are outside the contrastive path)."""

    def __init__(self, *, dim, image_size, patch_size, temporal_size, temporal_patch_size, transformer_blocks=8,
                 discr_base_dim=16, dim_head=64, heads=8, channels=1, use_vgg_and_gan=True, vgg=None,
                 discr_attn_res_layers=(16,), use_hinge_loss=True, attn_dropout=0.0, ff_dropout=0.0,
                 use_flash_attention=True, use_seg=False, **kwargs):
        super().__init__()
        assert channels == 1, "CT volumes are single channel"
        assert use_flash_attention, "every launcher builds CTViT3D with use_flash_attention=True (run_train.py:47)"
        assert not use_seg, "segmentation heads are outside the contrastive hot path"
        assert dim_head == 32, "libctk's q/k normalisation epilogue is specialised for dim_head 32 (run_train.py:45)"
        self.image_size = pair(image_size)
        self.patch_size = pair(patch_size)
        ph, pw = self.patch_size
        self.temporal_patch_size = temporal_patch_size
        self.dim, self.heads, self.dim_head = dim, heads, dim_head
        self.transformer_blocks = transformer_blocks
        self.ff_inner = int(4 * (2 / 3) * dim)
        ih, iw = self.image_size
        assert ih % ph == 0 and iw % pw == 0
        n_t, n_h, n_w = temporal_size // temporal_patch_size, ih // ph, iw // pw
        self.patch_voxel_nums = ph * pw * temporal_patch_size
        self.pos_embed = nn.Parameter(sincos_pos_embed_3d(dim, (n_t, n_h, n_w))[None], requires_grad=False)
        self.spatial_rel_pos_bias = ContinuousPositionBias(dim=dim, heads=heads)     # decoder-side; unused here
        pdim = channels * ph * pw * temporal_patch_size
        self.to_patch_emb = nn.Sequential(nn.Identity(), nn.LayerNorm(pdim), nn.Linear(pdim, dim), nn.LayerNorm(dim))
        self.enc_3D = Transformer(dim, depth=transformer_blocks, dim_head=dim_head, heads=heads,
                                  attn_dropout=attn_dropout, ff_dropout=ff_dropout, peg=False, attn_num_null_kv=2)
        self.to_pixels = nn.Sequential(nn.Linear(dim, pdim), nn.Identity())

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
        pe = self.to_patch_emb
        return [pe[1].weight, pe[1].bias, pe[2].weight, pe[2].bias, pe[3].weight, pe[3].bias] + _layer_params(self.enc_3D)

    def forward(self, video, mask=None, return_recons=False, return_recons_only=False, return_discr_loss=False,
                apply_grad_penalty=True, return_only_codebook_ids=False, return_encoded_tokens=False):
        assert video.ndim in {4, 5}
        if video.ndim == 4:
            video = video[:, :, None]
            assert mask is None
        b, c, f, *image_dims = video.shape
        assert tuple(image_dims) == self.image_size
        assert mask is None, "frame masks are not used on the CT-CLIP path"
        if not return_encoded_tokens:
            raise NotImplementedError("only the encoder branch (return_encoded_tokens=True) of CTViT3D is on the "
                                      "CT-CLIP hot path (ctvit3d.py:485-486)")
        assert video.is_cuda, "CTViT3D runs on sm_100a only: move the module and its input to a CUDA device"
        video = video.contiguous().float()
        params = self._flat_params()
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _CTViT3DEncode.apply(self, video, *params)
        return _forward(self, video, [p.detach() for p in params], save=False)[0]
