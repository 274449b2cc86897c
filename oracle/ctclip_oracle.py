"""CPU oracle for the CT-CLIP training hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker or the timed CPU baseline; the product package
(vit_exp_b200) never imports it and has no CPU path.

It is a plain functional restatement (torch on CPU, fp32 or fp64 according to the dtype of the
inputs) of the reference's algorithm, written from the reference's semantics; each function cites
the reference file:line it follows (paths relative to the upstream repo root; `attention.py`,
`ctvit.py` live in transformer_maskgit/transformer_maskgit/, `ct_clip.py`, `distributed.py` in
CT_CLIP/ct_clip/).  Parameters are passed as a flat dict keyed by the reference's state-dict names,
so a reference checkpoint feeds it directly.

Pinning (see oracle/make_golden.py and tests/test_oracle_cpu.py):
  * the reference's only fixed vectors (demo_tests/test_loss_type.py:14-15) with the known answers
    recorded in SURVEY.md section 8c;
  * golden fixtures under tests/golden/ produced by importing the *real* reference modules in the
    build container (stub recipe of SURVEY.md appendix D) on seeded inputs.
Exception - parity UNPINNED: VectorQuantize comes from vector-quantize-pytorch==1.1.2
(transformer_maskgit/setup.py:20), which is neither vendored nor installed; `vq_cosine` restates
its published CosineSimCodebook algorithm from memory and nothing in the reference exercises it.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# ------------------------------------------------------------------------------------------
# primitives
# ------------------------------------------------------------------------------------------
def l2norm(t: torch.Tensor) -> torch.Tensor:
    """attention.py:28-29 / ct_clip.py:70-71 : F.normalize(dim=-1), eps 1e-12."""
    return t / t.norm(dim=-1, keepdim=True).clamp_min(1e-12)


def layer_norm(x, gamma, beta=None, eps: float = 1e-5):
    """attention.py:34-41 (gamma, zero beta buffer) and nn.LayerNorm (attention.py:53, ctvit.py:172,174)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    y = (x - mu) / torch.sqrt(var + eps) * gamma
    return y if beta is None else y + beta


def gelu_exact(x):
    """attention.py:48 F.gelu default = erf form."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def feed_forward(x, p: Params, pre: str):
    """attention.py:50-58: LayerNorm -> Linear(dim, 2*inner, no bias) -> GEGLU (first half value,
    second half gate, :46-48) -> Linear(inner, dim, no bias)."""
    h = layer_norm(x, p[pre + "0.weight"], p[pre + "0.bias"])
    h = h @ p[pre + "1.weight"].T
    val, gate = h.chunk(2, dim=-1)
    h = gelu_exact(gate) * val
    return h @ p[pre + "4.weight"].T


def peg(x, shape: Tuple[int, int, int, int], w, b):
    """attention.py:62-90 with causal=True (ctvit.py:184): flat tokens are *reshaped* to `shape`
    (b, t, h, w) regardless of their semantic order, zero-padded (2,0) on the first grid axis and
    (1,1) on the other two, then a depthwise 3x3x3 conv (cross-correlation) + bias."""
    orig = x.shape
    d = x.shape[-1]
    v = x.reshape(*shape, d).permute(0, 4, 1, 2, 3)
    v = F.pad(v, (1, 1, 1, 1, 2, 0), value=0.0)
    v = F.conv3d(v, w, b, groups=d)
    return v.permute(0, 2, 3, 4, 1).reshape(orig)


def cpb_bias(p: Params, pre: str, gh: int, gw: int):
    """attention.py:363-382 ContinuousPositionBias (num_dims 2, layers 2, log_dist): all (gh*gw)^2
    relative offsets -> sign*log(|.|+1) -> MLP with LeakyReLU(0.1) -> (heads, i, j)."""
    dt, dev = p[pre + "net.0.0.weight"].dtype, p[pre + "net.0.0.weight"].device
    ys, xs = torch.meshgrid(torch.arange(gh, device=dev), torch.arange(gw, device=dev), indexing="ij")
    grid = torch.stack([ys, xs]).reshape(2, -1).T                      # (gh*gw, 2)
    rel = (grid[:, None, :] - grid[None, :, :]).to(dt)
    rel = torch.sign(rel) * torch.log(rel.abs() + 1)
    h = F.leaky_relu(rel @ p[pre + "net.0.0.weight"].T + p[pre + "net.0.0.bias"], 0.1)
    h = F.leaky_relu(h @ p[pre + "net.1.0.weight"].T + p[pre + "net.1.0.bias"], 0.1)
    h = h @ p[pre + "net.2.weight"].T + p[pre + "net.2.bias"]
    return h.permute(2, 0, 1)                                          # (heads, i, j)


def attention(x, p: Params, pre: str, heads: int, attn_bias=None, scale: float = 8.0):
    """attention.py:133-187 self-attention path (no context, no mask, non-causal, empty null_kv).
    NOTE the reference takes kv_input = x BEFORE normalising (attention.py:145-147): q is projected
    from LayerNorm(x) but k, v from the raw residual stream.  Then per-head l2norm * q_scale /
    k_scale -> sim * 8 (+ bias) -> softmax -> attn @ v -> to_out."""
    b, n, _ = x.shape
    xn = layer_norm(x, p[pre + "norm.gamma"])
    q = xn @ p[pre + "to_q.weight"].T
    k, v = (x @ p[pre + "to_kv.weight"].T).chunk(2, dim=-1)
    q, k, v = (t.reshape(b, n, heads, -1).permute(0, 2, 1, 3) for t in (q, k, v))
    q = l2norm(q) * p[pre + "q_scale"]
    k = l2norm(k) * p[pre + "k_scale"]
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * scale
    if attn_bias is not None:
        sim = sim + attn_bias
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", attn, v)
    out = out.permute(0, 2, 1, 3).reshape(b, n, -1)
    return out @ p[pre + "to_out.weight"].T


def transformer(x, p: Params, pre: str, depth: int, heads: int, video_shape, attn_bias=None):
    """attention.py:441-452: per layer x = peg(x)+x; x = attn(x)+x; x = ff(x)+x; final LayerNorm."""
    for l in range(depth):
        lp = f"{pre}layers.{l}."
        x = peg(x, video_shape, p[lp + "0.dsconv.weight"], p[lp + "0.dsconv.bias"]) + x
        x = attention(x, p, lp + "1.", heads, attn_bias) + x
        x = feed_forward(x, p, lp + "3.") + x
    return layer_norm(x, p[pre + "norm_out.gamma"])


def patch_embed(video, p: Params, patch: int, tpatch: int):
    """ctvit.py:170-175: 'b c (t pt)(h p1)(w p2) -> b t h w (c pt p1 p2)', LayerNorm, Linear, LayerNorm."""
    b, c, D, H, W = video.shape
    t, h, w = D // tpatch, H // patch, W // patch
    x = video.reshape(b, c, t, tpatch, h, patch, w, patch).permute(0, 2, 4, 6, 1, 3, 5, 7)
    x = x.reshape(b, t, h, w, c * tpatch * patch * patch)
    x = layer_norm(x, p["to_patch_emb.1.weight"], p["to_patch_emb.1.bias"])
    x = x @ p["to_patch_emb.2.weight"].T + p["to_patch_emb.2.bias"]
    return layer_norm(x, p["to_patch_emb.3.weight"], p["to_patch_emb.3.bias"])


def ctvit_encode(tokens, p: Params, spatial_depth: int, temporal_depth: int, heads: int):
    """ctvit.py:282-307: spatial stack on '(b t) (h w) d' with the continuous position bias, then
    temporal stack on '(b h w) t d' (PEG still reshapes with video_shape=(b,t,h,w): ctvit.py:289,303)."""
    b, t, h, w, d = tokens.shape
    video_shape = (b, t, h, w)
    x = tokens.reshape(b * t, h * w, d)
    bias = cpb_bias(p, "spatial_rel_pos_bias.", h, w)
    x = transformer(x, p, "enc_spatial_transformer.", spatial_depth, heads, video_shape, bias)
    x = x.reshape(b, t, h, w, d).permute(0, 2, 3, 1, 4).reshape(b * h * w, t, d)
    x = transformer(x, p, "enc_temporal_transformer.", temporal_depth, heads, video_shape, None)
    return x.reshape(b, h, w, t, d).permute(0, 3, 1, 2, 4)


def vq_cosine(x, embed, training: bool = False, cluster_size=None, decay: float = 0.8):
    """VectorQuantize(dim, codebook_size, use_cosine_sim=True) forward, vector-quantize-pytorch 1.1.2
    (call site ctvit.py:188,403).  PARITY UNPINNED - restated from the published algorithm:
    flatten = l2norm(x); dist = flatten @ l2norm(embed)^T; ind = argmax; quantize = embed[ind];
    training: EMA of cluster_size and of the l2-normalised per-code means (codes with no hits keep
    their value), straight-through estimator, commitment loss ignored by the caller
    (threshold_ema_dead_code defaults to 0 in VectorQuantize -> no code expiry).
    Returns (quantize, ind, new_embed, new_cluster_size)."""
    shape = x.shape
    flat = l2norm(x.reshape(-1, shape[-1]))
    en = l2norm(embed)
    ind = (flat @ en.T).argmax(dim=-1)
    quant = embed[ind].reshape(shape)
    new_embed, new_cs = embed, cluster_size
    if training:
        C = embed.shape[0]
        bins = torch.bincount(ind, minlength=C).to(x.dtype)
        new_cs = cluster_size * decay + bins * (1 - decay)
        zero = bins == 0
        esum = torch.zeros_like(embed).index_add_(0, ind, flat)
        enorm = l2norm(esum / bins.masked_fill(zero, 1.0)[:, None])
        enorm = torch.where(zero[:, None], en, enorm)
        new_embed = embed * decay + enorm * (1 - decay)
    return quant, ind.reshape(shape[:-1]), new_embed, new_cs


def ctvit_forward(video, p: Params, *, patch: int, tpatch: int, spatial_depth: int, temporal_depth: int,
                  heads: int, vq: bool = True, return_pre_vq: bool = False):
    """ctvit.py:353-412 with return_encoded_tokens=True (eval-mode VQ)."""
    tokens = patch_embed(video, p, patch, tpatch)
    enc = ctvit_encode(tokens, p, spatial_depth, temporal_depth, heads)
    if not vq:
        return enc
    quant, ind, _, _ = vq_cosine(enc, p["vq._codebook.embed"][0] if p["vq._codebook.embed"].dim() == 3
                                 else p["vq._codebook.embed"])
    return (quant, enc, ind) if return_pre_vq else quant


# ------------------------------------------------------------------------------------------
# CTViT3D: joint 3-D transformer, ctvit3d.py:175-520
# ------------------------------------------------------------------------------------------
def flash_attention(x, p: Params, pre: str, heads: int):
    """attention.py:189-284 `FlashAttention` as CTViT3D uses it (no context / mask / causal; `attn_bias` is ignored,
    attention.py:257): q from LayerNorm(x), k and v from the raw x; learned null key/values are PREPENDED
    ('h (n r) d -> b h n r d', r = 2: even rows are keys, odd rows values), l2norm + q_scale / k_scale are applied
    AFTER the concat (so the null keys are normalised too), then F.scaled_dot_product_attention with its default
    scale 1/sqrt(dim_head) - not the cosine-attention scale 8 of `Attention`."""
    b, n, _ = x.shape
    xn = layer_norm(x, p[pre + "norm.gamma"])
    q = xn @ p[pre + "to_q.weight"].T
    k, v = (x @ p[pre + "to_kv.weight"].T).chunk(2, dim=-1)
    q, k, v = (t.reshape(b, n, heads, -1).permute(0, 2, 1, 3) for t in (q, k, v))
    null = p[pre + "null_kv"]                                            # (heads, 2 * n_null, dh)
    nk, nv = null.reshape(heads, -1, 2, null.shape[-1]).unbind(dim=-2)
    k = torch.cat([nk.expand(b, *nk.shape), k], dim=-2)
    v = torch.cat([nv.expand(b, *nv.shape), v], dim=-2)
    q = l2norm(q) * p[pre + "q_scale"]
    k = l2norm(k) * p[pre + "k_scale"]
    sim = torch.einsum("bhid,bhjd->bhij", q, k) / math.sqrt(q.shape[-1])
    out = torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v)
    out = out.permute(0, 2, 1, 3).reshape(b, n, -1)
    return out @ p[pre + "to_out.weight"].T


def sincos_pos_embed_3d(dim: int, grid):
    """ctvit3d.py:122-173 fixed 3-D sin/cos table (n_t*n_h*n_w, dim), float32 numpy semantics restated in torch
    float64 then cast.  Reproduces the reference's construction literally: np.meshgrid(t, w, h) in its default 'xy'
    indexing yields arrays of shape (n_w, n_t, n_h) that are then *reshaped* (not transposed) to (n_t, n_w, n_h)
    (ctvit3d.py:131-135), so unless n_t == n_w the three coordinate channels are a scrambled - but fixed and
    deterministic - function of the token index; a third of the channels encodes each of them as [sin | cos] with
    frequencies 10000^(-i / (dim/6))."""
    n_t, n_h, n_w = grid
    assert dim % 6 == 0
    gt = torch.arange(n_t, dtype=torch.float32)
    gh = torch.arange(n_h, dtype=torch.float32)
    gw = torch.arange(n_w, dtype=torch.float32)
    # numpy 'xy' meshgrid of (t, w, h): out[k][j, i, l] with x = t (index i), y = w (index j), z = h (index l)
    T = gt[None, :, None].expand(n_w, n_t, n_h)
    W = gw[:, None, None].expand(n_w, n_t, n_h)
    H = gh[None, None, :].expand(n_w, n_t, n_h)
    chans = [c.reshape(-1) for c in (T, W, H)]             # the reshape to (1, n_t, n_w, n_h) keeps memory order
    d3 = dim // 3
    omega = torch.arange(d3 // 2, dtype=torch.float32) / (d3 / 2.0)
    omega = 1.0 / 10000 ** omega
    parts = []
    for pos in chans:
        out = pos[:, None] * omega[None, :]
        parts += [torch.sin(out), torch.cos(out)]
    return torch.cat(parts, dim=1)


def ctvit3d_forward(video, p: Params, *, patch: int, tpatch: int, blocks: int, heads: int):
    """ctvit3d.py:455-520 with return_encoded_tokens=True: patch embedding (ctvit3d.py:240-245, same as CTViT),
    + pos_embed, `blocks` x [x = flash_attn(x) + x; x = ff(x) + x] over ALL t*h*w tokens jointly (no PEG,
    ctvit3d.py:247-258), final LayerNorm (attention.py:452); no vector quantisation."""
    tokens = patch_embed(video, p, patch, tpatch)
    b, t, h, w, d = tokens.shape
    x = tokens.reshape(b, t * h * w, d) + p["pos_embed"]
    for l in range(blocks):
        lp = f"enc_3D.layers.{l}."
        x = flash_attention(x, p, lp + "1.", heads) + x
        x = feed_forward(x, p, lp + "3.") + x
    x = layer_norm(x, p["enc_3D.norm_out.gamma"])
    return x.reshape(b, t, h, w, d)


# ------------------------------------------------------------------------------------------
# contrastive head
# ------------------------------------------------------------------------------------------
def image_latent(enc_image, W):
    """ct_clip.py:1280-1297,1316: project every token with to_visual_latent, mean over tokens, l2norm."""
    B = enc_image.shape[0]
    lat = (enc_image.reshape(-1, enc_image.shape[-1]) @ W.T).reshape(B, -1, W.shape[0]).mean(dim=1)
    return l2norm(lat)


def text_latent(enc_text, W):
    """ct_clip.py:1309-1316: CLS token -> to_text_latent -> l2norm."""
    return l2norm(enc_text[:, 0, :] @ W.T)


def clip_loss_reference_form(text_lat, image_lat, log_temp, bs_single_gpu: int):
    """ct_clip.py:1320-1382 on already-gathered latents, literally: sim * exp(temp) -> exp ->
    diagonal / row sums of both orientations -> -log(pos+1e-20) + log(denom+1e-20) -> mean -> /2 /bs."""
    temp = log_temp.exp()
    t2i = text_lat @ image_lat.T * temp
    i2t = t2i.T
    t2i_e, i2t_e = t2i.exp(), i2t.exp()
    t2i_pos, i2t_pos = torch.diagonal(t2i_e), torch.diagonal(i2t_e)
    t2i_den, i2t_den = t2i_e.sum(dim=-1), i2t_e.sum(dim=-1)
    t2i_loss = (-torch.log(t2i_pos + 1e-20) + torch.log(t2i_den + 1e-20)).mean()
    i2t_loss = (-torch.log(i2t_pos + 1e-20) + torch.log(i2t_den + 1e-20)).mean()
    return (t2i_loss + i2t_loss) / 2 / bs_single_gpu


def clip_loss_open_clip(text_lat, image_lat, logit_scale):
    """demo_tests/clip_loss.py:104-128 (world_size 1): mean of the two cross-entropies."""
    li = logit_scale * image_lat @ text_lat.T
    lt = li.T
    labels = torch.arange(li.shape[0])
    return (F.cross_entropy(li, labels) + F.cross_entropy(lt, labels)) / 2


def clip_loss_and_local_grads(T, I, log_temp, b_local: int, rank: int):
    """Loss plus what rank `rank` keeps after AllGather.backward (distributed.py:18-20: the local
    chunk of the gradient, no reduction) and the full d(log temperature)."""
    T = T.detach().clone().requires_grad_(True)
    I = I.detach().clone().requires_grad_(True)
    lt = log_temp.detach().clone().requires_grad_(True)
    loss = clip_loss_reference_form(T, I, lt, b_local)
    loss.backward()
    sl = slice(rank * b_local, (rank + 1) * b_local)
    return {"loss": loss.detach(), "dlog_temp": lt.grad, "dT_local": T.grad[sl], "dI_local": I.grad[sl]}


def pooled_latent_fwd_bwd(x, W, dlat):
    """Mean-pool first, then project (exactly equal to ct_clip.py:1290-1297 because the Linear is
    bias-free), l2norm; gradients by autograd."""
    x = x.detach().clone().requires_grad_(True)
    W = W.detach().clone().requires_grad_(True)
    pooled = x.mean(dim=1)
    pooled.retain_grad()
    lat = l2norm(pooled @ W.T)
    (lat * dlat).sum().backward()
    return {"pooled": pooled.detach(), "latent": lat.detach(), "dW": W.grad, "dpooled": pooled.grad, "dx": x.grad}


def ctclip_loss(enc_text, enc_image, p: Params, b_local: Optional[int] = None):
    """ct_clip.py:1252-1388 single-process head: latents -> loss (accelerator.gather == identity)."""
    tl = text_latent(enc_text, p["to_text_latent.weight"])
    il = image_latent(enc_image, p["to_visual_latent.weight"])
    return clip_loss_reference_form(tl, il, p["temperature"], b_local or tl.shape[0]), tl, il


def forward_infer_logits(text_lat, image_lat, log_temp):
    """ct_clip.py:842-855: einsum('b d, b d -> b') * exp(temperature) with text b broadcast on image b=1."""
    return (text_lat * image_lat).sum(dim=-1) * log_temp.exp()


# ------------------------------------------------------------------------------------------
# legacy pooling of the original CT-CLIP checkpoints: CTCLIP.forward_old
# ------------------------------------------------------------------------------------------
def image_embeds_legacy(enc_image):
    """ct_clip.py:1549,1566: mean over axis 1 of the encoded tokens (B, t, h, w, C), then flatten (h w C)."""
    return enc_image.mean(dim=1).reshape(enc_image.shape[0], -1)


def forward_old_latents(enc_text, enc_image, p: Params, text_valid_mask):
    """ct_clip.py:1583-1626: CLS row / legacy image embedding, rows selected by `text_valid_mask` (B, 1) BEFORE the
    projections (:1593-1594), to_text_latent / to_visual_latent (in_features = h*w*C), l2norm."""
    keep = text_valid_mask.squeeze(1).bool()
    text_embeds = enc_text[:, 0, :][keep, :]
    image_embeds = image_embeds_legacy(enc_image)[keep, :]
    return l2norm(text_embeds @ p["to_text_latent.weight"].T), l2norm(image_embeds @ p["to_visual_latent.weight"].T)


def forward_old_similarity(enc_text, enc_image, p: Params, text_valid_mask):
    """ct_clip.py:1655-1657 (return_loss=False): einsum('b d, b d -> b') * exp(temperature) over the valid rows."""
    tl, il = forward_old_latents(enc_text, enc_image, p, text_valid_mask)
    return (tl * il).sum(dim=-1) * p["temperature"].exp()


def forward_old_loss(enc_text, enc_image, p: Params, text_valid_mask):
    """ct_clip.py:1661-1768 single process, no multiview / MLM / SSL terms (weights 0 -> cl_loss_weight 1,
    :1757-1761) and seg_loss 0: the same symmetric loss as the new forward over the VALID rows, divided by their
    count (`bs_single_gpu = text_latents.shape[0]` is taken after the selection, :1661)."""
    tl, il = forward_old_latents(enc_text, enc_image, p, text_valid_mask)
    return clip_loss_reference_form(tl, il, p["temperature"], tl.shape[0]), tl, il
