"""Text tower on libctk: the forward and hand-written backward of the HF `BertModel` the
reference passes as `text_encoder` (CT_CLIP/ct_clip/ct_clip.py:1271 `self.text_transformer(text.input_ids,
attention_mask=text.attention_mask)`; constructed in scripts/run_train.py:2,143-154).

The caller's `BertModel` stays the parameter holder (its state-dict keys are the checkpoint format); this module
only *reads its parameters* and runs the arithmetic of `BertEmbeddings`, `BertLayer` x depth (post-LayerNorm:
attention -> dense + residual -> LayerNorm -> intermediate dense + erf GELU -> dense + residual -> LayerNorm)
through the same tcgen05 GEMM (`ctk_gemm_bf16`, epilogues BF16 / RESID_F32 / F32 / GELU / GELU_BWD / ATOMIC_F32) and
LayerNorm kernels the image encoder uses (bias gradients are dY^T 1 products on the same GEMM).  The pooler is not evaluated (CTCLIP reads `[0][:, 0, :]` only,
ct_clip.py:1273,1313), so `pooler.dense.*` receives no gradient, exactly as in the reference.  Dropout (training
mode, hidden 0.1 / attention 0.1 in CXR-BERT): the three hidden dropouts use torch's dropout kernel (the residual add
then leaves the GEMM epilogue), the attention-probability dropout is SDPA's `dropout_p`; masks are statistically,
not bitwise, those of the stock module.

Data layout: M = B*L token rows, hidden H; fp32 residual stream, bf16 GEMM operands written by the producing
kernel, packed `qkv` bf16 [M, 3H] (q | k | v, heads contiguous inside each third).

Attention core (softmax(q k^T / sqrt(dh) + key mask) v, head dim 64): `torch.nn.functional.
scaled_dot_product_attention` on strided views of the packed buffer - library code, like the HF module's own
`sdpa` path; an in-tree d = 64 kernel is the next step.

STATUS: opt-in (`config["ctk_text_tower"]` / `CTK_TEXT_TOWER=1`).  The host logic is checked on CPU against HF
autograd with emulated kernels (tests/test_text_tower_cpu.py); the GELU epilogues have not run on hardware yet.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import ops

OPERAND_DTYPE = torch.bfloat16      # tests set this to float32 to check the host math exactly on CPU
N_EMB = 5
N_PER_LAYER = 16


def unsupported_reason(bert, training: bool) -> Optional[str]:
    """None when `bert` is a BertModel this path reproduces; otherwise why not."""
    cfg = getattr(bert, "config", None)
    if cfg is None or not all(hasattr(bert, a) for a in ("embeddings", "encoder")):
        return "not a BertModel"
    if getattr(cfg, "hidden_act", "gelu") != "gelu":
        return f"hidden_act {cfg.hidden_act!r} (only the erf GELU is fused)"
    if getattr(cfg, "position_embedding_type", "absolute") not in (None, "absolute"):
        return "relative position embeddings"
    if getattr(cfg, "is_decoder", False) or getattr(cfg, "add_cross_attention", False):
        return "decoder / cross-attention configuration"
    H, heads = cfg.hidden_size, cfg.num_attention_heads
    if H % 64 or H > 1024 or cfg.intermediate_size % 32 or H % heads:
        return f"hidden size {H} / intermediate {cfg.intermediate_size}: LayerNorm needs H % 64 == 0, H <= 1024"
    return None


def flat_params(bert) -> List[torch.Tensor]:
    e = bert.embeddings
    out = [e.word_embeddings.weight, e.position_embeddings.weight, e.token_type_embeddings.weight,
           e.LayerNorm.weight, e.LayerNorm.bias]
    for layer in bert.encoder.layer:
        a, o = layer.attention.self, layer.attention.output
        out += [a.query.weight, a.query.bias, a.key.weight, a.key.bias, a.value.weight, a.value.bias,
                o.dense.weight, o.dense.bias, o.LayerNorm.weight, o.LayerNorm.bias,
                layer.intermediate.dense.weight, layer.intermediate.dense.bias,
                layer.output.dense.weight, layer.output.dense.bias, layer.output.LayerNorm.weight,
                layer.output.LayerNorm.bias]
    return out


class _Shape:
    def __init__(self, bert, input_ids):
        cfg = bert.config
        self.B, self.L = input_ids.shape
        self.M = self.B * self.L
        self.H = cfg.hidden_size
        self.heads = cfg.num_attention_heads
        self.dh = self.H // self.heads
        self.I = cfg.intermediate_size
        self.depth = len(bert.encoder.layer)
        self.eps = float(cfg.layer_norm_eps)
        # BertEmbeddings / BertSelfOutput / BertOutput dropout and the attention-probability dropout: active in
        # training mode only (CXR-BERT ships with 0.1 / 0.1)
        self.pad_idx = bert.embeddings.word_embeddings.padding_idx      # nn.Embedding(padding_idx=pad_token_id)
        self.p_hidden = float(cfg.hidden_dropout_prob) if bert.training else 0.0
        self.p_attn = float(cfg.attention_probs_dropout_prob) if bert.training else 0.0


def _operand(w: torch.Tensor) -> torch.Tensor:
    return ops.cast_bf16(w.contiguous())


def _operand_t(w: torch.Tensor) -> torch.Tensor:
    return ops.transpose_cast_bf16(w.contiguous())


def _dropout(x: torch.Tensor, p: float):
    """(x * mask / (1 - p), mask) - torch's own dropout kernel and Philox stream, like nn.Dropout in the HF module"""
    return torch.native_dropout(x, p, True)


def _dropout_bwd(g: torch.Tensor, mask: torch.Tensor, p: float) -> torch.Tensor:
    return g * mask * (1.0 / (1.0 - p))


def _prep_layer(lp: List[torch.Tensor], need_bwd: bool) -> Dict[str, torch.Tensor]:
    (wq, bq, wk, bk, wv, bv, wo, bo, g1, b1, wi, bi, wo2, bo2, g2, b2) = lp
    wqkv = torch.cat([wq, wk, wv], dim=0)             # [3H, H] fp32: one projection GEMM instead of three
    d = dict(wqkv=_operand(wqkv), bqkv=torch.cat([bq, bk, bv]).float(), wo=_operand(wo), wi=_operand(wi),
             wo2=_operand(wo2))
    if need_bwd:
        d.update(wqkv_t=_operand_t(wqkv), wo_t=_operand_t(wo), wi_t=_operand_t(wi), wo2_t=_operand_t(wo2))
    return d


def _attention(qkv: torch.Tensor, s: _Shape, key_mask: Optional[torch.Tensor], need_bwd: bool):
    """qkv [M, 3H] -> (context [M, H], closure for the backward).  BertSelfAttention: scale 1/sqrt(dh), additive
    key-padding mask broadcast over heads and queries."""
    q5 = qkv.view(s.B, s.L, 3, s.heads, s.dh)
    q, k, v = (q5[:, :, i].transpose(1, 2) for i in range(3))            # [B, heads, L, dh] views
    mask4 = None if key_mask is None else key_mask.view(s.B, 1, 1, s.L)
    if need_bwd:
        with torch.enable_grad():
            q, k, v = (t.detach().requires_grad_(True) for t in (q, k, v))
            o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask4, dropout_p=s.p_attn, scale=s.dh ** -0.5)
    else:
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask4, dropout_p=s.p_attn, scale=s.dh ** -0.5)
    ctx = o.detach().transpose(1, 2).reshape(s.M, s.H).contiguous()
    return ctx, ((q, k, v, o) if need_bwd else None)


def _attention_bwd(saved, dctx: torch.Tensor, s: _Shape) -> torch.Tensor:
    q, k, v, o = saved
    do = dctx.view(s.B, s.L, s.heads, s.dh).transpose(1, 2)
    dq, dk, dv = torch.autograd.grad(o, (q, k, v), do)
    dqkv = torch.empty(s.B, s.L, 3, s.heads, s.dh, dtype=dctx.dtype, device=dctx.device)
    for i, t in enumerate((dq, dk, dv)):
        dqkv[:, :, i].copy_(t.transpose(1, 2))
    return dqkv.view(s.M, 3 * s.H)


def _key_mask(attention_mask: Optional[torch.Tensor], s: _Shape) -> Optional[torch.Tensor]:
    """bool [B, L] (True = attend) or None when the caller passed no mask.  The tokenizer pads reports to 512
    (CTCLIPTrainer.py:562), so real batches do carry padding; the mask is always handed to the attention kernel - no
    host-side 'is it all ones' shortcut (that needs a device sync per batch, and caching its answer by tensor address
    is wrong as soon as the allocator reuses the address)."""
    if attention_mask is None:
        return None
    return attention_mask.to(torch.bool).reshape(s.B, s.L)


def _forward(s: _Shape, params: List[torch.Tensor], input_ids, token_type_ids, key_mask, save: bool):
    word, pos, typ, ge, be = params[:N_EMB]
    od = OPERAND_DTYPE
    dev = input_ids.device
    M, H, I = s.M, s.H, s.I
    # BertEmbeddings: word + position (absolute, 0..L-1) + token type, LayerNorm, dropout
    e = F.embedding(input_ids, word)
    e = e + pos[: s.L].unsqueeze(0)
    e = e + (typ[0] if token_type_ids is None else F.embedding(token_type_ids, typ))
    e = e.reshape(M, H).float().contiguous()
    ph = s.p_hidden
    m0 = None
    if ph > 0:
        _, x, _, mue, rse = ops.layernorm_fwd(e, ge, be, want_bf16=False, want_f32=True, eps=s.eps)
        x, m0 = _dropout(x, ph)
        xb = ops.cast_bf16(x)
    else:
        xb, x, _, mue, rse = ops.layernorm_fwd(e, ge, be, want_bf16=True, want_f32=True, eps=s.eps)
    saved = []
    for li in range(s.depth):
        lp = params[N_EMB + li * N_PER_LAYER: N_EMB + (li + 1) * N_PER_LAYER]
        (wq, bq, wk, bk, wv, bv, wo, bo, g1, b1, wi, bi, wo2, bo2, g2, b2) = lp
        w = _prep_layer(lp, save)
        qkv = torch.empty(M, 3 * H, dtype=od, device=dev)
        ops.gemm(xb, w["wqkv"], ops.EPI_BF16, qkv, M=M, N=3 * H, K=H, bias=w["bqkv"])
        ctxb, attn_saved = _attention(qkv, s, key_mask, save)
        y1 = torch.empty(M, H, dtype=torch.float32, device=dev)
        m1 = m2 = None
        if ph > 0:      # BertSelfOutput: dense -> dropout -> + input (the residual add leaves the epilogue)
            ops.gemm(ctxb, w["wo"], ops.EPI_F32, y1, M=M, N=H, K=H, bias=bo)
            y1, m1 = _dropout(y1, ph)
            y1 = y1.add_(x)
        else:
            ops.gemm(ctxb, w["wo"], ops.EPI_RESID_F32, y1, M=M, N=H, K=H, bias=bo, resid=x)  # BertSelfOutput
        x1b, x1, _, mu1, rs1 = ops.layernorm_fwd(y1, g1, b1, want_bf16=True, want_f32=True, eps=s.eps)
        U = torch.empty(M, I, dtype=od, device=dev)
        G = torch.empty(M, I, dtype=od, device=dev)
        ops.gemm(x1b, w["wi"], ops.EPI_GELU, U, M=M, N=I, K=H, bias=bi, aux0=G, ld_aux0=I)       # BertIntermediate
        y2 = torch.empty(M, H, dtype=torch.float32, device=dev)
        if ph > 0:      # BertOutput: dense -> dropout -> + input
            ops.gemm(G, w["wo2"], ops.EPI_F32, y2, M=M, N=H, K=I, bias=bo2)
            y2, m2 = _dropout(y2, ph)
            y2 = y2.add_(x1)
        else:
            ops.gemm(G, w["wo2"], ops.EPI_RESID_F32, y2, M=M, N=H, K=I, bias=bo2, resid=x1)     # BertOutput
        x2b, x2, _, mu2, rs2 = ops.layernorm_fwd(y2, g2, b2, want_bf16=True, want_f32=True, eps=s.eps)
        if save:
            saved.append(dict(w=w, xb=xb, attn=attn_saved, ctxb=ctxb, y1=y1, mu1=mu1, rs1=rs1, x1b=x1b, U=U, G=G,
                              y2=y2, mu2=mu2, rs2=rs2, m1=m1, m2=m2))
        x, xb = x2, x2b
    emb_saved = dict(e=e, mu=mue, rs=rse, m0=m0) if save else None
    return x.view(s.B, s.L, H), saved, emb_saved


def _backward(s: _Shape, params: List[torch.Tensor], input_ids, token_type_ids, saved, emb_saved, dout):
    od = OPERAND_DTYPE
    M, H, I = s.M, s.H, s.I
    dev = dout.device
    f32 = dict(dtype=torch.float32, device=dev)
    grads: List[Optional[torch.Tensor]] = [None] * len(params)
    g = dout.reshape(M, H).float().contiguous()                  # d(last_hidden_state)
    ones = torch.ones(M, 8, dtype=od, device=dev)

    def bias_grad(dyb: torch.Tensor, n: int) -> torch.Tensor:
        """column sums of dY [M, n] as a token-contraction on the tensor cores, dY^T 1: the MN-major split-K product
        of the weight gradients with a ones operand streams dY through TMA at full rate (a plain column-sum kernel
        keeps too few loads in flight for [4096, 3072] matrices)."""
        out = torch.zeros(n, 8, **f32)
        ops.gemm(dyb, ones, ops.EPI_ATOMIC_F32, out, M=n, N=8, K=M, mn_major=True, ldc=8)
        return out[:, 0]
    for li in range(s.depth - 1, -1, -1):
        base = N_EMB + li * N_PER_LAYER
        (wq, bq, wk, bk, wv, bv, wo, bo, g1, b1, wi, bi, wo2, bo2, g2, b2) = params[base: base + N_PER_LAYER]
        sv = saved[li]
        w = sv["w"]
        # ---- x2 = LN(y2), y2 = G Wo2^T + bo2 + x1
        dg2, db2 = torch.zeros(H, **f32), torch.zeros(H, **f32)
        dy2b = torch.empty(M, H, dtype=od, device=dev)
        dy2 = ops.layernorm_bwd(g, sv["y2"], g2, sv["mu2"], sv["rs2"], dg2, db2, dx_bf16=dy2b)
        if sv["m2"] is not None:         # the dense branch sees the masked gradient, the residual branch dy2 itself
            dy2b = ops.cast_bf16(_dropout_bwd(dy2, sv["m2"], s.p_hidden))
        dbo2 = bias_grad(dy2b, H)
        dwo2 = torch.zeros(H, I, **f32)
        ops.gemm(dy2b, sv["G"], ops.EPI_ATOMIC_F32, dwo2, M=H, N=I, K=M, mn_major=True, ldc=I)
        # ---- G = gelu(U), U = x1 Wi^T + bi
        dU = torch.empty(M, I, dtype=od, device=dev)
        ops.gemm(dy2b, w["wo2_t"], ops.EPI_GELU_BWD, dU, M=M, N=I, K=H, aux0=sv["U"], ld_aux0=I)
        dbi = bias_grad(dU, I)
        dwi = torch.zeros(I, H, **f32)
        ops.gemm(dU, sv["x1b"], ops.EPI_ATOMIC_F32, dwi, M=I, N=H, K=M, mn_major=True, ldc=H)
        dx1 = torch.empty(M, H, **f32)
        ops.gemm(dU, w["wi_t"], ops.EPI_RESID_F32, dx1, M=M, N=H, K=I, resid=dy2)      # + residual branch of y2
        # ---- x1 = LN(y1), y1 = ctx Wo^T + bo + x
        dg1, db1 = torch.zeros(H, **f32), torch.zeros(H, **f32)
        dy1b = torch.empty(M, H, dtype=od, device=dev)
        dy1 = ops.layernorm_bwd(dx1, sv["y1"], g1, sv["mu1"], sv["rs1"], dg1, db1, dx_bf16=dy1b)
        if sv["m1"] is not None:
            dy1b = ops.cast_bf16(_dropout_bwd(dy1, sv["m1"], s.p_hidden))
        dbo = bias_grad(dy1b, H)
        dwo = torch.zeros(H, H, **f32)
        ops.gemm(dy1b, sv["ctxb"], ops.EPI_ATOMIC_F32, dwo, M=H, N=H, K=M, mn_major=True, ldc=H)
        dctx = torch.empty(M, H, dtype=od, device=dev)
        ops.gemm(dy1b, w["wo_t"], ops.EPI_BF16, dctx, M=M, N=H, K=H)
        # ---- attention core and the packed q|k|v projection
        dqkv = _attention_bwd(sv["attn"], dctx, s)
        dbqkv = bias_grad(dqkv, 3 * H)
        dwqkv = torch.zeros(3 * H, H, **f32)
        ops.gemm(dqkv, sv["xb"], ops.EPI_ATOMIC_F32, dwqkv, M=3 * H, N=H, K=M, mn_major=True, ldc=H)
        gx = torch.empty(M, H, **f32)
        ops.gemm(dqkv, w["wqkv_t"], ops.EPI_RESID_F32, gx, M=M, N=H, K=3 * H, resid=dy1)  # + residual branch of y1
        g = gx
        grads[base: base + N_PER_LAYER] = [dwqkv[:H], dbqkv[:H], dwqkv[H:2 * H], dbqkv[H:2 * H], dwqkv[2 * H:],
                                           dbqkv[2 * H:], dwo, dbo, dg1, db1, dwi, dbi, dwo2, dbo2, dg2, db2]
        saved[li] = None
    # ---- embeddings
    word, pos, typ, ge, be = params[:N_EMB]
    dge, dbe = torch.zeros(H, **f32), torch.zeros(H, **f32)
    if emb_saved["m0"] is not None:
        g = _dropout_bwd(g, emb_saved["m0"], s.p_hidden)
    de = ops.layernorm_bwd(g, emb_saved["e"], ge, emb_saved["mu"], emb_saved["rs"], dge, dbe)
    dword = torch.zeros_like(word, dtype=torch.float32).index_add_(0, input_ids.reshape(-1), de)
    if s.pad_idx is not None:
        dword[s.pad_idx].zero_()                   # nn.Embedding never updates its padding row
    dpos = torch.zeros_like(pos, dtype=torch.float32)
    dpos[: s.L] = de.view(s.B, s.L, H).sum(0)
    dtyp = torch.zeros_like(typ, dtype=torch.float32)
    if token_type_ids is None:
        dtyp[0] = de.sum(0)
    else:
        dtyp.index_add_(0, token_type_ids.reshape(-1), de)
    grads[:N_EMB] = [dword, dpos, dtyp, dge, dbe]
    return grads


class _BertEncode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bert, input_ids, attention_mask, token_type_ids, *params):
        s = _Shape(bert, input_ids)
        need_bwd = any(ctx.needs_input_grad[4:])       # grad mode is off inside forward(); encode() checked it
        plist = [p.detach() for p in params]
        with torch.autocast(device_type=input_ids.device.type, enabled=False):
            out, saved, emb_saved = _forward(s, plist, input_ids, token_type_ids, _key_mask(attention_mask, s), need_bwd)
        ctx.state = (s, plist, input_ids, token_type_ids, saved, emb_saved) if need_bwd else None
        return out

    @staticmethod
    def forward_no_grad(bert, input_ids, attention_mask, token_type_ids, params):
        s = _Shape(bert, input_ids)
        with torch.autocast(device_type=input_ids.device.type, enabled=False):
            return _forward(s, [p.detach() for p in params], input_ids, token_type_ids,
                            _key_mask(attention_mask, s), False)[0]

    @staticmethod
    def backward(ctx, dout):
        assert ctx.state is not None, "text tower: backward without saved activations"
        s, plist, input_ids, token_type_ids, saved, emb_saved = ctx.state
        ctx.state = None
        with torch.autocast(device_type=dout.device.type, enabled=False):
            grads = _backward(s, plist, input_ids, token_type_ids, saved, emb_saved, dout)
        grads = [gr if need else None for gr, need in zip(grads, ctx.needs_input_grad[4:])]
        return (None, None, None, None, *grads)


def encode(bert, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
           token_type_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
    """last_hidden_state fp32 [B, L, H] of `bert` (a HF BertModel) through libctk; differentiable with respect to
    the module's parameters.  Raises if the configuration is one this path does not reproduce."""
    why = unsupported_reason(bert, bert.training)
    if why is not None:
        raise NotImplementedError(f"ctk text tower: {why}")
    assert input_ids.dim() == 2 and input_ids.shape[1] <= bert.config.max_position_embeddings
    params = flat_params(bert)
    if not (torch.is_grad_enabled() and any(p.requires_grad for p in params)):
        return _BertEncode.forward_no_grad(bert, input_ids, attention_mask, token_type_ids, params)
    return _BertEncode.apply(bert, input_ids, attention_mask, token_type_ids, *params)
