"""Text tower on libctk: the forward and hand-written backward of the HF `BertModel` the
reference passes as `text_encoder` (CT_CLIP/ct_clip/ct_clip.py:1271 `self.text_transformer(text.input_ids,
attention_mask=text.attention_mask)`; constructed in scripts/run_train.py:2,143-154).

The caller's `BertModel` stays the parameter holder (its state-dict keys are the checkpoint format); this module
only *reads its parameters* and runs the arithmetic of `BertEmbeddings`, `BertLayer` x depth (post-LayerNorm:
attention -> dense + residual -> LayerNorm -> intermediate dense + erf GELU -> dense + residual -> LayerNorm)
through libctk: `ctk_bert_embed_fwd/bwd` (table gathers / scatters), the tcgen05 GEMM (`ctk_gemm_bf16`, epilogues BF16 /
RESID_F32 / F32 / GELU / GELU_BWD / ATOMIC_F32; bias gradients are dY^T 1 products on the same GEMM), `ctk_layernorm_*`
and `ctk_mha_fwd/bwd` (softmax(q k^T / sqrt(dh) + key-padding mask) v, head dim 64, attention-probability dropout from a
counter-based hash whose seed lives in device memory).  The pooler is not evaluated (CTCLIP reads `[0][:, 0, :]` only,
ct_clip.py:1273,1313), so `pooler.dense.*` receives no gradient, exactly as in the reference.  The three hidden dropouts
(training mode, 0.1 in CXR-BERT) use torch's dropout kernel (the residual add then leaves the GEMM epilogue); masks are
statistically, not bitwise, those of the stock module.

Data layout: M = B*L token rows, hidden H; fp32 residual stream, bf16 GEMM operands written by the producing
kernel, packed `qkv` bf16 [M, 3H] (q | k | v, heads contiguous inside each third).

Launch path: a training-shape forward / backward is ~200 / ~330 kernel launches with static shapes; after two eager
calls with the same key (shapes, parameter addresses, dropout configuration) both are captured into CUDA graphs and
replayed (`_TowerGraph`, same mechanics as the image encoder's `_EncoderGraph`): the token ids / mask are copied into
static buffers, the caller receives a private copy of the hidden states.  `CTK_TEXT_GRAPHS=0` keeps eager launches.
"""
from __future__ import annotations

import os
import weakref
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
    if H // heads != 64:
        return f"head dim {H // heads} (ctk_mha_fwd is instantiated for 64, BERT-base / CXR-BERT)"
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
    H = wq.shape[0]
    # one projection GEMM instead of three: q | k | v weights cast straight into the row blocks of one bf16 operand
    wqkv = torch.empty(3 * H, H, dtype=OPERAND_DTYPE, device=wq.device)
    for i, w in enumerate((wq, wk, wv)):
        ops.cast_bf16(w.contiguous(), out=wqkv[i * H:(i + 1) * H])
    # the input-gradient products of the backward pass read these same [out, in] operands MN-major
    # (ops.gemm(b_mn_major=True)): no transposed weight copies
    return dict(wqkv=wqkv, bqkv=torch.cat([bq, bk, bv]).float(), wo=_operand(wo), wi=_operand(wi), wo2=_operand(wo2))


def _key_mask(attention_mask: Optional[torch.Tensor], s: _Shape) -> Optional[torch.Tensor]:
    """uint8 [B, L] (non-zero = attend) or None when the caller passed no mask.  The tokenizer pads reports to 512
    (CTCLIPTrainer.py:562), so real batches do carry padding; the mask is always handed to the attention kernel - no
    host-side 'is it all ones' shortcut (that needs a device sync per batch, and caching its answer by tensor address
    is wrong as soon as the allocator reuses the address)."""
    if attention_mask is None:
        return None
    return (attention_mask.reshape(s.B, s.L) != 0).to(torch.uint8).contiguous()


# attention-probability dropout: the kernels hash (seed, report*head, query, key); the seed is a device counter so that
# CUDA-graph replays draw fresh masks.  One counter per device, started from torch's CPU generator (torch.manual_seed).
_SEED: Dict[torch.device, torch.Tensor] = {}


def _next_seed(dev: torch.device) -> torch.Tensor:
    """int64 [1] device tensor holding this forward's seed (a private copy: the backward reads it again)"""
    st = _SEED.get(dev)
    if st is None:
        st = _SEED[dev] = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).to(dev)
    seed = st.clone()
    st.add_(1)
    return seed


def _forward(s: _Shape, params: List[torch.Tensor], input_ids, token_type_ids, key_mask, save: bool):
    """key_mask: uint8 [B, L] or None.  Returns (hidden fp32 [B, L, H], saved per layer, saved embeddings)."""
    word, pos, typ, ge, be = params[:N_EMB]
    od = OPERAND_DTYPE
    dev = input_ids.device
    M, H, I = s.M, s.H, s.I
    # BertEmbeddings: word + token type + position (absolute, 0..L-1), LayerNorm, dropout
    e = ops.bert_embed_fwd(input_ids.contiguous(), token_type_ids, word.contiguous(), pos.contiguous(), typ.contiguous())
    seed = _next_seed(dev) if s.p_attn > 0 else None
    scale = s.dh ** -0.5
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
        # BertSelfAttention core: scale 1/sqrt(dh), key-padding mask broadcast over heads and queries, dropout on the
        # probabilities (training mode)
        ctxb, lse = ops.mha_fwd(qkv, key_mask, s.B, s.L, s.heads, scale, s.p_attn, seed, li)
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
            saved.append(dict(w=w, xb=xb, qkv=qkv, lse=lse, ctxb=ctxb, y1=y1, mu1=mu1, rs1=rs1, x1b=x1b, U=U, G=G,
                              y2=y2, mu2=mu2, rs2=rs2, m1=m1, m2=m2))
        x, xb = x2, x2b
    emb_saved = dict(e=e, mu=mue, rs=rse, m0=m0, key_mask=key_mask, seed=seed) if save else None
    return x.view(s.B, s.L, H), saved, emb_saved


def _backward(s: _Shape, params: List[torch.Tensor], input_ids, token_type_ids, saved, emb_saved, dout):
    od = OPERAND_DTYPE
    M, H, I = s.M, s.H, s.I
    dev = dout.device
    f32 = dict(dtype=torch.float32, device=dev)
    grads: List[Optional[torch.Tensor]] = [None] * len(params)
    g = dout.reshape(M, H).float().contiguous()                  # d(last_hidden_state)
    ones = torch.ones(M, 8, dtype=od, device=dev)
    # zero-initialised accumulation targets (split-K weight gradients, LayerNorm parameter gradients, the [n, 8] bias
    # gradient products): slices of one pre-filled buffer instead of ~200 fill kernels
    arena = ops.ZeroArena(ops.ZeroArena.room(*(p.numel() for p in params[N_EMB:]),
                                             *([8 * H, 8 * I, 8 * H, 8 * 3 * H] * s.depth), H, H), dev)
    zeros = arena.take

    def bias_grad(dyb: torch.Tensor, n: int) -> torch.Tensor:
        """column sums of dY [M, n] as a token-contraction on the tensor cores, dY^T 1: the MN-major split-K product
        of the weight gradients with a ones operand streams dY through TMA at full rate (a plain column-sum kernel
        keeps too few loads in flight for [4096, 3072] matrices)."""
        out = zeros(n, 8)
        ops.gemm(dyb, ones, ops.EPI_ATOMIC_F32, out, M=n, N=8, K=M, mn_major=True, ldc=8)
        return out[:, 0]
    for li in range(s.depth - 1, -1, -1):
        base = N_EMB + li * N_PER_LAYER
        (wq, bq, wk, bk, wv, bv, wo, bo, g1, b1, wi, bi, wo2, bo2, g2, b2) = params[base: base + N_PER_LAYER]
        sv = saved[li]
        w = sv["w"]
        # ---- x2 = LN(y2), y2 = G Wo2^T + bo2 + x1
        dg2, db2 = zeros(H), zeros(H)
        dy2b = torch.empty(M, H, dtype=od, device=dev)
        dy2 = ops.layernorm_bwd(g, sv["y2"], g2, sv["mu2"], sv["rs2"], dg2, db2, dx_bf16=dy2b)
        if sv["m2"] is not None:         # the dense branch sees the masked gradient, the residual branch dy2 itself
            dy2b = ops.cast_bf16(_dropout_bwd(dy2, sv["m2"], s.p_hidden))
        dbo2 = bias_grad(dy2b, H)
        dwo2 = zeros(H, I)
        ops.gemm(dy2b, sv["G"], ops.EPI_ATOMIC_F32, dwo2, M=H, N=I, K=M, mn_major=True, ldc=I)
        # ---- G = gelu(U), U = x1 Wi^T + bi
        dU = torch.empty(M, I, dtype=od, device=dev)
        ops.gemm(dy2b, w["wo2"], ops.EPI_GELU_BWD, dU, M=M, N=I, K=H, aux0=sv["U"], ld_aux0=I, b_mn_major=True)
        dbi = bias_grad(dU, I)
        dwi = zeros(I, H)
        ops.gemm(dU, sv["x1b"], ops.EPI_ATOMIC_F32, dwi, M=I, N=H, K=M, mn_major=True, ldc=H)
        dx1 = torch.empty(M, H, **f32)
        ops.gemm(dU, w["wi"], ops.EPI_RESID_F32, dx1, M=M, N=H, K=I, resid=dy2, b_mn_major=True)  # + residual branch of y2
        # ---- x1 = LN(y1), y1 = ctx Wo^T + bo + x
        dg1, db1 = zeros(H), zeros(H)
        dy1b = torch.empty(M, H, dtype=od, device=dev)
        dy1 = ops.layernorm_bwd(dx1, sv["y1"], g1, sv["mu1"], sv["rs1"], dg1, db1, dx_bf16=dy1b)
        if sv["m1"] is not None:
            dy1b = ops.cast_bf16(_dropout_bwd(dy1, sv["m1"], s.p_hidden))
        dbo = bias_grad(dy1b, H)
        dwo = zeros(H, H)
        ops.gemm(dy1b, sv["ctxb"], ops.EPI_ATOMIC_F32, dwo, M=H, N=H, K=M, mn_major=True, ldc=H)
        dctx = torch.empty(M, H, dtype=od, device=dev)
        ops.gemm(dy1b, w["wo"], ops.EPI_BF16, dctx, M=M, N=H, K=H, b_mn_major=True)
        # ---- attention core and the packed q|k|v projection
        dqkv = ops.mha_bwd(sv["qkv"], emb_saved["key_mask"], sv["ctxb"], dctx, sv["lse"], s.B, s.L, s.heads, s.dh ** -0.5,
                           s.p_attn, emb_saved["seed"], li)
        dbqkv = bias_grad(dqkv, 3 * H)
        dwqkv = zeros(3 * H, H)
        ops.gemm(dqkv, sv["xb"], ops.EPI_ATOMIC_F32, dwqkv, M=3 * H, N=H, K=M, mn_major=True, ldc=H)
        gx = torch.empty(M, H, **f32)
        ops.gemm(dqkv, w["wqkv"], ops.EPI_RESID_F32, gx, M=M, N=H, K=3 * H, resid=dy1, b_mn_major=True)  # + residual of y1
        g = gx
        grads[base: base + N_PER_LAYER] = [dwqkv[:H], dbqkv[:H], dwqkv[H:2 * H], dbqkv[H:2 * H], dwqkv[2 * H:],
                                           dbqkv[2 * H:], dwo, dbo, dg1, db1, dwi, dbi, dwo2, dbo2, dg2, db2]
        saved[li] = None
    # ---- embeddings
    word, pos, typ, ge, be = params[:N_EMB]
    dge, dbe = zeros(H), zeros(H)
    if emb_saved["m0"] is not None:
        g = _dropout_bwd(g, emb_saved["m0"], s.p_hidden)
    de = ops.layernorm_bwd(g, emb_saved["e"], ge, emb_saved["mu"], emb_saved["rs"], dge, dbe)
    # nn.Embedding never updates its padding row (pad_idx); positions >= L and unused token types stay zero
    dword, dpos, dtyp = ops.bert_embed_bwd(de, input_ids.contiguous(), token_type_ids, word.shape, pos.shape, typ.shape,
                                           s.pad_idx)
    grads[:N_EMB] = [dword, dpos, dtyp, dge, dbe]
    return grads


# ------------------------------------------------------------------------------------------------
# CUDA-graph replay of the training-shape forward / backward
# ------------------------------------------------------------------------------------------------
class _TowerGraph:
    """Graphs of one (shape, parameter addresses, dropout configuration) of the tower.  Eagerly the tower is ~200
    forward + ~330 backward launches through ctypes (~15 ms of host time per step, which made the whole train step
    launch-bound on the host); replayed, it is two graph launches.  The token ids / mask / incoming gradient are
    copied into static buffers; saved activations live in the graphs' private pool."""
    WARMUP = 2

    def __init__(self):
        self.calls = 0
        self.fwd = None
        self.bwd = None
        self.pool = None
        self.failed = False
        self.owner = None        # weakref to the token of the forward whose activations sit in the static buffers


class _Token:
    __slots__ = ("done", "__weakref__")

    def __init__(self):
        self.done = False


def _graphs_enabled() -> bool:
    return os.environ.get("CTK_TEXT_GRAPHS", "1") != "0"


_CAPTURE_STREAM: Dict[torch.device, "torch.cuda.Stream"] = {}


def _capture_stream(dev: torch.device):
    """Graph kernel nodes keep the priority of the stream they were CAPTURED on.  The tower runs next to the image
    encoder's persistent kernels and must win the SMs that free up (ct_clip.py: _encode_both), so its graphs are
    captured on a high-priority stream (CTK_TEXT_STREAM_PRIORITY=0: default priority)."""
    st = _CAPTURE_STREAM.get(dev)
    if st is None:
        prio = -1 if os.environ.get("CTK_TEXT_STREAM_PRIORITY", "1") != "0" else 0
        st = _CAPTURE_STREAM[dev] = torch.cuda.Stream(device=dev, priority=prio)
    return st


def _graph_forward(bert, s: _Shape, plist, input_ids, attention_mask, token_type_ids):
    """Replay (capturing on first use) the forward graph for these inputs.  Returns None when this call must run on
    eager launches, else dict(out, eg, token)."""
    if not (input_ids.is_cuda and _graphs_enabled() and ops.GEMM_PROFILE is None
            and not torch.cuda.is_current_stream_capturing()):
        return None
    from .transformer_maskgit import _capture
    key = (s.B, s.L, attention_mask is not None, token_type_ids is not None, s.p_hidden, s.p_attn,
           tuple(p.data_ptr() for p in plist))
    graphs = bert.__dict__.setdefault("_ctk_graphs", {})
    eg = graphs.get(key)
    if eg is None:
        if len(graphs) >= 4:                       # shapes keep changing: stay eager
            graphs.clear()
        eg = graphs[key] = _TowerGraph()
    eg.calls += 1
    prev = eg.owner() if eg.owner is not None else None
    if eg.failed or eg.calls <= _TowerGraph.WARMUP or (prev is not None and not prev.done):
        return None
    if eg.fwd is None:
        try:
            dev = input_ids.device
            eg.pool = torch.cuda.graph_pool_handle()
            st = dict(ids=input_ids.clone(),
                      mask=None if attention_mask is None else _key_mask(attention_mask, s).clone(),
                      tt=None if token_type_ids is None else token_type_ids.clone())
            if s.p_attn > 0:
                _next_seed(dev)                    # create the device counter outside the capture
            g, outs, n = _capture(lambda: _forward(s, plist, st["ids"], st["tt"], st["mask"], True), eg.pool,
                                  _capture_stream(dev))
            eg.fwd = dict(graph=g, outs=outs, launches=n, st=st)
        except Exception as e:                                                  # pragma: no cover
            import warnings
            warnings.warn(f"ctk text tower: CUDA-graph capture failed ({e}); staying on eager launches")
            eg.failed = True
            torch.cuda.synchronize()
            return None
    f = eg.fwd
    st = f["st"]
    st["ids"].copy_(input_ids)
    if st["mask"] is not None:
        st["mask"].copy_(attention_mask.reshape(s.B, s.L) != 0)
    if st["tt"] is not None:
        st["tt"].copy_(token_type_ids)
    f["graph"].replay()
    ops.GRAPH_LAUNCHES += f["launches"]
    token = _Token()
    eg.owner = weakref.ref(token)
    return dict(out=f["outs"][0].clone(), eg=eg, token=token)     # private copy: the next replay rewrites the buffer


def _graph_backward(eg: _TowerGraph, token: _Token, s: _Shape, plist, dout):
    from .transformer_maskgit import _capture
    f = eg.fwd
    if eg.bwd is None:
        dstat = dout.reshape(s.B, s.L, s.H).float().contiguous().clone()
        _, saved, emb_saved = f["outs"]
        st = f["st"]
        # _backward drops its references to the saved activations layer by layer: hand it a copy of the list
        g, grads, n = _capture(lambda: _backward(s, plist, st["ids"], st["tt"], list(saved), emb_saved, dstat), eg.pool,
                               _capture_stream(dstat.device))
        eg.bwd = dict(graph=g, grads=grads, launches=n, dout=dstat)
    b = eg.bwd
    b["dout"].copy_(dout.reshape(s.B, s.L, s.H))
    b["graph"].replay()
    ops.GRAPH_LAUNCHES += b["launches"]
    token.done = True
    return b["grads"]


class _BertEncode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bert, input_ids, attention_mask, token_type_ids, *params):
        s = _Shape(bert, input_ids)
        need_bwd = any(ctx.needs_input_grad[4:])       # grad mode is off inside forward(); encode() checked it
        plist = [p.detach() for p in params]
        with torch.autocast(device_type=input_ids.device.type, enabled=False):
            launched = _graph_forward(bert, s, plist, input_ids, attention_mask, token_type_ids) if need_bwd else None
            if launched is not None:
                ctx.state = ("graph", launched["eg"], launched["token"], s, plist)
                return launched["out"]
            out, saved, emb_saved = _forward(s, plist, input_ids, token_type_ids, _key_mask(attention_mask, s), need_bwd)
        ctx.state = ("eager", s, plist, input_ids, token_type_ids, saved, emb_saved) if need_bwd else None
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
        state, ctx.state = ctx.state, None
        with torch.autocast(device_type=dout.device.type, enabled=False):
            if state[0] == "graph":
                _, eg, token, s, plist = state
                grads = _graph_backward(eg, token, s, plist, dout)
            else:
                _, s, plist, input_ids, token_type_ids, saved, emb_saved = state
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
