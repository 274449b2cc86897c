"""Host logic of the image encoder and the contrastive head on CPU: `_encode_forward` / `_encode_backward`
(vit_exp_b200/transformer_maskgit.py) and `_ClipHead` (vit_exp_b200/ct_clip.py) run with the libctk entry points
replaced by the torch doubles of tests/emulated_ops.py, and are compared with autograd through the pinned oracle
(oracle/ctclip_oracle.py).  This checks everything the Python side decides - operand routing, weight packing
(GEGLU interleave, LayerNorm(4000) folding), row permutations between the stacks, the hand-derived backward chain
and the order of the returned gradients - without a GPU.  The kernels are checked on the GPU suite.
"""
import pytest
import torch

import emulated_ops as E
from oracle import ctclip_oracle as O
from vit_exp_b200 import ct_clip as CC
from vit_exp_b200 import transformer_maskgit as TM


class _Ops:
    """namespace handed to the modules in place of vit_exp_b200.ops"""
    def __getattr__(self, name):
        full = {"gemm": E.gemm_full, "cast_bf16": E.cast_bf16_full, "transpose_cast_bf16": E.transpose_cast_bf16_full,
                "layernorm_fwd": E.layernorm_fwd_full, "layernorm_bwd": E.layernorm_bwd_full}
        return full[name] if name in full else getattr(E, name)


@pytest.fixture
def doubles(monkeypatch):
    o = _Ops()
    monkeypatch.setattr(TM, "ops", o)
    monkeypatch.setattr(CC, "ops", o)
    E.OPERAND = torch.float32
    _orig_empty = torch.empty

    def _empty(*a, **k):                       # the modules allocate bf16 operand buffers: fp32 in the exact-math mode
        if k.get("dtype") is torch.bfloat16:
            k["dtype"] = torch.float32
        return _orig_empty(*a, **k)
    monkeypatch.setattr(torch, "empty", _empty)
    _orig_zeros = torch.zeros

    def _zeros(*a, **k):
        if k.get("dtype") is torch.bfloat16:
            k["dtype"] = torch.float32
        return _orig_zeros(*a, **k)
    monkeypatch.setattr(torch, "zeros", _zeros)
    yield o
    E.OPERAND = torch.bfloat16


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _vit(seed=0):
    torch.manual_seed(seed)
    vit = TM.CTViT(dim=64, codebook_size=32, image_size=(8, 12), patch_size=(4, 4), temporal_patch_size=2,
                   spatial_depth=2, temporal_depth=2, dim_head=32, heads=2)
    with torch.no_grad():                    # move every parameter off its symmetric initial value
        for n, p in vit.named_parameters():
            if n.endswith(("gamma", "q_scale", "k_scale")) or ".0.weight" in n and p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.1)
            elif n.endswith("bias"):
                p.normal_(0, 0.05)
    return vit


def _oracle_params(vit):
    return {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in vit.state_dict().items()}


def _oracle_tokens(video, p, vit):
    return O.ctvit_forward(video, p, patch=vit.patch_size[0], tpatch=vit.temporal_patch_size,
                           spatial_depth=vit.spatial_depth, temporal_depth=vit.temporal_depth, heads=vit.heads, vq=False)


def test_encoder_forward_and_backward_match_oracle(doubles):
    vit = _vit()
    video = torch.rand(2, 1, 6, 8, 12, generator=torch.Generator().manual_seed(1))     # t,h,w = 3,2,3
    params = [p.detach() for p in vit._flat_params()]
    out, ind, pre_vq, ctx = TM._encode_forward(vit, video, params, save=True, training=False)
    p = _oracle_params(vit)
    ref = _oracle_tokens(video, p, vit)
    B, t, h, w, d = ref.shape
    assert pre_vq.shape == (B * t * h * w, d)
    assert _rel(pre_vq.view(B, t, h, w, d), ref) < 1e-4
    # VQ: straight-through output = nearest code (cosine) of the oracle's tokens
    q_ref, ind_ref, _, _ = O.vq_cosine(ref.detach(), vit.vq._codebook.embed[0])
    assert torch.equal(ind.reshape(-1), ind_ref.reshape(-1)) and _rel(out, q_ref) < 1e-6

    dtok = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2))
    (ref * dtok).sum().backward()
    grads = TM._encode_backward(vit, params, ctx, dtok)
    by_param = {id(q): n for n, q in vit.named_parameters()}
    assert len(grads) == len(params)
    for q, g in zip(vit._flat_params(), grads):
        n = by_param[id(q)]
        assert g is not None and tuple(g.shape) == tuple(q.shape), n
        if n == "spatial_rel_pos_bias.net.2.bias":      # a per-head constant shifts every logit of a row: true gradient 0
            assert float(g.abs().max()) < 1e-5
            continue
        assert _rel(g, p[n].grad) < 5e-4, (n, _rel(g, p[n].grad))


def test_mean_pool_gradient_is_consumed_unmaterialised(doubles):
    """_ClipHead hands the encoder a stride-0 expand of dpooled / n_tok; the result must equal the dense gradient."""
    vit = _vit(seed=3)
    video = torch.rand(2, 1, 6, 8, 12, generator=torch.Generator().manual_seed(4))
    params = [p.detach() for p in vit._flat_params()]
    dpool = torch.randn(2, 64, generator=torch.Generator().manual_seed(5))
    n_tok = 3 * 2 * 3
    dense = (dpool / n_tok).view(2, 1, 1, 1, 64).expand(2, 3, 2, 3, 64)
    _, _, _, ctx1 = TM._encode_forward(vit, video, params, save=True, training=False)
    g_bcast = TM._encode_backward(vit, params, ctx1, dense)                        # recognised as a broadcast
    _, _, _, ctx2 = TM._encode_forward(vit, video, params, save=True, training=False)
    g_dense = TM._encode_backward(vit, params, ctx2, dense.contiguous())
    for a, b in zip(g_bcast, g_dense):
        assert _rel(a, b) < 1e-4


def test_clip_head_matches_oracle(doubles):
    g = torch.Generator().manual_seed(6)
    B, n_tok, dim, dt, dl = 4, 18, 64, 48, 32
    tokens = torch.randn(B, 3, 2, 3, dim, generator=g).requires_grad_()
    cls = torch.randn(B, dt, generator=g).requires_grad_()
    wt = (torch.randn(dl, dt, generator=g) * 0.2).requires_grad_()
    wv = (torch.randn(dl, dim, generator=g) * 0.2).requires_grad_()
    temp = torch.tensor(1.3, requires_grad=True)
    acc = CC.TorchDistAccelerator()
    loss, tl, il = CC._ClipHead.apply(cls, tokens, wt, wv, temp, acc)
    loss.backward()
    got = [t.grad.clone() for t in (cls, tokens, wt, wv, temp)]
    for t in (cls, tokens, wt, wv, temp):
        t.grad = None
    enc_text = cls[:, None, :]
    ref, tl_ref, il_ref = O.ctclip_loss(enc_text, tokens, {"to_text_latent.weight": wt, "to_visual_latent.weight": wv,
                                                            "temperature": temp}, b_local=B)
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    assert _rel(tl, tl_ref) < 1e-5 and _rel(il, il_ref) < 1e-5
    for a, t in zip(got, (cls, tokens, wt, wv, temp)):
        assert _rel(a, t.grad) < 1e-4


def test_training_forward_updates_codebook_like_oracle(doubles):
    """training-mode VQ (ctvit.py:403): EMA of cluster sizes and code vectors, decay 0.8 (parity unpinned: oracle header)."""
    vit = _vit(seed=7)
    video = torch.rand(2, 1, 6, 8, 12, generator=torch.Generator().manual_seed(8))
    params = [p.detach() for p in vit._flat_params()]
    embed0 = vit.vq._codebook.embed[0].clone()
    cs0 = vit.vq._codebook.cluster_size[0].clone()
    _, ind, pre_vq, _ = TM._encode_forward(vit, video, params, save=False, training=True)
    _, ind_ref, embed_ref, cs_ref = O.vq_cosine(pre_vq, embed0, training=True, cluster_size=cs0, decay=vit.vq.decay)
    assert torch.equal(ind.reshape(-1), ind_ref.reshape(-1))
    assert _rel(vit.vq._codebook.cluster_size[0], cs_ref) < 1e-6
    assert _rel(vit.vq._codebook.embed[0], embed_ref) < 1e-5


def test_bf16_operands_stay_within_the_stated_tolerances(monkeypatch):
    """bf16 operand copies / fp32 accumulation and residual stream, as the kernels compute: pre-VQ tokens within 2e-2
    relative L2 of the fp32 oracle, parameter gradients within 5e-2."""
    o = _Ops()
    monkeypatch.setattr(TM, "ops", o)
    E.OPERAND = torch.bfloat16
    vit = _vit(seed=9)
    video = torch.rand(2, 1, 6, 8, 12, generator=torch.Generator().manual_seed(10))
    params = [p.detach() for p in vit._flat_params()]
    _, _, pre_vq, ctx = TM._encode_forward(vit, video, params, save=True, training=False)
    p = _oracle_params(vit)
    ref = _oracle_tokens(video, p, vit)
    assert _rel(pre_vq.view(ref.shape), ref) < 2e-2
    dtok = torch.randn(ref.shape, generator=torch.Generator().manual_seed(11))
    (ref * dtok).sum().backward()
    grads = TM._encode_backward(vit, params, ctx, dtok)
    by_param = {id(q): n for n, q in vit.named_parameters()}
    for q, g in zip(vit._flat_params(), grads):
        n = by_param[id(q)]
        if n == "spatial_rel_pos_bias.net.2.bias":
            continue
        assert _rel(g, p[n].grad) < 5e-2, (n, _rel(g, p[n].grad))
