"""Host logic of the libctk text tower (vit_exp_b200/text_tower.py) against the module it replaces.

The kernels are replaced by the torch doubles of tests/emulated_ops.py, so what is checked here is the
orchestration: operand routing, the hand-derived backward chain of BertEmbeddings / BertLayer, mask handling and
gradient bookkeeping, compared with HF `BertModel` (the reference's text encoder, ct_clip.py:1271) under autograd.
The kernels themselves are checked on the GPU (tests/test_text_tower_gpu.py).
"""
import pytest
import torch

import emulated_ops
from vit_exp_b200 import text_tower


def _tiny_bert(seed=0, layers=2, hidden=128, heads=2, inter=256):
    from transformers import BertConfig, BertModel
    torch.manual_seed(seed)
    cfg = BertConfig(vocab_size=97, hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads,
                     intermediate_size=inter, max_position_embeddings=40, hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0)
    bert = BertModel(cfg)
    with torch.no_grad():        # random biases / LayerNorm affine so that every gradient path is exercised
        for n, p in bert.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.05)
            elif "LayerNorm.weight" in n:
                p.add_(torch.randn_like(p) * 0.1)
    return bert.train()


def _inputs(B=3, L=16, padded=True, seed=1):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, 97, (B, L), generator=g)
    ids[1, 3] = 0                                  # the pad token id: nn.Embedding(padding_idx=0) gives its row no gradient
    mask = torch.ones(B, L, dtype=torch.int64)
    if padded:
        mask[0, L - 5:] = 0
        mask[2, L - 1:] = 0
    return ids, mask


def _objective(hidden, seed=2):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(hidden.shape[-1], generator=g)
    v = torch.randn(hidden.shape, generator=g)
    return (hidden[:, 0, :] * w).sum() + 0.05 * (hidden * v).sum()      # CLS read (ct_clip.py:1313) + all tokens


@pytest.fixture
def emulated(monkeypatch):
    monkeypatch.setattr(text_tower, "ops", emulated_ops)
    emulated_ops.CALLS.clear()
    yield emulated_ops
    emulated_ops.OPERAND = torch.bfloat16
    text_tower.OPERAND_DTYPE = torch.bfloat16


def _set_operand(dtype):
    emulated_ops.OPERAND = dtype
    text_tower.OPERAND_DTYPE = dtype


def _rel(a, b):
    a, b = a.detach(), b.detach()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


@pytest.mark.parametrize("padded", [True, False])
def test_forward_backward_match_hf_exact_operands(emulated, padded):
    """fp32 operands: the host math (forward and every parameter gradient) must agree with HF to fp32 round-off."""
    _set_operand(torch.float32)
    bert = _tiny_bert()
    ids, mask = _inputs(padded=padded)
    ref = bert(ids, attention_mask=mask)[0]
    _objective(ref).backward()
    ref_grads = {n: p.grad.clone() for n, p in bert.named_parameters() if p.grad is not None}
    bert.zero_grad(set_to_none=True)

    out = text_tower.encode(bert, ids, mask)
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert _rel(out, ref.detach()) < 1e-5
    _objective(out).backward()
    for n, p in bert.named_parameters():
        if n.startswith("pooler"):
            assert p.grad is None, "the pooler is not on the path (ct_clip.py:1273 reads [0] only)"
            continue
        assert p.grad is not None, n
        if n.endswith("key.bias"):       # softmax is invariant to a key bias: the true gradient is 0 (round-off only)
            assert float(p.grad.abs().max()) < 1e-6
            continue
        assert _rel(p.grad, ref_grads[n]) < 2e-4, (n, _rel(p.grad, ref_grads[n]))


def test_bf16_operands_within_stated_tolerance(emulated):
    """bf16 operand copies, fp32 accumulation / residual stream (what the kernels do): hidden states within 2e-2
    relative L2, parameter gradients within 5e-2."""
    _set_operand(torch.bfloat16)
    bert = _tiny_bert(seed=3)
    ids, mask = _inputs(seed=4)
    ref = bert(ids, attention_mask=mask)[0]
    _objective(ref).backward()
    ref_grads = {n: p.grad.clone() for n, p in bert.named_parameters() if p.grad is not None}
    bert.zero_grad(set_to_none=True)
    out = text_tower.encode(bert, ids, mask)
    assert _rel(out, ref.detach()) < 2e-2
    _objective(out).backward()
    for n, g in ref_grads.items():
        if n.endswith("key.bias"):
            continue
        assert _rel(dict(bert.named_parameters())[n].grad, g) < 5e-2, n


def test_launch_plan_and_no_grad_path(emulated):
    """per layer: 4 forward GEMMs (QKV, out-proj+residual, intermediate+GELU, output+residual); nothing is saved and
    no transposed weights are built when no gradient is needed."""
    _set_operand(torch.float32)
    bert = _tiny_bert(layers=3)
    ids, mask = _inputs()
    with torch.no_grad():
        out = text_tower.encode(bert, ids, mask)
    assert not out.requires_grad
    epi = [e for name, e in emulated.CALLS if name == "gemm"]
    assert epi == [emulated.EPI_BF16, emulated.EPI_RESID_F32, emulated.EPI_GELU, emulated.EPI_RESID_F32] * 3
    assert not any(name == "transpose_cast_bf16" for name, _ in emulated.CALLS)
    # frozen tower (config fix_text_encoder, ct_clip.py:654-658): same values, no graph
    for p in bert.parameters():
        p.requires_grad = False
    out2 = text_tower.encode(bert.eval(), ids, mask)
    assert not out2.requires_grad and torch.equal(out, out2)


def test_token_type_ids_and_partial_requires_grad(emulated):
    _set_operand(torch.float32)
    bert = _tiny_bert(seed=5)
    ids, mask = _inputs(seed=6)
    tt = torch.zeros_like(ids)
    tt[:, 8:] = 1
    for p in bert.embeddings.parameters():           # frozen embeddings: those gradients must come back as None
        p.requires_grad = False
    ref = bert(ids, attention_mask=mask, token_type_ids=tt)[0]
    _objective(ref).backward()
    ref_g = bert.encoder.layer[0].attention.self.query.weight.grad.clone()
    bert.zero_grad(set_to_none=True)
    out = text_tower.encode(bert, ids, mask, tt)
    assert _rel(out, ref.detach()) < 1e-5
    _objective(out).backward()
    assert bert.embeddings.word_embeddings.weight.grad is None
    assert _rel(bert.encoder.layer[0].attention.self.query.weight.grad, ref_g) < 2e-4


def test_unsupported_configurations_are_refused():
    from transformers import BertConfig, BertModel
    odd = BertModel(BertConfig(vocab_size=50, hidden_size=96, num_hidden_layers=1, num_attention_heads=2,
                               intermediate_size=256, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0))
    assert text_tower.unsupported_reason(odd, training=True) is not None
    with pytest.raises(NotImplementedError):
        text_tower.encode(odd, torch.zeros(1, 4, dtype=torch.long))
    relu = BertModel(BertConfig(vocab_size=50, hidden_size=128, num_hidden_layers=1, num_attention_heads=2,
                                intermediate_size=256, hidden_act="relu"))
    assert "hidden_act" in text_tower.unsupported_reason(relu, training=False)
    assert text_tower.unsupported_reason(torch.nn.Linear(2, 2), training=False) == "not a BertModel"
    stock = BertModel(BertConfig(vocab_size=50, hidden_size=128, num_hidden_layers=1, num_attention_heads=2,
                                 intermediate_size=256))                     # HF defaults: dropout 0.1 / 0.1
    assert text_tower.unsupported_reason(stock, training=True) is None


def test_hidden_dropout_matches_hf_with_shared_masks(emulated, monkeypatch):
    """training mode, hidden_dropout_prob 0.2: with the SAME keep-masks fed to HF's nn.Dropout and to the tower (call
    order: embeddings, then per layer BertSelfOutput, BertOutput) outputs and every gradient agree - i.e. dropout sits
    at the reference's three places, scaled by 1 / (1 - p), and the residual branches bypass it."""
    _set_operand(torch.float32)
    from transformers import BertConfig, BertModel
    torch.manual_seed(11)
    bert = BertModel(BertConfig(vocab_size=97, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                intermediate_size=256, max_position_embeddings=40, hidden_dropout_prob=0.2,
                                attention_probs_dropout_prob=0.0)).train()
    calls = {"n": 0}

    def keep_mask(numel, p):
        calls["n"] += 1
        return torch.rand(numel, generator=torch.Generator().manual_seed(1000 + calls["n"])) >= p

    def f_dropout(x, p=0.5, training=True, inplace=False):
        if not training or p == 0:
            return x
        return x * keep_mask(x.numel(), p).view(x.shape) / (1 - p)

    def tower_dropout(x, p):
        m = keep_mask(x.numel(), p).view(x.shape)
        return x * m / (1 - p), m
    monkeypatch.setattr(torch.nn.functional, "dropout", f_dropout)
    monkeypatch.setattr(text_tower, "_dropout", tower_dropout)
    ids, mask = _inputs(seed=12)
    calls["n"] = 0
    ref = bert(ids, attention_mask=mask)[0]
    assert calls["n"] == 1 + 2 * 2
    _objective(ref).backward()
    ref_grads = {n: p.grad.clone() for n, p in bert.named_parameters() if p.grad is not None}
    bert.zero_grad(set_to_none=True)
    calls["n"] = 0
    out = text_tower.encode(bert, ids, mask)
    assert calls["n"] == 5 and _rel(out, ref.detach()) < 1e-5
    _objective(out).backward()
    for n, g in ref_grads.items():
        if n.endswith("key.bias"):
            continue
        assert _rel(dict(bert.named_parameters())[n].grad, g) < 2e-4, n
    # eval mode: no dropout call at all
    calls["n"] = 0
    with torch.no_grad():
        text_tower.encode(bert.eval(), ids, mask)
    assert calls["n"] == 0


def test_attention_dropout_is_passed_to_the_attention_core(emulated):
    """training mode hands attention_probs_dropout_prob and a device-resident seed to ctk_mha_fwd / ctk_mha_bwd (the
    backward regenerates the same mask: same seed tensor, same per-layer offset); eval mode passes 0."""
    _set_operand(torch.float32)
    from transformers import BertConfig, BertModel
    bert = BertModel(BertConfig(vocab_size=97, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                intermediate_size=256, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.3)).train()
    ids, mask = _inputs()
    out = text_tower.encode(bert, ids, mask)
    _objective(out).backward()
    assert all(torch.isfinite(p.grad).all() for n, p in bert.named_parameters() if p.grad is not None)
    with torch.no_grad():
        text_tower.encode(bert.eval(), ids, mask)
    seen = [(n, p) for n, p in emulated.CALLS if n.startswith("mha_")]
    assert seen == [("mha_fwd", 0.3)] * 2 + [("mha_bwd", 0.3)] * 2 + [("mha_fwd", 0.0)] * 2
    # two training forwards draw different masks (the device counter advances)
    a = text_tower.encode(bert.train(), ids, mask)
    b = text_tower.encode(bert, ids, mask)
    assert not torch.equal(a, b)


def test_attention_dropout_gradients_match_autograd(emulated):
    """with the kernels' hash mask restated in torch, the tower's hand-written backward equals autograd through the
    same masked attention (dropout sits on the probabilities, after the softmax, scaled by 1 / (1 - p))"""
    _set_operand(torch.float32)
    from transformers import BertConfig, BertModel
    torch.manual_seed(3)
    bert = BertModel(BertConfig(vocab_size=97, hidden_size=128, num_hidden_layers=1, num_attention_heads=2,
                                intermediate_size=256, max_position_embeddings=40, hidden_dropout_prob=0.0,
                                attention_probs_dropout_prob=0.25)).train()
    ids, mask = _inputs()
    text_tower._SEED.clear()
    torch.manual_seed(7)
    out = text_tower.encode(bert, ids, mask)
    _objective(out).backward()
    got = bert.encoder.layer[0].attention.self.value.weight.grad.clone()
    bert.zero_grad(set_to_none=True)
    # reference: HF module with its attention dropout replaced by the same hash mask
    seed0 = int(next(iter(text_tower._SEED.values())).item()) - 1
    keep = emulated.mha_keep_mask(emulated.mha_seed(seed0, 0), ids.shape[0] * 2, ids.shape[1], 0.25)
    keep = keep.view(ids.shape[0], 2, ids.shape[1], ids.shape[1])
    att = bert.encoder.layer[0].attention.self
    x = bert.embeddings(input_ids=ids)
    q, k, v = (lin(x).view(ids.shape[0], ids.shape[1], 2, 64).transpose(1, 2) for lin in (att.query, att.key, att.value))
    sc = (q @ k.transpose(-1, -2)) / 8.0
    sc = sc.masked_fill(mask.view(ids.shape[0], 1, 1, -1) == 0, float("-inf"))
    ctx = ((torch.softmax(sc, -1) * keep / 0.75) @ v).transpose(1, 2).reshape(ids.shape[0], ids.shape[1], 128)
    layer = bert.encoder.layer[0]
    h1 = layer.attention.output.LayerNorm(layer.attention.output.dense(ctx) + x)
    ref = layer.output.LayerNorm(layer.output.dense(layer.intermediate(h1)) + h1)
    assert _rel(out, ref.detach()) < 1e-5
    _objective(ref).backward()
    assert _rel(got, att.value.weight.grad) < 2e-4


def test_product_path_has_no_cpu_fallback():
    """without the doubles the real ops run: CPU tensors must be refused, not silently computed."""
    bert = _tiny_bert()
    ids, mask = _inputs()
    with pytest.raises((AssertionError, RuntimeError, OSError)):
        text_tower.encode(bert, ids, mask)
