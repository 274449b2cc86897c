"""libctk text tower on the GPU: the GELU GEMM epilogues and the full BertModel forward/backward against HF fp32.

Validated on a B200 in round 2.
Tolerances: bf16 operands / fp32 accumulation -> hidden states 2e-2 relative L2, parameter
gradients 5e-2.
"""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, dev, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev)


def _rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def _gelu(u):
    return 0.5 * u * (1.0 + torch.erf(u / math.sqrt(2.0)))


def _gelu_grad(u):
    return 0.5 * (1.0 + torch.erf(u / math.sqrt(2.0))) + u * torch.exp(-0.5 * u * u) / math.sqrt(2.0 * math.pi)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 352, 192), (4096, 3072, 768), (512, 96, 128)])
def test_gemm_gelu_epilogue(cuda_dev, M, N, K):
    from vit_exp_b200 import ops
    a = _rand((M, K), cuda_dev, 1, 0.5).bfloat16()
    b = _rand((N, K), cuda_dev, 2, 0.2).bfloat16()
    bias = _rand((N,), cuda_dev, 3)
    U = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    G = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    ops.gemm(a, b, ops.EPI_GELU, U, M=M, N=N, K=K, bias=bias, aux0=G, ld_aux0=N)
    u = a.float() @ b.float().T + bias
    assert torch.isfinite(U.float()).all() and torch.isfinite(G.float()).all()
    assert _rel(U, u) < 4e-3 and _rel(G, _gelu(u)) < 4e-3


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 352, 192), (4096, 3072, 768)])
def test_gemm_gelu_bwd_epilogue(cuda_dev, M, N, K):
    from vit_exp_b200 import ops
    a = _rand((M, K), cuda_dev, 4, 0.5).bfloat16()          # dY
    b = _rand((N, K), cuda_dev, 5, 0.2).bfloat16()          # W^T rows
    U = _rand((M, N), cuda_dev, 6, 1.5).bfloat16()
    dU = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    ops.gemm(a, b, ops.EPI_GELU_BWD, dU, M=M, N=N, K=K, aux0=U, ld_aux0=N)
    ref = (a.float() @ b.float().T) * _gelu_grad(U.float())
    assert torch.isfinite(dU.float()).all()
    assert _rel(dU, ref) < 4e-3


def _bert(dev, hidden, heads, layers, inter, seed=0):
    from transformers import BertConfig, BertModel
    torch.manual_seed(seed)
    bert = BertModel(BertConfig(vocab_size=30522, hidden_size=hidden, num_hidden_layers=layers,
                                num_attention_heads=heads, intermediate_size=inter, hidden_dropout_prob=0.0,
                                attention_probs_dropout_prob=0.0))
    with torch.no_grad():
        for n, p in bert.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.05)
    return bert.to(dev).train()


@pytest.mark.parametrize("hidden,heads,layers,inter,B,L,padded", [(256, 4, 2, 1024, 2, 128, True),
                                                                  (768, 12, 2, 3072, 2, 512, False),
                                                                  (768, 12, 1, 3072, 3, 200, True)])
def test_text_tower_matches_hf(cuda_dev, hidden, heads, layers, inter, B, L, padded):
    from vit_exp_b200 import text_tower
    bert = _bert(cuda_dev, hidden, heads, layers, inter)
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(0, 30522, (B, L), generator=g).to(cuda_dev)
    mask = torch.ones(B, L, dtype=torch.int64, device=cuda_dev)
    if padded:
        mask[0, L - L // 3:] = 0
    w = _rand((hidden,), cuda_dev, 8)

    ref = bert(ids, attention_mask=mask)[0]
    (ref[:, 0, :] * w).sum().backward()
    ref_grads = {n: p.grad.clone() for n, p in bert.named_parameters() if p.grad is not None}
    bert.zero_grad(set_to_none=True)

    out = text_tower.encode(bert, ids, mask)
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert _rel(out[:, 0, :], ref[:, 0, :]) < 2e-2
    (out[:, 0, :] * w).sum().backward()
    for n, gr in ref_grads.items():
        if n.endswith("key.bias") or n.startswith("pooler"):
            continue
        got = dict(bert.named_parameters())[n].grad
        assert got is not None, n
        assert _rel(got, gr) < 5e-2, (n, _rel(got, gr))


def test_text_tower_training_mode_dropout(cuda_dev):
    """CXR-BERT's dropout (0.1 / 0.1) in training mode: runs, finite, and is stochastic; eval mode is deterministic and
    equals the dropout-free tower."""
    from transformers import BertConfig, BertModel
    from vit_exp_b200 import text_tower
    torch.manual_seed(0)
    bert = BertModel(BertConfig(vocab_size=30522, hidden_size=256, num_hidden_layers=2, num_attention_heads=4,
                                intermediate_size=1024)).to(cuda_dev).train()
    ids = torch.randint(0, 30522, (2, 128), generator=torch.Generator().manual_seed(1)).to(cuda_dev)
    mask = torch.ones(2, 128, dtype=torch.int64, device=cuda_dev)
    a = text_tower.encode(bert, ids, mask)
    b = text_tower.encode(bert, ids, mask)
    assert torch.isfinite(a).all() and not torch.equal(a, b)
    a[:, 0, :].square().sum().backward()
    assert all(torch.isfinite(p.grad).all() for n, p in bert.named_parameters() if p.grad is not None)
    e1 = text_tower.encode(bert.eval(), ids, mask)
    e2 = text_tower.encode(bert, ids, mask)
    assert torch.equal(e1, e2)
    ref = bert(ids, attention_mask=mask)[0]
    assert _rel(e1[:, 0, :], ref[:, 0, :]) < 2e-2


def test_ctclip_step_with_ctk_text_tower(cuda_dev):
    """CTCLIP(config={'ctk_text_tower': True}) gives the same loss as the stock text encoder path (1e-3 relative,
    the north star's loss tolerance) on identical weights and inputs."""
    from types import SimpleNamespace

    from vit_exp_b200.ct_clip import CTCLIP, TorchDistAccelerator
    from vit_exp_b200.transformer_maskgit import CTViT
    torch.manual_seed(0)
    vit = CTViT(dim=128, codebook_size=256, image_size=40, patch_size=20, temporal_patch_size=10, spatial_depth=1,
                temporal_depth=1, dim_head=32, heads=4)
    bert = _bert(cuda_dev, 256, 4, 2, 1024)
    losses = []
    for flag in (False, True):
        clip = CTCLIP(image_encoder=vit, text_encoder=bert, dim_text=256, dim_image=128, dim_latent=64,
                      config={"ctk_text_tower": flag}).to(cuda_dev).train()
        vit.eval()          # frozen codebook: a training-mode forward runs the VQ EMA update, which would change the
                            # image tokens between the two passes of this comparison
        torch.manual_seed(1)
        with torch.no_grad():
            clip.to_text_latent.weight.normal_(0, 0.05)
            clip.to_visual_latent.weight.normal_(0, 0.05)
        g = torch.Generator().manual_seed(3)
        batch = {"data_type": ["imagereport"] * 4,
                 "image": torch.rand(4, 1, 20, 40, 40, generator=g).to(cuda_dev),
                 "text": SimpleNamespace(input_ids=torch.randint(0, 30522, (4, 64), generator=g).to(cuda_dev),
                                         attention_mask=torch.ones(4, 64, dtype=torch.int64, device=cuda_dev))}
        loss, ld = clip(batch, device=cuda_dev, accelerator=TorchDistAccelerator())
        loss.backward()
        losses.append(float(loss))
        clip.zero_grad(set_to_none=True)
    assert abs(losses[0] - losses[1]) <= 1e-3 * abs(losses[0]) + 1e-6


# ----------------------------------------------------------------------------- attention core (ctk_mha_fwd / ctk_mha_bwd)
@pytest.mark.parametrize("B,L,heads,dh,p_drop,padded,n_null", [
    (2, 512, 12, 64, 0.0, True, 0), (3, 200, 4, 64, 0.0, True, 0), (1, 64, 2, 64, 0.0, False, 0), (2, 37, 3, 64, 0.0, True, 0),
    (2, 256, 4, 64, 0.1, True, 0), (1, 130, 2, 64, 0.5, False, 0),
    (2, 576, 8, 32, 0.0, False, 2), (1, 1500, 4, 32, 0.0, False, 2), (2, 100, 2, 32, 0.0, True, 5), (1, 256, 2, 32, 0.1, False, 0),
    (1, 200, 2, 64, 0.0, False, 3)])
def test_mha_forward_backward(cuda_dev, B, L, heads, dh, p_drop, padded, n_null):
    """against fp32 softmax attention on the same bf16-rounded projections (head dim 64: text tower; head dim 32 with
    learned null key/value pairs: CTViT3D's FlashAttention, attention.py:240-260); the dropout mask of the reference is
    the torch restatement of the kernels' hash (tests/emulated_ops.mha_keep_mask, pinned to csrc/mha_dropout.cuh on the
    CPU).  bf16 probabilities / outputs: 1e-2 relative L2 forward, 2e-2 backward; lse 1e-3 absolute."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    import emulated_ops as E
    from vit_exp_b200 import ops
    H = heads * dh
    qkv = (_rand((B * L, 3 * H), cuda_dev, 21) * (1.2 if dh == 64 else 2.0)).bfloat16()
    dout = _rand((B * L, H), cuda_dev, 22).bfloat16()
    mask = torch.ones(B, L, dtype=torch.uint8, device=cuda_dev)
    if padded:
        mask[0, L - L // 3:] = 0
        mask[B - 1, L // 2] = 0                       # a hole in the middle as well
    nk = nv = None
    if n_null:
        nk = _rand((heads, n_null, dh), cuda_dev, 23, 1.5).bfloat16()
        nv = _rand((heads, n_null, dh), cuda_dev, 24).bfloat16()
    seed = torch.tensor([987654321012345], dtype=torch.int64, device=cuda_dev)
    scale = dh ** -0.5
    ctx, lse = ops.mha_fwd(qkv, mask, B, L, heads, scale, p_drop, seed, 5, null_k=nk, null_v=nv)
    res = ops.mha_bwd(qkv, mask, ctx, dout, lse, B, L, heads, scale, p_drop, seed, 5, null_k=nk, null_v=nv)
    dqkv = res if n_null == 0 else res[0]
    torch.cuda.synchronize()
    x = qkv.cpu().float().requires_grad_(True)
    leaves = [x]
    nkc = nvc = None
    if n_null:
        nkc, nvc = nk.cpu().float().requires_grad_(True), nv.cpu().float().requires_grad_(True)
        leaves += [nkc, nvc]
    ref, ref_lse = E._mha_math(x, mask.cpu(), B, L, heads, scale, p_drop, seed.cpu(), 5, nkc, nvc)
    grefs = torch.autograd.grad(ref, leaves, dout.cpu().float())
    gref = grefs[0]
    assert torch.isfinite(ctx.float()).all() and torch.isfinite(dqkv.float()).all()
    assert _rel(ctx.cpu(), ref) < 1e-2, _rel(ctx.cpu(), ref)
    assert (lse.cpu() - ref_lse).abs().max().item() < 1e-3
    for name, sl in (("dq", slice(0, H)), ("dk", slice(H, 2 * H)), ("dv", slice(2 * H, 3 * H))):
        err = _rel(dqkv.cpu()[:, sl], gref[:, sl])
        assert err < 2e-2, (name, err)
    if n_null:
        assert res[1].shape == (B, heads, n_null, dh)
        assert _rel(res[1].sum(0).cpu(), grefs[1]) < 2e-2 and _rel(res[2].sum(0).cpu(), grefs[2]) < 2e-2
    # masked keys receive exactly zero gradient
    dead = (mask.cpu() == 0).reshape(-1)
    if dead.any():
        assert dqkv.cpu()[dead][:, H:].abs().max().item() == 0.0


def test_mha_long_sequence_with_null_pairs_vs_sdpa(cuda_dev):
    """CTViT3D's real shape: 13 824 tokens + 2 null pairs, 8 heads x 32 (attention.py:250-260).  Checker: torch's SDPA in
    fp32 on the GPU (a library call used as the test's reference only), forward and all gradients."""
    import torch.nn.functional as F
    from vit_exp_b200 import ops
    B, L, heads, dh, n_null = 1, 13824, 8, 32, 2
    H = heads * dh
    qkv = (_rand((B * L, 3 * H), cuda_dev, 31) * 1.5).bfloat16()
    dout = _rand((B * L, H), cuda_dev, 32).bfloat16()
    nk = _rand((heads, n_null, dh), cuda_dev, 33, 1.5).bfloat16()
    nv = _rand((heads, n_null, dh), cuda_dev, 34).bfloat16()
    ctx, lse = ops.mha_fwd(qkv, None, B, L, heads, 1.0, null_k=nk, null_v=nv)
    dqkv, dnk, dnv = ops.mha_bwd(qkv, None, ctx, dout, lse, B, L, heads, 1.0, null_k=nk, null_v=nv)
    x = qkv.float().requires_grad_(True)
    nkf, nvf = nk.float().requires_grad_(True), nv.float().requires_grad_(True)
    q, k, v = (x.view(B, L, 3, heads, dh)[:, :, t].transpose(1, 2) for t in range(3))
    K = torch.cat([nkf[None].expand(B, -1, -1, -1), k], dim=2)
    V = torch.cat([nvf[None].expand(B, -1, -1, -1), v], dim=2)
    ref = F.scaled_dot_product_attention(q, K, V, scale=1.0).transpose(1, 2).reshape(B * L, H)
    gx, gnk, gnv = torch.autograd.grad(ref, (x, nkf, nvf), dout.float())
    assert _rel(ctx, ref) < 1e-2
    assert _rel(dqkv, gx) < 2e-2
    assert _rel(dnk.sum(0), gnk) < 2e-2 and _rel(dnv.sum(0), gnv) < 2e-2


def test_mha_is_deterministic_and_seed_sensitive(cuda_dev):
    from vit_exp_b200 import ops
    B, L, heads = 2, 128, 4
    qkv = _rand((B * L, 3 * heads * 64), cuda_dev, 31).bfloat16()
    s1 = torch.tensor([5], dtype=torch.int64, device=cuda_dev)
    a, _ = ops.mha_fwd(qkv, None, B, L, heads, 0.125, 0.1, s1, 0)
    b, _ = ops.mha_fwd(qkv, None, B, L, heads, 0.125, 0.1, s1, 0)
    c, _ = ops.mha_fwd(qkv, None, B, L, heads, 0.125, 0.1, s1, 1)
    s1.add_(1)
    d, _ = ops.mha_fwd(qkv, None, B, L, heads, 0.125, 0.1, s1, 0)
    assert torch.equal(a, b) and not torch.equal(a, c) and not torch.equal(a, d)


def test_bert_embeddings_forward_backward(cuda_dev):
    from vit_exp_b200 import ops
    g = torch.Generator().manual_seed(41)
    V, P, T, H, B, L = 1000, 64, 2, 256, 3, 40
    word, pos, typ = (torch.randn(n, H, generator=g).to(cuda_dev) for n in (V, P, T))
    ids = torch.randint(0, V, (B, L), generator=g).to(cuda_dev)
    ids[0, :5] = 0
    tt = torch.randint(0, T, (B, L), generator=g).to(cuda_dev)
    de = torch.randn(B * L, H, generator=g).to(cuda_dev)
    for token_type in (None, tt):
        e = ops.bert_embed_fwd(ids, token_type, word, pos, typ)
        ref = word[ids] + (typ[0] if token_type is None else typ[token_type]) + pos[:L]
        assert torch.equal(e.view(B, L, H), ref)
        dword, dpos, dtyp = ops.bert_embed_bwd(de, ids, token_type, word.shape, pos.shape, typ.shape, 0)
        rw = torch.zeros_like(word).index_add_(0, ids.reshape(-1), de)
        rw[0].zero_()
        rp = torch.zeros_like(pos)
        rp[:L] = de.view(B, L, H).sum(0)
        rt = torch.zeros_like(typ)
        if token_type is None:
            rt[0] = de.sum(0)
        else:
            rt.index_add_(0, token_type.reshape(-1), de)
        assert _rel(dword, rw) < 1e-6 and _rel(dpos, rp) < 1e-6 and _rel(dtyp, rt) < 1e-5
        assert dword[0].abs().max().item() == 0.0


@pytest.mark.parametrize("dropout", [0.0, 0.1])
def test_text_tower_graph_replay_matches_eager(cuda_dev, dropout):
    """after two eager calls the tower's forward / backward are replayed from CUDA graphs (text_tower._TowerGraph):
    same outputs and gradients as eager launches on fresh inputs (bitwise without dropout: the kernels are deterministic
    except for the split-K weight-gradient atomics), fresh dropout masks on every replay."""
    import os
    from transformers import BertConfig, BertModel
    from vit_exp_b200 import ops, text_tower
    torch.manual_seed(0)
    bert = BertModel(BertConfig(vocab_size=30522, hidden_size=256, num_hidden_layers=2, num_attention_heads=4,
                                intermediate_size=1024, hidden_dropout_prob=dropout,
                                attention_probs_dropout_prob=dropout)).to(cuda_dev).train()
    w = _rand((256,), cuda_dev, 8)
    g = torch.Generator().manual_seed(5)
    batches = [(torch.randint(0, 30522, (2, 96), generator=g).to(cuda_dev), torch.ones(2, 96, dtype=torch.int64, device=cuda_dev))
               for _ in range(5)]
    for ids, mask in batches:
        mask[1, 70:] = 0

    def run(ids, mask):
        bert.zero_grad(set_to_none=True)
        out = text_tower.encode(bert, ids, mask)
        (out[:, 0, :] * w).sum().backward()
        return out.detach().clone(), {n: p.grad.detach().clone() for n, p in bert.named_parameters() if p.grad is not None}
    n0 = ops.GRAPH_LAUNCHES
    outs = [run(*b) for b in batches]                         # calls 3.. are graph replays
    assert ops.GRAPH_LAUNCHES > n0, "the tower never switched to graph replay"
    if dropout == 0.0:
        os.environ["CTK_TEXT_GRAPHS"] = "0"
        try:
            for (out, grads), b in list(zip(outs, batches))[2:]:
                ref_out, ref_grads = run(*b)
                assert torch.equal(out, ref_out)
                for n, gr in ref_grads.items():
                    assert _rel(grads[n], gr) < 1e-5, n
        finally:
            os.environ.pop("CTK_TEXT_GRAPHS")
    else:
        a = text_tower.encode(bert, *batches[0])
        b = text_tower.encode(bert, *batches[0])
        assert torch.isfinite(a).all() and not torch.equal(a, b)            # replays draw fresh masks
        assert all(torch.isfinite(v).all() for _, grads in outs for v in grads.values())
