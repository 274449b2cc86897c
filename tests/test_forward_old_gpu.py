"""SURVEY 8f row 5 - the pooling of the original CT-CLIP checkpoints, `CTCLIP.forward_old` (ct_clip.py:1392-1778;
`dim_image = 294912`, scripts/run_zero_shot_latent.py:26-31) - on the GPU through the C ABI:

  * `ctk_mean_pool_fwd` on a short axis of wide rows (mean over the 24 frames of 24*24*512 floats) and `ctk_latent_fwd` /
    `ctk_latent_bwd` with 294 912 input features, against fp64 torch; repeatability of the wide projection (fixed
    summation order, no atomics);
  * the whole module against the oracle pinned to the REAL reference `forward_old` (tests/golden/forward_old_golden.pt,
    oracle/make_golden_legacy.py) and against the reference's recorded outputs: latents, similarity, loss, gradients;
  * the head at full size (tokens 4 x 24 x 24 x 24 x 512, one report masked out) against autograd through the oracle.
"""
import os
from types import SimpleNamespace

import pytest
import torch

from oracle import ctclip_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "forward_old_golden.pt")


def rel_l2(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


@pytest.mark.parametrize("B,n,dim", [(3, 24, 24 * 24 * 512), (2, 5, 2052), (1, 256, 2048)])
def test_mean_over_frames(cuda_dev, B, n, dim):
    from vit_exp_b200 import ops
    x = torch.randn(B, n, dim, device=cuda_dev, generator=torch.Generator(cuda_dev).manual_seed(B * n))
    got = ops.mean_pool(x)
    ref = x.double().mean(1)
    assert got.shape == (B, dim)
    assert (got.double() - ref).abs().max().item() < 1e-6


@pytest.mark.parametrize("B,din,dl,pad", [(8, 294912, 512, 0), (1, 294912, 512, 0), (11, 5004, 10, 0), (3, 4096, 7, 8),
                                          (2, 1028, 5, 0)])
def test_latent_wide_fwd_bwd(cuda_dev, B, din, dl, pad):
    from vit_exp_b200 import ops
    g = torch.Generator(cuda_dev).manual_seed(din + B)
    xs = torch.randn(B, din + pad, device=cuda_dev, generator=g)
    x = xs[:, :din]                                        # rows may be strided views (stride a multiple of 4)
    W = torch.randn(dl, din, device=cuda_dev, generator=g) * din ** -0.5
    lat, rn = ops.latent_fwd(x, W)
    lat2, rn2 = ops.latent_fwd(x, W)
    assert torch.equal(lat, lat2) and torch.equal(rn, rn2)                    # fixed summation order
    xd = x.double().clone().requires_grad_()
    Wd = W.double().clone().requires_grad_()
    raw = xd @ Wd.t()
    ref = torch.nn.functional.normalize(raw, dim=-1, eps=1e-12)
    assert rel_l2(lat, ref) < 1e-5
    assert rel_l2(rn, 1.0 / raw.norm(dim=-1)) < 1e-5
    dlat = torch.randn(B, dl, device=cuda_dev, generator=g)
    (ref * dlat.double()).sum().backward()
    dW, dx = ops.latent_bwd(dlat, lat, rn, x.contiguous(), W)
    assert rel_l2(dW, Wd.grad) < 1e-5
    assert rel_l2(dx, xd.grad) < 1e-5


def _tiny_clip(dev, g):
    from transformers import BertConfig, BertModel
    from vit_exp_b200.ct_clip import CTCLIP
    from vit_exp_b200.transformer_maskgit import CTViT
    sd = g["state_dict"]
    vit = CTViT(dim=64, codebook_size=64, image_size=20, patch_size=10, temporal_patch_size=5, spatial_depth=2,
                temporal_depth=1, dim_head=32, heads=2)
    vit.load_state_dict({k[len("visual_transformer."):]: v for k, v in sd.items() if k.startswith("visual_transformer.")},
                        strict=False)
    bert = BertModel(BertConfig(vocab_size=100, hidden_size=48, num_hidden_layers=1, num_attention_heads=2,
                                intermediate_size=64, max_position_embeddings=32,
                                hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0))
    bert.load_state_dict(g["bert_state_dict"])
    clip = CTCLIP(image_encoder=vit, text_encoder=bert, dim_text=48, dim_image=2 * 2 * 64, dim_latent=32, config={})
    clip.load_state_dict({k: v for k, v in sd.items() if not k.startswith("visual_transformer.")}, strict=False)
    clip = clip.to(dev)
    text = SimpleNamespace(input_ids=g["ids"].to(dev), attention_mask=torch.ones_like(g["ids"]).to(dev))
    return clip, vit, bert, text


def test_forward_old_matches_reference(cuda_dev, gold):
    """latents / similarity / loss / gradients of the whole module (tests/forward_old_checks.py explains how near-tie code
    flips of the bf16 encoder are kept apart from the path under test)."""
    import forward_old_checks
    from vit_exp_b200.ct_clip import TorchDistAccelerator
    clip, vit, bert, text = _tiny_clip(cuda_dev, gold)
    forward_old_checks.run(clip, vit, bert, text, gold["video"].to(cuda_dev), gold, TorchDistAccelerator(), cuda_dev)


def test_forward_old_head_full_size_vs_oracle(cuda_dev):
    """Encoded tokens of the full-size encoder (24 x 24 x 24 x 512 per volume), Linear(294912 -> 512), report 2 of 4 masked
    out: loss and every gradient of the head against autograd through the oracle restatement (fp32 on the CPU)."""
    from vit_exp_b200.ct_clip import TorchDistAccelerator, _ClipHead
    g = torch.Generator().manual_seed(41)
    B, t, h, w, dim, dt, dl = 4, 24, 24, 24, 512, 768, 512
    tokens = torch.randn(B, t, h, w, dim, generator=g)
    cls = torch.randn(B, dt, generator=g)
    wt = torch.randn(dl, dt, generator=g) * dt ** -0.5
    wv = torch.randn(dl, h * w * dim, generator=g) * (h * w * dim) ** -0.5 * t ** 0.5      # unit-scale raw latents
    temp = torch.tensor(1.1)
    valid_mask = torch.tensor([[1.], [1.], [0.], [1.]])
    dev = [v.to(cuda_dev).requires_grad_() for v in (cls, tokens, wt, wv, temp)]
    valid = torch.nonzero(valid_mask[:, 0]).squeeze(1).to(cuda_dev)
    loss, tl, il = _ClipHead.apply(*dev, TorchDistAccelerator(), True, valid)
    loss.backward()
    ref_in = [v.clone().requires_grad_() for v in (cls, tokens, wt, wv, temp)]
    p = {"to_text_latent.weight": ref_in[2], "to_visual_latent.weight": ref_in[3], "temperature": ref_in[4]}
    ref, tl_ref, il_ref = orc.forward_old_loss(ref_in[0][:, None, :], ref_in[1], p, valid_mask)
    ref.backward()
    assert tl.shape == (3, dl) and rel_l2(tl, tl_ref) < 2e-5 and rel_l2(il, il_ref) < 2e-5
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())
    for name, a, b in zip(("cls", "tokens", "w_text", "w_vis", "temperature"), dev, ref_in):
        assert rel_l2(a.grad, b.grad) < 2e-4, name
    assert dev[1].grad[2].abs().max().item() == 0 and dev[0].grad[2].abs().max().item() == 0      # the masked-out report
