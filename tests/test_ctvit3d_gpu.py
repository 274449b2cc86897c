"""Drop-in CTViT3D on the GPU against the oracle pinned to the reference module (bf16 operands / fp32 accumulation:
tokens 2e-2 relative L2, parameter gradients 5e-2).  Uses libctk kernels validated for CTViT (at
dim 768) plus library SDPA for the joint attention core.

Validated on a B200 in round 2.
"""
import os
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _model(blocks=2, heads=4):
    from vit_exp_b200.ctvit3d import CTViT3D
    torch.manual_seed(0)
    m = CTViT3D(dim=768, image_size=80, patch_size=20, temporal_size=40, temporal_patch_size=10,
                transformer_blocks=blocks, dim_head=32, heads=heads)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith(("gamma", "q_scale", "k_scale")):
                p.mul_(1 + 0.2 * torch.randn(p.shape, generator=g))
    return m


def test_tokens_and_gradients_vs_oracle(cuda_dev):
    from oracle import ctclip_oracle as O
    m = _model()
    g = torch.Generator().manual_seed(2)
    video = torch.rand(2, 1, 40, 80, 80, generator=g)
    p = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and k != "pos_embed")
         for k, v in m.state_dict().items()}
    ref = O.ctvit3d_forward(video, p, patch=20, tpatch=10, blocks=2, heads=4)
    dy = torch.randn(ref.shape, generator=g)
    (ref * dy).sum().backward()
    m = m.to(cuda_dev)
    out = m(video.to(cuda_dev), return_encoded_tokens=True)
    assert out.shape == ref.shape and _rel(out, ref) < 2e-2
    (out * dy.to(cuda_dev)).sum().backward()
    for n, q in m.named_parameters():
        if n.startswith(("spatial_rel_pos_bias", "to_pixels")) or n.endswith("context_norm.gamma") or n == "pos_embed":
            assert q.grad is None, n
            continue
        assert q.grad is not None, n
        assert _rel(q.grad, p[n].grad) < 5e-2, (n, _rel(q.grad, p[n].grad))


def test_ctclip_step_with_ctvit3d(cuda_dev):
    from oracle import ctclip_oracle as O
    from vit_exp_b200.ct_clip import CTCLIP, TorchDistAccelerator
    m = _model(blocks=1)

    class _Text(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = torch.nn.Embedding(50, 96)

        def forward(self, input_ids, attention_mask=None):
            return (self.emb(input_ids),)

    clip = CTCLIP(image_encoder=m, text_encoder=_Text(), dim_text=96, dim_image=768, dim_latent=64).to(cuda_dev).train()
    g = torch.Generator().manual_seed(3)
    video = torch.rand(4, 1, 40, 80, 80, generator=g)
    ids = torch.randint(0, 50, (4, 8), generator=g)
    batch = {"data_type": ["imagereport"] * 4, "image": video.to(cuda_dev),
             "text": SimpleNamespace(input_ids=ids.to(cuda_dev), attention_mask=None)}
    loss, ld = clip(batch, device=cuda_dev, accelerator=TorchDistAccelerator())
    loss.backward()
    p = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    tokens = O.ctvit3d_forward(video, p, patch=20, tpatch=10, blocks=1, heads=4)
    ref, _, _ = O.ctclip_loss(clip.text_transformer.emb.weight.detach().cpu()[ids], tokens,
                              {"to_text_latent.weight": clip.to_text_latent.weight.detach().cpu(),
                               "to_visual_latent.weight": clip.to_visual_latent.weight.detach().cpu(),
                               "temperature": clip.temperature.detach().cpu()})
    assert abs(float(loss) - float(ref)) <= 1e-3 * abs(float(ref))               # north-star loss tolerance
    assert m.enc_3D.layers[0][1].null_kv.grad is not None
