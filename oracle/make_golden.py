"""Generate tests/golden/*.pt by running the REAL reference modules (build container only).

    python -m oracle.make_golden

Fixtures are small (tiny CTViT / CTCLIP configurations, seeded inputs) so they can be committed;
tests/test_oracle_cpu.py checks oracle/ctclip_oracle.py against them, which pins the oracle to the
reference's own code. The VectorQuantize call is served by a stand-in (parity unpinned there).
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def tiny_ctvit(CTViT, seed, dim=64, heads=2, depth=(2, 1), codebook=64):
    torch.manual_seed(seed)
    vit = CTViT(dim=dim, codebook_size=codebook, image_size=20, patch_size=10, temporal_patch_size=5,
                spatial_depth=depth[0], temporal_depth=depth[1], dim_head=32, heads=heads)
    # move parameters off their init values so gamma/beta/scales/bias paths are all exercised
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in vit.named_parameters():
            if p.numel() == 0:
                continue
            if n.endswith(("gamma", "q_scale", "k_scale")) or ".0.weight" in n and p.dim() == 1 \
                    or n in ("to_patch_emb.1.weight", "to_patch_emb.3.weight"):
                p.mul_(1 + 0.2 * torch.randn(p.shape, generator=g))
            elif p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    return vit


def main():
    os.makedirs(OUT, exist_ok=True)
    from oracle.ref_import import FakeAccelerator, import_reference
    CTViT, CTCLIP, ref_attn, AllGather = import_reference()
    torch.set_grad_enabled(True)

    # ---- 1. loss known answers on the reference's fixed vectors (demo_tests/test_loss_type.py:14-15)
    sys.path.insert(0, "/root/reference/demo_tests")
    import types
    timm = types.ModuleType("timm"); timm_loss = types.ModuleType("timm.loss")
    timm_loss.LabelSmoothingCrossEntropy = object
    sys.modules.setdefault("timm", timm); sys.modules.setdefault("timm.loss", timm_loss)
    from clip_loss import ClipLoss
    x = torch.tensor([[0.1, 0.2], [0.3, 0.4], [0.5, 0.6], [0.7, 0.8]])
    y = torch.tensor([[0.2, 0.3], [0.3, 0.6], [0.4, 0.9], [0.2, 0.5]])
    open_clip = ClipLoss(local_loss=False, gather_with_grad=True, cache_labels=False, rank=0, world_size=1,
                         use_horovod=False, smoothing=0.)
    g = torch.Generator().manual_seed(0)
    T8 = torch.nn.functional.normalize(torch.randn(8, 512, generator=g), dim=-1)
    I8 = torch.nn.functional.normalize(torch.randn(8, 512, generator=g), dim=-1)
    loss_fix = {
        "x": x, "y": y,
        "open_clip_xy": open_clip(x, y, 1.0).detach(),             # image_features=x, text_features=y
        "T8": T8, "I8": I8,
        "open_clip_8": open_clip(I8, T8, torch.tensor(1.0).exp()).detach(),
    }

    # ---- 2. tiny CTViT through the real reference code
    vit = tiny_ctvit(CTViT, seed=0).eval()
    gv = torch.Generator().manual_seed(10)
    video = torch.rand(2, 1, 15, 20, 20, generator=gv)
    video[1, :, 10:] = -1.0                                          # padded slab (data.py:99)
    caps = {}
    hooks = [
        vit.to_patch_emb.register_forward_hook(lambda m, i, o: caps.__setitem__("patch_tokens", o.detach().clone())),
        vit.spatial_rel_pos_bias.register_forward_hook(lambda m, i, o: caps.__setitem__("attn_bias", o.detach().clone())),
        vit.enc_spatial_transformer.register_forward_hook(lambda m, i, o: caps.__setitem__("spatial_out", o.detach().clone())),
        vit.enc_spatial_transformer.layers[0][0].register_forward_hook(lambda m, i, o: caps.__setitem__("peg0_out", o.detach().clone())),
        vit.enc_spatial_transformer.layers[0][1].register_forward_hook(lambda m, i, o: caps.__setitem__("attn0_out", o.detach().clone())),
        vit.enc_spatial_transformer.layers[0][3].register_forward_hook(lambda m, i, o: caps.__setitem__("ff0_out", o.detach().clone())),
        vit.enc_temporal_transformer.layers[0][0].register_forward_hook(lambda m, i, o: caps.__setitem__("tpeg0_out", o.detach().clone())),
        vit.enc_temporal_transformer.register_forward_hook(lambda m, i, o: caps.__setitem__("temporal_out", o.detach().clone())),
    ]
    with torch.no_grad():
        tokens = vit(video, return_encoded_tokens=True)
    for h in hooks:
        h.remove()
    skip = ("to_pixels", "to_patch_emb_first_frame")
    vit_fix = {"state_dict": {k: v.clone() for k, v in vit.state_dict().items() if not k.startswith(skip)}, "video": video,
               "tokens": tokens.detach(), "pre_vq": vit.vq.pre_vq.reshape(tokens.shape), **caps,
               "cfg": dict(dim=64, heads=2, spatial_depth=2, temporal_depth=1, patch=10, tpatch=5, codebook=64)}

    # gradients of the encoder (train mode, straight-through VQ) for a fixed cotangent
    vit.train()
    vit.zero_grad()
    out = vit(video, return_encoded_tokens=True)
    cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(11))
    (out * cot).sum().backward()
    vit_fix["cotangent"] = cot
    vit_fix["grads"] = {k: p.grad.clone() for k, p in vit.named_parameters() if p.grad is not None and p.numel() > 0}

    # ---- 3. CTCLIP head + loss through the real reference forward (tiny BERT as text encoder)
    from transformers import BertConfig, BertModel
    torch.manual_seed(3)
    bert = BertModel(BertConfig(vocab_size=100, hidden_size=48, num_hidden_layers=1, num_attention_heads=2,
                                intermediate_size=64, max_position_embeddings=32,
                                hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0))
    vit2 = tiny_ctvit(CTViT, seed=5)
    clip = CTCLIP(image_encoder=vit2, text_encoder=bert, dim_text=48, dim_image=64, dim_latent=32, config={})
    clip.train()
    B = 4
    gi = torch.Generator().manual_seed(12)
    ids = torch.randint(0, 100, (B, 16), generator=gi)
    text = SimpleNamespace(input_ids=ids, attention_mask=torch.ones_like(ids))
    vids = torch.rand(B, 1, 10, 20, 20, generator=gi)
    batch = {"data_type": ["imagereport"] * B, "text": text, "image": vids}
    enc_text = bert(ids, attention_mask=text.attention_mask)[0].detach()
    with torch.no_grad():
        clip.eval(); enc_image = vit2(vids, return_encoded_tokens=True).detach(); clip.train()
    loss, ld = clip(batch, device=torch.device("cpu"), accelerator=FakeAccelerator())
    loss.backward()
    clip_fix = {
        "state_dict": {k: v.clone() for k, v in clip.state_dict().items() if not k.startswith("text_transformer.") and "to_pixels" not in k and "first_frame" not in k},
        "enc_text": enc_text, "enc_image": enc_image, "video": vids, "loss": loss.detach(), "cl_loss": ld["cl_loss"],
        "grad_to_text_latent": clip.to_text_latent.weight.grad.clone(),
        "grad_to_visual_latent": clip.to_visual_latent.weight.grad.clone(),
        "grad_temperature": clip.temperature.grad.clone(),
        "grad_patch_weight": vit2.to_patch_emb[2].weight.grad.clone(),
    }
    # zero-shot logits (ct_clip.py:792-855) for two prompts against volume 0
    clip.eval()
    with torch.no_grad():
        t2 = SimpleNamespace(input_ids=ids[:2], attention_mask=torch.ones_like(ids[:2]))
        clip_fix["forward_infer"] = clip.forward_infer(t2, vids[:1]).detach()

    # ---- 4. AllGather backward convention (distributed.py:9-20) on a 2-rank fake gather
    class TwoRank:
        num_processes, process_index = 2, 1
        def __init__(self, other): self.other = other
        def gather(self, x): return torch.cat([self.other, x], dim=0)
    a = torch.randn(3, 4, generator=torch.Generator().manual_seed(13), requires_grad=True)
    other = torch.randn(3, 4, generator=torch.Generator().manual_seed(14))
    gathered = AllGather.apply(a, TwoRank(other))
    w = torch.arange(24.0).reshape(6, 4)
    (gathered * w).sum().backward()
    gather_fix = {"a": a.detach(), "other": other, "gathered": gathered.detach(), "w": w, "grad_a": a.grad.clone()}

    torch.save({"loss": loss_fix, "ctvit": vit_fix, "ctclip": clip_fix, "allgather": gather_fix},
               os.path.join(OUT, "reference_golden.pt"))
    sz = os.path.getsize(os.path.join(OUT, "reference_golden.pt"))
    print(f"wrote tests/golden/reference_golden.pt ({sz/1e6:.2f} MB)")
    print("open_clip_xy", loss_fix["open_clip_xy"].item(), "cl_loss", ld["cl_loss"])


if __name__ == "__main__":
    main()
