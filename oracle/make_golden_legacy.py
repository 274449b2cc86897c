"""Generate tests/golden/forward_old_golden.pt by running the REAL reference `CTCLIP.forward_old`
(CT_CLIP/ct_clip/ct_clip.py:1392-1778; build container only).

    python -m oracle.make_golden_legacy

`forward_old` is the pooling of the original CT-CLIP checkpoints (scripts/run_zero_shot_latent.py:26-31,
`dim_image = 294912`): encoded tokens (B, t, h, w, C) -> mean over axis 1 -> flatten (h w C) -> to_visual_latent,
with the rows of both towers selected by `text_valid_mask` before the projections (ct_clip.py:1549-1565,1593-1594,1614).

Three entry modes of the reference are exercised on a tiny model whose token grid has t != h (so the averaged axis
is pinned): `return_latents=True`, the pairwise similarity (`return_loss=False`) and the loss + gradients.
The reference's loss branch reads `seg_loss` unconditionally (ct_clip.py:1766), so it only runs with `use_seg=True`;
the fixture passes `use_seg=True` with an all-zero `seg_valid_mask`, which takes the reference's own
"no volume to segment -> seg_loss = 0." branch (ct_clip.py:1523-1525) and leaves the contrastive loss alone.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    from oracle.make_golden import tiny_ctvit
    from oracle.ref_import import FakeAccelerator, import_reference
    CTViT, CTCLIP, _, _ = import_reference()
    from transformers import BertConfig, BertModel
    torch.set_grad_enabled(True)
    torch.manual_seed(21)
    bert = BertModel(BertConfig(vocab_size=100, hidden_size=48, num_hidden_layers=1, num_attention_heads=2,
                                intermediate_size=64, max_position_embeddings=32,
                                hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0))
    vit = tiny_ctvit(CTViT, seed=22)
    # video (B, 1, 15, 20, 20) with patches 10 x 10 x 5 -> tokens (B, t=3, h=2, w=2, 64): dim_image = 2*2*64
    clip = CTCLIP(image_encoder=vit, text_encoder=bert, dim_text=48, dim_image=2 * 2 * 64, dim_latent=32, config={})
    clip.train()
    B = 5
    g = torch.Generator().manual_seed(23)
    ids = torch.randint(0, 100, (B, 16), generator=g)
    text = SimpleNamespace(input_ids=ids, attention_mask=torch.ones_like(ids))
    vids = torch.rand(B, 1, 15, 20, 20, generator=g)
    dev = torch.device("cpu")
    valid_all = torch.ones(B, 1)
    valid_some = torch.tensor([[1.], [0.], [1.], [1.], [0.]])
    sink = io.StringIO()                                  # forward_old prints shapes (ct_clip.py:1596,1715)

    enc_text = bert(ids, attention_mask=text.attention_mask)[0].detach()
    with torch.no_grad():
        clip.eval()
        enc_image = vit(vids, return_encoded_tokens=True).detach()
        with contextlib.redirect_stdout(sink):
            tl, il, enc_send = clip.forward_old(text, vids, dev, return_latents=True, text_valid_mask=valid_all)
            tl_s, il_s, _ = clip.forward_old(text, vids, dev, return_latents=True, text_valid_mask=valid_some)
            sim = clip.forward_old(text, vids, dev, text_valid_mask=valid_some)
        clip.train()
    assert torch.equal(enc_send, enc_image)

    fix = {
        "state_dict": {k: v.clone() for k, v in clip.state_dict().items()
                       if not k.startswith("text_transformer.") and "to_pixels" not in k and "first_frame" not in k},
        "bert_state_dict": {k: v.clone() for k, v in bert.state_dict().items()},
        "ids": ids, "video": vids, "enc_text": enc_text, "enc_image": enc_image,
        "valid_some": valid_some,
        "text_latents": tl.detach(), "image_latents": il.detach(),
        "text_latents_some": tl_s.detach(), "image_latents_some": il_s.detach(),
        "similarity_some": sim.detach(),
    }
    # loss + gradients (train mode: straight-through VQ), rows 0, 2, 3 valid
    clip.zero_grad()
    with contextlib.redirect_stdout(sink):
        loss, ld = clip.forward_old(text, vids, dev, return_loss=True, return_loss_dict=True, use_seg=True,
                                    seg_mask=torch.zeros(B, 1, 15, 20, 20), seg_valid_mask=torch.zeros(B, 1),
                                    text_valid_mask=valid_some, accelerator=FakeAccelerator())
    loss.backward()
    fix.update({
        "loss": loss.detach(), "cl_loss": ld["cl_loss"],
        "grad_to_text_latent": clip.to_text_latent.weight.grad.clone(),
        "grad_to_visual_latent": clip.to_visual_latent.weight.grad.clone(),
        "grad_temperature": clip.temperature.grad.clone(),
        "grad_patch_weight": vit.to_patch_emb[2].weight.grad.clone(),
        "grad_word_embeddings": bert.embeddings.word_embeddings.weight.grad.clone(),
    })
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "forward_old_golden.pt")
    torch.save(fix, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB); loss {loss.item():.6f}, "
          f"similarity {[round(v, 4) for v in sim.tolist()]}")


if __name__ == "__main__":
    main()
