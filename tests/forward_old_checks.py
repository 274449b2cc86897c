"""Shared body of the module-level `forward_old` checks (tests/test_forward_old_gpu.py on the GPU; tests/test_forward_old_cpu.py
runs the same body on the torch doubles with bf16 operands).

The encoder computes with bf16 operands, so a token whose two best codes are almost tied may land on the other code
than in the fp32 reference (the search is exact on the encoder's OWN output).  One such flip replaces a
whole code vector, and on this fixture (3 frames x 4 tokens per volume) it moves the image latent by more than any
tolerance meant for rounding.  So the head is checked against the oracle fed with the tokens the module itself
produced - which pins pooling axis, row selection, projections, loss and every gradient - and the tokens are checked
against the reference's by code agreement; where all codes agree the outputs are also compared with the reference's
recorded ones directly.
"""
import torch

from oracle import ctclip_oracle as orc


def rel_l2(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()


def run(clip, vit, bert, text, video, g, accelerator, dev, min_agree=0.9):
    from transformers import BertConfig, BertModel
    p = {k: v.clone() for k, v in g["state_dict"].items()}
    B = video.shape[0]
    ones = torch.ones(B, 1)
    some = g["valid_some"]
    clip.eval()
    tl, il, enc = clip.forward_old(text, video, dev, return_latents=True, text_valid_mask=ones)
    assert enc.shape == g["enc_image"].shape
    ours = enc.detach().float().cpu()
    same = (ours.reshape(-1, ours.shape[-1]) == g["enc_image"].reshape(-1, ours.shape[-1])).all(-1)
    agree = same.float().mean().item()
    assert agree >= min_agree, f"only {agree:.2f} of the tokens carry the reference's code"
    # every token is a codebook row, bit for bit (the quantised output is a gather)
    embed = p["visual_transformer.vq._codebook.embed"][0]
    assert all((embed == row).all(-1).any() for row in ours.reshape(-1, ours.shape[-1]))
    tl_ref, il_ref = orc.forward_old_latents(g["enc_text"], ours, p, ones)
    assert rel_l2(tl, tl_ref) < 1e-4 and rel_l2(il, il_ref) < 1e-4
    assert rel_l2(tl, g["text_latents"]) < 1e-4
    if agree == 1.0:
        assert rel_l2(il, g["image_latents"]) < 2e-2
    tl_s, il_s, _ = clip.forward_old(text, video, dev, return_latents=True, text_valid_mask=some.to(video.device))
    tl_ref, il_ref = orc.forward_old_latents(g["enc_text"], ours, p, some)
    assert tl_s.shape == g["text_latents_some"].shape
    assert rel_l2(tl_s, tl_ref) < 1e-4 and rel_l2(il_s, il_ref) < 1e-4
    sim = clip.forward_old(text, video, dev, text_valid_mask=some)
    sim_ref = orc.forward_old_similarity(g["enc_text"], ours, p, some)
    assert sim.shape == g["similarity_some"].shape and (sim.cpu() - sim_ref).abs().max().item() < 1e-4
    if agree == 1.0:
        ref = g["similarity_some"]
        assert (sim.cpu() - ref).abs().max().item() < 2e-2 * ref.abs().max().item() + 1e-3
    _, il2 = clip.latents(text, video)                     # pooling follows the width of to_visual_latent
    assert rel_l2(il2, il) < 1e-6

    # loss + gradients (train mode: straight-through VQ; the first training forward quantises like the eval one)
    clip.train()
    loss, ld = clip.forward_old(text, video, dev, return_loss=True, return_loss_dict=True, text_valid_mask=some,
                                accelerator=accelerator)
    assert ld["loss_total"] == ld["cl_loss"] and abs(ld["cl_loss"] - loss.item()) < 1e-7
    loss.backward()
    # reference chain: HF BERT (fp32, CPU) -> oracle head on OUR tokens -> oracle encoder backward (straight-through)
    ref_bert = BertModel(BertConfig(vocab_size=100, hidden_size=48, num_hidden_layers=1, num_attention_heads=2,
                                    intermediate_size=64, max_position_embeddings=32,
                                    hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0))
    ref_bert.load_state_dict(g["bert_state_dict"])
    for k in ("to_text_latent.weight", "to_visual_latent.weight", "temperature"):
        p[k].requires_grad_()
    tok_leaf = ours.clone().requires_grad_()
    enc_text = ref_bert(g["ids"], attention_mask=torch.ones_like(g["ids"]))[0]
    ref_loss, _, _ = orc.forward_old_loss(enc_text, tok_leaf, p, some)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-3 * abs(ref_loss.item())
    if agree == 1.0:
        assert abs(loss.item() - g["loss"].item()) <= 1e-3 * abs(g["loss"].item())
    assert rel_l2(clip.to_text_latent.weight.grad, p["to_text_latent.weight"].grad) < 1e-3
    assert rel_l2(clip.to_visual_latent.weight.grad, p["to_visual_latent.weight"].grad) < 1e-3
    gt = p["temperature"].grad.item()
    assert abs(clip.temperature.grad.item() - gt) <= 1e-3 * abs(gt) + 1e-6
    assert rel_l2(bert.embeddings.word_embeddings.weight.grad, ref_bert.embeddings.word_embeddings.weight.grad) < 1e-3
    vp = {k[len("visual_transformer."):]: v.clone().requires_grad_(k == "visual_transformer.to_patch_emb.2.weight")
          for k, v in g["state_dict"].items() if k.startswith("visual_transformer.")}
    _, pre, _ = orc.ctvit_forward(g["video"], vp, patch=10, tpatch=5, spatial_depth=2, temporal_depth=1, heads=2,
                                  return_pre_vq=True)
    (pre * tok_leaf.grad.reshape(pre.shape)).sum().backward()
    assert rel_l2(vit.to_patch_emb[2].weight.grad, vp["to_patch_emb.2.weight"].grad) < 8e-2
    return agree
