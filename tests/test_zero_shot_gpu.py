"""ZeroShotScorer on the GPU against the reference's forward_infer + apply_softmax math (oracle restatement).
Uses the encoder, mean-pool, latent projection and pair-logit kernels; validated on a B200 in round 2."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_scorer_vs_oracle(cuda_dev):
    from oracle import ctclip_oracle as O
    from vit_exp_b200.ct_clip import CTCLIP
    from vit_exp_b200.transformer_maskgit import CTViT
    from vit_exp_b200.zero_shot import ZeroShotScorer
    torch.manual_seed(0)
    vit = CTViT(dim=128, codebook_size=256, image_size=40, patch_size=20, temporal_patch_size=10, spatial_depth=1,
                temporal_depth=1, dim_head=32, heads=4)

    class _Text(torch.nn.Module):
        def forward(self, input_ids, attention_mask=None):
            return (emb[int(input_ids[0, 0])],)

    g = torch.Generator().manual_seed(1)
    emb = [torch.randn(2, 8, 96, generator=g).to(cuda_dev) for _ in range(18)]
    clip = CTCLIP(image_encoder=vit, text_encoder=_Text(), dim_text=96, dim_image=128, dim_latent=64).to(cuda_dev)
    sc = ZeroShotScorer(clip)
    sc.prepare(text_embeds=[(e,) for e in emb])
    vols = [torch.rand(1, 1, 20, 40, 40, generator=g).to(cuda_dev) for _ in range(3)]
    probs = sc.run(3, lambda i: vols[i])
    assert probs.shape == (3, 18)
    for i, v in enumerate(vols):
        tokens = clip.visual_transformer(v, return_encoded_tokens=True).float()
        il = O.image_latent(tokens.cpu().double(), clip.to_visual_latent.weight.detach().cpu().double())
        for p in range(18):
            tl = O.text_latent(emb[p].cpu().double(), clip.to_text_latent.weight.detach().cpu().double())
            want = torch.softmax(O.forward_infer_logits(tl, il, clip.temperature.detach().cpu().double()), 0)[0]
            assert abs(float(probs[i, p]) - float(want)) < 1e-4


def test_eval_graph_replay_matches_eager(cuda_dev):
    """CUDA-graph replay of the no-grad eval forward (CTViT.eval_graphs, on by default): bit-identical to eager launches, across
    different input volumes and after a parameter update (the graph reads the fp32 masters each replay)."""
    from vit_exp_b200.transformer_maskgit import CTViT
    torch.manual_seed(0)
    vit = CTViT(dim=128, codebook_size=256, image_size=40, patch_size=20, temporal_patch_size=10, spatial_depth=1,
                temporal_depth=1, dim_head=32, heads=4).to(cuda_dev).eval()
    g = torch.Generator().manual_seed(1)
    vols = [torch.rand(1, 1, 20, 40, 40, generator=g).to(cuda_dev) for _ in range(5)]
    with torch.no_grad():
        vit.eval_graphs = False
        eager = [vit(v, return_encoded_tokens=True).clone() for v in vols]
        vit.eval_graphs = True
        replay = [vit(v, return_encoded_tokens=True).clone() for v in vols]        # calls 1-2 eager, 3 captures, 4-5 replay
        assert len(vit._graphs) == 1 and next(iter(vit._graphs.values())).fwd is not None
        for a, b in zip(eager, replay):
            assert torch.equal(a, b)
        vit.to_patch_emb[2].weight.mul_(1.01)
        vit.eval_graphs = False
        want = vit(vols[0], return_encoded_tokens=True).clone()
        vit.eval_graphs = True
        assert torch.equal(vit(vols[0], return_encoded_tokens=True), want)
