"""Host logic of the sharded zero-shot scorer (vit_exp_b200/zero_shot.py): shares, the padded gather with ragged
shares on 2 gloo ranks, the 2-way softmax convention of scripts/zero_shot.py:83-96,564-568, the prompt list."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_prompts_and_softmax_convention():
    from vit_exp_b200 import zero_shot as Z
    pairs = Z.prompt_pairs()
    assert len(pairs) == 18 and pairs[2] == ("Cardiomegaly is present.", "Cardiomegaly is not present.")
    logits = torch.tensor([2.0, 1.0, -1.0, 3.0])                       # (present, absent) x 2 pathologies
    p = Z.probs_from_logits(logits)
    ref = torch.stack([torch.softmax(logits[0:2], 0)[0], torch.softmax(logits[2:4], 0)[0]])   # apply_softmax(...)[0]
    assert torch.allclose(p, ref)


@pytest.mark.parametrize("n,world", [(10, 4), (3, 8), (16, 8), (3039, 8), (1, 1)])
def test_shard_bounds_partition(n, world):
    from vit_exp_b200.zero_shot import shard_bounds
    spans = [shard_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out, n):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vit_exp_b200 import zero_shot as Z

    class _Scorer(Z.ZeroShotScorer):             # the per-volume arithmetic is libctk's (GPU suite); here: volume i -> row i
        def score_many(self, volumes):
            return volumes.reshape(volumes.shape[0], -1)[:, :18] * 2.0

    sc = _Scorer(clip=None)
    sc.prompt_latents = torch.zeros(36, 4)
    loaded = []

    def load(i):
        loaded.append(i)
        return torch.full((1, 1, 2, 3, 3), float(i))
    res = sc.run(n, load, batch_size=2)
    torch.save(dict(res=res, loaded=loaded), f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [5, 4, 1])
def test_two_rank_sharded_run(tmp_path, n):
    world = 2
    out = str(tmp_path / "zs")
    mp.spawn(_worker, args=(world, _free_port(), out, n), nprocs=world, join=True)
    want = torch.arange(n, dtype=torch.float32)[:, None].expand(n, 18) * 2.0
    seen = []
    for r in range(world):
        d = torch.load(f"{out}.{r}")
        assert torch.equal(d["res"], want)               # every rank holds the full matrix, in volume order
        seen += d["loaded"]
    assert sorted(seen) == list(range(n))                # every volume scored exactly once


def test_scorer_matches_forward_infer_math(monkeypatch):
    """ZeroShotScorer (prompt latents cached, one pair_logits launch per volume) against the reference's per-pathology
    forward_infer + apply_softmax (oracle restatement), with torch doubles standing in for the kernels."""
    import emulated_ops as E
    from oracle import ctclip_oracle as O
    from vit_exp_b200 import ct_clip as CC
    from vit_exp_b200 import zero_shot as Z
    monkeypatch.setattr(CC, "ops", E)
    monkeypatch.setattr(Z, "ops", E)
    g = torch.Generator().manual_seed(0)
    tokens = torch.randn(1, 3, 2, 3, 64, generator=g)

    class _Vit(torch.nn.Module):
        def forward(self, video, return_encoded_tokens=False):
            return tokens

    class _Text(torch.nn.Module):
        def forward(self, input_ids, attention_mask=None):
            return (emb[int(input_ids[0, 0])],)

    emb = [torch.randn(2, 8, 48, generator=g) for _ in range(18)]
    clip = CC.CTCLIP(image_encoder=_Vit(), text_encoder=_Text(), dim_text=48, dim_image=64, dim_latent=32)
    with torch.no_grad():
        clip.temperature.fill_(0.7)
    sc = Z.ZeroShotScorer(clip)
    sc.prepare(text_embeds=[(e,) for e in emb])
    got = sc.score(torch.zeros(1, 1, 6, 8, 12))
    many = sc.score_many(torch.zeros(1, 1, 6, 8, 12))
    assert many.shape == (1, 18) and torch.equal(many[0], got)
    il = O.image_latent(tokens, clip.to_visual_latent.weight.detach())
    want = []
    for e in emb:
        tl = O.text_latent(e, clip.to_text_latent.weight.detach())
        want.append(torch.softmax(O.forward_infer_logits(tl, il, clip.temperature.detach()), dim=0)[0])
    assert torch.allclose(got, torch.stack(want), atol=1e-6)
