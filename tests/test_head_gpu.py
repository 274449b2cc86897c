"""Contrastive head kernels (pool, latent projection, fused InfoNCE fwd+bwd) vs the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lat(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1)


def test_clip_loss_reference_fixture(cuda_dev):
    """demo_tests/test_loss_type.py:14-15 fixed vectors; known answers from SURVEY.md 8c."""
    from vit_exp_b200 import ops
    x = torch.tensor([[.1, .2], [.3, .4], [.5, .6], [.7, .8]], device=cuda_dev)
    y = torch.tensor([[.2, .3], [.3, .6], [.4, .9], [.2, .5]], device=cuda_dev)
    out, _ = ops.clip_loss_fwd_bwd(x, y, torch.zeros(1, device=cuda_dev), b_local=4, row0=0)
    assert abs(out[0].item() - 0.34436899) < 2e-6
    out, _ = ops.clip_loss_fwd_bwd(x, y, torch.ones(1, device=cuda_dev), b_local=4, row0=0)
    assert abs(out[0].item() - 0.35895544) < 2e-6


@pytest.mark.parametrize("N,W,rank", [(8, 1, 0), (64, 8, 3), (256, 4, 1), (1000, 8, 7), (4096, 8, 5)])
def test_clip_loss_vs_oracle(cuda_dev, N, W, rank):
    from oracle import ctclip_oracle as orc
    from vit_exp_b200 import ops
    d = 512
    B = N // W
    T, I = _lat(N, d, 0), _lat(N, d, 1)
    lt = torch.tensor(1.0)
    ref = orc.clip_loss_and_local_grads(T.double(), I.double(), lt.double(), B, rank)
    out, dl = ops.clip_loss_fwd_bwd(T.to(cuda_dev), I.to(cuda_dev), lt.reshape(1).to(cuda_dev), b_local=B, row0=rank * B)
    out, dl = out.cpu().double(), dl.cpu().double()
    assert abs(out[0] - ref["loss"]) / abs(ref["loss"]) < 1e-5
    assert abs(out[1] - ref["dlog_temp"]) <= 1e-5 * abs(ref["dlog_temp"]) + 1e-7
    for got, want in ((dl[0], ref["dT_local"]), (dl[1], ref["dI_local"])):
        assert (got - want).abs().max() <= 1e-4 * want.abs().max() + 1e-9


def test_pool_latent_fwd_bwd(cuda_dev):
    from oracle import ctclip_oracle as orc
    from vit_exp_b200 import ops
    g = torch.Generator().manual_seed(0)
    B, n, dim, dl = 3, 1000, 512, 512
    x = torch.randn(B, n, dim, generator=g)
    W = torch.randn(dl, dim, generator=g) * dim ** -0.5
    dlat = torch.randn(B, dl, generator=g)
    ref = orc.pooled_latent_fwd_bwd(x.double(), W.double(), dlat.double())
    xc, Wc = x.to(cuda_dev), W.to(cuda_dev)
    pooled = ops.mean_pool(xc)
    lat, rn = ops.latent_fwd(pooled, Wc)
    dW, dpool = ops.latent_bwd(dlat.to(cuda_dev), lat, rn, pooled, Wc)
    assert (pooled.cpu().double() - ref["pooled"]).abs().max() < 1e-5
    assert (lat.cpu().double() - ref["latent"]).abs().max() < 1e-5
    assert (dW.cpu().double() - ref["dW"]).abs().max() <= 1e-4 * ref["dW"].abs().max()
    assert (dpool.cpu().double() - ref["dpooled"]).abs().max() <= 1e-4 * ref["dpooled"].abs().max()


def test_latent_strided_cls_rows(cuda_dev):
    """text branch: CLS rows enc_text[:, 0, :] (ct_clip.py:1309) are read in place (row stride L*768)."""
    from vit_exp_b200 import ops
    g = torch.Generator().manual_seed(1)
    enc = torch.randn(4, 16, 768, generator=g).to(cuda_dev)
    W = (torch.randn(512, 768, generator=g) * 768 ** -0.5).to(cuda_dev)
    lat, _ = ops.latent_fwd(enc[:, 0, :], W)
    ref = torch.nn.functional.normalize(enc[:, 0, :] @ W.T, dim=-1)
    assert (lat - ref).abs().max().item() < 1e-5


def test_pair_logits(cuda_dev):
    from vit_exp_b200 import ops
    t, i = _lat(36, 512, 0).to(cuda_dev), _lat(1, 512, 1).to(cuda_dev)
    lt = torch.tensor([1.0], device=cuda_dev)
    out = ops.pair_logits(t, i[0].contiguous(), lt)
    assert (out - (t @ i[0]) * lt.exp()).abs().max().item() < 1e-5


# ----------------------------------------------------------------------------- optimizer tail
@pytest.mark.gpu
@pytest.mark.parametrize("wd,max_norm", [(0.0, 0.5), (0.0, None), (0.01, 1e-3)])
def test_fused_clip_adam_matches_torch(cuda_dev, wd, max_norm):
    """FusedClipAdam == clip_grad_norm_ + torch.optim.Adam / AdamW (optimizer.py:14-24, CTCLIPTrainer.py:711-715)
    over tensors of awkward sizes (unaligned tails, > 1 chunk, a parameter without gradient), 3 steps."""
    import torch
    from vit_exp_b200.optim import FusedClipAdam
    g = torch.Generator().manual_seed(11)
    shapes = [(3,), (1,), (257, 129), (70001,), (64, 1024), (5, 7, 11)]
    ref_p = [torch.nn.Parameter(torch.randn(*s, generator=g).to(cuda_dev)) for s in shapes] + \
            [torch.nn.Parameter(torch.randn(9, generator=g).to(cuda_dev))]                     # never gets a gradient
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    ref_opt = (torch.optim.AdamW if wd > 0 else torch.optim.Adam)(ref_p, lr=1e-2, betas=(0.9, 0.99), weight_decay=wd)
    our_opt = FusedClipAdam(our_p, lr=1e-2, betas=(0.9, 0.99), weight_decay=wd, max_grad_norm=max_norm)
    for step in range(3):
        grads = [torch.randn(*s, generator=g).to(cuda_dev) * (10.0 if step == 1 else 0.01) for s in shapes]
        for p, q, gr in zip(ref_p, our_p, grads):
            p.grad = gr.clone()
            q.grad = gr.clone()
        norm_ref = None
        if max_norm is not None:
            norm_ref = torch.nn.utils.clip_grad_norm_([p for p in ref_p if p.grad is not None], max_norm)
        ref_opt.step()
        norm = our_opt.step()
        if max_norm is not None:
            assert abs(norm.item() - norm_ref.item()) <= 1e-5 * norm_ref.item()
        for p, q in zip(ref_p, our_p):
            assert torch.allclose(p, q, rtol=2e-5, atol=1e-7), (step, tuple(p.shape), (p - q).abs().max().item())
    assert torch.equal(ref_p[-1], our_p[-1])


@pytest.mark.gpu
def test_fused_clip_adam_state_dict_round_trip(cuda_dev):
    """resume: state_dict() -> a fresh optimizer -> load_state_dict() -> step continues with the right bias correction
    (the step count lives in the per-parameter state like torch.optim.Adam's, so the two are interchangeable), and a
    parameter whose first gradient arrives late keeps its own step count."""
    import torch
    from vit_exp_b200.optim import FusedClipAdam
    g = torch.Generator().manual_seed(5)
    shapes = [(33,), (130, 17), (4099,)]
    ref_p = [torch.nn.Parameter(torch.randn(*s, generator=g).to(cuda_dev)) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    kw = dict(lr=1e-2, betas=(0.9, 0.99))
    ref_opt = torch.optim.Adam(ref_p, **kw)
    our_opt = FusedClipAdam(our_p, **kw)

    def both_step(ro, oo, active):
        for i, (p, q) in enumerate(zip(ref_p, our_p)):
            gr = torch.randn(*p.shape, generator=g).to(cuda_dev) if i in active else None
            p.grad = None if gr is None else gr.clone()
            q.grad = None if gr is None else gr.clone()
        ro.step()
        oo.step()
    both_step(ref_opt, our_opt, {0, 1})                 # parameter 2 joins one step late
    both_step(ref_opt, our_opt, {0, 1, 2})
    sd = our_opt.state_dict()
    assert float(sd["state"][0]["step"]) == 2.0 and float(sd["state"][2]["step"]) == 1.0
    resumed = FusedClipAdam(our_p, **kw)
    resumed.load_state_dict(sd)
    both_step(ref_opt, resumed, {0, 1, 2})
    for p, q in zip(ref_p, our_p):
        assert torch.allclose(p, q, rtol=2e-5, atol=1e-7), (tuple(p.shape), (p - q).abs().max().item())
    # torch's own Adam state loads into FusedClipAdam (same keys: step / exp_avg / exp_avg_sq)
    swapped = FusedClipAdam(our_p, **kw)
    for q, p in zip(our_p, ref_p):
        q.data.copy_(p.data)
    import copy
    tsd = copy.deepcopy(ref_opt.state_dict())      # load_state_dict keeps same-dtype/device tensors by reference
    swapped.load_state_dict({"state": tsd["state"], "param_groups": resumed.state_dict()["param_groups"]})
    both_step(ref_opt, swapped, {0, 1, 2})
    for p, q in zip(ref_p, our_p):
        assert torch.allclose(p, q, rtol=2e-5, atol=1e-7)
