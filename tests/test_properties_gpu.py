"""Size-independent properties at BASELINE sizes and edge cases (through the C ABI)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import ctclip_oracle as orc

pytestmark = pytest.mark.gpu


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _lat(n, d, seed, dev):
    return F.normalize(torch.randn(n, d, generator=_g(seed)), dim=-1).to(dev)


def test_loss_single_sample_is_zero(cuda_dev):
    """B = 1 per process and one process: the contrastive loss is identically 0."""
    from vit_exp_b200 import ops
    out, dl = ops.clip_loss_fwd_bwd(_lat(1, 512, 0, cuda_dev), _lat(1, 512, 1, cuda_dev), torch.ones(1, device=cuda_dev), 1, 0)
    assert abs(out[0].item()) < 1e-7 and abs(out[1].item()) < 1e-7
    assert dl.abs().max().item() < 1e-7


def test_loss_permutation_invariance_and_gradient_sums(cuda_dev):
    """N = 4096 (BASELINE config 5 maximum): permuting the pairs leaves the loss unchanged; rows of
    G sum to zero so sum_i dT_i . T_i + dI_i . I_i == d loss / d log-temperature."""
    from vit_exp_b200 import ops
    N, d = 4096, 512
    T, I = _lat(N, d, 0, cuda_dev), _lat(N, d, 1, cuda_dev)
    lt = torch.ones(1, device=cuda_dev)
    out, dl = ops.clip_loss_fwd_bwd(T, I, lt, N, 0)
    perm = torch.randperm(N, generator=_g(2)).to(cuda_dev)
    out_p, _ = ops.clip_loss_fwd_bwd(T[perm].contiguous(), I[perm].contiguous(), lt, N, 0)
    assert abs(out[0].item() - out_p[0].item()) < 1e-6 * abs(out[0].item())
    # d loss/d tau = sum_ij G_ij S_ij = sum_i <dT_i, T_i>  (Euler: the loss is a function of s*T only through S)
    lhs = (dl[0] * T).sum().item()
    assert abs(lhs - out[1].item()) <= 1e-4 * abs(out[1].item()) + 1e-7
    # rank slices of a W=8 run tile the W=1 gradients exactly (AllGather backward = local slice)
    B = N // 8
    _, dl3 = ops.clip_loss_fwd_bwd(T, I, lt, B, 3 * B)
    assert torch.allclose(dl3[0], dl[0][3 * B:4 * B] * 8, rtol=1e-3, atol=1e-10)    # loss carries 1/b_local


def test_loss_stable_beyond_reference_overflow(cuda_dev):
    """log-temperature 4.4: reference still finite -> equal; 6.0: the reference's unstabilised exp
    overflows fp32 (ct_clip.py:1358) while the max-subtracted kernel matches exact arithmetic."""
    from vit_exp_b200 import ops
    N, d = 64, 512
    T = _lat(N, d, 0, cuda_dev)
    # nearly matched pairs: diagonal logits reach ~exp(log_temp)
    I = F.normalize(T + 0.05 * torch.randn(N, d, generator=_g(9)).to(cuda_dev), dim=-1)
    for lt, ref_finite in ((4.4, True), (6.0, False)):
        out, _ = ops.clip_loss_fwd_bwd(T, I, torch.full((1,), lt, device=cuda_dev), 8, 0)
        ref32 = orc.clip_loss_reference_form(T.cpu(), I.cpu(), torch.tensor(lt), 8)
        exact = orc.clip_loss_open_clip(T.cpu().double(), I.cpu().double(), torch.tensor(lt).double().exp()) / 8
        assert torch.isfinite(ref32).item() == ref_finite
        assert torch.isfinite(out[0]).item()
        assert abs(out[0].item() - exact.item()) <= 1e-4 * abs(exact.item()) + 1e-7


def test_peg_is_causal_along_axis0(cuda_dev):
    """changing plane a0 = k must not change outputs of planes < k (attention.py:80-82 causal pad)."""
    from vit_exp_b200 import ops
    shape, dim = (1, 24, 24, 24), 512
    n = 24 ** 3
    x = torch.randn(n, dim, generator=_g(1)).to(cuda_dev)
    w = (torch.randn(dim, 27, generator=_g(2)) * 0.2).to(cuda_dev)
    b = torch.zeros(dim, device=cuda_dev)
    y0 = ops.peg_fwd(x, w, b, shape)
    x2 = x.clone()
    x2.view(24, 24 * 24, dim)[17] += 1.0
    y1 = ops.peg_fwd(x2, w, b, shape)
    d = (y1 - y0).view(24, -1).abs().max(dim=1).values
    assert d[:17].max().item() == 0.0 and d[17].item() > 0 and d[19].item() > 0 and d[20:].max().item() == 0.0


def test_layernorm_permutation_round_trip(cuda_dev):
    """spatial->temporal->spatial token transposition (ctvit.py:301,305) is the identity."""
    from vit_exp_b200 import ops
    B, t, hw, dim = 2, 24, 576, 512
    x = torch.randn(B * t * hw, dim, generator=_g(3)).to(cuda_dev)
    one = torch.ones(dim, device=cuda_dev)
    _, y, _, _, _ = ops.layernorm_fwd(x, one, None, want_bf16=False, want_f32=True, perm_outer=t, perm_inner=hw)
    ref = F.layer_norm(x, (dim,)).view(B, t, hw, dim).transpose(1, 2).reshape(-1, dim)
    assert (y - ref).abs().max().item() < 1e-5
    _, z, _, _, _ = ops.layernorm_fwd(y, one, None, want_bf16=False, want_f32=True, perm_outer=hw, perm_inner=t)
    ref2 = F.layer_norm(F.layer_norm(x, (dim,)), (dim,))
    assert (z - ref2).abs().max().item() < 1e-5


def test_vq_idempotent_on_codebook_rows(cuda_dev):
    """quantising codebook rows returns the rows themselves (full 8192 x 512 codebook)."""
    from vit_exp_b200 import ops
    C, dim = 8192, 512
    embed = (F.normalize(torch.randn(C, dim, generator=_g(4)), dim=-1) * 1.01).to(cuda_dev).contiguous()
    ind, quant, _ = ops.vq_search(embed, embed)
    assert torch.equal(ind.cpu(), torch.arange(C))
    assert torch.equal(quant, embed)


def test_attention_rows_are_convex_combinations(cuda_dev):
    """softmax(.) v at full spatial size: with v == const per head the output equals that constant."""
    from vit_exp_b200 import ops
    nseq, L, heads = 4, 576, 8
    inner = heads * 32
    g = _g(5)
    q = F.normalize(torch.randn(nseq * L, heads, 32, generator=g), dim=-1) * 8
    k = F.normalize(torch.randn(nseq * L, heads, 32, generator=g), dim=-1)
    vconst = torch.randn(heads, 32, generator=g)
    v = vconst.expand(nseq * L, heads, 32)
    qkv = torch.cat([q.reshape(-1, inner), k.reshape(-1, inner), v.reshape(-1, inner)], dim=1).bfloat16().to(cuda_dev)
    table = torch.randn(heads, 47, 47, generator=g).to(cuda_dev)
    out, lse = ops.attn_fwd(qkv, table, nseq, L, heads, 24, 24)
    ref = vconst.bfloat16().float().reshape(1, inner)
    assert (out.float().cpu() - ref).abs().max().item() < 2e-2
    assert torch.isfinite(lse).all()


def test_ctvit_batch_independence(cuda_dev):
    """encoder output of a volume does not depend on what else is in the batch (data-parallel path)."""
    from vit_exp_b200.transformer_maskgit import CTViT
    torch.manual_seed(0)
    vit = CTViT(dim=128, codebook_size=512, image_size=80, patch_size=20, temporal_patch_size=10, spatial_depth=1,
                temporal_depth=1, dim_head=32, heads=4).to(cuda_dev).eval()
    v = torch.rand(3, 1, 40, 80, 80, generator=_g(6)).to(cuda_dev)
    with torch.no_grad():
        _, _, pre_all = vit.encode_with_aux(v)
        _, _, pre_one = vit.encode_with_aux(v[1:2].contiguous())
    n = pre_one.shape[0]
    assert torch.equal(pre_all[n:2 * n], pre_one)


def test_bias_table_gradient_tcgen05_variant(cuda_dev):
    """The tcgen05 bias-gradient kernel (default; dS summed over slices by an identity MMA into TMEM) agrees with
    the mma.sync kernel (CTK_DBIAS_TC=0); the switch is read once per process, hence the subprocess."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, torch, torch.nn.functional as F
sys.path.insert(0, %r)
from vit_exp_b200 import ops
g = torch.Generator().manual_seed(3)
nseq, L, heads = 5, 576, 8
inner = heads * 32
q = F.normalize(torch.randn(nseq * L, heads, 32, generator=g), dim=-1) * 8
k = F.normalize(torch.randn(nseq * L, heads, 32, generator=g), dim=-1)
v = torch.randn(nseq * L, heads, 32, generator=g)
qkv = torch.cat([q.reshape(-1, inner), k.reshape(-1, inner), v.reshape(-1, inner)], dim=1).bfloat16().cuda()
table = torch.randn(heads, 47, 47, generator=g).cuda()
dout = torch.randn(nseq * L, inner, generator=g).bfloat16().cuda()
out, lse = ops.attn_fwd(qkv, table, nseq, L, heads, 24, 24)
dtable = torch.zeros_like(table)
dqkv = ops.attn_bwd(qkv, table, out, dout, lse, dtable, nseq, L, heads, 24, 24)
torch.cuda.synchronize()
torch.save({"dtable": dtable.cpu(), "dqkv": dqkv.float().cpu()}, sys.argv[1])
''' % root
    import tempfile
    res = {}
    with tempfile.TemporaryDirectory() as td:
        for flag in ("0", "1"):
            path = os.path.join(td, f"r{flag}.pt")
            env = dict(os.environ, CTK_DBIAS_TC=flag)
            subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=300)
            res[flag] = torch.load(path)
    a, b = res["0"]["dtable"], res["1"]["dtable"]
    assert torch.isfinite(b).all()
    assert ((a - b).norm() / a.norm()).item() < 1e-2        # bf16 dS in the tcgen05 variant, fp32 dS in the default
    assert torch.equal(res["0"]["dqkv"], res["1"]["dqkv"])


def test_cuda_graph_replay_matches_eager(cuda_dev):
    """After two eager steps the encoder forward/backward are replayed from CUDA graphs; tokens, loss-side gradient
    flow and parameter gradients must match the eager launches step for step (same kernels, same order)."""
    from vit_exp_b200.transformer_maskgit import CTViT
    res = {}
    for graphs in (False, True):
        torch.manual_seed(0)
        vit = CTViT(dim=128, codebook_size=256, image_size=80, patch_size=20, temporal_patch_size=10, spatial_depth=1,
                    temporal_depth=1, dim_head=32, heads=4).to(cuda_dev).eval()      # eval: the codebook stays fixed
        vit.cuda_graphs = graphs
        vids = [torch.rand(2, 1, 30, 80, 80, generator=_g(10 + i)).to(cuda_dev) for i in range(2)]
        w = torch.randn(2, 3, 4, 4, 128, generator=_g(20)).to(cuda_dev)
        outs = []
        for step in range(5):
            for q in vit.parameters():
                q.grad = None
            tokens = vit(vids[step % 2], return_encoded_tokens=True)
            (tokens * w).sum().backward()
            outs.append((tokens.detach().clone(), {n: q.grad.detach().clone() for n, q in vit.named_parameters()
                                                    if q.grad is not None}))
        res[graphs] = outs
        if graphs:
            assert any(eg.fwd is not None and eg.bwd for eg in vit._graphs.values()), "graphs were never captured"
    # two forwards in flight (micro-batches) before any backward: the second one must not clobber the first one's
    # saved activations (it falls back to eager launches)
    torch.manual_seed(0)
    vit = CTViT(dim=128, codebook_size=256, image_size=80, patch_size=20, temporal_patch_size=10, spatial_depth=1,
                temporal_depth=1, dim_head=32, heads=4).to(cuda_dev).eval()
    vids = [torch.rand(2, 1, 30, 80, 80, generator=_g(10 + i)).to(cuda_dev) for i in range(2)]
    w = torch.randn(2, 3, 4, 4, 128, generator=_g(20)).to(cuda_dev)
    for _ in range(4):                                        # reach the replay regime
        for q in vit.parameters():
            q.grad = None
        (vit(vids[0], return_encoded_tokens=True) * w).sum().backward()
    ref = {}
    for k in range(2):
        for q in vit.parameters():
            q.grad = None
        (vit(vids[k], return_encoded_tokens=True) * w).sum().backward()
        ref[k] = {n: q.grad.detach().clone() for n, q in vit.named_parameters() if q.grad is not None}
    for q in vit.parameters():
        q.grad = None
    la = (vit(vids[0], return_encoded_tokens=True) * w).sum()
    lb = (vit(vids[1], return_encoded_tokens=True) * w).sum()
    la.backward()
    ga = {n: q.grad.detach().clone() for n, q in vit.named_parameters() if q.grad is not None}
    for q in vit.parameters():
        q.grad = None
    lb.backward()
    gb = {n: q.grad.detach().clone() for n, q in vit.named_parameters() if q.grad is not None}
    for got, want in ((ga, ref[0]), (gb, ref[1])):
        for n in want:
            d = (got[n].double() - want[n].double()).norm() / want[n].double().norm().clamp_min(1e-30)
            assert d.item() < 1e-4, n
    for step in range(5):
        t0, g0 = res[False][step]
        t1, g1 = res[True][step]
        assert torch.equal(t0, t1), f"tokens differ at step {step}"
        assert g0.keys() == g1.keys()
        for n in g0:
            d = (g1[n].double() - g0[n].double()).norm() / g0[n].double().norm().clamp_min(1e-30)
            assert d.item() < 1e-4, (step, n)                                 # atomics order only
