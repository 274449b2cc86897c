"""Per-kernel parity of the encoder ops (through the C ABI) against the CPU oracle / fp64 autograd."""
import pytest
import torch
import torch.nn.functional as F

from oracle import ctclip_oracle as orc

pytestmark = pytest.mark.gpu


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _rel(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-20)).item()


# ----------------------------------------------------------------------------- patch gather + LN stats
@pytest.mark.parametrize("B,D,H,W,pt,p", [(2, 15, 20, 20, 5, 10), (1, 20, 80, 80, 10, 20), (1, 10, 60, 60, 10, 20)])
def test_patch_norm(cuda_dev, B, D, H, W, pt, p):
    from vit_exp_b200 import ops
    video = torch.rand(B, 1, D, H, W, generator=_g(0))
    video[0, :, : D // 2, : H // 2] = -1.0                       # padded region (data.py:99)
    xhat, mean, rstd = ops.patch_norm_fwd(video.to(cuda_dev), pt, p, p)
    t, h, w = D // pt, H // p, W // p
    x = video.double().reshape(B, 1, t, pt, h, p, w, p).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(B * t * h * w, -1)
    mu = x.mean(dim=1)
    var = ((x - mu[:, None]) ** 2).mean(dim=1)
    ref = (x - mu[:, None]) / torch.sqrt(var + 1e-5)[:, None]
    K = pt * p * p
    assert (xhat[:, :K].double().cpu() - ref).abs().max().item() < 2e-2       # bf16 storage
    assert (xhat[:, K:].float().abs().max().item() if xhat.shape[1] > K else 0.0) == 0.0
    assert (mean.double().cpu() - mu).abs().max().item() < 1e-6
    assert _rel(rstd, 1 / torch.sqrt(var + 1e-5)) < 1e-5


# ----------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("rows,dim,po,pi", [(96, 64, 0, 0), (2 * 3 * 4, 64, 3, 4), (13824, 512, 24, 576), (1000, 768, 0, 0)])
def test_layernorm_fwd_bwd(cuda_dev, rows, dim, po, pi):
    from vit_exp_b200 import ops
    x = torch.randn(rows, dim, generator=_g(1)) * 2 + 0.5
    gamma = 1 + 0.2 * torch.randn(dim, generator=_g(2))
    beta = 0.1 * torch.randn(dim, generator=_g(3))
    xd = x.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    y = orc.layer_norm(xd, gd, bd)
    if pi:
        yperm = y.reshape(-1, po, pi, dim).transpose(1, 2).reshape(rows, dim)
    else:
        yperm = y
    ob, of, raw, mean, rstd = ops.layernorm_fwd(x.to(cuda_dev), gamma.to(cuda_dev), beta.to(cuda_dev), want_f32=True,
                                                want_raw=True, perm_outer=po, perm_inner=pi)
    assert _rel(of, yperm) < 1e-5
    assert _rel(ob, yperm) < 8e-3
    assert _rel(raw, x) < 8e-3
    dy = torch.randn(rows, dim, generator=_g(4))
    (yperm * dy.double()).sum().backward()
    dgamma = torch.zeros(dim, device=cuda_dev)
    dbeta = torch.zeros(dim, device=cuda_dev)
    dx = ops.layernorm_bwd(dy.to(cuda_dev), x.to(cuda_dev), gamma.to(cuda_dev), mean, rstd, dgamma, dbeta,
                           perm_outer=po, perm_inner=pi)
    assert _rel(dx, xd.grad) < 1e-4
    assert _rel(dgamma, gd.grad) < 1e-4
    assert _rel(dbeta, bd.grad) < 1e-4
    # accumulate + bf16 dy
    dx2 = torch.ones(rows, dim, device=cuda_dev)
    ops.layernorm_bwd(dy.to(cuda_dev).bfloat16(), x.to(cuda_dev), gamma.to(cuda_dev), mean, rstd,
                      torch.zeros(dim, device=cuda_dev), None, dx=dx2, accum=True, perm_outer=po, perm_inner=pi)
    assert _rel(dx2 - 1, xd.grad) < 2e-2


def test_layernorm_bwd_broadcast(cuda_dev):
    """gradient of the token mean-pool: every token of a volume receives dpooled / n_tokens."""
    from vit_exp_b200 import ops
    B, n, dim = 2, 48, 64
    x = torch.randn(B * n, dim, generator=_g(1))
    gamma = 1 + 0.2 * torch.randn(dim, generator=_g(2))
    dp = torch.randn(B, dim, generator=_g(3))
    xd, gd = x.double().requires_grad_(True), gamma.double().requires_grad_(True)
    y = orc.layer_norm(xd, gd)
    (y.reshape(B, n, dim).mean(dim=1) * dp.double()).sum().backward()
    _, _, _, mean, rstd = ops.layernorm_fwd(x.to(cuda_dev), gamma.to(cuda_dev), None)
    dgamma = torch.zeros(dim, device=cuda_dev)
    dx = ops.layernorm_bwd(dp.to(cuda_dev), x.to(cuda_dev), gamma.to(cuda_dev), mean, rstd, dgamma, None,
                           bcast_rows=n, dy_scale=1.0 / n)
    assert _rel(dx, xd.grad) < 1e-4
    assert _rel(dgamma, gd.grad) < 1e-4


# ----------------------------------------------------------------------------- PEG
# axis-2 extents <= 32 take the packed-fp32 channel-pair kernel, longer ones the one-channel-per-lane kernel
@pytest.mark.parametrize("shape,dim", [((2, 3, 2, 2), 64), ((1, 5, 4, 6), 64), ((1, 24, 24, 24), 512), ((2, 3, 5, 7), 32),
                                       ((1, 2, 3, 40), 32)])
def test_peg_fwd_bwd(cuda_dev, shape, dim):
    from vit_exp_b200 import ops
    n = shape[0] * shape[1] * shape[2] * shape[3]
    x = torch.randn(n, dim, generator=_g(1))
    w = torch.randn(dim, 1, 3, 3, 3, generator=_g(2)) * 0.2
    b = torch.randn(dim, generator=_g(3)) * 0.1
    xd, wd, bd = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    y = orc.peg(xd, shape, wd, bd) + xd
    dy = torch.randn(n, dim, generator=_g(4))
    (y * dy.double()).sum().backward()
    xc, wc, bc = x.to(cuda_dev), w.to(cuda_dev).reshape(dim, 27).contiguous(), b.to(cuda_dev)
    yg = ops.peg_fwd(xc, wc, bc, shape)
    assert _rel(yg, y) < 1e-5
    dw = torch.zeros(dim, 27, device=cuda_dev)
    db = torch.zeros(dim, device=cuda_dev)
    dx = ops.peg_bwd(dy.to(cuda_dev), xc, wc, shape, dw, db)
    assert _rel(dx, xd.grad) < 1e-5
    assert _rel(dw, wd.grad.reshape(dim, 27)) < 1e-4
    assert _rel(db, bd.grad) < 1e-4


# ----------------------------------------------------------------------------- continuous position bias
@pytest.mark.parametrize("gh,gw,dim,heads", [(2, 2, 64, 2), (24, 24, 512, 8), (3, 5, 64, 4)])
def test_cpb_fwd_bwd(cuda_dev, gh, gw, dim, heads):
    from vit_exp_b200 import ops
    g = _g(5)
    p = {"net.0.0.weight": torch.randn(dim, 2, generator=g) * 0.5, "net.0.0.bias": torch.randn(dim, generator=g) * 0.1,
         "net.1.0.weight": torch.randn(dim, dim, generator=g) * dim ** -0.5, "net.1.0.bias": torch.randn(dim, generator=g) * 0.1,
         "net.2.weight": torch.randn(heads, dim, generator=g) * dim ** -0.5, "net.2.bias": torch.randn(heads, generator=g) * 0.1}
    pd = {k: v.double().requires_grad_(True) for k, v in p.items()}
    bias = orc.cpb_bias(pd, "", gh, gw)                                    # (heads, L, L) as the reference builds it
    cot = torch.randn(bias.shape, generator=_g(6)).double()
    (bias * cot).sum().backward()
    c = {k: v.to(cuda_dev) for k, v in p.items()}
    table, h0, h1 = ops.cpb_fwd(c["net.0.0.weight"], c["net.0.0.bias"], c["net.1.0.weight"], c["net.1.0.bias"],
                                c["net.2.weight"], c["net.2.bias"], gh, gw)
    # gather the full bias from the table and compare with the reference construction
    ys, xs = torch.meshgrid(torch.arange(gh), torch.arange(gw), indexing="ij")
    ys, xs = ys.reshape(-1), xs.reshape(-1)
    oy = ys[:, None] - ys[None, :] + gh - 1
    ox = xs[:, None] - xs[None, :] + gw - 1
    full = table.cpu()[:, oy, ox]
    assert _rel(full, bias) < 1e-5
    # scatter the cotangent into table space (what the attention backward produces)
    dtable = torch.zeros(heads, (2 * gh - 1) * (2 * gw - 1), dtype=torch.float64)
    dtable.index_add_(1, (oy * (2 * gw - 1) + ox).reshape(-1), cot.reshape(heads, -1))
    grads = ops.cpb_bwd(dtable.float().to(cuda_dev).contiguous(), c["net.0.0.weight"], c["net.1.0.weight"],
                        c["net.2.weight"], h0, h1, gh, gw)
    names = ["net.0.0.weight", "net.0.0.bias", "net.1.0.weight", "net.1.0.bias", "net.2.weight", "net.2.bias"]
    for got, n in zip(grads, names):
        assert _rel(got, pd[n].grad) < 2e-4, n


# ----------------------------------------------------------------------------- attention core
def _attn_ref(qkv, table, nseq, L, heads, gh, gw):
    """fp64 reference on the same (bf16-rounded) packed qkv."""
    inner = heads * 32
    q, k, v = qkv.double().reshape(nseq, L, 3, heads, 32).permute(2, 0, 3, 1, 4)
    sim = torch.einsum("bhid,bhjd->bhij", q, k)
    if table is not None:
        ys, xs = torch.meshgrid(torch.arange(gh), torch.arange(gw), indexing="ij")
        ys, xs = ys.reshape(-1), xs.reshape(-1)
        sim = sim + table[:, ys[:, None] - ys[None, :] + gh - 1, xs[:, None] - xs[None, :] + gw - 1]
    out = torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v)
    return out.permute(0, 2, 1, 3).reshape(nseq * L, inner), torch.logsumexp(sim, dim=-1)


# (nseq, L, heads, gh, gw, qmul): 576-token cases run the tcgen05 kernels (9 slices: CTAs whose tile range starts
# inside an item; qmul 40: logits far above the Cauchy-Schwarz fast-path bound -> exact row maximum), 24-token
# cases the TMA-ring + warp-MMA kernels (more sequences than resident CTAs; 2 / 4 / 8 heads), the rest the
# legacy mma.sync / SIMT kernels
@pytest.mark.parametrize("nseq,L,heads,gh,gw,qmul", [(3, 4, 2, 2, 2, 8.0), (2, 576, 8, 24, 24, 8.0), (9, 576, 8, 24, 24, 8.0),
                                                      (3, 576, 4, 24, 24, 40.0), (5, 24, 8, 0, 0, 8.0), (700, 24, 8, 0, 0, 8.0),
                                                      (333, 24, 4, 0, 0, 8.0), (50, 24, 2, 0, 0, 8.0), (4, 3, 2, 0, 0, 8.0),
                                                      (2, 100, 4, 10, 10, 8.0), (2, 40, 2, 0, 0, 8.0)])
def test_attention_fwd_bwd(cuda_dev, nseq, L, heads, gh, gw, qmul):
    from vit_exp_b200 import ops
    inner = heads * 32
    g = _g(7)
    q = F.normalize(torch.randn(nseq * L, heads, 32, generator=g), dim=-1) * qmul * (1 + 0.1 * torch.randn(32, generator=g))
    k = F.normalize(torch.randn(nseq * L, heads, 32, generator=g), dim=-1) * (1 + 0.1 * torch.randn(32, generator=g))
    v = torch.randn(nseq * L, heads, 32, generator=g)
    qkv = torch.cat([q.reshape(-1, inner), k.reshape(-1, inner), v.reshape(-1, inner)], dim=1).bfloat16()
    has_bias = gh > 0
    table = torch.randn(heads, 2 * gh - 1, 2 * gw - 1, generator=g) if has_bias else None
    qd = qkv.double().requires_grad_(True)
    td = table.double().requires_grad_(True) if has_bias else None
    ref, lse_ref = _attn_ref(qd, td, nseq, L, heads, gh, gw)
    dout = torch.randn(nseq * L, inner, generator=g).bfloat16()
    (ref * dout.double()).sum().backward()

    qc = qkv.to(cuda_dev)
    tc = table.to(cuda_dev).contiguous() if has_bias else None
    out, lse = ops.attn_fwd(qc, tc, nseq, L, heads, gh, gw)
    assert _rel(out, ref) < 1.5e-2
    assert (lse.double().cpu() - lse_ref.detach()).abs().max().item() < 2e-3
    dtable = torch.zeros_like(tc) if has_bias else None
    # backward consumes the reference forward output (isolates the backward kernels)
    dqkv = ops.attn_bwd(qc, tc, ref.detach().float().bfloat16().to(cuda_dev), dout.to(cuda_dev), lse, dtable, nseq, L, heads, gh, gw)
    assert _rel(dqkv[:, :inner], qd.grad[:, :inner]) < 2.5e-2
    assert _rel(dqkv[:, inner:2 * inner], qd.grad[:, inner:2 * inner]) < 2.5e-2
    assert _rel(dqkv[:, 2 * inner:], qd.grad[:, 2 * inner:]) < 2.5e-2
    if has_bias:
        assert _rel(dtable, td.grad) < 2e-2


def test_qknorm_bwd(cuda_dev):
    from vit_exp_b200 import ops
    rows, heads, alpha = 200, 4, 8.0
    inner = heads * 32
    g = _g(8)
    raw = torch.randn(rows, 3 * inner, generator=g)
    qs = (1 + 0.2 * torch.randn(32, generator=g))
    ks = (1 + 0.2 * torch.randn(32, generator=g))
    rd, qsd, ksd = raw.double().requires_grad_(True), qs.double().requires_grad_(True), ks.double().requires_grad_(True)
    qn = orc.l2norm(rd[:, :inner].reshape(rows, heads, 32)) * qsd * alpha
    kn = orc.l2norm(rd[:, inner:2 * inner].reshape(rows, heads, 32)) * ksd
    packed = torch.cat([qn.reshape(rows, inner), kn.reshape(rows, inner), rd[:, 2 * inner:]], dim=1)
    dy = torch.randn(rows, 3 * inner, generator=g).bfloat16()
    (packed * dy.double()).sum().backward()
    rn = 1.0 / torch.cat([raw[:, :inner].reshape(rows, heads, 32).norm(dim=-1),
                          raw[:, inner:2 * inner].reshape(rows, heads, 32).norm(dim=-1)], dim=1)
    dq = dy.clone().to(cuda_dev)
    dqs, dks = torch.zeros(32, device=cuda_dev), torch.zeros(32, device=cuda_dev)
    ops.qknorm_bwd_(dq, packed.detach().float().bfloat16().to(cuda_dev), rn.to(cuda_dev).contiguous(), qs.to(cuda_dev),
                    ks.to(cuda_dev), alpha, dqs, dks, heads)
    assert _rel(dq[:, :2 * inner], rd.grad[:, :2 * inner]) < 2e-2
    assert torch.equal(dq[:, 2 * inner:].cpu(), dy[:, 2 * inner:])
    assert _rel(dqs, qsd.grad) < 1e-2
    assert _rel(dks, ksd.grad) < 1e-2


# ----------------------------------------------------------------------------- VQ
def test_vq_search_gather_ema(cuda_dev):
    from vit_exp_b200 import ops
    rows, dim, C = 1000, 512, 8192
    g = _g(9)
    x = torch.randn(rows, dim, generator=g)
    embed = F.normalize(torch.randn(C, dim, generator=g), dim=-1) * (1 + 0.01 * torch.randn(C, 1, generator=g))
    cs = torch.rand(C, generator=g)
    q_ref, ind_ref, emb_ref, cs_ref = orc.vq_cosine(x, embed, training=True, cluster_size=cs)
    xc, ec = x.to(cuda_dev), embed.to(cuda_dev).contiguous()
    ind, quant, xf = ops.vq_search(xc, ec, want_xn_f32=True)
    # index work is exact: the kernel's pick equals the oracle's fp32 arg-max; a row may differ only where the two best
    # codes tie to fp32 rounding (their fp64 similarities closer than 1e-6), which the fp32 oracle cannot order either
    sim = F.normalize(x.double(), dim=-1) @ F.normalize(embed.double(), dim=-1).T
    diff = (ind.cpu() != ind_ref).nonzero()[:, 0]
    assert diff.numel() <= 2, diff.numel()
    for r in diff.tolist():
        assert abs(sim[r, ind[r].item()] - sim[r, ind_ref[r]]) < 1e-6, (r, sim[r, ind[r].item()], sim[r, ind_ref[r]])
    assert (sim.max(dim=1).values - sim.gather(1, ind.cpu()[:, None])[:, 0]).max().item() < 1e-6
    assert torch.equal(quant.cpu(), embed[ind.cpu()])
    # EMA update driven by the oracle's indices (so both sides update the same codes)
    csc = cs.to(cuda_dev).clone()
    emc = ec.clone()
    ops.vq_ema_update_(xf, ind_ref.to(cuda_dev), csc, emc)
    assert _rel(csc, cs_ref) < 1e-5
    assert _rel(emc, emb_ref) < 1e-4


@pytest.mark.parametrize("rows,dim,C", [(4096, 512, 8192), (777, 64, 200), (33, 128, 128)])
def test_vq_search_exact_on_clustered_tokens(cuda_dev, rows, dim, C):
    """Adversarial for a bf16 search: tokens sit next to SEVERAL near-duplicate codes (cosine gaps 1e-4 .. 1e-2, far below
    the bf16 resolution of the similarities), including runs of near-duplicates inside one 128-code block and ragged
    codebook sizes.  The selected code must be the fp64 arg-max up to fp32 ties."""
    from vit_exp_b200 import ops
    g = _g(21)
    base = F.normalize(torch.randn(C // 4 + 1, dim, generator=g), dim=-1)
    embed = base[torch.arange(C) // 4] + torch.randn(C, dim, generator=g) * torch.logspace(-4, -2, C)[torch.randperm(C, generator=g)][:, None]
    embed = embed * (1 + 0.05 * torch.randn(C, 1, generator=g))          # codebook rows are not unit length
    x = 3.0 * embed[torch.randint(0, C, (rows,), generator=g)] + 1e-3 * torch.randn(rows, dim, generator=g)
    ind, quant, _ = ops.vq_search(x.to(cuda_dev), embed.to(cuda_dev).contiguous())
    sim = F.normalize(x.double(), dim=-1) @ F.normalize(embed.double(), dim=-1).T
    picked = sim.gather(1, ind.cpu()[:, None])[:, 0]
    assert (sim.max(dim=1).values - picked).max().item() < 1e-6
    _, ind_ref, _, _ = orc.vq_cosine(x, embed)
    assert (ind.cpu() == ind_ref).float().mean().item() > 0.999
    assert torch.equal(quant.cpu(), embed[ind.cpu()])
