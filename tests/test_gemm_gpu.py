"""tcgen05 GEMM (ctk_gemm_bf16) against fp32 matmuls of the same bf16-rounded operands."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, dev, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev)


def _relerr(got, ref):
    return ((got.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-20)).item()


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 512), (256, 512, 512), (300, 352, 200),
                                    (13824, 512, 1408), (13824, 2816, 512)])
def test_gemm_f32_kmajor(cuda_dev, M, N, K):
    from vit_exp_b200 import ops
    a = _rand((M, K), cuda_dev, 1).bfloat16()
    b = _rand((N, K), cuda_dev, 2).bfloat16()
    bias = _rand((N,), cuda_dev, 3)
    c = torch.full((M, N), float("nan"), device=cuda_dev)
    ops.gemm(a, b, ops.EPI_F32, c, M=M, N=N, K=K, bias=bias)
    ref = a.float() @ b.float().T + bias
    assert _relerr(c, ref) < 2e-5


def test_gemm_bf16_alpha(cuda_dev):
    from vit_exp_b200 import ops
    M, N, K = 384, 768, 256
    a = _rand((M, K), cuda_dev, 1).bfloat16()
    b = _rand((N, K), cuda_dev, 2).bfloat16()
    c = torch.empty(M, N, dtype=torch.bfloat16, device=cuda_dev)
    ops.gemm(a, b, ops.EPI_BF16, c, M=M, N=N, K=K, alpha=0.5)
    ref = 0.5 * (a.float() @ b.float().T)
    assert _relerr(c, ref) < 8e-3


def test_gemm_resid_inplace(cuda_dev):
    from vit_exp_b200 import ops
    M, N, K = 512, 512, 256
    a = _rand((M, K), cuda_dev, 1).bfloat16()
    b = _rand((N, K), cuda_dev, 2).bfloat16()
    x = _rand((M, N), cuda_dev, 3)
    ref = a.float() @ b.float().T + x
    ops.gemm(a, b, ops.EPI_RESID_F32, x, M=M, N=N, K=K, resid=x)
    assert _relerr(x, ref) < 2e-5


def test_gemm_padded_lda(cuda_dev):
    """K = 1365 logical columns inside a 1408-pitch buffer (FeedForward inner dim, attention.py:51)."""
    from vit_exp_b200 import ops
    M, N, K, ld = 256, 512, 1365, 1408
    a = torch.zeros(M, ld, dtype=torch.bfloat16, device=cuda_dev)
    b = torch.zeros(N, ld, dtype=torch.bfloat16, device=cuda_dev)
    a[:, :K] = _rand((M, K), cuda_dev, 1).bfloat16()
    b[:, :K] = _rand((N, K), cuda_dev, 2).bfloat16()
    c = torch.empty(M, N, device=cuda_dev)
    ops.gemm(a, b, ops.EPI_F32, c, M=M, N=N, K=ld)
    ref = a.float() @ b.float().T
    assert _relerr(c, ref) < 2e-5


def test_gemm_geglu_fwd_bwd(cuda_dev):
    from vit_exp_b200 import ops
    M, dim, inner, inner_pad = 384, 512, 300, 384
    x = _rand((M, dim), cuda_dev, 1).bfloat16()
    w1 = _rand((2 * inner, dim), cuda_dev, 2, scale=dim ** -0.5)
    w1p, w1t, row_map = ops.pack_ff_w1(w1, inner, inner_pad)
    U = torch.empty(M, 2 * inner_pad, dtype=torch.bfloat16, device=cuda_dev)
    H = torch.empty(M, inner_pad, dtype=torch.bfloat16, device=cuda_dev)
    ops.gemm(x, w1p, ops.EPI_GEGLU, U, M=M, N=2 * inner_pad, K=dim, aux0=H, ld_aux0=inner_pad)
    y = x.float() @ w1.bfloat16().float().T
    val, gate = y[:, :inner], y[:, inner:]
    h_ref = torch.nn.functional.gelu(gate) * val
    assert _relerr(H[:, :inner], h_ref) < 1e-2
    assert H[:, inner:].abs().max().item() == 0.0
    # interleaved pre-activations: block k holds value units [128k,128k+128) then their gates
    Ub = U.float().view(M, inner_pad // 128, 2, 128)
    assert _relerr(Ub[:, :, 0].reshape(M, -1)[:, :inner], val) < 1e-2
    assert _relerr(Ub[:, :, 1].reshape(M, -1)[:, :inner], gate) < 1e-2
    # row_map points back at the source rows
    rm = row_map.cpu().view(inner_pad // 128, 2, 128)
    assert rm[0, 0, 5].item() == 5 and rm[0, 1, 5].item() == inner + 5 and rm[-1, 0, -1].item() == -1
    assert torch.equal(w1t, w1p.T.contiguous())

    # backward epilogue: dH = dY W2  ->  dU
    w2 = _rand((dim, inner), cuda_dev, 4, scale=inner ** -0.5)
    w2t = ops.transpose_cast_bf16(w2, ld=dim)            # [inner, dim]  (B operand: N=inner_pad rows)
    w2t_pad = torch.zeros(inner_pad, dim, dtype=torch.bfloat16, device=cuda_dev)
    w2t_pad[:inner] = w2t[:inner]
    dy = _rand((M, dim), cuda_dev, 5).bfloat16()
    dU = torch.empty_like(U)
    ops.gemm(dy, w2t_pad, ops.EPI_GEGLU_BWD, dU, M=M, N=inner_pad, K=dim, aux0=U, ld_aux0=2 * inner_pad)
    dh = dy.float() @ w2.bfloat16().float()              # [M, inner]
    Uv = Ub[:, :, 0].reshape(M, -1)[:, :inner]
    Ug = Ub[:, :, 1].reshape(M, -1)[:, :inner].clone().requires_grad_(True)
    gl = torch.nn.functional.gelu(Ug)
    dgl, = torch.autograd.grad(gl, Ug, torch.ones_like(gl))
    dv_ref = dh * gl.detach()
    dg_ref = dh * Uv * dgl
    dUb = dU.float().view(M, inner_pad // 128, 2, 128)
    assert _relerr(dUb[:, :, 0].reshape(M, -1)[:, :inner], dv_ref) < 1e-2
    assert _relerr(dUb[:, :, 1].reshape(M, -1)[:, :inner], dg_ref) < 1e-2


def test_gemm_qkv(cuda_dev):
    """q from LayerNorm(x), k/v from raw x (attention.py:145-149) -> two GEMMs into one packed buffer."""
    from vit_exp_b200 import ops
    M, dim, heads = 256, 512, 8
    inner = heads * 32
    xn = _rand((M, dim), cuda_dev, 1).bfloat16()
    xr = _rand((M, dim), cuda_dev, 5).bfloat16()
    wq = _rand((inner, dim), cuda_dev, 2, scale=dim ** -0.5).bfloat16()
    wkv = _rand((2 * inner, dim), cuda_dev, 6, scale=dim ** -0.5).bfloat16()
    qs = _rand((32,), cuda_dev, 3).abs() + 0.5
    ks = _rand((32,), cuda_dev, 4).abs() + 0.5
    out = torch.empty(M, 3 * inner, dtype=torch.bfloat16, device=cuda_dev)
    rn = torch.empty(M, 2 * heads, device=cuda_dev)
    ops.gemm(xn, wq, ops.EPI_QKV, out, M=M, N=inner, K=dim, aux0=rn, ld_aux0=2 * heads, vec0=qs, alpha=8.0,
             i0=inner, i1=0)
    ops.gemm(xr, wkv, ops.EPI_QKV, out, M=M, N=2 * inner, K=dim, aux0=rn, ld_aux0=2 * heads, vec0=ks, alpha=1.0,
             i0=inner, i1=inner)
    q = xn.float() @ wq.float().T
    k, v = (xr.float() @ wkv.float().T).split(inner, dim=1)
    qn = torch.nn.functional.normalize(q.view(M, heads, 32), dim=-1) * qs * 8.0
    kn = torch.nn.functional.normalize(k.view(M, heads, 32), dim=-1) * ks
    assert _relerr(out[:, :inner], qn.reshape(M, -1)) < 1e-2
    assert _relerr(out[:, inner:2 * inner], kn.reshape(M, -1)) < 1e-2
    assert _relerr(out[:, 2 * inner:], v) < 1e-2
    rn_ref = 1.0 / torch.cat([q.view(M, heads, 32).norm(dim=-1), k.view(M, heads, 32).norm(dim=-1)], dim=1)
    assert _relerr(rn, rn_ref) < 1e-4


@pytest.mark.parametrize("Kt,Mo,No,split", [(64, 128, 256, 1), (1024, 256, 512, 0), (13824, 512, 1408, 0), (4000, 300, 1365, 3)])
def test_gemm_wgrad_mnmajor(cuda_dev, Kt, Mo, No, split):
    """dW[Mo, No] = dY[Kt, Mo]^T X[Kt, No]: both operands MN-major, split-K with fp32 atomics."""
    from vit_exp_b200 import ops
    ldm, ldn = (Mo + 7) // 8 * 8, (No + 7) // 8 * 8
    dy = torch.zeros(Kt, ldm, dtype=torch.bfloat16, device=cuda_dev)
    x = torch.zeros(Kt, ldn, dtype=torch.bfloat16, device=cuda_dev)
    dy[:, :Mo] = _rand((Kt, Mo), cuda_dev, 1).bfloat16()
    x[:, :No] = _rand((Kt, No), cuda_dev, 2).bfloat16()
    out = torch.zeros(Mo, No, device=cuda_dev)
    perm = torch.randperm(Mo, generator=torch.Generator().manual_seed(0)).to(torch.int32).to(cuda_dev)
    ops.gemm(dy, x, ops.EPI_ATOMIC_F32, out, M=Mo, N=No, K=Kt, mn_major=True, row_map=perm, alpha=1.0,
             split_k=split, ldc=No)
    ref = torch.zeros_like(out)
    ref[perm.long()] = dy[:, :Mo].float().T @ x[:, :No].float()
    assert _relerr(out, ref) < 5e-5


def test_gemm_argmax(cuda_dev):
    from vit_exp_b200 import ops
    M, N, K = 512, 8192, 512
    a = torch.nn.functional.normalize(_rand((M, K), cuda_dev, 1), dim=-1).bfloat16()
    b = torch.nn.functional.normalize(_rand((N, K), cuda_dev, 2), dim=-1).bfloat16()
    best = torch.zeros(M, dtype=torch.int64, device=cuda_dev)
    ops.gemm(a, b, ops.EPI_ARGMAX, best, M=M, N=N, K=K, ldc=0)
    idx = (0xFFFFFFFF - (best & 0xFFFFFFFF)).long()
    sim = a.float() @ b.float().T
    ref_idx = sim.argmax(dim=1)
    picked = sim.gather(1, idx[:, None])[:, 0]
    # fp32 accumulation order may differ from torch's: accept equal-valued picks
    assert (picked >= sim.max(dim=1).values - 1e-6).all()
    assert (idx == ref_idx).float().mean().item() > 0.99


@pytest.mark.parametrize("M,N,K,epi", [(256, 512, 256, "bf16"), (300, 352, 200, "bf16"), (13824, 512, 2816, "bf16"),
                                       (4096, 768, 3072, "resid"), (520, 256, 512, "resid"), (128, 96, 64, "bf16")])
def test_gemm_kmajor_a_mnmajor_b(cuda_dev, M, N, K, epi):
    """D = A[M, K] W[K, N] with the weight used as stored ([out = K, in = N], N contiguous): the input-gradient
    products of both towers (ops.gemm(b_mn_major=True)) - no transposed weight copy.  Ragged M / N / K included."""
    from vit_exp_b200 import ops
    a = _rand((M, K), cuda_dev, 1).bfloat16()
    w = _rand((K, N), cuda_dev, 2, 0.3).bfloat16()
    ref = a.float() @ w.float()
    if epi == "bf16":
        c = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
        ops.gemm(a, w, ops.EPI_BF16, c, M=M, N=N, K=K, b_mn_major=True)
        assert _relerr(c, ref) < 8e-3
    else:
        x = _rand((M, N), cuda_dev, 3)
        c = torch.full((M, N), float("nan"), device=cuda_dev)
        ops.gemm(a, w, ops.EPI_RESID_F32, c, M=M, N=N, K=K, resid=x, b_mn_major=True)
        assert _relerr(c, ref + x) < 2e-5


def test_gemm_mnmajor_b_strided_a_and_padded_weight(cuda_dev):
    """the two odd operand views of the encoder backward: A = the k|v columns of the packed dqkv buffer (lda = 3*inner),
    and W2 [dim, ff_pad] whose pad columns are zero (GEGLU_BWD reads it MN-major with N = ff_pad)."""
    from vit_exp_b200 import ops
    M, inner, dim = 1000, 256, 512
    dqkv = _rand((M, 3 * inner), cuda_dev, 4).bfloat16()
    wkv = _rand((2 * inner, dim), cuda_dev, 5, 0.2).bfloat16()
    g = _rand((M, dim), cuda_dev, 6)
    ref = dqkv[:, inner:].float() @ wkv.float() + g
    ops.gemm(dqkv[:, inner:], wkv, ops.EPI_RESID_F32, g, M=M, N=dim, K=2 * inner, lda=3 * inner, resid=g, b_mn_major=True)
    assert _relerr(g, ref) < 2e-5
