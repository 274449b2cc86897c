"""dev tool: temporal (24-token) attention kernels vs the fp64 reference, plus timing at bench size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from vit_exp_b200 import ops
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_ops_gpu import _attn_ref
dev = torch.device("cuda:0")
rel = lambda a, b: ((a.double().cpu() - b).norm() / b.norm()).item()

def case(nseq, heads, seed=0):
    g = torch.Generator().manual_seed(seed)
    L, inner = 24, heads * 32
    q = F.normalize(torch.randn(nseq * L, heads, 32, generator=g), dim=-1) * 8 * (1 + 0.1 * torch.randn(32, generator=g))
    k = F.normalize(torch.randn(nseq * L, heads, 32, generator=g), dim=-1) * (1 + 0.1 * torch.randn(32, generator=g))
    v = torch.randn(nseq * L, heads, 32, generator=g)
    qkv = torch.cat([q.reshape(-1, inner), k.reshape(-1, inner), v.reshape(-1, inner)], dim=1).bfloat16()
    qd = qkv.double().requires_grad_(True)
    ref, lse_ref = _attn_ref(qd, None, nseq, L, heads, 0, 0)
    dout = torch.randn(nseq * L, inner, generator=g).bfloat16()
    (ref * dout.double()).sum().backward()
    qc = qkv.to(dev)
    out, lse = ops.attn_fwd(qc, None, nseq, L, heads)
    dqkv = ops.attn_bwd(qc, None, ref.detach().float().bfloat16().to(dev), dout.to(dev), lse, None, nseq, L, heads)
    torch.cuda.synchronize()
    gq = qd.grad
    print(f"nseq {nseq} heads {heads}: out {rel(out, ref.detach()):.2e} lse {(lse.double().cpu() - lse_ref.detach()).abs().max().item():.1e}"
          f" | dq {rel(dqkv[:, :inner], gq[:, :inner]):.2e} dk {rel(dqkv[:, inner:2*inner], gq[:, inner:2*inner]):.2e}"
          f" dv {rel(dqkv[:, 2*inner:], gq[:, 2*inner:]):.2e}", flush=True)

BENCH_ONLY = os.environ.get('BENCH_ONLY') == '1'
for a in () if BENCH_ONLY else ((1, 8), (5, 8), (700, 8), (333, 4), (50, 2)):
    case(*a)

def timeit(fn, iters=2 if os.environ.get('BENCH_ONLY') == '1' else 20, warm=1 if os.environ.get('BENCH_ONLY') == '1' else 3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
B = 8
nseq, L = 576 * B, 24
M = nseq * L
qkv = torch.randn(M, 768, device=dev).bfloat16()
qkv[:, :512] = F.normalize(qkv[:, :512].float().view(M, 16, 32), dim=-1).view(M, 512).bfloat16()
qkv[:, :256] *= 8
out, lse = ops.attn_fwd(qkv, None, nseq, L, 8)
dout = torch.randn(M, 256, device=dev).bfloat16()
tf = timeit(lambda: ops.attn_fwd(qkv, None, nseq, L, 8))
tb = timeit(lambda: ops.attn_bwd(qkv, None, out, dout, lse, None, nseq, L, 8))
gb_f = (M * 768 * 2 + M * 256 * 2) / 1e9
gb_b = (M * 768 * 2 * 2 + 2 * M * 256 * 2) / 1e9
print(f"temporal B=8: fwd {tf:.3f} ms ({gb_f / tf * 1e3:.0f} GB/s)  bwd {tb:.3f} ms ({gb_b / tb * 1e3:.0f} GB/s)  "
      f"({'legacy' if os.environ.get('CTK_ATTN_LEGACY') == '1' else 'tma+mma'})")
