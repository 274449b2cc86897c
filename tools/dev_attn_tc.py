"""dev tool: tcgen05 spatial attention forward/backward vs the fp64 reference, plus timing at bench size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from vit_exp_b200 import ops
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_ops_gpu import _attn_ref
dev = torch.device("cuda:0")
rel = lambda a, b: ((a.double().cpu() - b).norm() / b.norm()).item()

def case(nseq, heads, qmul=8.0, seed=0, bwd=True):
    g = torch.Generator().manual_seed(seed)
    L, inner = 576, heads * 32
    q = F.normalize(torch.randn(nseq * L, heads, 32, generator=g), dim=-1) * qmul * (1 + 0.1 * torch.randn(32, generator=g))
    k = F.normalize(torch.randn(nseq * L, heads, 32, generator=g), dim=-1) * (1 + 0.1 * torch.randn(32, generator=g))
    v = torch.randn(nseq * L, heads, 32, generator=g)
    qkv = torch.cat([q.reshape(-1, inner), k.reshape(-1, inner), v.reshape(-1, inner)], dim=1).bfloat16()
    table = torch.randn(heads, 47, 47, generator=g)
    qd = qkv.double().requires_grad_(True)
    td = table.double().requires_grad_(True)
    ref, lse_ref = _attn_ref(qd, td, nseq, L, heads, 24, 24)
    dout = torch.randn(nseq * L, inner, generator=g).bfloat16()
    (ref * dout.double()).sum().backward()
    qc, tc = qkv.to(dev), table.to(dev)
    out, lse = ops.attn_fwd(qc, tc, nseq, L, heads, 24, 24)
    torch.cuda.synchronize()
    msg = f"nseq {nseq} heads {heads} qmul {qmul}: out {rel(out, ref.detach()):.2e} lse {(lse.double().cpu() - lse_ref.detach()).abs().max().item():.1e}"
    if bwd:
        dtable = torch.zeros_like(tc)
        dqkv = ops.attn_bwd(qc, tc, ref.detach().float().bfloat16().to(dev), dout.to(dev), lse, dtable, nseq, L, heads, 24, 24)
        torch.cuda.synchronize()
        gq = qd.grad
        msg += (f" | dq {rel(dqkv[:, :inner], gq[:, :inner]):.2e} dk {rel(dqkv[:, inner:2*inner], gq[:, inner:2*inner]):.2e}"
                f" dv {rel(dqkv[:, 2*inner:], gq[:, 2*inner:]):.2e} dtable {rel(dtable, td.grad):.2e}")
    print(msg, flush=True)

BENCH_ONLY = os.environ.get('BENCH_ONLY') == '1'
for a in () if BENCH_ONLY else ((1, 1), (2, 8), (9, 8), (20, 4), (3, 8, 40.0)):
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
nseq, L = 24 * B, 576
M = nseq * L
qkv = torch.randn(M, 768, device=dev).bfloat16()
qkv[:, :512] = F.normalize(qkv[:, :512].float().view(M, 16, 32), dim=-1).view(M, 512).bfloat16()
qkv[:, :256] *= 8
table = torch.randn(8, 47, 47, device=dev)
out, lse = ops.attn_fwd(qkv, table, nseq, L, 8, 24, 24)
dout = torch.randn(M, 256, device=dev).bfloat16()
dtable = torch.zeros_like(table)
tf = timeit(lambda: ops.attn_fwd(qkv, table, nseq, L, 8, 24, 24))
tb = timeit(lambda: ops.attn_bwd(qkv, table, out, dout, lse, dtable, nseq, L, 8, 24, 24))
print(f"spatial B=8: fwd {tf:.3f} ms  bwd(all) {tb:.3f} ms  ({'legacy' if os.environ.get('CTK_ATTN_LEGACY') == '1' else 'tcgen05'})")
