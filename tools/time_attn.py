"""dev tool: time the spatial / temporal attention kernels at bench size (B=8)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_exp_b200 import ops
dev = torch.device("cuda:0")
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
B = 8
for name, nseq, L, gh in (("spatial", 24 * B, 576, 24), ("temporal", 576 * B, 24, 0)):
    M = nseq * L
    qkv = torch.randn(M, 768, device=dev).bfloat16()
    qkv[:, :512] = torch.nn.functional.normalize(qkv[:, :512].float().view(M, 16, 32), dim=-1).view(M, 512).bfloat16()
    qkv[:, :256] *= 8
    table = torch.randn(8, 47, 47, device=dev) if gh else None
    out, lse = ops.attn_fwd(qkv, table, nseq, L, 8, gh, gh)
    dout = torch.randn(M, 256, device=dev).bfloat16()
    dtable = torch.zeros_like(table) if gh else None
    tf = timeit(lambda: ops.attn_fwd(qkv, table, nseq, L, 8, gh, gh))
    tb = timeit(lambda: ops.attn_bwd(qkv, table, out, dout, lse, dtable, nseq, L, 8, gh, gh))
    print(f"{name}: fwd {tf:.3f} ms  bwd(all) {tb:.3f} ms", flush=True)
