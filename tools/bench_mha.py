"""dev tool: ctk_mha_fwd / ctk_mha_bwd against the library SDPA (bf16 flash) at the two shapes the reference uses:
text tower (B=8, L=512, 12 heads x 64) and CTViT3D (B=1, L=13824 + 2 null pairs, 8 heads x 32)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from vit_exp_b200 import ops

dev = torch.device("cuda:0")


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for name, B, L, heads, dh, n_null in (("text_tower", 8, 512, 12, 64, 0), ("ctvit3d", 1, 13824, 8, 32, 2)):
    H = heads * dh
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(B * L, 3 * H, generator=g).to(dev).bfloat16()
    dout = torch.randn(B * L, H, generator=g).to(dev).bfloat16()
    nk = torch.randn(heads, n_null, dh, generator=g).to(dev).bfloat16() if n_null else None
    nv = torch.randn(heads, n_null, dh, generator=g).to(dev).bfloat16() if n_null else None
    scale = dh ** -0.5
    ctx, lse = ops.mha_fwd(qkv, None, B, L, heads, scale, null_k=nk, null_v=nv)
    ms_f = timeit(lambda: ops.mha_fwd(qkv, None, B, L, heads, scale, null_k=nk, null_v=nv))
    ms_b = timeit(lambda: ops.mha_bwd(qkv, None, ctx, dout, lse, B, L, heads, scale, null_k=nk, null_v=nv))
    q, k, v = (qkv.view(B, L, 3, heads, dh)[:, :, t].transpose(1, 2).contiguous().requires_grad_(True) for t in range(3))
    do = dout.view(B, L, heads, dh).transpose(1, 2).contiguous()
    ms_lf = timeit(lambda: F.scaled_dot_product_attention(q, k, v, scale=scale))
    o = F.scaled_dot_product_attention(q, k, v, scale=scale)
    ms_lb = timeit(lambda: torch.autograd.grad(o, (q, k, v), do, retain_graph=True))
    gf = 4.0 * B * heads * L * (L + n_null) * dh / 1e9
    print(json.dumps({"shape": name, "B": B, "L": L, "heads": heads, "dh": dh, "n_null": n_null, "fwd_gflop": round(gf, 1),
                      "ctk_fwd_ms": round(ms_f, 3), "ctk_bwd_ms": round(ms_b, 3), "ctk_fwd_tflops": round(gf / ms_f, 1),
                      "ctk_bwd_tflops": round(2.5 * gf / ms_b, 1), "sdpa_fwd_ms": round(ms_lf, 3), "sdpa_bwd_ms": round(ms_lb, 3),
                      "note": "library SDPA timed without the null pairs (it needs a concatenated K/V copy for them)"}), flush=True)
