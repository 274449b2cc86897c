import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from vit_exp_b200 import ops
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_ops_gpu import _attn_ref
dev = torch.device("cuda:0")
nseq, heads, L = int(sys.argv[1]), int(sys.argv[2]), 576
inner = heads * 32
g = torch.Generator().manual_seed(0)
q = F.normalize(torch.randn(nseq * L, heads, 32, generator=g), dim=-1) * 8
k = F.normalize(torch.randn(nseq * L, heads, 32, generator=g), dim=-1)
v = torch.randn(nseq * L, heads, 32, generator=g)
qkv = torch.cat([q.reshape(-1, inner), k.reshape(-1, inner), v.reshape(-1, inner)], dim=1).bfloat16()
table = torch.randn(heads, 47, 47, generator=g)
qd = qkv.double().requires_grad_(True); td = table.double().requires_grad_(True)
ref, lse_ref = _attn_ref(qd, td, nseq, L, heads, 24, 24)
dout = torch.randn(nseq * L, inner, generator=g).bfloat16()
(ref * dout.double()).sum().backward()
qc, tc = qkv.to(dev), table.to(dev)
out, lse = ops.attn_fwd(qc, tc, nseq, L, heads, 24, 24)
dtable = torch.zeros_like(tc)
ops.attn_bwd(qc, tc, ref.detach().float().bfloat16().to(dev), dout.to(dev), lse, dtable, nseq, L, heads, 24, 24)
torch.cuda.synchronize()
got, want = dtable.double().cpu()[0], td.grad[0]
print("nan count", torch.isnan(got).sum().item(), "rel", ((got - want).norm() / want.norm()).item())
err = (got - want).abs()
print("rows (dy) with max err:", err.max(dim=1).values[:47].numpy().round(3))
print("cols (dx) with max err:", err.max(dim=0).values[:47].numpy().round(3))
print("want[20:27,20:27]\n", want[20:27, 20:27].numpy().round(2)); print("got\n", got[20:27, 20:27].numpy().round(2))
