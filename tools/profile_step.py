"""dev tool: per-kernel time breakdown of one bench train step (torch.profiler / CUPTI)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace
import bench
from vit_exp_b200.ct_clip import TorchDistAccelerator

B = int(os.environ.get("B", "8"))
dev = torch.device("cuda:0")
clip = bench.build_model(dev).train()
if os.environ.get("TEXT_STUB") == "1":            # isolate the image path: replace BERT by an embedding table
    class _Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = torch.nn.Embedding(30522, 768)
        def forward(self, input_ids, attention_mask=None):
            return (self.emb(input_ids),)
    clip.text_transformer = _Stub().to(dev)
bert = clip.text_transformer
orig = bert.forward
def fwd(*a, **k):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return orig(*a, **k)
bert.forward = fwd
params = [p for p in clip.parameters() if p.requires_grad]
opt = torch.optim.Adam(params, lr=1.25e-6, betas=(0.9, 0.99), fused=True)
acc = TorchDistAccelerator()
vid = torch.rand(B, 1, 240, 480, 480, device=dev)
ids = torch.randint(0, 30522, (B, 512), device=dev)
mask = torch.ones_like(ids)
def step():
    batch = {"data_type": ["imagereport"] * B, "text": SimpleNamespace(input_ids=ids, attention_mask=mask), "image": vid}
    loss, ld = clip(batch, device=dev, accelerator=acc)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(params, 0.5)
    opt.step(); opt.zero_grad(set_to_none=True)
for _ in range(2): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
    if e.device_type == torch.autograd.DeviceType.CUDA and t > 0:
        rows.append((t, e.count, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total device time {tot/1e3:.2f} ms over {sum(r[1] for r in rows)} kernels")
for t, c, k in rows[:int(os.environ.get('TOPN', '45'))]:
    print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}% x{c:<5d} {k[:110]}")

# ---- host-side cost per phase (no sync inside; CPU launch time only)
import time
def phase_times():
    batch = {"data_type": ["imagereport"] * B, "text": SimpleNamespace(input_ids=ids, attention_mask=mask), "image": vid}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    loss, ld = clip(batch, device=dev, accelerator=acc)
    t1 = time.perf_counter()
    loss.backward()
    t2 = time.perf_counter()
    torch.nn.utils.clip_grad_norm_(params, 0.5)
    t3 = time.perf_counter()
    opt.step(); opt.zero_grad(set_to_none=True)
    t4 = time.perf_counter()
    torch.cuda.synchronize(); t5 = time.perf_counter()
    return [1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0)]
for _ in range(2):
    print("host ms: fwd(+loss.item sync) %.1f  bwd %.1f  clip %.1f  adam %.1f  drain %.1f  total %.1f" % tuple(phase_times()))
