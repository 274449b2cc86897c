"""dev tool: coarse CUPTI timeline of one single-GPU train step: when each stream is busy, where the gaps are."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace
import bench
from vit_exp_b200.ct_clip import TorchDistAccelerator
from vit_exp_b200.optim import FusedClipAdam
B = 8
dev = torch.device("cuda:0")
clip = bench.build_model(dev).train()
bert = clip.text_transformer
orig = bert.forward
def fwd(*a, **k):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return orig(*a, **k)
bert.forward = fwd
params = [p for p in clip.parameters() if p.requires_grad]
opt = FusedClipAdam(params, lr=1.25e-6, betas=(0.9, 0.99), max_grad_norm=0.5)
acc = TorchDistAccelerator()
vid = torch.rand(B, 1, 240, 480, 480, device=dev)
ids = torch.randint(0, 30522, (B, 512), device=dev)
mask = torch.ones_like(ids)
def step():
    batch = {"data_type": ["imagereport"] * B, "text": SimpleNamespace(input_ids=ids, attention_mask=mask), "image": vid}
    loss, ld = clip(batch, device=dev, accelerator=acc)
    loss.backward()
    opt.step(); opt.zero_grad(set_to_none=True)
for _ in range(4): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.name and "Memset" not in e.name]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
ours = lambda n: "anonymous namespace" in n and "at::native" not in n
span = (evs[-1].time_range.end - t0) / 1e3
print(f"{len(evs)} kernels, span {span:.2f} ms")
# 1 ms bins: busy time of libctk kernels vs everything else
nb = int(span) + 1
binsA, binsB = [0.0] * nb, [0.0] * nb
for e in evs:
    s, t = (e.time_range.start - t0) / 1e3, (e.time_range.end - t0) / 1e3
    tgt = binsA if ours(e.name) else binsB
    b = int(s)
    while s < t and b < nb:
        seg = min(t, b + 1) - s
        tgt[b] += seg
        s = b + 1
        b += 1
print("ms bin : libctk busy | other (text tower, optimizer, torch) busy   [ms of kernel time inside each 1 ms bin]")
for b in range(nb):
    print(f"{b:3d}: {binsA[b]:5.2f} | {binsB[b]:5.2f}  {'#' * int(binsA[b] * 20)}{'.' * int(binsB[b] * 20)}")
