"""dev tool: fused clip+Adam (libctk) vs clip_grad_norm_ + torch Adam(fused) on the bench's parameter set."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vit_exp_b200.optim import FusedClipAdam
dev = torch.device("cuda:0")
clip = bench.build_model(dev)
params = [p for p in clip.parameters() if p.requires_grad]
print(len(params), "tensors", sum(p.numel() for p in params) / 1e6, "M parameters")
for p in params:
    p.grad = torch.randn_like(p) * 1e-3
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(iters): fn()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (t1 - t0) / iters * 1e3
ours = FusedClipAdam(params, lr=1.25e-6, betas=(0.9, 0.99), max_grad_norm=0.5)
ref = torch.optim.Adam(params, lr=1.25e-6, betas=(0.9, 0.99), fused=True)
def ref_step():
    torch.nn.utils.clip_grad_norm_(params, 0.5); ref.step()
g, h = timeit(ours.step); print(f"libctk fused clip+Adam: device {g:.3f} ms/step, host {h:.3f} ms/step")
g, h = timeit(ref_step); print(f"torch clip + Adam(fused): device {g:.3f} ms/step, host {h:.3f} ms/step")
