"""dev tool: CUPTI timeline of one DDP train step (torchrun, rank 0 reports): NCCL kernels vs compute."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from types import SimpleNamespace
import bench
from vit_exp_b200.ct_clip import TorchDistAccelerator
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = 8
clip = bench.build_model(dev).train()
bert = clip.text_transformer
orig = bert.forward
def fwd(*a, **k):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return orig(*a, **k)
bert.forward = fwd
model = torch.nn.parallel.DistributedDataParallel(clip, device_ids=[local], find_unused_parameters=True,
                                                  gradient_as_bucket_view=True, bucket_cap_mb=int(os.environ.get("BUCKET_MB", "25")))
params = [p for p in clip.parameters() if p.requires_grad]
opt = torch.optim.Adam(params, lr=1.25e-6, betas=(0.9, 0.99), fused=True)
acc = TorchDistAccelerator()
vid = torch.rand(B, 1, 240, 480, 480, device=dev)
ids = torch.randint(0, 30522, (B, 512), device=dev)
mask = torch.ones_like(ids)
def step():
    batch = {"data_type": ["imagereport"] * B, "text": SimpleNamespace(input_ids=ids, attention_mask=mask), "image": vid}
    loss, ld = model(batch, device=dev, accelerator=acc)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(params, 0.5)
    opt.step(); opt.zero_grad(set_to_none=True)
for _ in range(3): step()
torch.cuda.synchronize(); dist.barrier()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    tot = sum(e.time_range.end - e.time_range.start for e in evs)
    print(f"{len(evs)} device events, busy sum {tot/1e3:.2f} ms, span {(evs[-1].time_range.end - t0)/1e3:.2f} ms")
    for e in evs:
        n = e.name
        if "nccl" in n.lower() or "Adam" in n or "FusedOpt" in n or "multi_tensor" in n or "clip_loss" in n or "patch_norm" in n:
            print(f"  t={(e.time_range.start - t0)/1e3:8.3f} ms  dur={(e.time_range.end - e.time_range.start)/1e3:7.3f} ms  {n[:90]}")
dist.destroy_process_group()
