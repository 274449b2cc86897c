"""dev tool: why is the host-input (e2e) train step slower than the HBM-resident one when the H2D copy (1.77 GB at
~55 GB/s = 32 ms) is shorter than the step (42 ms)?  Times the same step under different copy arrangements.

  python tools/e2e_probe.py [--steps 6]
"""
import argparse
import json
import os
import sys
import time
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from vit_exp_b200.ct_clip import TorchDistAccelerator
from vit_exp_b200.optim import FusedClipAdam

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--text-tower", default="hf")
args = ap.parse_args()
B = 8
dev = torch.device("cuda:0")
cfg = {"defer_loss_read": True}
cfg["ctk_text_tower"] = args.text_tower == "ctk"
clip = bench.build_model(dev, config=cfg).train()
bert = clip.text_transformer
orig = bert.forward
def fwd(*a, **k):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return orig(*a, **k)
bert.forward = fwd
params = [p for p in clip.parameters() if p.requires_grad]
opt = FusedClipAdam(params, lr=1.25e-6, betas=(0.9, 0.99), max_grad_norm=0.5)
acc = TorchDistAccelerator()
g = torch.Generator().manual_seed(1)
host_vid = [torch.rand(B, 1, *bench.VOL, generator=g).pin_memory() for _ in range(2)]
host16 = [(v[:, 0] * 2 - 1).half().pin_memory() for v in host_vid]
ids = torch.randint(0, 30522, (B, 512), device=dev)
mask = torch.ones_like(ids)
NSLOT = 3
dev_vid = [torch.empty(B, 1, *bench.VOL, device=dev) for _ in range(NSLOT)]
dev16 = [torch.empty(B, *bench.VOL, dtype=torch.float16, device=dev) for _ in range(NSLOT)]
scratch = torch.empty(B, 1, *bench.VOL, device=dev)
copy_stream = torch.cuda.Stream()
for s in range(NSLOT):
    dev_vid[s].copy_(host_vid[s % 2])
host_ms = []


def step(image):
    t0 = time.perf_counter()
    batch = {"data_type": ["imagereport"] * B, "text": SimpleNamespace(input_ids=ids, attention_mask=mask), "image": image}
    loss, ld = clip(batch, device=dev, accelerator=acc)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    t1 = time.perf_counter()
    v = float(ld["cl_loss"])
    host_ms.append((1e3 * (t1 - t0), 1e3 * (time.perf_counter() - t1)))
    return v


def timed(name, fn, k=args.steps, **extra):
    torch.cuda.synchronize()
    host_ms.clear()
    sampler = bench.ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(k)
    e1.record()
    torch.cuda.synchronize()
    clk = sampler.stop()
    ms = e0.elapsed_time(e1) / k
    enq = sum(h[0] for h in host_ms) / max(len(host_ms), 1)
    wait = sum(h[1] for h in host_ms) / max(len(host_ms), 1)
    print(json.dumps(dict(variant=name, ms_per_step=round(ms, 2), host_enqueue_ms=round(enq, 2), host_loss_wait_ms=round(wait, 2),
                          sm_mhz=clk["sm_mhz"], reasons=clk["reasons"], **extra)), flush=True)


for i in range(4):
    step(dev_vid[i % 2])

timed("resident", lambda k: [step(dev_vid[i % 2]) for i in range(k)])


def resident_plus_bg(k, src, dst, nchunk=1):
    for i in range(k):
        with torch.cuda.stream(copy_stream):
            if nchunk == 1:
                dst.copy_(src[i % 2], non_blocking=True)
            else:
                for c in range(B):
                    dst[c].copy_(src[i % 2][c], non_blocking=True)
        step(dev_vid[i % 2])
    torch.cuda.current_stream().wait_stream(copy_stream)


timed("resident + independent background H2D of 1.77 GB per step", lambda k: resident_plus_bg(k, host_vid, scratch))
timed("resident + independent background H2D of 0.88 GB per step (fp16)", lambda k: resident_plus_bg(k, host16, dev16[0]))
timed("resident + background H2D 1.77 GB as 8 per-volume copies", lambda k: resident_plus_bg(k, host_vid, scratch, nchunk=8))


def d2d_bg(k):
    for i in range(k):
        with torch.cuda.stream(copy_stream):
            scratch.copy_(dev_vid[2], non_blocking=True)
        step(dev_vid[i % 2])
    torch.cuda.current_stream().wait_stream(copy_stream)


timed("resident + background D2D copy of 1.77 GB per step", d2d_bg)


def e2e(k, late=False, depth=1):
    done = [None] * NSLOT

    def h2d(slot, src):
        with torch.cuda.stream(copy_stream):
            if done[slot] is not None:
                copy_stream.wait_event(done[slot])
            dev_vid[slot].copy_(host_vid[src], non_blocking=True)
    h2d(0, 0)
    for i in range(k):
        torch.cuda.current_stream().wait_stream(copy_stream)
        if not late and i + 1 < k:
            h2d((i + 1) % NSLOT, (i + 1) % 2)
        step(dev_vid[i % NSLOT])
        ev = torch.cuda.Event()
        ev.record()
        done[i % NSLOT] = ev
        if late and i + 1 < k:
            h2d((i + 1) % NSLOT, (i + 1) % 2)


timed("e2e (bench pipeline: copy of batch i+1 issued before step i)", e2e)
timed("e2e, copy of batch i+1 issued after step i is enqueued", lambda k: e2e(k, late=True))


def e2e_events(k):
    """wait on a per-copy event instead of the whole copy stream: step i only needs batch i, not batch i+1's enqueue order"""
    done = [None] * NSLOT
    ready = [None] * NSLOT

    def h2d(slot, src):
        with torch.cuda.stream(copy_stream):
            if done[slot] is not None:
                copy_stream.wait_event(done[slot])
            dev_vid[slot].copy_(host_vid[src], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            ready[slot] = ev
    h2d(0, 0)
    h2d(1, 1)
    for i in range(k):
        torch.cuda.current_stream().wait_event(ready[i % NSLOT])
        if i + 2 < k:
            h2d((i + 2) % NSLOT, i % 2)
        step(dev_vid[i % NSLOT])
        ev = torch.cuda.Event()
        ev.record()
        done[i % NSLOT] = ev


timed("e2e, two batches in flight, per-copy events", e2e_events)
