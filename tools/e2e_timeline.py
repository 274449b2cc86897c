"""dev tool: GPU timeline of the host-input (e2e) pipeline from a CUPTI trace: per stream busy time, every H2D copy,
and the idle gaps of the main stream with the kernels around them.  python tools/e2e_timeline.py [--mode fp32|stored|resident]
"""
import argparse
import json
import os
import sys
import tempfile
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from vit_exp_b200 import ops
from vit_exp_b200.ct_clip import TorchDistAccelerator
from vit_exp_b200.optim import FusedClipAdam

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="fp32")
ap.add_argument("--steps", type=int, default=4)
args = ap.parse_args()
B = 8
dev = torch.device("cuda:0")
clip = bench.build_model(dev, config={"defer_loss_read": True}).train()
params = [p for p in clip.parameters() if p.requires_grad]
opt = FusedClipAdam(params, lr=1.25e-6, betas=(0.9, 0.99), max_grad_norm=0.5)
acc = TorchDistAccelerator()
g = torch.Generator().manual_seed(1)
host_vid = [torch.rand(B, 1, *bench.VOL, generator=g).pin_memory() for _ in range(2)]
stored = [(v[:, 0] * 2 - 1).half().pin_memory() for v in host_vid]
ids = torch.randint(0, 30522, (B, 512), device=dev)
mask = torch.ones_like(ids)
NSLOT = 3
dev_vid = [torch.empty(B, 1, *bench.VOL, device=dev) for _ in range(NSLOT)]
dev_st = [torch.empty(B, *bench.VOL, dtype=torch.float16, device=dev) for _ in range(NSLOT)]
copy_stream = torch.cuda.Stream()
prep_stream = torch.cuda.Stream()
for s in range(NSLOT):
    dev_vid[s].copy_(host_vid[s % 2])


def step(image):
    batch = {"data_type": ["imagereport"] * B, "text": SimpleNamespace(input_ids=ids, attention_mask=mask), "image": image}
    loss, ld = clip(batch, device=dev, accelerator=acc)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    return float(ld["cl_loss"])


def feed(slot, src):
    if args.mode == "fp32":
        dev_vid[slot].copy_(host_vid[src], non_blocking=True)
    elif args.mode == "stored":                      # bench.py's feed: one copy, preparation kernels on their own stream
        dev_st[slot].copy_(stored[src], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(copy_stream)
        with torch.cuda.stream(prep_stream):
            prep_stream.wait_event(ev)
            for b in range(B):
                ops.volume_prep(dev_st[slot][b], dev_vid[slot][b])
        copy_stream.wait_stream(prep_stream)


def pipeline(k):
    done = [None] * NSLOT

    def fill(slot, src):
        with torch.cuda.stream(copy_stream):
            if done[slot] is not None:
                copy_stream.wait_event(done[slot])
            feed(slot, src)
    if args.mode != "resident":
        fill(0, 0)
    for i in range(k):
        if args.mode != "resident":
            torch.cuda.current_stream().wait_stream(copy_stream)
            if i + 1 < k:
                fill((i + 1) % NSLOT, (i + 1) % 2)
        step(dev_vid[i % NSLOT])
        ev = torch.cuda.Event()
        ev.record()
        done[i % NSLOT] = ev


pipeline(5)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pipeline(args.steps)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), f"e2e_trace_{os.getpid()}.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
streams = {}
for e in ev:
    streams.setdefault(e["args"].get("stream"), []).append(e)
print(f"mode {args.mode}: {len(ev)} GPU activities over {(ev[-1]['ts'] + ev[-1]['dur'] - t0) / 1e3:.1f} ms, {args.steps} steps")
for sid, lst in sorted(streams.items(), key=lambda kv: -sum(x["dur"] for x in kv[1])):
    busy = sum(x["dur"] for x in lst) / 1e3
    print(f"  stream {sid}: {len(lst):5d} activities, busy {busy:8.2f} ms, first at {(lst[0]['ts'] - t0) / 1e3:7.2f} ms, "
          f"e.g. {lst[len(lst) // 2]['name'][:60]}")
print("H2D copies > 1 MB:")
for e in ev:
    if e["cat"] == "gpu_memcpy" and "HtoD" in e["name"] and e["args"].get("bytes", 0) > 1 << 20:
        gb = e["args"]["bytes"] / 1e9
        print(f"  at {(e['ts'] - t0) / 1e3:8.2f} ms  dur {e['dur'] / 1e3:7.2f} ms  {gb:.3f} GB  {gb / (e['dur'] / 1e6):6.1f} GB/s  stream {e['args'].get('stream')}")
print("small H2D / D2H copies (< 1 MB):")
for e in ev:
    if e["cat"] == "gpu_memcpy" and ("HtoD" in e["name"] or "DtoH" in e["name"]) and e["args"].get("bytes", 0) <= 1 << 20:
        print(f"  at {(e['ts'] - t0) / 1e3:8.2f} ms  dur {e['dur'] / 1e3:7.3f} ms  {e['args'].get('bytes')} B  {e['name'][:30]}  stream {e['args'].get('stream')}")
# idle gaps of the whole GPU (no activity on any stream) and of the busiest stream
def gaps(lst, label, thr=300.0):
    end = lst[0]["ts"]
    prev = lst[0]
    for e in lst:
        if e["ts"] - end > thr:
            print(f"  {label} gap {(e['ts'] - end) / 1e3:6.2f} ms at {(end - t0) / 1e3:8.2f} ms: after {prev['name'][:48]} -> before {e['name'][:48]}")
        if e["ts"] + e["dur"] > end:
            end = e["ts"] + e["dur"]
            prev = e
compute = [e for e in ev if not (e["cat"] == "gpu_memcpy" and "HtoD" in e["name"] and e["args"].get("bytes", 0) > 1 << 20)]
print("gaps with no kernel / small copy running anywhere (> 0.3 ms):")
gaps(compute, "all-stream")
