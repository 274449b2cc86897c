"""dev tool: GPU timeline of DDP train steps (torchrun, rank 0 reports): where the NCCL kernels sit relative to the two
towers' backward graphs and the optimizer, per stream busy time and all-stream idle gaps.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/ddp_timeline.py
"""
import json
import os
import sys
import tempfile
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from vit_exp_b200.ct_clip import TorchDistAccelerator
from vit_exp_b200.optim import FusedClipAdam

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = 8
clip = bench.build_model(dev, config={"defer_loss_read": True}).train()
DDP = torch.nn.parallel.DistributedDataParallel
mode = os.environ.get("DDP_MODE", "ignore-unused")
kw = dict(device_ids=[local], gradient_as_bucket_view=True, bucket_cap_mb=int(os.environ.get("BUCKET_MB", "25")))
if mode == "ignore-unused":
    DDP._set_params_and_buffers_to_ignore_for_model(clip, clip.unused_parameter_names())
    model = DDP(clip, find_unused_parameters=False, **kw)
else:
    model = DDP(clip, find_unused_parameters=True, static_graph=mode == "static-graph", **kw)
params = [p for p in clip.parameters() if p.requires_grad]
opt = FusedClipAdam(params, lr=1.25e-6, betas=(0.9, 0.99), max_grad_norm=0.5)
acc = TorchDistAccelerator()
vid = torch.rand(B, 1, 240, 480, 480, device=dev)
ids = torch.randint(0, 30522, (B, 512), device=dev)
mask = torch.ones_like(ids)


def step():
    batch = {"data_type": ["imagereport"] * B, "text": SimpleNamespace(input_ids=ids, attention_mask=mask), "image": vid}
    loss, ld = model(batch, device=dev, accelerator=acc)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    return float(ld["cl_loss"])


E2E = os.environ.get("E2E", "")           # "stored" / "fp32": bench.py's host-input pipeline instead of a resident batch
if E2E:
    from vit_exp_b200 import ops
    g = torch.Generator().manual_seed(1000 + rank)
    host_vid = [torch.rand(B, 1, 240, 480, 480, generator=g).pin_memory() for _ in range(2)]
    stored = [(v[:, 0] * 2 - 1).half().pin_memory() for v in host_vid]
    NSLOT = 3
    dev_vid = [torch.empty(B, 1, 240, 480, 480, device=dev) for _ in range(NSLOT)]
    dev_st = [torch.empty(B, 240, 480, 480, dtype=torch.float16, device=dev) for _ in range(NSLOT)]
    copy_stream, prep_stream = torch.cuda.Stream(), torch.cuda.Stream()
    done = [None] * NSLOT

    def fill(slot, src):
        with torch.cuda.stream(copy_stream):
            if done[slot] is not None:
                copy_stream.wait_event(done[slot])
            if E2E == "fp32":
                dev_vid[slot].copy_(host_vid[src], non_blocking=True)
                return
            dev_st[slot].copy_(stored[src], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
            with torch.cuda.stream(prep_stream):
                prep_stream.wait_event(ev)
                for b in range(B):
                    ops.volume_prep(dev_st[slot][b], dev_vid[slot][b])
            copy_stream.wait_stream(prep_stream)

    def run(k, first):
        global vid
        for i in range(first, first + k):
            torch.cuda.current_stream().wait_stream(copy_stream)
            fill((i + 1) % NSLOT, (i + 1) % 2)
            vid = dev_vid[i % NSLOT]
            step()
            ev = torch.cuda.Event()
            ev.record()
            done[i % NSLOT] = ev
    fill(0, 0)
    run(6, 0)
else:
    def run(k, first):
        for _ in range(k):
            step()
    run(5, 0)
torch.cuda.synchronize()
dist.barrier()
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(3, 6)
    torch.cuda.synchronize()
if rank == 0:
    path = os.path.join(tempfile.gettempdir(), f"ddp_trace_{os.getpid()}.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    span = (ev[-1]["ts"] + ev[-1]["dur"] - t0) / 1e3
    print(f"{len(ev)} GPU activities over {span:.1f} ms, 3 steps -> {span / 3:.2f} ms/step")
    streams = {}
    for e in ev:
        streams.setdefault(e["args"].get("stream"), []).append(e)
    for sid, lst in sorted(streams.items(), key=lambda kv: -sum(x["dur"] for x in kv[1])):
        print(f"  stream {sid}: {len(lst):5d} activities, busy {sum(x['dur'] for x in lst) / 1e3:8.2f} ms, e.g. {lst[len(lst) // 2]['name'][:70]}")
    print("H2D copies > 1 MB:")
    for e in ev:
        if e["cat"] == "gpu_memcpy" and "HtoD" in e["name"] and e["args"].get("bytes", 0) > 1 << 20:
            gb = e["args"]["bytes"] / 1e9
            print(f"  at {(e['ts'] - t0) / 1e3:8.2f} ms  dur {e['dur'] / 1e3:7.2f} ms  {gb:.3f} GB  {gb / (e['dur'] / 1e6):6.1f} GB/s")
    print("NCCL kernels and step markers:")
    for e in ev:
        n = e["name"]
        if "nccl" in n.lower() or "multi_adam" in n or "multi_sqnorm" in n or "patch_norm" in n or "clip_grad_tiles" in n:
            print(f"  t={(e['ts'] - t0) / 1e3:8.3f} ms dur={e['dur'] / 1e3:7.3f} ms stream {e['args'].get('stream')}  {n[:80]}")
    # per stream: busy segments (activities closer than 0.3 ms merged) with their three most frequent kernels
    import collections
    for sid, lst in sorted(streams.items(), key=lambda kv: -sum(x["dur"] for x in kv[1])):
        print(f"segments of stream {sid} (first 45 ms):")
        seg = None
        segs = []
        for e in lst:
            if seg is not None and e["ts"] - seg["end"] < 300:
                seg["end"] = max(seg["end"], e["ts"] + e["dur"])
                seg["names"][e["name"][:46]] += 1
            else:
                seg = {"start": e["ts"], "end": e["ts"] + e["dur"], "names": collections.Counter({e["name"][:46]: 1})}
                segs.append(seg)
        for g in segs:
            if (g["start"] - t0) / 1e3 > 45:
                break
            top = ", ".join(f"{n} x{c}" for n, c in g["names"].most_common(3))
            print(f"  {(g['start'] - t0) / 1e3:7.2f} - {(g['end'] - t0) / 1e3:7.2f} ms  {sum(g['names'].values()):4d} act  {top}")
    print("1 ms bins of the SECOND profiled step: DtoD copies (gradient -> bucket) | NCCL all-reduces | text-tower backward kernels (mha_bwd) | encoder GEMMs")
    starts = [e["ts"] for e in ev if "patch_norm" in e["name"]]
    t1 = starts[1] if len(starts) > 1 else t0
    bins = collections.defaultdict(lambda: [0, 0, 0, 0])
    for e in ev:
        if e["ts"] < t1:
            continue
        b = int((e["ts"] - t1) / 1e3)
        if b >= 46:
            break
        n = e["name"]
        if "DtoD" in n:
            bins[b][0] += 1
        elif "AllReduce" in n:
            bins[b][1] += 1
        elif "mha_bwd" in n:
            bins[b][2] += 1
        elif "gemm_kernel" in n:
            bins[b][3] += 1
    for b in range(46):
        print(f"  {b:3d} ms: {bins[b][0]:4d} | {bins[b][1]:3d} | {bins[b][2]:3d} | {bins[b][3]:3d}")
    try:
        info = model._get_ddp_logging_data()
        print("ddp buckets:", info.get("rebuilt_bucket_sizes") or info.get("bucket_sizes"), "| has_rebuilt_buckets", info.get("has_rebuilt_buckets"))
    except Exception as e:
        print("no ddp logging data:", e)
    # idle gaps: no activity on any stream
    end, prev = ev[0]["ts"], ev[0]
    print("all-stream idle gaps > 0.2 ms:")
    for e in ev:
        if e["ts"] - end > 200:
            print(f"  gap {(e['ts'] - end) / 1e3:6.2f} ms at {(end - t0) / 1e3:8.2f} ms: after {prev['name'][:50]} -> before {e['name'][:50]}")
        if e["ts"] + e["dur"] > end:
            end, prev = e["ts"] + e["dur"], e
    # compute-only gaps: time when only NCCL kernels are running
    comp = [e for e in ev if "nccl" not in e["name"].lower()]
    end, prev = comp[0]["ts"], comp[0]
    print("gaps with no compute kernel running (NCCL-only or idle) > 0.2 ms:")
    for e in comp:
        if e["ts"] - end > 200:
            print(f"  gap {(e['ts'] - end) / 1e3:6.2f} ms at {(end - t0) / 1e3:8.2f} ms: after {prev['name'][:50]} -> before {e['name'][:50]}")
        if e["ts"] + e["dur"] > end:
            end, prev = e["ts"] + e["dur"], e
dist.destroy_process_group()
