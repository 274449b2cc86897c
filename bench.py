#!/usr/bin/env python
"""bench.py - CT volumes/s of one CT-CLIP train step (BASELINE.json north star) on N B200s.

  python bench.py [--gpus N --steps K --warmup W]          our arm (libctk CUDA path)
  python bench.py --impl reference [...]                   the reference's CPU path (oracle port) on the host cores
  torchrun --nproc-per-node N bench.py --gpus N ...         one rank per GPU over NCCL

One step = forward + backward + all-gather of latents + DDP gradient all-reduce + grad-norm clip 0.5
+ Adam(lr 1.25e-6, betas (0.9, 0.99)) on a batch of synthetic CT-RATE-shaped volumes
(B, 1, 240, 480, 480) fp32 and (B, 512) token ids, random-init CTViT(dim 512, 4+4) and a random-init
BERT-base-shaped text encoder.  Per-GPU batch is fixed (weak scaling):
8 volumes / GPU -> global batch 64 at 8 GPUs, the configuration BASELINE.json quotes.

`value`  : volumes/s with the batch already resident in HBM (CUDA events, max over ranks).
`e2e`    : the same step driven through the public CTCLIP.forward API with HOST (pinned) inputs: the
           H2D copy of every step's batch and the D2H read of the loss are inside the timed region
           (the copy of step i+1 is issued on a side stream while step i computes).
`roofline`: all tcgen05 GEMM launches of a step, timed with CUDA events on the launching stream in an
           extra instrumented step; algorithmic FLOPs from SURVEY.md 8d.
`cpu_baseline`: the oracle port of the reference timed on this box's host cores (bounded sample).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VOL = (240, 480, 480)
TEXT_LEN = 512
PER_GPU_BATCH = 8
# algorithmic GFLOP per volume executed by ctk_gemm_bf16 launches (SURVEY.md 8d / BASELINE.md 4):
GF_PATCH = 56.62
GF_LAYER_GEMM = 3.62 + 7.25 + 3.62 + 38.65 + 19.32        # q, kv, out, FF1, FF2 (attention core excluded)
GF_VQ = 116.0
GF_GEMM_FWD = GF_PATCH + 8 * GF_LAYER_GEMM + GF_VQ
GF_GEMM_BWD = GF_PATCH + 2 * 8 * GF_LAYER_GEMM             # patch: wgrad only; layers: dgrad + wgrad
GF_GEMM_STEP = GF_GEMM_FWD + GF_GEMM_BWD
# --text-tower ctk: BERT-base at 512 tokens, per report and layer: QKV 1.812 + out 0.604 + intermediate 2.416 + output 2.416
# GF forward (attention core excluded), dgrad + wgrad in the backward pass
GF_TEXT_GEMM_STEP = 3 * 12 * (1.812 + 0.604 + 2.416 + 2.416)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The driver reads ONE JSON line from stdout.  Libraries write there too (NCCL prints its version banner on
# fd 1 when the first communicator is created), so file descriptor 1 is pointed at stderr for the whole run and
# the result line goes to a private duplicate of the original stdout.
_REAL_STDOUT = None


def protect_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def gemm_traffic_sample():
    """Per-launch DRAM bytes of the tcgen05 GEMM family from the committed `ncu --set full` captures
    (profiles/r2_ncu_gemm_{fwd,bwd}.csv: GEMM launches of the device-timed region of one B=8 train step; the round-1
    captures if those are missing - round 2's `--set full` run was lost to the copy-back size limit, profiles/README.md).
    Returns (mean bytes per launch, launches, the files used)."""
    import csv
    tot, n = 0.0, 0
    names = ("r2_ncu_gemm_fwd.csv", "r2_ncu_gemm_bwd.csv")
    if not all(os.path.exists(os.path.join(ROOT, "profiles", x)) for x in names):
        names = ("r1_ncu_gemm_b8_fwd.csv", "r1_ncu_gemm_b8_bwd.csv")
    for name in names:
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(open(path)))
        hdr, unit = rows[0], rows[1]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            if not r[0].startswith("gemm_kernel"):
                continue
            tot += float(r[ir]) * scale.get(unit[ir], 1.0) + float(r[iw]) * scale.get(unit[iw], 1.0)
            n += 1
    return (tot / n, n, names) if n else (None, 0, ())


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1349.9), d.get("hbm_gbs", 6547.8), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.th.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def build_model(dev, seed=0, config=None, text_dropout=0.0):
    import torch
    from transformers import BertConfig, BertModel
    from vit_exp_b200.ct_clip import CTCLIP
    from vit_exp_b200.transformer_maskgit import CTViT
    torch.manual_seed(seed)
    vit = CTViT(dim=512, codebook_size=8192, image_size=480, patch_size=20, temporal_patch_size=10,
                spatial_depth=4, temporal_depth=4, dim_head=32, heads=8)          # run_train.py:56-66
    bert = BertModel(BertConfig(vocab_size=30522, hidden_dropout_prob=text_dropout, attention_probs_dropout_prob=text_dropout))
    clip = CTCLIP(image_encoder=vit, text_encoder=bert, dim_text=768, dim_image=512, dim_latent=512, config=dict(config or {}))
    return clip.to(dev)


class AutocastText:
    """runs the (stock PyTorch) text encoder under bf16 autocast; everything else is libctk."""

    def __init__(self, bert):
        self.bert = bert

    def __call__(self, input_ids, attention_mask=None):
        import torch
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return self.bert(input_ids, attention_mask=attention_mask)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from vit_exp_b200 import _lib, ops
    from vit_exp_b200.ct_clip import TorchDistAccelerator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = True
    lib = _lib.load()
    B = args.batch_per_gpu

    cfg = {} if args.sync_loss_read else {"defer_loss_read": True}
    cfg["ctk_text_tower"] = args.text_tower == "ctk"
    clip = build_model(dev, seed=0, config=cfg, text_dropout=args.text_dropout)
    clip.train()
    bert = clip.text_transformer
    model = clip
    if world > 1:
        # The reference trainer wraps with find_unused_parameters=True (CTCLIPTrainer.py:318) because the pooler, the
        # *_extra projections, the first-frame / pixel heads and the cross-attention norms never get a gradient.  With
        # that flag DDP all-reduces a used-parameter bitmap after every backward pass and reads it back on the host: a
        # sync that leaves the GPU idle for the host time of optimizer.step() (profiles/r2_ddp_timeline_n2.txt).
        #   --ddp ignore-unused (default): those parameters (CTCLIP.unused_parameter_names(), a fixed set) are handed to
        #       DDP as ignored, and DDP runs with find_unused_parameters=False: no bitmap, no sync, buckets all-reduced
        #       as soon as they fill.
        #   --ddp find-unused: the reference's setting, unchanged.
        #   --ddp static-graph: find_unused_parameters=True + static_graph=True (no sync either, but DDP then launches
        #       every all-reduce at the end of the backward pass).
        DDP = torch.nn.parallel.DistributedDataParallel
        kw = dict(device_ids=[local], gradient_as_bucket_view=True, bucket_cap_mb=args.bucket_mb)
        if args.ddp == "ignore-unused":
            DDP._set_params_and_buffers_to_ignore_for_model(clip, clip.unused_parameter_names())
            model = DDP(clip, find_unused_parameters=False, **kw)
        else:
            model = DDP(clip, find_unused_parameters=True, static_graph=args.ddp == "static-graph", **kw)
    params = [p for p in clip.parameters() if p.requires_grad]
    if args.optimizer == "fused":        # clip_grad_norm_(0.5) + Adam in two libctk launches (vit_exp_b200/optim.py)
        from vit_exp_b200.optim import FusedClipAdam
        opt = FusedClipAdam(params, lr=1.25e-6, betas=(0.9, 0.99), max_grad_norm=0.5)   # optimizer.py:14,23-24
    else:
        opt = torch.optim.Adam(params, lr=1.25e-6, betas=(0.9, 0.99), fused=True)
    acc = TorchDistAccelerator()

    # synthetic CT-RATE-shaped host batch in pinned memory (two buffers -> consecutive steps differ)
    g = torch.Generator().manual_seed(1000 + rank)
    host_vid = [torch.rand(B, 1, *VOL, generator=g).pin_memory() for _ in range(2)]
    host_ids = [torch.randint(0, 30522, (B, TEXT_LEN), generator=g).pin_memory() for _ in range(2)]
    NSLOT = 3            # device slots: the copy of step i+1 needs a slot that step i-1 may still be reading
    dev_vid = [torch.empty(B, 1, *VOL, device=dev) for _ in range(NSLOT)]
    dev_ids = [torch.empty(B, TEXT_LEN, dtype=torch.int64, device=dev) for _ in range(NSLOT)]
    mask = torch.ones(B, TEXT_LEN, dtype=torch.int64, device=dev)
    copy_stream = torch.cuda.Stream()

    def text_forward_patch():
        # bf16 autocast for the stock-PyTorch text tower only
        orig = bert.forward
        def fwd(*a, **k):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return orig(*a, **k)
        bert.forward = fwd
    text_forward_patch()

    def h2d(slot, src=None):
        src = slot % 2 if src is None else src
        with torch.cuda.stream(copy_stream):
            dev_vid[slot].copy_(host_vid[src], non_blocking=True)
            dev_ids[slot].copy_(host_ids[src], non_blocking=True)

    def step(slot, image=None):
        batch = {"data_type": ["imagereport"] * B,
                 "text": SimpleNamespace(input_ids=dev_ids[slot], attention_mask=mask),
                 "image": dev_vid[slot] if image is None else image}
        loss, ld = model(batch, device=dev, accelerator=acc, return_loss=True, return_loss_dict=True)
        loss.backward()
        if args.optimizer != "fused":
            torch.nn.utils.clip_grad_norm_(params, 0.5)                        # CTCLIPTrainer.py:711-712
        opt.step()
        opt.zero_grad(set_to_none=True)
        return float(ld["cl_loss"])                   # D2H read of the loss, every step (deferred: after the step is enqueued)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(k)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), out

    # inputs resident for the device-timed loop
    for s in range(2):
        h2d(s)
    copy_stream.synchronize()
    log(f"[rank {rank}] warm-up {args.warmup} steps, B={B}/GPU")
    loss_val = None
    for i in range(args.warmup):
        loss_val = step(i % 2)

    # ---- value: K steps, inputs in HBM (each volume batch 1.77 GB >> 126 MB L2) -----------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = lib.ctk_launch_count() + ops.GRAPH_LAUNCHES
    # CTK_BENCH_PROFILER_RANGE=1: cudaProfilerStart/Stop around the device-timed loop, so that
    # `ncu --profile-from-start off ...` lists exactly the launches of the timed region (tools/ncu_round2.sh)
    prof_range = os.environ.get("CTK_BENCH_PROFILER_RANGE") == "1"
    if prof_range:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    ms_dev, _ = timed(lambda k: [step(i % 2) for i in range(k)], args.steps)
    if prof_range:
        torch.cuda.profiler.stop()
        # a profiler run only wants the launches of the timed region: no e2e / baseline legs under the profiler, and the
        # line says so (a number printed by a run under ncu is never a bench value)
        if rank == 0:
            emit({"metric": "CT volumes/s, CT-CLIP train step", "under_profiler": True, "n_gpus": world, "steps": args.steps,
                  "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
                  "gpu_launches": int(lib.ctk_launch_count() + ops.GRAPH_LAUNCHES - n0)})
        if world > 1:
            dist.destroy_process_group()
        return
    launches = (lib.ctk_launch_count() + ops.GRAPH_LAUNCHES - n0)      # direct launches + launches replayed from CUDA graphs
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host inputs, transfer of step i+1 overlapped with step i, loss read back every step ---
    # Three device slots: the batch of step i+1 is copied (side stream) into the slot step i-2 used, which is free as
    # soon as step i-2 has finished, i.e. when step i-1 starts; the copy therefore has two step times of slack.
    #   e2e           : the reference's STORED format on the host - float16 `arr_0` of the *_fp16 datasets, values in
    #                   [-1, 1] - shipped as stored, one volume at a time, and turned into the loader's fp32
    #                   (B, 1, 240, 480, 480) tensor on the device by ctk_volume_prep (scripts/data.py:49-111, bit-exact:
    #                   tests/test_volume_prep_gpu.py), each volume's kernel right behind its copy.
    #   e2e_fp32_host : the loader's fp32 result on the host (what the reference's DataLoader hands over), 2x the bytes.
    # The pipeline is software-pipelined by one batch, as a training loop's prefetching loader is: batch 0 is staged
    # (and has landed) before the clock starts, and inside the timed region every step i issues the transfer of batch
    # i+1 - including the last one, whose batch is for the step after the clock stops.  The timed region therefore
    # holds K steps, K host->device batch transfers and K loss read-backs, without charging the one-off pipeline fill
    # (one transfer with nothing to overlap) to a K of 5.
    def pipeline(k, feed, prime_only=False):
        done = pipeline.done
        def fill(slot, src):
            with torch.cuda.stream(copy_stream):
                if done[slot] is not None:
                    copy_stream.wait_event(done[slot])          # last reader of that slot
                feed(slot, src)
        if prime_only:
            pipeline.done = done = [None] * NSLOT
            fill(0, 0)
            return None
        last = None
        for i in range(k):
            torch.cuda.current_stream().wait_stream(copy_stream)
            fill((i + 1) % NSLOT, (i + 1) % 2)
            last = step(i % NSLOT)
            ev = torch.cuda.Event()
            ev.record()
            done[i % NSLOT] = ev
        return last
    pipeline.done = [None] * NSLOT

    def timed_e2e(feed):
        pipeline(0, feed, prime_only=True)                  # stage batch 0; `timed` synchronises before starting the clock
        return timed(lambda k: pipeline(k, feed), args.steps)

    def feed_fp32(slot, src):
        dev_vid[slot].copy_(host_vid[src], non_blocking=True)
        dev_ids[slot].copy_(host_ids[src], non_blocking=True)

    stored = [(v[:, 0] * 2 - 1).half().pin_memory() for v in host_vid]          # (B, D, H, W) stored arrays
    dev_stored = [torch.empty(B, *VOL, dtype=torch.float16, device=dev) for _ in range(NSLOT)]

    prep_stream = torch.cuda.Stream()

    def feed_stored(slot, src):
        # called under copy_stream.  ONE copy of the whole stored batch (at 4-8 GPUs the host delivers ~25 % less per GPU
        # when the same bytes arrive as eight per-volume copies), then the preparation kernels on their own stream behind
        # an event: on the copy stream they would wait for SMs behind the train step's persistent kernels and hold up
        # the next transfer with them.
        dev_ids[slot].copy_(host_ids[src], non_blocking=True)
        dev_stored[slot].copy_(stored[src], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(copy_stream)
        with torch.cuda.stream(prep_stream):
            prep_stream.wait_event(ev)
            for b in range(B):
                ops.volume_prep(dev_stored[slot][b], dev_vid[slot][b])
        copy_stream.wait_stream(prep_stream)            # "batch staged" = copy and preparation done

    pipeline(0, feed_stored, prime_only=True)
    pipeline(2, feed_stored)                                # first use of the prep kernel / slots
    ms_e2e, loss_val = timed_e2e(feed_stored)
    ms_e2e32, _ = timed_e2e(feed_fp32)
    del stored, dev_stored

    # ---- informational: the reference's own loss read (loss.item() inside forward: a host sync between forward and
    # backward, ct_clip.py:1384) instead of the deferred read the timed loops use
    ms_sync = None
    if not args.sync_loss_read:
        clip.config["defer_loss_read"] = False
        step(0)
        ms_sync, _ = timed(lambda k: [step(i % 2) for i in range(k)], args.steps)
        clip.config["defer_loss_read"] = True

    # ---- roofline: instrument every tcgen05 GEMM launch of one more step --------------------------
    # (eager launches, and the two towers run one after the other on one stream: with the text tower on its side stream
    # the events around a GEMM would also count the time its CTAs wait for SMs held by the other tower's kernels)
    # The events around a launch must not also measure the host: the host needs ~45 ms to enqueue an eager step, about as
    # long as the GPU needs to run it, so the GPU is first given two ordinary (graph-replayed, 7 ms of host time) steps
    # to chew on and the instrumented step is enqueued behind them.  (A device-side sleep instead lets the clocks drop.)
    big = B > 16
    if big:
        # large per-GPU batches: drop the CUDA graphs first (their private pools hold one full set of saved activations,
        # 1.7 GB per volume) before the eager step allocates its own
        import gc
        clip.visual_transformer._graphs.clear()
        getattr(bert, "__dict__", {}).pop("_ctk_graphs", None)
        gc.collect()
        torch.cuda.empty_cache()
    else:
        step(0)
        step(1)
    clip.overlap_text_encoder = False          # towers one after the other on one stream: no second stream competing for SMs
    ops.GEMM_PROFILE = []
    step(0)
    torch.cuda.synchronize()
    prof = ops.GEMM_PROFILE
    ops.GEMM_PROFILE = None
    clip.overlap_text_encoder = True
    gemm_ms = sum(a.elapsed_time(b) for a, b, _ in prof)
    by_epi = {}
    for a, b, tag in prof:
        by_epi[tag] = by_epi.get(tag, 0.0) + a.elapsed_time(b)
    peak_tf, peak_hbm, peak_src = measured_peaks()
    gf_step = GF_GEMM_STEP + (GF_TEXT_GEMM_STEP if args.text_tower == "ctk" else 0.0)
    achieved_tf = gf_step * B / gemm_ms                                        # GFLOP / ms == TFLOP/s
    traffic, traffic_n, traffic_files = gemm_traffic_sample()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    vols = B * world * args.steps
    line = {
        "metric": "CT volumes/s, CT-CLIP train step", "value": vols / (ms_dev / 1e3), "unit": "volumes/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "impl": "ours",
        "config": {"workload": "ctclip_train_step: CTViT(dim512, 4 spatial + 4 temporal, heads 8x32, patch 20x20x10 of "
                               "480x480x240) + random-init BERT-base text tower + all-gathered InfoNCE + clip 0.5 + Adam",
                   "per_gpu_batch": B, "global_batch": B * world, "text_len": TEXT_LEN, "parallelism": f"dp{world}",
                   "l2": "inputs larger than L2 (1.77 GB of volumes per step); two alternating batches",
                   "e2e_pipeline": "host batch (pinned, stored float16 volumes) -> H2D (copy stream) + ctk_volume_prep per volume "
                                   "(second side stream) into one of 3 device slots while the previous step computes; loss read back "
                                   "every step; software-pipelined by one batch: the timed region holds K steps and the K "
                                   "transfers of batches 1..K (batch 0 is staged before the clock starts)",
                   "text_tower": (("BertModel parameters through libctk (vit_exp_b200/text_tower.py)"
                                   if args.text_tower == "ctk" else "stock PyTorch BertModel under bf16 autocast")
                                  + f", hidden / attention dropout {args.text_dropout} (CXR-BERT ships 0.1; the CPU arm runs without)"),
                   "loss_read": "loss.item() inside forward" if args.sync_loss_read else
                                "config['defer_loss_read']: async D2H copy, read every step after the step is enqueued",
                   "launch": "encoder and text tower forward/backward replayed from CUDA graphs" if clip.visual_transformer.cuda_graphs
                             else "eager launches",
                   "ddp": None if world == 1 else f"{args.ddp}, gradient_as_bucket_view=True, bucket_cap_mb={args.bucket_mb}"},
        "e2e": {"value": vols / (ms_e2e / 1e3), "unit": "volumes/s",
                "h2d_bytes_per_step": int(host_vid[0].numel() * 2 + host_ids[0].numel() * 8), "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps,
                "host_format": "float16 stored arrays (arr_0 of the *_fp16 datasets), prepared on the device by "
                               "ctk_volume_prep = scripts/data.py:49-111 npz_to_tensor, bit-exact"},
        "e2e_fp32_host": {"value": vols / (ms_e2e32 / 1e3), "unit": "volumes/s", "ms_per_step": ms_e2e32 / args.steps,
                          "h2d_bytes_per_step": int(host_vid[0].numel() * 4 + host_ids[0].numel() * 8),
                          "note": "the loader's fp32 result shipped from the host (twice the PCIe bytes of `e2e`)"},
        "value_sync_loss_read": None if ms_sync is None else {
            "value": vols / (ms_sync / 1e3), "unit": "volumes/s", "ms_per_step": ms_sync / args.steps,
            "note": "reference behaviour: loss.item() inside forward (host sync between forward and backward)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "gemm_kernel<EPI, major> (tcgen05 128x256x64, all launches of one step)",
                     "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                     "peak_source": f"{peak_src} bf16_tflops_sustained", "traffic": traffic,
                     "traffic_note": f"mean dram__bytes_read+write per launch over {traffic_n} GEMM launches of one B=8 step "
                                     f"(ncu --set full: {', '.join('profiles/' + f for f in traffic_files) or 'no capture found'}; "
                                     "the round-1 captures predate the packed-fp32 epilogues and the mixed-major "
                                     "input-gradient products, which changed launch times but not operand bytes)",
                     "measurement_note": "event pairs serialise the boundary between consecutive kernels (~7 us on each of the "
                                         "mostly short launches): CUPTI kernel time of the same 315 launches is 19.6 ms = 911 "
                                         "TF/s = 0.675 of peak (profiles/r2_step_profile_cupti.txt, profiles/r2_launches_summary.md)",
                     "gemm_ms_per_step": gemm_ms, "gemm_launches_per_step": len(prof),
                     "gemm_share_of_step": gemm_ms / (ms_dev / args.steps),
                     "algorithmic_gflop_per_volume": gf_step, "ms_by_epilogue": by_epi},
        "loss": loss_val,
    }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_train_step_baseline(max_steps=1, warmup=1)       # 1 untimed + 1 timed step (~8 s each)
    if not args.no_torch_eager and world == 1:
        del clip, model, opt, dev_vid, host_vid
        torch.cuda.empty_cache()
        line["torch_eager_gpu"] = torch_eager_gpu_baseline(dev)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's train step on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_train_step_baseline(max_steps=1, warmup=0, budget_s=240.0):
    """One CT-CLIP train step (B=2: B=1 makes the contrastive loss identically 0) through the CPU
    oracle: CTViT fwd+bwd by autograd over oracle.ctclip_oracle, HF BertModel, loss, clip, Adam."""
    import torch
    from transformers import BertConfig, BertModel
    from oracle import ctclip_oracle as orc
    from vit_exp_b200.transformer_maskgit import CTViT
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    B = 2
    vit = CTViT(dim=512, codebook_size=8192, image_size=480, patch_size=20, temporal_patch_size=10,
                spatial_depth=4, temporal_depth=4, dim_head=32, heads=8)          # parameter container only
    bert = BertModel(BertConfig(vocab_size=30522, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0))
    wt = torch.nn.Parameter(torch.randn(512, 768) * 768 ** -0.5)
    wv = torch.nn.Parameter(torch.randn(512, 512) * 512 ** -0.5)
    temp = torch.nn.Parameter(torch.tensor(1.0))
    p = dict(vit.named_parameters())
    p.update(dict(vit.named_buffers()))
    train_params = [q for q in list(vit.parameters()) + list(bert.parameters()) + [wt, wv, temp] if q.numel() > 0]
    opt = torch.optim.Adam(train_params, lr=1.25e-6, betas=(0.9, 0.99))
    g = torch.Generator().manual_seed(0)
    video = torch.rand(B, 1, *VOL, generator=g)
    ids = torch.randint(0, 30522, (B, TEXT_LEN), generator=g)

    def one():
        enc_text = bert(ids, attention_mask=torch.ones_like(ids))[0]
        enc = orc.ctvit_forward(video, p, patch=20, tpatch=10, spatial_depth=4, temporal_depth=4, heads=8, vq=False)
        q, _, _, _ = orc.vq_cosine(enc.detach(), p["vq._codebook.embed"][0])
        tokens = enc + (q - enc).detach()                                       # straight-through
        loss, _, _ = orc.ctclip_loss(enc_text, tokens, {"to_text_latent.weight": wt, "to_visual_latent.weight": wv,
                                                        "temperature": temp})
        loss.backward()
        torch.nn.utils.clip_grad_norm_(train_params, 0.5)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss.item()

    times = []
    t_begin = time.time()
    for i in range(warmup + max_steps):
        t0 = time.time()
        loss = one()
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
        log(f"[cpu] step {i} {dt:.1f}s loss {loss:.4f}")
        if time.time() - t_begin + dt > budget_s:
            break
    if not times:
        times = [dt]
    sec = sum(times) / len(times)
    return {"value": B / sec, "unit": "volumes/s", "cores": cores, "kind": "port",
            "sample": f"{len(times)} full train step(s) of B=2 volumes (1x240x480x480) + 2x512-token reports, fp32, "
                      f"torch CPU oracle port (autograd), {sec:.1f} s/step", "steps_timed": len(times),
            "ms_per_step": sec * 1e3}


def torch_eager_gpu_baseline(dev, B=4, steps=3):
    """Informational: the same train step as plain PyTorch eager ops on THIS GPU - the oracle port of the reference
    modules (oracle/ctclip_oracle.py: einops-style gather, LayerNorm / Linear / softmax / conv3d ATen kernels, autograd)
    plus the stock HF BertModel, clip_grad_norm_ and torch.optim.Adam(fused=True).  This is the bar SURVEY 2a names
    ("stock PyTorch eager running the reference modules" on the B200); fp32 as the reference runs by default, and under
    bf16 autocast.  B = 4 volumes per step (fp32 autograd keeps the 576 x 576 attention matrices of every layer)."""
    import torch
    from transformers import BertConfig, BertModel
    from oracle import ctclip_oracle as orc
    from vit_exp_b200.transformer_maskgit import CTViT
    out = {"per_step_batch": B, "note": "oracle port (plain torch ops + autograd) + HF BertModel on cuda; informational"}
    torch.manual_seed(0)
    vit = CTViT(dim=512, codebook_size=8192, image_size=480, patch_size=20, temporal_patch_size=10,
                spatial_depth=4, temporal_depth=4, dim_head=32, heads=8)          # parameter container only
    is_param = {k for k, q in vit.named_parameters() if q.numel() > 0}
    p = {k: v.detach().to(dev).requires_grad_(k in is_param) for k, v in vit.state_dict().items()}
    bert = BertModel(BertConfig(vocab_size=30522, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)).to(dev).train()
    wt = torch.nn.Parameter(torch.randn(512, 768, device=dev) * 768 ** -0.5)
    wv = torch.nn.Parameter(torch.randn(512, 512, device=dev) * 512 ** -0.5)
    temp = torch.nn.Parameter(torch.tensor(1.0, device=dev))
    train = [q for q in p.values() if q.requires_grad] + list(bert.parameters()) + [wt, wv, temp]
    opt = torch.optim.Adam(train, lr=1.25e-6, betas=(0.9, 0.99), fused=True)
    g = torch.Generator().manual_seed(0)
    video = torch.rand(B, 1, *VOL, generator=g).to(dev)
    ids = torch.randint(0, 30522, (B, TEXT_LEN), generator=g).to(dev)
    mask = torch.ones_like(ids)

    def one(autocast):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            enc_text = bert(ids, attention_mask=mask)[0]
            enc = orc.ctvit_forward(video, p, patch=20, tpatch=10, spatial_depth=4, temporal_depth=4, heads=8, vq=False)
        enc = enc.float()
        q, _, _, _ = orc.vq_cosine(enc.detach(), p["vq._codebook.embed"][0])
        tokens = enc + (q - enc).detach()
        loss, _, _ = orc.ctclip_loss(enc_text.float(), tokens, {"to_text_latent.weight": wt, "to_visual_latent.weight": wv,
                                                               "temperature": temp})
        loss.backward()
        torch.nn.utils.clip_grad_norm_(train, 0.5)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss.item()

    for name, autocast in (("fp32", False), ("bf16_autocast", True)):
        try:
            torch.backends.cuda.matmul.allow_tf32 = False
            one(autocast)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = one(autocast)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": B / (ms / 1e3), "unit": "volumes/s", "ms_per_step": ms, "loss": loss}
        except Exception as e:                       # e.g. out of memory: report, do not fail the bench line
            out[name] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
            torch.cuda.empty_cache()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # every step is ~7-8 s of host time: W warm-up + K timed steps, cut short by the time budget (the line reports what ran)
    warm = min(args.warmup, 3)
    res = cpu_train_step_baseline(max_steps=args.steps, warmup=warm, budget_s=280.0)
    line = {
        "metric": "CT volumes/s, CT-CLIP train step", "value": res["value"], "unit": "volumes/s", "impl": "reference",
        "n_gpus": args.gpus, "steps": res["steps_timed"], "warmup": warm, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ctclip_train_step (CPU: bounded sample, B=2 volumes per step; the Python reference "
                               "cannot travel to the GPU box, so the pinned oracle port runs in its place)",
                   "per_gpu_batch": 2, "global_batch": 2, "text_len": TEXT_LEN, "parallelism": "cpu"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=PER_GPU_BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-eager", action="store_true",
                    help="skip the informational torch_eager_gpu leg (oracle port + HF BertModel as plain PyTorch eager ops on this GPU)")
    ap.add_argument("--bucket-mb", type=int, default=25, help="DDP gradient bucket size (MB)")
    ap.add_argument("--ddp", default="ignore-unused", choices=["ignore-unused", "find-unused", "static-graph"],
                    help="how DDP deals with the parameters that never get a gradient (see run_ours)")
    ap.add_argument("--sync-loss-read", action="store_true",
                    help="CTCLIP returns cl_loss via loss.item() inside forward (reference behaviour: a host sync between "
                         "forward and backward); default: config['defer_loss_read'], the same value read back at the end of the step")
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"],
                    help="fused: libctk clip+Adam (2 launches); torch: clip_grad_norm_ + torch.optim.Adam(fused=True)")
    ap.add_argument("--text-dropout", type=float, default=0.1,
                    help="hidden / attention dropout of the random-init BERT-base text tower: 0.1 as CXR-BERT's config ships "
                         "and the reference trains with (the tower is in train mode); 0 switches it off")
    ap.add_argument("--text-tower", default="ctk", choices=["hf", "ctk"],
                    help="ctk: the BertModel's forward/backward run through libctk (vit_exp_b200/text_tower.py, CUDA-graph "
                         "replay); hf: the module runs as passed (stock PyTorch under bf16 autocast)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)          # timing rule: >= 3 warm-up steps (they also cover the CUDA-graph captures)
    protect_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
