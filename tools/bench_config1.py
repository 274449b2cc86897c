"""BASELINE config 1: CLIP symmetric InfoNCE loss (demo_tests/clip_loss.py style) on synthetic 512-d latents,
batch 8, single process on CPU - the reference's own CPU-runnable case.

    python tools/bench_config1.py            # CPU arm only (runs anywhere)
    python tools/bench_config1.py --gpu      # + libctk's fused loss kernel on cuda:0 for the same inputs

CPU arm = the pinned oracle restatements of the two reference formulations (open_clip-style ClipLoss,
demo_tests/clip_loss.py:104-128, and CT-CLIP's exp / diag / log form, ct_clip.py:1332-1382), forward + backward with
autograd, fp32, 1000 iterations after 100 warm-up iterations, on all host cores torch uses.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from oracle import ctclip_oracle as orc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--iters", type=int, default=1000)
    args = ap.parse_args()
    N, d = 8, 512
    T = F.normalize(torch.randn(N, d, generator=torch.Generator().manual_seed(0)), dim=-1)
    I = F.normalize(torch.randn(N, d, generator=torch.Generator().manual_seed(1)), dim=-1)
    lt = torch.tensor(1.0)                                   # logit_scale = e^1

    def run(fn):
        t = T.clone().requires_grad_(True)
        i = I.clone().requires_grad_(True)
        s = lt.clone().requires_grad_(True)
        loss = fn(t, i, s)
        loss.backward()
        return float(loss.detach())

    forms = {"open_clip ClipLoss (demo_tests/clip_loss.py:104-128)": lambda t, i, s: orc.clip_loss_open_clip(t, i, s.exp()),
             "CT-CLIP exp/diag/log (ct_clip.py:1332-1382), / bs_single_gpu": lambda t, i, s: orc.clip_loss_reference_form(t, i, s, N)}
    for name, fn in forms.items():
        for _ in range(100):
            run(fn)
        t0 = time.perf_counter()
        for _ in range(args.iters):
            val = run(fn)
        us = (time.perf_counter() - t0) / args.iters * 1e6
        print(json.dumps({"config": 1, "arm": "cpu", "form": name, "N": N, "d": d, "loss": val, "us_per_iter_fwd_bwd": us,
                          "cores": torch.get_num_threads(), "iters": args.iters}), flush=True)
    if args.gpu:
        from vit_exp_b200 import ops
        dev = torch.device("cuda:0")
        Tg, Ig, ltg = T.to(dev), I.to(dev), lt.reshape(1).to(dev)
        for _ in range(100):
            out, _ = ops.clip_loss_fwd_bwd(Tg, Ig, ltg, b_local=N, row0=0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            out, dl = ops.clip_loss_fwd_bwd(Tg, Ig, ltg, b_local=N, row0=0)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"config": 1, "arm": "libctk ctk_clip_loss_fwd_bwd (fused fwd + bwd)", "N": N, "d": d,
                          "loss": float(out[0]), "us_per_iter_fwd_bwd": e0.elapsed_time(e1) / args.iters * 1e3,
                          "iters": args.iters}), flush=True)


if __name__ == "__main__":
    main()
