"""dev tool: device time of the forward_old pooling head at full size: mean over 24 frames of
24*24*512 floats, Linear(294912 -> 512) + l2norm forward / backward, B = 8.  CUDA events, L2 flushed between launches.
    python tools/bench_legacy_pool.py > gpurun_out/legacy_pool.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vit_exp_b200 import ops

dev = torch.device("cuda:0")
B, t, width, dl = 8, 24, 24 * 24 * 512, 512
tokens = torch.randn(B, t, width, device=dev)
W = torch.randn(dl, width, device=dev) * width ** -0.5
dlat = torch.randn(B, dl, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


pooled = ops.mean_pool(tokens)
lat, rn = ops.latent_fwd(pooled, W)
res = {
    "mean_frames_ms": timed(lambda: ops.mean_pool(tokens)),
    "mean_frames_bytes": tokens.numel() * 4 + pooled.numel() * 4,
    "latent_fwd_wide_ms": timed(lambda: ops.latent_fwd(pooled, W)),
    "latent_fwd_wide_bytes": W.numel() * 4,
    "latent_bwd_wide_ms": timed(lambda: ops.latent_bwd(dlat, lat, rn, pooled, W)),
    "latent_bwd_wide_bytes": W.numel() * 4 * 2,
}
for k in ("mean_frames", "latent_fwd_wide", "latent_bwd_wide"):
    res[k + "_GBps"] = res[k + "_bytes"] / res[k + "_ms"] / 1e6
print(json.dumps(res))
