"""Side benchmarks for the other BASELINE.json configs (not the driver's bench line):
  config 2: CTViT encoder forward, one 1x240x480x480 volume (and B=8), ms and volumes/s
  config 4: zero-shot scoring of volumes against 36 prompt latents (per-volume latency)
  config 5: contrastive-loss sweep N in {256..4096}: fused loss fwd+bwd, microseconds
Run on one B200:  python tools/bench_configs.py > gpurun_out/configs.jsonl
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_exp_b200 import ops
from vit_exp_b200.transformer_maskgit import CTViT

dev = torch.device("cuda:0")


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


torch.manual_seed(0)
vit = CTViT(dim=512, codebook_size=8192, image_size=480, patch_size=20, temporal_patch_size=10, spatial_depth=4,
            temporal_depth=4, dim_head=32, heads=8).to(dev).eval()
for B in (1, 8):
    vid = torch.rand(B, 1, 240, 480, 480, device=dev)
    with torch.no_grad():
        ms = timeit(lambda: vit(vid, return_encoded_tokens=True), iters=10)
    gf = 673.4 + 116.0
    print(json.dumps({"config": 2, "workload": "CTViT encoder forward (eval, incl. VQ)", "batch": B, "ms": ms,
                      "volumes_per_s": B / ms * 1e3, "algorithmic_tflops": gf * B / ms}), flush=True)

# config 4: per-volume zero-shot: encoder + pooled latent + 36 prompt logits
wv = torch.randn(512, 512, device=dev) * 512 ** -0.5
tl = torch.nn.functional.normalize(torch.randn(36, 512, device=dev), dim=-1)
lt = torch.ones(1, device=dev)
vid = torch.rand(1, 1, 240, 480, 480, device=dev)
def zero_shot():
    with torch.no_grad():
        tok = vit(vid, return_encoded_tokens=True)
        pooled = ops.mean_pool(tok.reshape(1, -1, 512))
        il, _ = ops.latent_fwd(pooled, wv)
        logits = ops.pair_logits(tl, il[0].contiguous(), lt)
        return logits.view(18, 2).softmax(dim=-1)[:, 0]
ms = timeit(zero_shot, iters=10)
print(json.dumps({"config": 4, "workload": "zero-shot 18 pathologies x 2 prompts per volume", "ms_per_volume": ms,
                  "volumes_per_s_per_gpu": 1e3 / ms}), flush=True)

# config 5: loss sweep (single rank holds all N rows; W=8 ranks would each run the same full N x N pass)
for N in (256, 512, 1024, 2048, 4096):
    T = torch.nn.functional.normalize(torch.randn(N, 512, device=dev), dim=-1)
    I = torch.nn.functional.normalize(torch.randn(N, 512, device=dev), dim=-1)
    for W in (1, 8):
        b = N // W
        ms = timeit(lambda: ops.clip_loss_fwd_bwd(T, I, lt, b_local=b, row0=0), iters=20)
        flops = 2.0 * N * N * 512 * (1 + 2.0 / W + 2.0 / W)     # logits + local stripes recompute + dT/dI products
        print(json.dumps({"config": 5, "N": N, "world": W, "b_local": b, "us": ms * 1e3,
                          "fp32_tflops": flops / ms / 1e9}), flush=True)
