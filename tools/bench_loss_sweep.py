"""BASELINE config 5: contrastive-loss scaling sweep - global batch N in {256..4096}, latent dim 512: the packed
all-gather of the l2-normalised latents (NCCL over NVLink) + similarity + fused symmetric cross-entropy forward and
backward (dT_local, dI_local, d log-temperature), at 1/2/4/8 GPUs.

    python tools/bench_loss_sweep.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
           tools/bench_loss_sweep.py
    CTK_CLIP_LOSS_TC=0 ... (force the fp32 SIMT tiles; the tensor-core path is the default for N >= 1024)

One JSON line per N from rank 0: microseconds per step (device-timed, max over ranks), split into gather and loss.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from vit_exp_b200 import ops


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    d = 512
    lt = torch.ones(1, device=dev)
    for N in (256, 512, 1024, 2048, 4096):
        B = N // world
        g = torch.Generator().manual_seed(rank)
        tl = torch.nn.functional.normalize(torch.randn(B, d, generator=g), dim=-1).to(dev)
        il = torch.nn.functional.normalize(torch.randn(B, d, generator=g), dim=-1).to(dev)
        packed = torch.cat([tl, il], dim=1).contiguous()                    # one gather instead of two (ct_clip.py:1329-1330)
        gathered = torch.empty(N, 2 * d, device=dev)

        def gather():
            if world > 1:
                dist.all_gather_into_tensor(gathered, packed)
            else:
                gathered.copy_(packed)
            return gathered[:, :d].contiguous(), gathered[:, d:].contiguous()

        def step():
            T, I = gather()
            return ops.clip_loss_fwd_bwd(T, I, lt, b_local=B, row0=rank * B)

        def timeit(fn, iters=20, warm=5):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return ms.item() * 1e3

        us_step, us_gather = timeit(step), timeit(gather)
        out, _ = step()
        if rank == 0:
            flops = 2.0 * N * N * d * (1 + 2.0 / world)          # full N x N similarity + this rank's two gradient products
            print(json.dumps({"config": 5, "N": N, "n_gpus": world, "b_local": B, "us_per_step": us_step,
                              "us_gather": us_gather, "loss": float(out[0]), "algorithmic_tflops": flops / us_step * 1e-6,
                              "tensor_core_path": os.environ.get("CTK_CLIP_LOSS_TC", "1") != "0" and N >= 1024}),
                  flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
