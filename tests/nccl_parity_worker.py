"""Rank body of tests/test_nccl_parity_gpu.py (launched by torchrun, one rank per GPU, NCCL).

Every rank builds the same seeded CT-CLIP (full CT-CLIP geometry: 480x480x240 volumes, patch 20x20x10, dim 512,
heads 8 x 32, codebook 8192; one spatial + one temporal layer so that the CPU oracle finishes in seconds), wraps it in
DistributedDataParallel(find_unused_parameters=True) exactly as the reference trainer does (CTCLIPTrainer.py:318), runs
CTCLIP.forward on ITS slice of the global batch, backward, and writes loss / gradients / code indices to disk.
"""
import os
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
B = 2                   # volumes per rank


class Text(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.emb = torch.nn.Embedding(64, 768)
        self.unused = torch.nn.Linear(3, 3)          # like BERT's pooler: never reached (find_unused_parameters)

    def forward(self, input_ids, attention_mask=None):
        return (self.emb(input_ids),)


def build(world):
    from vit_exp_b200.ct_clip import CTCLIP
    from vit_exp_b200.transformer_maskgit import CTViT
    torch.manual_seed(0)
    vit = CTViT(dim=512, codebook_size=8192, image_size=480, patch_size=20, temporal_patch_size=10, spatial_depth=1,
                temporal_depth=1, dim_head=32, heads=8)
    clip = CTCLIP(image_encoder=vit, text_encoder=Text(), dim_text=768, dim_image=512, dim_latent=512, config={})
    with torch.no_grad():
        clip.temperature.fill_(0.9)
    g = torch.Generator().manual_seed(1)
    video = torch.rand(world * B, 1, 240, 480, 480, generator=g)
    ids = torch.randint(0, 64, (world * B, 4), generator=g)
    return clip.eval(), video, ids           # eval: frozen codebook (the training-mode EMA is not a gradient path)


def main():
    out = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from vit_exp_b200.ct_clip import TorchDistAccelerator
    clip, video, ids = build(world)
    clip = clip.to(dev)
    for q in clip.parameters():
        q.requires_grad_(True)
    model = torch.nn.parallel.DistributedDataParallel(clip, device_ids=[local], find_unused_parameters=True)
    sl = slice(rank * B, (rank + 1) * B)
    with torch.no_grad():
        _, ind, _ = clip.visual_transformer.encode_with_aux(video[sl].to(dev))
    batch = {"data_type": ["imagereport"] * B, "image": video[sl].to(dev),
             "text": SimpleNamespace(input_ids=ids[sl].to(dev), attention_mask=None)}
    loss, ld = model(batch, device=dev, accelerator=TorchDistAccelerator())
    loss.backward()
    torch.cuda.synchronize()
    grads = {n: (p.grad.detach().float().cpu() if p.grad is not None else None) for n, p in clip.named_parameters()}
    torch.save(dict(loss=float(loss.detach()), cl_loss=float(ld["cl_loss"]), grads=grads, ind=ind.cpu()), f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
