"""Two-rank DDP train step on CPU (gloo) through the REAL module code - `CTCLIP.forward` -> `_CTViTEncode` ->
`_ClipHead` -> `loss.backward()` under `DistributedDataParallel(find_unused_parameters=True)` - with the libctk entry
points replaced by the torch doubles of tests/emulated_ops.py (fp32 operands).

Checks the distributed gradient convention of SURVEY 8a: every rank evaluates the full N x N loss on the gathered
latents, keeps only its own rows of dT / dI (distributed.py:18-20, no reduction), and DDP averages parameter
gradients, so each encoder / projection parameter ends up with (1/W) * d(loss_global)/d(theta) - and the temperature,
which every rank differentiates through the full replicated loss, with the undivided gradient.  The single-process
answer comes from autograd through the pinned oracle on the concatenated batch.
"""
import os
import socket
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, WORLD = 2, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build():
    """identical model on every rank / in the checker (seeded)"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emulated_ops as E
    from vit_exp_b200 import ct_clip as CC
    from vit_exp_b200 import transformer_maskgit as TM

    class _Ops:
        def __getattr__(self, name):
            full = {"gemm": E.gemm_full, "cast_bf16": E.cast_bf16_full, "transpose_cast_bf16": E.transpose_cast_bf16_full,
                    "layernorm_fwd": E.layernorm_fwd_full, "layernorm_bwd": E.layernorm_bwd_full}
            return full[name] if name in full else getattr(E, name)

    saved = (E.OPERAND, TM.ops, CC.ops, torch.empty, torch.zeros)
    E.OPERAND = torch.float32
    TM.ops = CC.ops = _Ops()
    _empty, _zeros = torch.empty, torch.zeros

    def _f32(fn):
        def wrapped(*a, **k):
            if k.get("dtype") is torch.bfloat16:
                k["dtype"] = torch.float32
            return fn(*a, **k)
        return wrapped
    torch.empty, torch.zeros = _f32(_empty), _f32(_zeros)

    class _CpuViT(TM.CTViT):
        """the public entry asserts a CUDA input (there is no CPU path); the doubles stand in for the kernels here"""
        def encode_with_aux(self, video, _launched=None):
            return TM._CTViTEncode.apply(self, video.contiguous().float(), self.training, None, *self._flat_params())

    class _Text(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = torch.nn.Embedding(40, 24)
            self.unused = torch.nn.Linear(3, 3)          # like BERT's pooler: never reached (find_unused_parameters)

        def forward(self, input_ids, attention_mask=None):
            return (self.emb(input_ids),)

    torch.manual_seed(0)
    vit = _CpuViT(dim=64, codebook_size=32, image_size=(8, 8), patch_size=(4, 4), temporal_patch_size=2,
                  spatial_depth=1, temporal_depth=1, dim_head=32, heads=2)
    vit.cuda_graphs = False
    clip = CC.CTCLIP(image_encoder=vit, text_encoder=_Text(), dim_text=24, dim_image=64, dim_latent=16,
                     config={"overlap_text_encoder": False})
    with torch.no_grad():
        clip.temperature.fill_(0.9)
    g = torch.Generator().manual_seed(1)
    video = torch.rand(WORLD * B, 1, 4, 8, 8, generator=g)
    ids = torch.randint(0, 40, (WORLD * B, 5), generator=g)
    def restore():                               # the checker runs inside the pytest process: undo the patches there
        E.OPERAND, TM.ops, CC.ops, torch.empty, torch.zeros = saved
    # eval: frozen codebook (the training-mode EMA is not a gradient path)
    return clip.eval(), video, ids, CC, restore


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    clip, video, ids, CC, _ = _build()
    model = torch.nn.parallel.DistributedDataParallel(clip, find_unused_parameters=True)
    sl = slice(rank * B, (rank + 1) * B)
    batch = {"data_type": ["imagereport"] * B, "image": video[sl],
             "text": SimpleNamespace(input_ids=ids[sl], attention_mask=None)}
    loss, ld = model(batch, device=None, accelerator=CC.TorchDistAccelerator())
    loss.backward()
    grads = {n: (p.grad.clone() if p.grad is not None else None) for n, p in clip.named_parameters()}
    torch.save(dict(loss=float(loss.detach()), cl_loss=float(ld["cl_loss"]), grads=grads), f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ddp_step_matches_global_gradient(tmp_path):
    out = str(tmp_path / "ddp")
    mp.spawn(_worker, args=(WORLD, _free_port(), out), nprocs=WORLD, join=True)
    res = [torch.load(f"{out}.{r}", weights_only=False) for r in range(WORLD)]

    # single-process answer: the oracle on the concatenated batch, loss / bs_single_gpu (ct_clip.py:1379)
    from oracle import ctclip_oracle as O
    clip, video, ids, _, restore = _build()
    restore()                                    # only the seeded model and inputs are needed here; the oracle computes
    vit = clip.visual_transformer
    p = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in vit.state_dict().items()}
    emb = clip.text_transformer.emb.weight.detach().clone().requires_grad_()
    wt = clip.to_text_latent.weight.detach().clone().requires_grad_()
    wv = clip.to_visual_latent.weight.detach().clone().requires_grad_()
    temp = clip.temperature.detach().clone().requires_grad_()
    enc = O.ctvit_forward(video, p, patch=4, tpatch=2, spatial_depth=1, temporal_depth=1, heads=2, vq=False)
    q, _, _, _ = O.vq_cosine(enc.detach(), p["vq._codebook.embed"][0].detach())
    tokens = enc + (q - enc).detach()                                   # straight-through (ctvit.py:403)
    loss, _, _ = O.ctclip_loss(emb[ids], tokens, {"to_text_latent.weight": wt, "to_visual_latent.weight": wv,
                                                  "temperature": temp}, b_local=B)
    loss.backward()
    want = {"text_transformer.emb.weight": emb.grad, "to_text_latent.weight": wt.grad,
            "to_visual_latent.weight": wv.grad, "temperature": temp.grad}
    want.update({"visual_transformer." + k: v.grad for k, v in p.items() if v.requires_grad and v.grad is not None})

    for r in range(WORLD):
        assert abs(res[r]["loss"] - float(loss.detach())) < 1e-5 * abs(float(loss.detach()))          # every rank sees the global loss
        assert abs(res[r]["cl_loss"] - res[r]["loss"]) < 1e-7
    checked = 0
    for n, g0 in res[0]["grads"].items():
        g1 = res[1]["grads"][n]
        if g0 is None:
            assert g1 is None and (n not in want or want[n].abs().max() == 0), n
            continue
        assert torch.equal(g0, g1), n                                               # DDP left both ranks with the average
        if n == "visual_transformer.spatial_rel_pos_bias.net.2.bias":               # softmax shift invariance: zero
            continue
        # the temperature enters the replicated loss directly on every rank (ct_clip.py:1343-1347), so each rank holds
        # its FULL gradient and DDP's mean leaves it undivided - W times larger, relative to every other parameter,
        # than in a single-process run; that is the reference's behaviour and it is reproduced
        ref = want[n] if n == "temperature" else want[n] / WORLD
        err = float((g0 - ref).norm() / ref.norm().clamp_min(1e-30))
        assert err < 1e-3, (n, err)
        checked += 1
    assert checked > 30
