"""world_size-2 gloo test of the host-side multi-rank logic (no GPU): the accelerator shim gathers
rank-major, AllGather's backward keeps only the local slice (distributed.py:18-20), and the loss each
rank evaluates on the gathered latents reproduces the single-process oracle."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ctclip_oracle as orc
    from vit_exp_b200.ct_clip import AllGather, TorchDistAccelerator
    acc = TorchDistAccelerator()
    assert acc.num_processes == world and acc.process_index == rank
    B, d = 3, 16
    g = torch.Generator().manual_seed(0)
    T_all = orc.l2norm(torch.randn(world * B, d, generator=g).double())
    I_all = orc.l2norm(torch.randn(world * B, d, generator=g).double())
    lt = torch.tensor(0.3).double()
    tl = T_all[rank * B:(rank + 1) * B].clone().requires_grad_(True)
    il = I_all[rank * B:(rank + 1) * B].clone().requires_grad_(True)
    ltp = lt.clone().requires_grad_(True)
    Tg, Ig = AllGather.apply(tl, acc), AllGather.apply(il, acc)
    ok_gather = torch.equal(Tg.detach(), T_all) and torch.equal(Ig.detach(), I_all)
    loss = orc.clip_loss_reference_form(Tg, Ig, ltp, B)
    loss.backward()
    ref = orc.clip_loss_and_local_grads(T_all, I_all, lt, B, rank)
    res = dict(ok_gather=ok_gather,
               loss_err=abs(loss.item() - ref["loss"].item()),
               dT_err=(tl.grad - ref["dT_local"]).abs().max().item(),
               dI_err=(il.grad - ref["dI_local"]).abs().max().item(),
               dtemp_err=abs(ltp.grad.item() - ref["dlog_temp"].item()))
    torch.save(res, f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_and_loss(tmp_path):
    world = 2
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(f"{out}.{r}")
        assert res["ok_gather"]
        assert res["loss_err"] < 1e-12 and res["dT_err"] < 1e-12 and res["dI_err"] < 1e-12 and res["dtemp_err"] < 1e-12
