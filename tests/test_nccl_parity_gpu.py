"""Two-GPU NCCL parity (torchrun, one rank per GPU): the distributed gradient convention of the reference on hardware.

Every rank evaluates the full N x N loss on the all-gathered latents, keeps only its own rows of dT / dI
(distributed.py:18-20 - no reduction over ranks) and DDP averages parameter gradients, so every encoder / projection
parameter ends up with (1/W) d(loss_global)/d(theta) while the temperature, differentiated by every rank through the
full replicated loss (ct_clip.py:1343-1347), keeps its undivided gradient.  The answer comes from autograd through the
CPU oracle on the concatenated batch (codes taken from the ranks: the straight-through forward value of the tokens is
embed[ind]).  Needs >= 2 GPUs (`gpurun --gpus 2`); on a 1-GPU box it reports itself as skipped.
"""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_gpu_nccl_step_matches_global_gradient(tmp_path):
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device: there is no CPU path")
    if torch.cuda.device_count() < WORLD:
        pytest.skip("needs 2 GPUs (run under `gpurun --gpus 2`); the gloo twin tests/test_ddp_step_cpu.py covers the host logic")
    out = str(tmp_path / "nccl")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={WORLD}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "nccl_parity_worker.py"), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = [torch.load(f"{out}.{k}", weights_only=False) for k in range(WORLD)]

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import nccl_parity_worker as W
    from oracle import ctclip_oracle as O
    clip, video, ids = W.build(WORLD)
    vit = clip.visual_transformer
    is_param = {k for k, _ in vit.named_parameters()}
    p = {k: v.detach().clone().requires_grad_(k in is_param and v.numel() > 0) for k, v in vit.state_dict().items()}
    emb = clip.text_transformer.emb.weight.detach().clone().requires_grad_()
    wt = clip.to_text_latent.weight.detach().clone().requires_grad_()
    wv = clip.to_visual_latent.weight.detach().clone().requires_grad_()
    temp = clip.temperature.detach().clone().requires_grad_()
    enc = O.ctvit_forward(video, p, patch=20, tpatch=10, spatial_depth=1, temporal_depth=1, heads=8, vq=False)
    ind = torch.cat([res[k]["ind"].reshape(W.B, -1) for k in range(WORLD)], dim=0)            # rank-major, like the gather
    quant = p["vq._codebook.embed"][0].detach()[ind.reshape(-1)].reshape(enc.shape)
    tokens = enc + (quant - enc).detach()
    loss, _, _ = O.ctclip_loss(emb[ids], tokens, {"to_text_latent.weight": wt, "to_visual_latent.weight": wv,
                                                  "temperature": temp}, b_local=W.B)
    loss.backward()
    want = {"text_transformer.emb.weight": emb.grad, "to_text_latent.weight": wt.grad,
            "to_visual_latent.weight": wv.grad, "temperature": temp.grad}
    want.update({"visual_transformer." + k: v.grad for k, v in p.items() if v.requires_grad and v.grad is not None})

    ref = float(loss.detach())
    for k in range(WORLD):
        assert abs(res[k]["loss"] - ref) <= 1e-3 * abs(ref), (k, res[k]["loss"], ref)   # every rank sees the GLOBAL loss
        assert abs(res[k]["cl_loss"] - res[k]["loss"]) < 1e-6
    assert abs(res[0]["loss"] - res[1]["loss"]) <= 1e-6 * abs(ref)
    errs, checked = {}, 0
    gmax = max(v.abs().max().item() for v in want.values())
    for n, g0 in res[0]["grads"].items():
        g1 = res[1]["grads"][n]
        if g0 is None:
            assert g1 is None and (n not in want or want[n].abs().max() == 0), n
            continue
        assert torch.equal(g0, g1), n                                   # the all-reduce left both ranks with the same average
        if n not in want or g0.numel() == 0:
            continue
        # temperature: every rank holds the FULL gradient, DDP's mean leaves it undivided (reference behaviour)
        r = want[n] if n == "temperature" else want[n] / WORLD
        if r.abs().max().item() < 1e-7 * gmax:                          # cpb net.2.bias: softmax shift invariance -> 0
            assert g0.abs().max().item() < 1e-3 * gmax, n
            continue
        errs[n] = float((g0.double() - r.double()).norm() / r.double().norm().clamp_min(1e-30))
        checked += 1
    print("worst gradient errors:", sorted(errs.items(), key=lambda kv: -kv[1])[:6])
    bad = {k: v for k, v in errs.items() if v > 5e-2}
    assert not bad, bad
    assert errs["temperature"] < 2e-2 and errs["to_text_latent.weight"] < 2e-2 and errs["to_visual_latent.weight"] < 2e-2
    assert checked > 30
