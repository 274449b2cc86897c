"""Import the REAL reference modules from /root/reference (build container only).

Used by oracle/make_golden.py to pin the oracle; never at test/bench run time on the GPU box
(the reference does not travel). Recipe verified in SURVEY.md appendix D: stub the unrelated
third-party imports, provide a VectorQuantize stand-in (vector-quantize-pytorch is not installed),
and replace the one method that hard-codes device='cuda' (attention.py:366) by a device-agnostic
copy so the spatial stack runs on CPU.
"""
from __future__ import annotations

import sys
import types

import torch
from torch import nn

REF_ROOT = "/root/reference"


class _VQStandIn(nn.Module):
    """Stand-in with the call signature of vector_quantize_pytorch.VectorQuantize (ctvit.py:188,403).
    Follows oracle.ctclip_oracle.vq_cosine (eval-mode semantics, state kept in `_codebook.*` buffers)."""

    def __init__(self, dim, codebook_size, use_cosine_sim=True, **kw):
        super().__init__()
        cb = nn.Module()
        embed = torch.nn.functional.normalize(
            torch.nn.init.kaiming_uniform_(torch.empty(1, codebook_size, dim)), dim=-1)
        cb.register_buffer("initted", torch.tensor([True]))
        cb.register_buffer("cluster_size", torch.zeros(1, codebook_size))
        cb.register_buffer("embed", embed)
        self._codebook = cb
        self.pre_vq = None

    def forward(self, x, mask=None):
        from oracle.ctclip_oracle import vq_cosine
        self.pre_vq = x.detach().clone()
        q, ind, _, _ = vq_cosine(x, self._codebook.embed[0])
        if self.training:
            q = x + (q - x).detach()
        return q, ind, torch.zeros(1)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns (CTViT, CTCLIP, attention_module, AllGather)."""
    import transformers  # noqa: F401  must precede the accelerate stub
    from transformers import BertTokenizer

    for p in (f"{REF_ROOT}/transformer_maskgit", f"{REF_ROOT}/CT_CLIP"):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "vector_quantize_pytorch" not in sys.modules:
        _stub("segmentation_models_pytorch")
        _stub("segmentation_models_pytorch.losses", TverskyLoss=type("TverskyLoss", (nn.Module,), {}))
        _stub("matplotlib")
        _stub("matplotlib.pyplot")
        _stub("nibabel")
        _stub("accelerate", Accelerator=object, DistributedDataParallelKwargs=object)
        _stub("ema_pytorch", EMA=object)
        _stub("wandb")
        _stub("vector_quantize_pytorch", VectorQuantize=_VQStandIn)
    BertTokenizer.from_pretrained = classmethod(lambda cls, *a, **k: None)   # ct_clip.py:650 needs the hub

    from transformer_maskgit import attention as ref_attention
    from transformer_maskgit.ctvit import CTViT
    from ct_clip.ct_clip import CTCLIP
    from ct_clip.distributed import AllGather

    def cpb_forward(self, *dimensions, device=torch.device("cpu")):
        # device-agnostic copy of attention.py:363-382 (the original forces device='cuda')
        dev = self.net[0][0].weight.device
        positions = [torch.arange(d, device=dev) for d in dimensions]
        grid = torch.stack(torch.meshgrid(*positions, indexing="ij"))
        grid = grid.reshape(grid.shape[0], -1).T
        rel_pos = grid[:, None, :] - grid[None, :, :]
        if self.log_dist:
            rel_pos = torch.sign(rel_pos) * torch.log(rel_pos.abs() + 1)
        rel_pos = rel_pos.to(torch.float32)
        for layer in self.net:
            rel_pos = layer(rel_pos.float())
        return rel_pos.permute(2, 0, 1)

    ref_attention.ContinuousPositionBias.forward = cpb_forward
    return CTViT, CTCLIP, ref_attention, AllGather


class FakeAccelerator:
    """accelerator protocol of distributed.py:11-14 for a single process."""
    num_processes = 1
    process_index = 0

    def gather(self, x):
        return x
