"""Optimizer tail of the CT-CLIP train step on libctk.

The reference trainer does `accelerator.clip_grad_norm_(params, max_grad_norm)` followed by
`optim.step()` with Adam (AdamW when weight_decay > 0): CTCLIPTrainer.py:711-715, optimizer.py:14-24.
`FusedClipAdam` does both in two kernel launches over a device table of tensor pointers.
"""
from __future__ import annotations

from typing import Iterable, Optional

import numpy as np
import torch

from . import _lib
from .ops import _stream, check


class FusedClipAdam(torch.optim.Optimizer):
    """torch.optim.Adam / AdamW semantics (no amsgrad) with clip_grad_norm_(max_grad_norm, 2.0) folded in.

    * fp32 CUDA parameters and gradients only (the train step's master weights); parameters whose
      `.grad` is None are skipped, as torch does.
    * `step()` returns the total gradient norm before clipping as a 0-d device tensor (what
      clip_grad_norm_ returns); nothing is synchronised with the host.
    * the step count lives in `state[p]["step"]` (torch.optim.Adam's key; a plain int here), so
      state_dict() / load_state_dict() round-trip the bias correction and the state is interchangeable with
      torch.optim.Adam's; parameters that joined later (first gradient at a later step) get their own launch.
    * `write_clipped_grads=True` also writes the scaled gradients back to `.grad` (clip_grad_norm_ does;
      costs one more pass of writes and is off by default).
    """

    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_grad_norm: Optional[float] = None, decoupled_weight_decay: Optional[bool] = None,
                 write_clipped_grads: bool = False):
        if decoupled_weight_decay is None:
            decoupled_weight_decay = weight_decay > 0          # optimizer.py:20-24: AdamW iff wd > 0
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled_weight_decay)
        super().__init__(params, defaults)
        self.max_grad_norm = max_grad_norm
        self.write_clipped_grads = write_clipped_grads
        self._lib = _lib.load()
        self._chunk = self._lib.ctk_opt_chunk_elems()
        self._sq = None
        self._host = None
        self._dev = None

    def _rows(self, group):
        """{step: [(p, g, m, v, numel)]} of the group's parameters that have a gradient; advances their step counts"""
        by_step = {}
        for p in group["params"]:
            g = p.grad
            if g is None:
                continue
            assert p.is_cuda and p.dtype == torch.float32 and g.dtype == torch.float32, "FusedClipAdam: fp32 CUDA tensors only"
            assert p.is_contiguous() and g.is_contiguous()
            st = self.state[p]
            if not st:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            # a plain number (torch.optim.Adam accepts one when it loads a state dict; a state loaded from torch's Adam
            # arrives as a 0-d tensor and is converted here): a tensor increment per parameter cost 1.5 ms of host time
            step = st["step"] = int(st["step"]) + 1
            by_step.setdefault(step, []).append(
                (p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()))
        return by_step

    def _with_chunks(self, rows):
        out, c0 = [], 0
        for r in rows:
            out.append(r + (c0,))
            c0 += (r[4] + self._chunk - 1) // self._chunk
        return out, c0

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None, "FusedClipAdam does not take a closure"
        # one launch per (param group, step count); in a normal run every parameter shares one step count
        groups = [(g, step, rows) for g in self.param_groups for step, rows in sorted(self._rows(g).items())]
        if not groups:
            return None
        dev = self.param_groups[0]["params"][0].device
        # tables: [all tensors (for the norm)] + [one per group when there are several], uploaded in one copy
        parts = [self._with_chunks([r for _, _, rows in groups for r in rows])]
        if len(groups) > 1:
            parts += [self._with_chunks(rows) for _, _, rows in groups]
        host = np.asarray([r for rows, _ in parts for r in rows], dtype=np.int64).reshape(-1, 6)
        s = _stream()
        n = host.shape[0]
        if self._host is None or self._host[0].shape[0] < n:
            # two pinned staging buffers (alternating) and one device table
            self._host = [torch.empty(2 * n + 16, 6, dtype=torch.int64).pin_memory() for _ in range(2)]
            self._dev = torch.empty(self._host[0].shape, dtype=torch.int64, device=dev)
            self._sq = torch.zeros(1, dtype=torch.float32, device=dev)
            self._last = None
            self._flip = 0
            self._staged = [None, None]          # event recorded after the copy kernel that read staging buffer i
        if self._last is None or self._last.shape != host.shape or not np.array_equal(self._last, host):
            # addresses changed (first step, re-allocated gradients): upload with a kernel that reads the pinned
            # buffer over PCIe - a DMA copy would queue behind the input batch's H2D transfer on the copy engine
            self._flip ^= 1
            hb = self._host[self._flip]
            if self._staged[self._flip] is not None:
                # the copy kernel that last read this buffer may still be queued when the host runs steps ahead
                self._staged[self._flip].synchronize()
            hb[:n].copy_(torch.from_numpy(host))
            check(self._lib.ctk_copy_from_pinned(self._dev.data_ptr(), hb.data_ptr(), n * 48, s), "ctk_copy_from_pinned")
            ev = torch.cuda.Event()
            ev.record()
            self._staged[self._flip] = ev
            self._last = host
        clip = self.max_grad_norm is not None and self.max_grad_norm > 0
        row_bytes = 6 * 8
        if clip:
            rows, nchunks = parts[0]
            check(self._lib.ctk_multi_sqnorm(self._dev.data_ptr(), len(rows), nchunks, self._sq.data_ptr(), s), "ctk_multi_sqnorm")
        off = 0 if len(groups) == 1 else len(parts[0][0])
        for k, (group, step, _) in enumerate(groups):
            rows, nchunks = parts[0] if len(groups) == 1 else parts[1 + k]
            b1, b2 = group["betas"]
            check(self._lib.ctk_multi_adam(self._dev.data_ptr() + off * row_bytes, len(rows), nchunks, self._sq.data_ptr(),
                                           float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                           float(group["weight_decay"]), int(bool(group["decoupled"])), step,
                                           float(self.max_grad_norm) if clip else 0.0, int(self.write_clipped_grads), s),
                  "ctk_multi_adam")
            off += len(rows)
        return self._sq.sqrt().reshape(()) if clip else None
