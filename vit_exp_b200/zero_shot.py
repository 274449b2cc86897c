"""Zero-shot 18-pathology scoring (BASELINE config 4; reference: scripts/zero_shot.py:387-611 `CTClipInferenceFast`,
CT_CLIP/ct_clip/ct_clip.py:792-855 `forward_infer`), sharded over the GPUs of one box.

What the reference does per volume: encoder once (zero_shot.py:550), then for each of 18 pathologies
`forward_infer` with the buffered text / image embeddings (zero_shot.py:563) - which re-projects all 13 824 tokens and
re-pools them 18 times (ct_clip.py:820-831) - and a softmax over the two prompts "... is present." /
"... is not present." keeping P(present) (zero_shot.py:564-568).

Here: the 36 prompt latents are computed once (`prepare`), a volume costs one encoder pass, one mean-pool, one
latent projection and ONE `ctk_pair_logits` launch over all 36 prompts (pooling first is exact because
`to_visual_latent` is bias-free); the 2-way softmax of 36 numbers stays in torch.  Volumes are independent, so ranks
take contiguous shares and the (V/W, 18) blocks meet in one all-gather ("replicas + trivial gather") -
there is no other collective.  No CPU path: the arithmetic is libctk's.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import ops

# zero_shot.py:482-488
PATHOLOGIES = ["Medical material", "Arterial wall calcification", "Cardiomegaly", "Pericardial effusion",
               "Coronary artery wall calcification", "Hiatal hernia", "Lymphadenopathy", "Emphysema", "Atelectasis",
               "Lung nodule", "Lung opacity", "Pulmonary fibrotic sequela", "Pleural effusion",
               "Mosaic attenuation pattern", "Peribronchial thickening", "Consolidation", "Bronchiectasis",
               "Interlobular septal thickening"]


def prompt_pairs(pathologies: Sequence[str] = PATHOLOGIES) -> List[Tuple[str, str]]:
    """zero_shot.py:490: the two prompts of every pathology, positive first."""
    return [(f"{p} is present.", f"{p} is not present.") for p in pathologies]


def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous share [lo, hi) of rank `rank`; the first n_items % world ranks take one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_rows(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All ranks' row blocks (sizes given by `shard_bounds`) concatenated in rank order -> [n_items, ...] everywhere.
    One padded all-gather; a single process returns `local` unchanged."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        assert local.shape[0] == n_items
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n_items, world, rank)
    assert local.shape[0] == hi - lo, f"rank {rank}: {local.shape[0]} rows, expected {hi - lo}"
    cap = -(-n_items // world)
    buf = local.new_zeros((cap,) + tuple(local.shape[1:]))
    buf[: hi - lo] = local
    out = local.new_empty((world * cap,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, buf, group=group)
    parts = []
    for r in range(world):
        a, b = shard_bounds(n_items, world, r)
        parts.append(out[r * cap: r * cap + (b - a)])
    return torch.cat(parts, dim=0)


def probs_from_logits(logits: torch.Tensor) -> torch.Tensor:
    """logits [..., 2 * n_path] ordered (present, absent) per pathology -> P(present) [..., n_path]
    (zero_shot.py:83-96,564-568: softmax over the pair, element 0)."""
    pairs = logits.reshape(*logits.shape[:-1], -1, 2)
    return torch.softmax(pairs.float(), dim=-1)[..., 0]


class ZeroShotScorer:
    """`clip`: a `vit_exp_b200.ct_clip.CTCLIP`.  `prepare` takes the tokenised prompts (one object with
    `.input_ids` / `.attention_mask` of shape [2, L] per pathology, as zero_shot.py:491 builds them) or precomputed
    text-encoder outputs, and caches the [2 * n_path, dl] latents."""

    def __init__(self, clip, group=None):
        self.clip = clip
        self.group = group
        self.prompt_latents: Optional[torch.Tensor] = None

    @torch.no_grad()
    def prepare(self, text_tokens: Optional[Sequence] = None, text_embeds: Optional[Sequence] = None) -> torch.Tensor:
        self.clip.eval()                                            # zero_shot.py:536
        lat = []
        n = len(text_tokens) if text_tokens is not None else len(text_embeds)
        for i in range(n):
            tl, _ = self.clip.latents(text=None if text_tokens is None else text_tokens[i],
                                      buffer_text_embed=None if text_embeds is None else text_embeds[i])
            assert tl.shape[0] == 2, "two prompts per pathology"
            lat.append(tl)
        self.prompt_latents = torch.cat(lat, dim=0).contiguous()
        return self.prompt_latents

    @torch.no_grad()
    def score(self, volume: torch.Tensor) -> torch.Tensor:
        """volume (1, 1, D, H, W) -> P(present) [n_path]"""
        return self.score_many(volume)[0]

    @torch.no_grad()
    def score_many(self, volumes: torch.Tensor) -> torch.Tensor:
        """volumes (B, 1, D, H, W) -> P(present) [B, n_path].  Volumes are independent, so one encoder pass over B of
        them (the encoder runs ~1.7x faster per volume at B = 8 than at B = 1, profiles/r1_side_configs.jsonl) gives
        exactly the per-volume results of the reference's batch-1 loop (zero_shot.py:543-550)."""
        assert self.prompt_latents is not None, "call prepare() first"
        _, il = self.clip.latents(image=volumes)
        lt = self.clip.temperature.detach().reshape(1).float()
        rows = [probs_from_logits(ops.pair_logits(self.prompt_latents, il[b].contiguous(), lt)) for b in range(il.shape[0])]
        return torch.stack(rows)

    @torch.no_grad()
    def run(self, n_volumes: int, load: Callable[[int], torch.Tensor], batch_size: int = 1, prefetch: bool = True) -> torch.Tensor:
        """Scores volumes [0, n_volumes): this rank calls `load(i)` (-> (1, 1, D, H, W)) for its contiguous share only,
        `batch_size` volumes per encoder pass; every rank returns the full [n_volumes, n_path] matrix in volume order.
        `prefetch`: `load` of the next batch (host->device copy, `data.npz_to_tensor`) is issued on a side stream while
        the current batch is scored, so the transfer hides behind the encoder (same results)."""
        on = dist.is_available() and dist.is_initialized()
        world = dist.get_world_size(self.group) if on else 1
        rank = dist.get_rank(self.group) if on else 0
        lo, hi = shard_bounds(n_volumes, world, rank)
        n_path = self.prompt_latents.shape[0] // 2
        dev = self.prompt_latents.device
        local = torch.empty(hi - lo, n_path, dtype=torch.float32, device=dev)
        bs = max(1, batch_size)
        side = torch.cuda.Stream(device=dev) if (prefetch and dev.type == "cuda") else None

        def fetch(i0):
            i1 = min(hi, i0 + bs)
            if side is None:
                vols = [load(i) for i in range(i0, i1)]
                return (vols[0] if len(vols) == 1 else torch.cat(vols, dim=0)), None
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                vols = [load(i) for i in range(i0, i1)]
                vols = vols[0] if len(vols) == 1 else torch.cat(vols, dim=0)
                ev = torch.cuda.Event()
                ev.record()
            return vols, ev

        nxt = fetch(lo) if hi > lo else None
        for i0 in range(lo, hi, bs):
            i1 = min(hi, i0 + bs)
            vols, ev = nxt
            nxt = fetch(i1) if i1 < hi else None
            if ev is not None:
                torch.cuda.current_stream(dev).wait_event(ev)
                vols.record_stream(torch.cuda.current_stream(dev))
            local[i0 - lo: i1 - lo] = self.score_many(vols)
        return gather_rows(local, n_volumes, self.group)
