"""Drop-in `CTCLIP` (reference: CT_CLIP/ct_clip/ct_clip.py, CT_CLIP/ct_clip/distributed.py).

Keeps the constructor keywords the launchers use, the attribute names, `forward(batch, device=,
accelerator=, **kw) -> (loss, {'cl_loss': float})`, `forward_infer`, `load`, and the state-dict
keys.  The contrastive head (ct_clip.py:1280-1388) - token mean-pool, latent projections,
l2-normalise, all-gather across ranks, scaled similarity, symmetric log-softmax cross entropy and
its backward - runs in libctk.so.  The text encoder is whatever module the caller passes
(HF BertModel in the reference): a BertModel's arithmetic runs through libctk (vit_exp_b200/text_tower.py, the module
stays the parameter holder); any other module - or `config["ctk_text_tower"] = False` / CTK_TEXT_TOWER=0 - runs as passed.
"""
from __future__ import annotations

import copy
import os
from pathlib import Path

import torch
import torch.distributed as dist
from torch import nn

from . import ops, text_tower


class TorchDistAccelerator:
    """Minimal object with the accelerator protocol the reference uses (distributed.py:11-14):
    gather / num_processes / process_index, backed by torch.distributed (NCCL over NVLink)."""

    def __init__(self, group=None):
        self.group = group
        self.on = dist.is_available() and dist.is_initialized()
        self.num_processes = dist.get_world_size(group) if self.on else 1
        self.process_index = dist.get_rank(group) if self.on else 0

    def gather(self, x: torch.Tensor) -> torch.Tensor:
        if self.num_processes == 1:
            return x
        out = torch.empty((x.shape[0] * self.num_processes,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous(), group=self.group)
        return out


class AllGather(torch.autograd.Function):
    """distributed.py:9-20: forward = accelerator.gather (rank-major concat); backward = the local
    chunk of the incoming gradient, NO reduction over ranks."""

    @staticmethod
    def forward(ctx, x, accelerator):
        ctx.num_processes = accelerator.num_processes
        ctx.process_index = accelerator.process_index
        return accelerator.gather(x)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.chunk(ctx.num_processes, dim=0)[ctx.process_index], None


class _ClipHead(torch.autograd.Function):
    """Fused contrastive head. Inputs: CLS rows of the text encoder output, encoded image tokens
    (B, t, h, w, d), the two projection weights and the log-temperature.

    `frames_pool` selects the pooling of `CTCLIP.forward_old` (ct_clip.py:1549,1566: mean over the t axis only, the
    (h w d) rest flattened into the projection's input) instead of the mean over all tokens (ct_clip.py:1297);
    `valid` (int64 indices, or None) keeps those rows of both towers before the projections (ct_clip.py:1593-1594)."""

    @staticmethod
    def forward(ctx, text_cls, tokens, w_text, w_vis, temperature, accelerator, frames_pool=False, valid=None):
        B = text_cls.shape[0]
        n_pool = tokens.shape[1] if frames_pool else tokens.numel() // (B * tokens.shape[-1])
        width = tokens.numel() // (B * n_pool)
        text_cls = text_cls.float()
        if text_cls.stride(-1) != 1:
            text_cls = text_cls.contiguous()
        pooled = ops.mean_pool(tokens.reshape(B, n_pool, width).contiguous())  # ct_clip.py:1297 (pool first) / :1549
        if valid is not None:
            text_cls, pooled = text_cls.index_select(0, valid), pooled.index_select(0, valid)
        il, rn_i = ops.latent_fwd(pooled, w_vis)                               # ct_clip.py:1290,1316
        tl, rn_t = ops.latent_fwd(text_cls, w_text)                            # ct_clip.py:1313-1316
        b_local = tl.shape[0]
        world, rank = accelerator.num_processes, accelerator.process_index
        if world > 1:
            packed = torch.cat([tl, il], dim=1)                                 # one gather instead of two
            g = accelerator.gather(packed)
            dl = tl.shape[1]
            T, I = g[:, :dl].contiguous(), g[:, dl:].contiguous()
        else:
            T, I = tl, il
        lt = temperature.detach().reshape(1).float()
        out, d_local = ops.clip_loss_fwd_bwd(T, I, lt, b_local=b_local, row0=rank * b_local)
        ctx.save_for_backward(text_cls, pooled, w_text, w_vis, tl, il, rn_t, rn_i, out, d_local, valid)
        ctx.tok_shape = tuple(tokens.shape)
        ctx.n_pool = n_pool
        ctx.frames_pool = frames_pool
        ctx.batch = B
        ctx.mark_non_differentiable(tl, il)
        return out[0].clone(), tl, il

    @staticmethod
    def backward(ctx, gloss, _gtl, _gil):
        text_cls, pooled, w_text, w_vis, tl, il, rn_t, rn_i, out, d_local, valid = ctx.saved_tensors
        d_local = d_local * gloss
        dwt, dcls = ops.latent_bwd(d_local[0].contiguous(), tl, rn_t, text_cls, w_text)
        dwv, dpool = ops.latent_bwd(d_local[1].contiguous(), il, rn_i, pooled, w_vis)
        if valid is not None:                     # rows that were not selected receive no gradient
            dcls = dcls.new_zeros(ctx.batch, dcls.shape[1]).index_copy_(0, valid, dcls)
            dpool = dpool.new_zeros(ctx.batch, dpool.shape[1]).index_copy_(0, valid, dpool)
        B = ctx.batch
        # mean-pool backward: every pooled-over position receives dpooled / n; left as a stride-0 expand so the
        # encoder backward can consume it without materialising B*n*dim floats
        if ctx.frames_pool:
            dtok = (dpool / ctx.n_pool).view(B, 1, *ctx.tok_shape[2:]).expand(ctx.tok_shape)
        else:
            dtok = (dpool / ctx.n_pool).view(B, *([1] * (len(ctx.tok_shape) - 2)), dpool.shape[1]).expand(ctx.tok_shape)
        dtemp = (out[1] * gloss).reshape(())
        return dcls, dtok, dwt, dwv, dtemp, None, None, None


class DeferredFloat:
    """A scalar that is on its way from the device: float(x), format(x), comparisons and arithmetic synchronise on
    first use.  Returned in the loss dict instead of `loss.item()` when `config["defer_loss_read"]` is set."""

    __slots__ = ("_host", "_event", "_value")

    def __init__(self, t: torch.Tensor):
        self._host = torch.empty((), dtype=torch.float32, pin_memory=True)
        self._host.copy_(t.detach().float(), non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record()
        self._value = None

    def __float__(self):
        if self._value is None:
            self._event.synchronize()
            self._value = float(self._host)
            self._host = None
        return self._value

    def item(self):
        return float(self)

    def __format__(self, spec):
        return format(float(self), spec)

    def __repr__(self):
        return repr(float(self))

    __str__ = __repr__

    def __add__(self, o): return float(self) + float(o)
    __radd__ = __add__
    def __sub__(self, o): return float(self) - float(o)
    def __rsub__(self, o): return float(o) - float(self)
    def __mul__(self, o): return float(self) * float(o)
    __rmul__ = __mul__
    def __truediv__(self, o): return float(self) / float(o)
    def __rtruediv__(self, o): return float(o) / float(self)
    def __neg__(self): return -float(self)
    def __abs__(self): return abs(float(self))
    def __lt__(self, o): return float(self) < float(o)
    def __le__(self, o): return float(self) <= float(o)
    def __gt__(self, o): return float(self) > float(o)
    def __ge__(self, o): return float(self) >= float(o)
    def __eq__(self, o): return float(self) == float(o)
    def __hash__(self): return hash(float(self))


class CTCLIP(nn.Module):
    """Reference constructor: ct_clip.py:467-660 (only the arguments reachable from the launchers
    are honoured; the x-clip default encoders, MLM / visual-SSL branches and segmentation heads are
    outside the contrastive hot path)."""

    def __init__(self, *, image_encoder=None, text_encoder=None, dim_text=512, dim_image=512, dim_latent=512,
                 use_all_token_embeds=False, downsample_image_embeds=False, decoupled_contrastive_learning=False,
                 extra_latent_projection=False, use_mlm=False, use_visual_ssl=False, visual_ssl=None,
                 text_encode_without_mask=False, config=None, tokenizer=None, **kwargs):
        super().__init__()
        assert image_encoder is not None and text_encoder is not None, \
            "pass image_encoder= and text_encoder= (every reference launcher does: run_train.py:143-154)"
        assert not use_all_token_embeds, "no more support for use_all_token_embeds"       # ct_clip.py:514
        assert not extra_latent_projection, "no more support for extra_latent_projection"  # ct_clip.py:515
        assert not (use_mlm or use_visual_ssl or visual_ssl is not None or downsample_image_embeds), \
            "MLM / visual-SSL / downsampled-latent branches are outside the contrastive hot path"
        assert not decoupled_contrastive_learning, \
            "the DCL branch builds a local-batch eye mask and is off in every launcher (ct_clip.py:1366-1368)"
        config = {} if config is None else config
        self.dtype = torch.float32
        self.config = config
        # Registration order (image tower first) only affects the ORDER of parameters()/state_dict(), not the keys.
        # DDP fills its gradient buckets in reverse registration order and all-reduces them strictly in that
        # order: with the text tower registered last, its buckets (80 % of the bytes) go first and their
        # all-reduces overlap the image encoder's backward; the encoder's gradients, which all become ready
        # at the very end of its single backward node, form the last buckets instead of blocking the queue.
        if os.environ.get("CTK_REFERENCE_PARAM_ORDER", "0") == "1":        # ct_clip.py:586-590 order
            self.text_transformer = text_encoder
            self.visual_transformer = image_encoder
        else:
            self.visual_transformer = image_encoder
            self.text_transformer = text_encoder
        self.text_encode_without_mask = text_encode_without_mask
        self.to_text_latent = nn.Linear(dim_text, dim_latent, bias=False)
        self.to_visual_latent = nn.Linear(dim_image, dim_latent, bias=False)
        self.temperature = nn.Parameter(torch.tensor(1.0))
        self.to_text_latent_extra = copy.deepcopy(self.to_text_latent)
        self.to_visual_latent_extra = copy.deepcopy(self.to_visual_latent)
        self.tokenizer = tokenizer       # the reference downloads CXR-BERT's tokenizer here (ct_clip.py:650)
        self.fix_text_encoder = config.get("fix_text_encoder", False)
        self.overlap_text_encoder = config.get("overlap_text_encoder", True)
        # a HF BertModel text encoder (what every reference launcher passes, run_train.py:143-154) runs through libctk
        # (vit_exp_b200/text_tower.py); config["ctk_text_tower"] = False / CTK_TEXT_TOWER=0 runs the module as passed
        self.ctk_text_tower = bool(config.get("ctk_text_tower", os.environ.get("CTK_TEXT_TOWER", "1") == "1"))
        self._text_tower_note = None
        self._side_stream = None
        self._enc_stream = None
        self._txt_stream = None
        if self.fix_text_encoder:
            for p in self.text_transformer.parameters():
                p.requires_grad = False
        for k in ("use_seg", "use_open_seg"):
            assert not config.get(k, False), f"{k}: segmentation heads are outside the contrastive hot path"

    # ------------------------------------------------------------------------------------------
    def load(self, path, check=True):
        path = Path(path)
        assert path.exists()
        pt = torch.load(str(path), map_location="cpu")
        self.load_state_dict({k[7:]: v for k, v in pt.items()})      # strip 'module.' (ct_clip.py:771)
        print(f"successfully loaded model from {path}")

    def _encode_text(self, text):
        if self.fix_text_encoder:
            self.text_transformer.eval()
        if self.ctk_text_tower:
            bert = self.text_transformer
            why = text_tower.unsupported_reason(bert, bert.training)
            if why is None:
                return text_tower.encode(bert, text.input_ids, text.attention_mask)
            if self._text_tower_note != why and why != "not a BertModel":
                # a BertModel this path does not reproduce: say so once; the stock module computes the same function
                self._text_tower_note = why
                print(f"[vit_exp_b200] libctk text tower not applicable ({why}); running the text encoder as passed",
                      flush=True)
        return self.text_transformer(text.input_ids, attention_mask=text.attention_mask)[0]

    def unused_parameter_names(self):
        """Names (relative to this module) of the parameters that never receive a gradient on the contrastive path -
        the reason the reference wraps the model with `find_unused_parameters=True` (CTCLIPTrainer.py:318; SURVEY
        appendix C): the first-frame / pixel heads and cross-attention norms of the image encoder, the text encoder's
        pooler, the `*_latent_extra` projections.  A caller may hand them to
        `DistributedDataParallel._set_params_and_buffers_to_ignore_for_model` and run DDP with
        `find_unused_parameters=False`, which removes DDP's per-step used-parameter bitmap all-reduce and host read."""
        vt = self.visual_transformer
        names = []
        if hasattr(vt, "_flat_params"):
            used = {id(p) for p in vt._flat_params()}
            names += [f"visual_transformer.{n}" for n, p in vt.named_parameters() if id(p) not in used]
        names += [f"text_transformer.{n}" for n, _ in self.text_transformer.named_parameters() if n.startswith("pooler.")]
        names += ["to_text_latent_extra.weight", "to_visual_latent_extra.weight"]
        return names

    def _text_on_ctk(self) -> bool:
        bert = self.text_transformer
        return self.ctk_text_tower and text_tower.unsupported_reason(bert, getattr(bert, "training", False)) is None

    def _encode_both(self, text, image):
        """Both towers of one step, on two streams (results are identical to running them one after the other).

        libctk text tower (the default for a HF BertModel): both towers are CUDA-graph replays, each on its own side
        stream - the text tower's with high priority - and the tower's autograd node is created last:
          * forward and backward of the two towers overlap (tensor-bound GEMMs of one next to the HBM-bound LayerNorm /
            PEG kernels of the other);
          * under DDP the text tower's gradients - 80 % of the bytes to all-reduce - are complete a few ms into the
            backward pass; their AccumulateGrad copies run on the caller's stream, which is idle during the backward
            pass, so their buckets travel while the encoder's backward computes.  (Tower on a side stream and the encoder
            on the caller's stream, round 1's arrangement: those copies queued behind the encoder's 22 ms backward graph.
            Tower on the caller's stream without priority: its backward was starved by the encoder's persistent kernels
            and finished with it.  Either way every all-reduce was exposed at the end of the step.)
        CTK_TOWER_STREAMS=serial runs both on the caller's stream.

        Text encoder run as passed (stock PyTorch, hundreds of small launches): on a high-priority side stream next to
        the image encoder; autograd replays each backward on the stream of its forward."""
        if not (self.overlap_text_encoder and image.is_cuda):
            enc_image = self.visual_transformer(image, return_encoded_tokens=True)
            return self._encode_text(text), enc_image
        if self._text_on_ctk():
            if os.environ.get("CTK_TOWER_STREAMS", "encoder_side") == "serial":
                enc_image = self.visual_transformer(image, return_encoded_tokens=True)
                return self._encode_text(text), enc_image
            cur = torch.cuda.current_stream()
            if self._enc_stream is None:
                self._enc_stream = torch.cuda.Stream(device=image.device)
                # high priority: whenever an SM frees up the tower's kernels are scheduled ahead of the encoder's
                # persistent ones.  Without it the tower's 3.5 ms backward is stretched over the encoder's whole 24 ms
                # backward pass and its gradient buckets leave as late as the encoder's (profiles/r2_ddp_timeline_n2.txt)
                prio = -1 if os.environ.get("CTK_TEXT_STREAM_PRIORITY", "1") != "0" else 0
                self._txt_stream = torch.cuda.Stream(device=image.device, priority=prio)
            enc, txt = self._enc_stream, self._txt_stream
            enc.wait_stream(cur)
            txt.wait_stream(cur)
            with torch.cuda.stream(enc):
                enc_image = self.visual_transformer(image, return_encoded_tokens=True)
            with torch.cuda.stream(txt):
                enc_text = self._encode_text(text)          # autograd node created last: its backward is launched first
            cur.wait_stream(enc)
            cur.wait_stream(txt)
            enc_image.record_stream(cur)
            enc_text.record_stream(cur)
            return enc_text, enc_image
        cur = torch.cuda.current_stream()
        if self._side_stream is None:
            # high priority: the tower's many small kernels are scheduled ahead of the encoder's persistent,
            # SM-filling kernels at every kernel boundary instead of queueing behind them
            prio = -1 if os.environ.get("CTK_TEXT_STREAM_PRIORITY", "1") != "0" else 0
            self._side_stream = torch.cuda.Stream(device=image.device, priority=prio)
        side = self._side_stream
        # If the encoder forward can be replayed from its CUDA graph it is enqueued FIRST (one launch), so the GPU
        # works on it while the host spends ~9 ms enqueueing the text tower; its autograd node is still created
        # AFTER the tower's nodes (vt(..., _launched=handle) below), so the encoder's backward is launched first too.
        ready = torch.cuda.Event()
        ready.record(cur)
        vt = self.visual_transformer
        handle = vt.encode_begin(image) if hasattr(vt, "encode_begin") else None
        side.wait_event(ready)
        with torch.cuda.stream(side):
            enc_text = self._encode_text(text)
        if handle is not None:
            enc_image = vt(image, return_encoded_tokens=True, _launched=handle)
        else:
            enc_image = vt(image, return_encoded_tokens=True)
        cur.wait_stream(side)
        enc_text.record_stream(cur)
        return enc_text, enc_image

    def forward(self, batch, device=None, accelerator=None, **kwargs):
        if batch["data_type"][0] == "imagereport":
            return self.forward_batch_image_report(batch, device=device, accelerator=accelerator, **kwargs)
        raise ValueError(f"Data type {batch['data_type']} is outside the contrastive hot path "
                         "(imageseg / imageopenseg need the CTViT3D segmentation heads)")

    def forward_batch_image_report(self, batch, device=None, accelerator=None, **kwargs):
        """ct_clip.py:1252-1388."""
        assert accelerator is not None, "accelerator is not provided"
        text, image = batch["text"], batch["image"]
        enc_text, enc_image = self._encode_both(text, image)                        # (B, L, dt), (B, t, h, w, C)
        loss, _, _ = _ClipHead.apply(enc_text[:, 0, :], enc_image, self.to_text_latent.weight,
                                     self.to_visual_latent.weight, self.temperature, accelerator)
        if self.config.get("defer_loss_read", False):
            # opt-in: same value, read back lazily (async copy to pinned memory now, synchronise on first use), so the
            # host can enqueue the backward pass while the forward still runs
            return loss, {"cl_loss": DeferredFloat(loss)}
        return loss, {"cl_loss": loss.item()}      # the reference also syncs here (ct_clip.py:1384)

    def _frames_pool(self, enc_image) -> bool:
        """True when `to_visual_latent` was built for the pooling of the original CT-CLIP checkpoints
        (`dim_image = h*w*C = 294912`, scripts/run_zero_shot_latent.py:26-31; ct_clip.py:1549,1566) rather than for
        the token mean (`dim_image = C`, ct_clip.py:1286-1297)."""
        din = self.to_visual_latent.in_features
        if din == enc_image.shape[-1]:
            return False
        per_frame = enc_image[0, 0].numel()
        assert din == per_frame, (f"to_visual_latent expects {din} inputs; the encoder yields tokens of width "
                                  f"{enc_image.shape[-1]} (token mean) or {per_frame} per frame (forward_old pooling)")
        return True

    @torch.no_grad()
    def latents(self, text=None, image=None, buffer_text_embed=None, buffer_image_embed=None):
        """l2-normalised (text_latents, image_latents); either side may be None.  The image pooling follows the
        width of `to_visual_latent` (see `_frames_pool`), so checkpoints of either generation load and score."""
        tl = il = None
        if text is not None or buffer_text_embed is not None:
            emb = buffer_text_embed if buffer_text_embed is not None else \
                self.text_transformer(text.input_ids, attention_mask=text.attention_mask)
            enc_text = emb[0]
            tl, _ = ops.latent_fwd(enc_text[:, 0, :].float(), self.to_text_latent.weight)
        if image is not None or buffer_image_embed is not None:
            enc = buffer_image_embed if buffer_image_embed is not None else \
                self.visual_transformer(image, return_encoded_tokens=True)
            il, _ = ops.latent_fwd(self._pool_tokens(enc), self.to_visual_latent.weight)
        return tl, il

    def _pool_tokens(self, enc):
        B, dim = enc.shape[0], enc.shape[-1]
        if self._frames_pool(enc):
            return ops.mean_pool(enc.reshape(B, enc.shape[1], -1).contiguous().float())      # ct_clip.py:1549,1566
        return ops.mean_pool(enc.reshape(B, -1, dim).contiguous().float())                   # ct_clip.py:1286-1297

    def forward_old(self, text, image, device=None, return_loss=False, return_loss_dict=False, return_encodings=False,
                    return_latents=False, use_seg=False, seg_mask=None, seg_valid_mask=None, text_valid_mask=None,
                    seg_weight=1.0, accelerator=None, freeze_image_encoder=False, freeze_text_encoder=False,
                    text_to_image=True, aug_text=None, aug_image=None):
        """ct_clip.py:1392-1778, the forward of the original CT-CLIP checkpoints: image embedding = mean over the t
        axis of the encoded tokens, flattened (h w C) (`dim_image = 294912` for the full-size encoder, :1549,1566),
        rows of both towers selected by `text_valid_mask` (B, 1) before the projections (:1593-1594).

          return_encodings -> (enc_text, image embedding)                               (:1572-1573)
          return_latents   -> (text_latents, image_latents, encoded tokens)             (:1638-1642)
          neither, no loss -> exp(temperature) * <text_latent_b, image_latent_b>        (:1655-1657)
          return_loss      -> the symmetric contrastive loss over the valid rows, divided by their count
                              (:1661-1768; with return_loss_dict also {'cl_loss', 'loss_total'})

        Outside this path, as in `forward`: segmentation (`use_seg`), multiview augmentations, MLM / visual SSL.  The
        reference's loss branch reads `seg_loss` even when `use_seg` is off (:1766, a NameError there); here that term
        is 0.  With at most one valid report the reference returns the segmentation loss alone (:1600-1608): an error
        here.  The three inference modes run without autograd."""
        assert not use_seg and seg_mask is None, "segmentation heads are outside the contrastive hot path"
        assert aug_text is None and aug_image is None, "multiview augmentations are outside the contrastive hot path"
        assert text_valid_mask is not None, "text_valid_mask (B, 1) is required (ct_clip.py:1593)"
        keep = text_valid_mask.reshape(text_valid_mask.shape[0], -1)[:, 0].bool()
        keep_host = keep if keep.device.type == "cpu" else keep.cpu()       # the reference's boolean indexing syncs too
        valid = None if bool(keep_host.all()) else torch.nonzero(keep_host).squeeze(1).to(image.device)
        if not return_loss:
            with torch.no_grad():                           # towers as in `latents` / `forward_infer`
                enc_text = self.text_transformer(text.input_ids, attention_mask=text.attention_mask)[0]
                enc_image = self.visual_transformer(image, return_encoded_tokens=True)
                assert self._frames_pool(enc_image), "forward_old needs dim_image = h*w*C (ct_clip.py:1566,1614)"
                embeds = ops.mean_pool(enc_image.reshape(enc_image.shape[0], enc_image.shape[1], -1).contiguous().float())
                if return_encodings:
                    return enc_text, embeds
                cls = enc_text[:, 0, :].float()
                if valid is not None:
                    cls, embeds = cls.index_select(0, valid), embeds.index_select(0, valid)
                tl, _ = ops.latent_fwd(cls, self.to_text_latent.weight)
                il, _ = ops.latent_fwd(embeds, self.to_visual_latent.weight)
                if return_latents:
                    return tl, il, enc_image
                lt = self.temperature.detach().reshape(1).float()
                return torch.cat([ops.pair_logits(tl[i:i + 1], il[i], lt) for i in range(tl.shape[0])])
        assert accelerator is not None, "accelerator is not provided"
        n_valid = keep_host.numel() if valid is None else valid.numel()
        if n_valid <= 1:
            raise ValueError("forward_old: the contrastive loss needs more than one valid report; the reference falls "
                             "back to the segmentation loss here (ct_clip.py:1600-1608), which is outside this path")
        enc_text, enc_image = self._encode_both(text, image)
        assert self._frames_pool(enc_image), "forward_old needs dim_image = h*w*C (ct_clip.py:1566,1614)"
        loss, _, _ = _ClipHead.apply(enc_text[:, 0, :], enc_image, self.to_text_latent.weight,
                                     self.to_visual_latent.weight, self.temperature, accelerator, True, valid)
        if not return_loss_dict:
            return loss
        value = DeferredFloat(loss) if self.config.get("defer_loss_read", False) else loss.item()
        return loss, {"cl_loss": value, "loss_total": value}

    def forward_infer(self, text, image, buffer_text_embed=None, buffer_image_embed=None):
        """ct_clip.py:792-855: exp(temperature) * <text latent_p, image latent> for each prompt p
        (text batch broadcast against a single volume)."""
        tl, il = self.latents(text, image, buffer_text_embed, buffer_image_embed)
        assert il.shape[0] == 1 or il.shape[0] == tl.shape[0]
        lt = self.temperature.detach().reshape(1).float()
        if il.shape[0] == 1:
            return ops.pair_logits(tl, il[0].contiguous(), lt)
        return torch.stack([ops.pair_logits(tl[i:i + 1], il[i].contiguous(), lt)[0] for i in range(tl.shape[0])])
