"""`CTCLIP.forward_old` (reference ct_clip.py:1392-1778) on CPU: the real module code of
vit_exp_b200/ct_clip.py with the libctk entry points replaced by the torch doubles of tests/emulated_ops.py, compared with
what the REAL reference `forward_old` produced (tests/golden/forward_old_golden.pt, oracle/make_golden_legacy.py) and with
autograd through the pinned oracle.  The kernels themselves are checked in tests/test_forward_old_gpu.py.
"""
import os
from types import SimpleNamespace

import pytest
import torch

import emulated_ops as E
from oracle import ctclip_oracle as O
from vit_exp_b200 import ct_clip as CC
from vit_exp_b200 import transformer_maskgit as TM

GOLD = os.path.join(os.path.dirname(__file__), "golden", "forward_old_golden.pt")


class _Ops:
    def __getattr__(self, name):
        full = {"gemm": E.gemm_full, "cast_bf16": E.cast_bf16_full, "transpose_cast_bf16": E.transpose_cast_bf16_full,
                "layernorm_fwd": E.layernorm_fwd_full, "layernorm_bwd": E.layernorm_bwd_full}
        return full[name] if name in full else getattr(E, name)


@pytest.fixture
def doubles(monkeypatch):
    o = _Ops()
    monkeypatch.setattr(TM, "ops", o)
    monkeypatch.setattr(CC, "ops", o)
    E.OPERAND = torch.float32
    for fn_name in ("empty", "zeros"):
        orig = getattr(torch, fn_name)

        def wrapped(*a, _orig=orig, **k):          # the modules allocate bf16 operand buffers: fp32 in the exact-math mode
            if k.get("dtype") is torch.bfloat16:
                k["dtype"] = torch.float32
            return _orig(*a, **k)
        monkeypatch.setattr(torch, fn_name, wrapped)
    yield o
    E.OPERAND = torch.bfloat16


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


class _CpuViT(TM.CTViT):
    """the public entry asserts a CUDA input (there is no CPU path); the doubles stand in for the kernels here"""
    def encode_with_aux(self, video, _launched=None):
        return TM._CTViTEncode.apply(self, video.contiguous().float(), self.training, None, *self._flat_params())


def _clip_from(g):
    from transformers import BertConfig, BertModel
    sd = g["state_dict"]
    vit = _CpuViT(dim=64, codebook_size=64, image_size=20, patch_size=10, temporal_patch_size=5, spatial_depth=2,
                  temporal_depth=1, dim_head=32, heads=2)
    vit.cuda_graphs = False
    vit.load_state_dict({k[len("visual_transformer."):]: v for k, v in sd.items() if k.startswith("visual_transformer.")},
                        strict=False)
    bert = BertModel(BertConfig(vocab_size=100, hidden_size=48, num_hidden_layers=1, num_attention_heads=2,
                                intermediate_size=64, max_position_embeddings=32,
                                hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0))
    bert.load_state_dict(g["bert_state_dict"])
    clip = CC.CTCLIP(image_encoder=vit, text_encoder=bert, dim_text=48, dim_image=2 * 2 * 64, dim_latent=32,
                     config={"overlap_text_encoder": False})
    clip.load_state_dict({k: v for k, v in sd.items() if not k.startswith("visual_transformer.")}, strict=False)
    text = SimpleNamespace(input_ids=g["ids"], attention_mask=torch.ones_like(g["ids"]))
    return clip, vit, bert, text


def test_forward_old_inference_modes_match_reference(doubles, gold):
    g = gold
    clip, vit, bert, text = _clip_from(g)
    clip.eval()
    ones = torch.ones(g["ids"].shape[0], 1)
    tl, il, enc = clip.forward_old(text, g["video"], None, return_latents=True, text_valid_mask=ones)
    assert enc.shape == g["enc_image"].shape and _rel(enc, g["enc_image"]) < 1e-4
    assert _rel(tl, g["text_latents"]) < 1e-5 and _rel(il, g["image_latents"]) < 1e-4
    tl, il, _ = clip.forward_old(text, g["video"], None, return_latents=True, text_valid_mask=g["valid_some"])
    assert tl.shape == g["text_latents_some"].shape
    assert _rel(tl, g["text_latents_some"]) < 1e-5 and _rel(il, g["image_latents_some"]) < 1e-4
    sim = clip.forward_old(text, g["video"], None, text_valid_mask=g["valid_some"])
    assert sim.shape == g["similarity_some"].shape and (sim - g["similarity_some"]).abs().max() < 1e-4
    enc_text, embeds = clip.forward_old(text, g["video"], None, return_encodings=True, text_valid_mask=ones)
    assert _rel(enc_text, g["enc_text"]) < 1e-5
    assert _rel(embeds, O.image_embeds_legacy(g["enc_image"])) < 1e-4
    # `latents` / `forward_infer` follow the width of to_visual_latent: a checkpoint of this generation scores too
    tl2, il2 = clip.latents(text, g["video"])
    assert _rel(tl2, g["text_latents"]) < 1e-5 and _rel(il2, g["image_latents"]) < 1e-4


def test_forward_old_loss_and_gradients_match_reference(doubles, gold):
    g = gold
    clip, vit, bert, text = _clip_from(g)
    clip.train()
    loss, ld = clip.forward_old(text, g["video"], None, return_loss=True, return_loss_dict=True,
                                text_valid_mask=g["valid_some"], accelerator=CC.TorchDistAccelerator())
    assert abs(loss.item() - g["loss"].item()) < 1e-5
    assert ld["cl_loss"] == ld["loss_total"] and abs(ld["cl_loss"] - g["cl_loss"]) < 1e-5
    loss.backward()
    assert _rel(clip.to_text_latent.weight.grad, g["grad_to_text_latent"]) < 1e-4
    assert _rel(clip.to_visual_latent.weight.grad, g["grad_to_visual_latent"]) < 1e-4
    assert abs(float(clip.temperature.grad) - float(g["grad_temperature"])) < 1e-5 * max(1.0, abs(float(g["grad_temperature"])))
    assert _rel(vit.to_patch_emb[2].weight.grad, g["grad_patch_weight"]) < 2e-3
    assert _rel(bert.embeddings.word_embeddings.weight.grad, g["grad_word_embeddings"]) < 1e-4
    # without return_loss_dict: the bare loss (ct_clip.py:1775-1778)
    clip.zero_grad()
    vit.eval()                                        # frozen codebook for the repeat
    again = clip.forward_old(text, g["video"], None, return_loss=True, text_valid_mask=g["valid_some"],
                             accelerator=CC.TorchDistAccelerator())
    assert torch.is_tensor(again) and again.dim() == 0


def test_forward_old_head_matches_oracle_all_rows_valid(doubles):
    """`_ClipHead` with the frame pooling, no row selection, t != h != w: loss and every gradient against autograd
    through the oracle's restatement."""
    g = torch.Generator().manual_seed(31)
    B, t, h, w, dim, dt, dl = 4, 5, 2, 3, 16, 24, 8
    tokens = torch.randn(B, t, h, w, dim, generator=g).requires_grad_()
    cls = torch.randn(B, dt, generator=g).requires_grad_()
    wt = (torch.randn(dl, dt, generator=g) * 0.2).requires_grad_()
    wv = (torch.randn(dl, h * w * dim, generator=g) * 0.1).requires_grad_()
    temp = torch.tensor(0.7, requires_grad=True)
    loss, tl, il = CC._ClipHead.apply(cls, tokens, wt, wv, temp, CC.TorchDistAccelerator(), True, None)
    loss.backward()
    got = [v.grad.clone() for v in (cls, tokens, wt, wv, temp)]
    for v in (cls, tokens, wt, wv, temp):
        v.grad = None
    p = {"to_text_latent.weight": wt, "to_visual_latent.weight": wv, "temperature": temp}
    ref, tl_ref, il_ref = O.forward_old_loss(cls[:, None, :], tokens, p, torch.ones(B, 1))
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    assert _rel(tl, tl_ref) < 1e-5 and _rel(il, il_ref) < 1e-5
    for a, v in zip(got, (cls, tokens, wt, wv, temp)):
        assert _rel(a, v.grad) < 1e-4


def test_forward_old_needs_two_valid_reports_and_the_mask(doubles, gold):
    clip, vit, bert, text = _clip_from(gold)
    one = torch.zeros(gold["ids"].shape[0], 1)
    one[2] = 1
    with pytest.raises(ValueError, match="more than one valid report"):
        clip.forward_old(text, gold["video"], None, return_loss=True, text_valid_mask=one,
                         accelerator=CC.TorchDistAccelerator())
    with pytest.raises(AssertionError, match="text_valid_mask"):
        clip.forward_old(text, gold["video"], None, return_loss=True, accelerator=CC.TorchDistAccelerator())
    with pytest.raises(AssertionError, match="segmentation"):
        clip.forward_old(text, gold["video"], None, use_seg=True, text_valid_mask=one)


def test_forward_old_with_bf16_operands(monkeypatch, gold):
    """The body of the GPU module test (tests/forward_old_checks.py) on the torch doubles with bf16 operands - the
    arithmetic the kernels perform: the checks must hold whether or not a near-tie code flips."""
    import forward_old_checks
    o = _Ops()
    monkeypatch.setattr(TM, "ops", o)
    monkeypatch.setattr(CC, "ops", o)
    assert E.OPERAND is torch.bfloat16
    clip, vit, bert, text = _clip_from(gold)
    agree = forward_old_checks.run(clip, vit, bert, text, gold["video"], gold, CC.TorchDistAccelerator(), None)
    assert agree >= 0.9


# ------------------------------------------------------------------------------------------
# two ranks (gloo): each rank masks out a different report; the gathered loss and the local gradients against the
# single-process oracle on the concatenated valid rows
# ------------------------------------------------------------------------------------------
def _two_rank_inputs():
    g = torch.Generator().manual_seed(77)
    world, B, t, h, w, dim, dt, dl = 2, 3, 4, 2, 3, 8, 12, 6
    return dict(world=world, B=B,
                tokens=torch.randn(world, B, t, h, w, dim, generator=g), cls=torch.randn(world, B, dt, generator=g),
                wt=torch.randn(dl, dt, generator=g) * 0.3, wv=torch.randn(dl, h * w * dim, generator=g) * 0.15,
                temp=torch.tensor(0.4), valid=[torch.tensor([0, 2]), torch.tensor([1, 2])])


def _two_rank_worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import emulated_ops
    from vit_exp_b200 import ct_clip
    ct_clip.ops = emulated_ops
    d = _two_rank_inputs()
    leaves = [v.clone().requires_grad_() for v in (d["cls"][rank], d["tokens"][rank], d["wt"], d["wv"], d["temp"])]
    loss, tl, il = ct_clip._ClipHead.apply(*leaves, ct_clip.TorchDistAccelerator(), True, d["valid"][rank])
    loss.backward()
    torch.save(dict(loss=loss.detach(), tl=tl, il=il, grads=[v.grad for v in leaves]), f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_forward_old_head_two_ranks(tmp_path):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "old")
    mp.spawn(_two_rank_worker, args=(2, port, out), nprocs=2, join=True)
    res = [torch.load(f"{out}.{r}", weights_only=False) for r in range(2)]
    d = _two_rank_inputs()
    # single process: every rank's inputs (and its own replica of the shared weights) as leaves of ONE global loss
    leaves = [[v.clone().requires_grad_() for v in (d["cls"][r], d["tokens"][r], d["wt"], d["wv"], d["temp"])] for r in range(2)]
    T, I = [], []
    for r in range(2):
        cls, tok, wt, wv, _ = leaves[r]
        mask = torch.zeros(d["B"], 1)
        mask[d["valid"][r]] = 1
        tl, il = O.forward_old_latents(cls[:, None, :], tok, {"to_text_latent.weight": wt, "to_visual_latent.weight": wv}, mask)
        assert _rel(res[r]["tl"], tl) < 1e-5 and _rel(res[r]["il"], il) < 1e-5
        T.append(tl)
        I.append(il)
    n_valid = len(d["valid"][0])
    for r in range(2):                       # every rank differentiates the replicated global loss through ITS temperature
        for v in leaves[0] + leaves[1]:
            v.grad = None
        loss = O.clip_loss_reference_form(torch.cat(T), torch.cat(I), leaves[r][4], n_valid)     # / bs_single_gpu (:1661,1747)
        loss.backward(retain_graph=True)
        assert abs(res[r]["loss"].item() - loss.item()) < 1e-6
        for name, got, ref in zip(("cls", "tokens", "w_text", "w_vis", "temperature"), res[r]["grads"], leaves[r]):
            assert _rel(got, ref.grad) < 1e-4, (r, name)
        masked = [i for i in range(d["B"]) if i not in d["valid"][r].tolist()]
        assert res[r]["grads"][1][masked].abs().max() == 0
