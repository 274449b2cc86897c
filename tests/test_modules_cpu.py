"""Host-side (no GPU) checks of the drop-in modules: state-dict compatibility with the reference,
constructor surface, C ABI symbol export, no-CPU-path behaviour."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "reference_golden.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def _tiny_vit():
    from vit_exp_b200.transformer_maskgit import CTViT
    return CTViT(dim=64, codebook_size=64, image_size=20, patch_size=10, temporal_patch_size=5, spatial_depth=2,
                 temporal_depth=1, dim_head=32, heads=2)


def test_ctvit_state_dict_matches_reference(gold):
    vit = _tiny_vit()
    ref = gold["ctvit"]["state_dict"]
    mine = vit.state_dict()
    skip = ("to_pixels", "to_patch_emb_first_frame")          # dropped from the fixture to keep it small
    mine_keys = {k for k in mine if not k.startswith(skip)}
    assert mine_keys == set(ref.keys())
    for k, v in ref.items():
        assert tuple(mine[k].shape) == tuple(v.shape), k
    missing, unexpected = vit.load_state_dict(ref, strict=False)
    assert not unexpected and all(k.startswith(skip) for k in missing)
    # buffers vs parameters as in the reference (beta is a buffer, gamma a parameter)
    names = dict(vit.named_parameters())
    assert "enc_spatial_transformer.layers.0.1.norm.gamma" in names
    assert "enc_spatial_transformer.layers.0.1.norm.beta" not in names
    assert names["enc_spatial_transformer.layers.0.1.null_kv"].shape == (2, 0, 32)


def test_ctvit_surface():
    vit = _tiny_vit()
    assert vit.image_size == (20, 20) and vit.patch_size == (10, 10) and vit.temporal_patch_size == 5
    assert vit.patch_height_width == (2, 2) and vit.image_num_tokens == 4
    with pytest.raises(AssertionError):
        vit(torch.zeros(1, 1, 5, 30, 30), return_encoded_tokens=True)       # wrong image size (ctvit.py:375)
    with pytest.raises(AssertionError):
        vit(torch.zeros(1, 1, 5, 20, 20), return_encoded_tokens=True)       # CPU tensor: there is no CPU path


def test_ctclip_state_dict_and_surface(gold):
    from vit_exp_b200.ct_clip import CTCLIP
    vit = _tiny_vit()
    text = torch.nn.Linear(4, 4)
    clip = CTCLIP(image_encoder=vit, text_encoder=text, dim_text=48, dim_image=64, dim_latent=32, config={},
                  num_text_tokens=123, visual_patch_dropout=0.5)            # extra reference kwargs are swallowed
    ref = {k: v for k, v in gold["ctclip"]["state_dict"].items() if not k.startswith("visual_transformer.")}
    mine = {k: v for k, v in clip.state_dict().items()
            if not k.startswith(("visual_transformer.", "text_transformer."))}
    assert set(mine) == set(ref)
    assert clip.temperature.shape == () and float(clip.temperature) == 1.0
    with pytest.raises(ValueError):
        clip({"data_type": ["imageseg"]})
    with pytest.raises(AssertionError):
        clip({"data_type": ["imagereport"], "text": None, "image": None}, accelerator=None)


def test_allgather_backward_convention(gold):
    from vit_exp_b200.ct_clip import AllGather
    g = gold["allgather"]

    class TwoRank:
        num_processes, process_index = 2, 1
        def gather(self, x):
            return torch.cat([g["other"], x], dim=0)
    a = g["a"].clone().requires_grad_(True)
    out = AllGather.apply(a, TwoRank())
    assert torch.equal(out, g["gathered"])
    (out * g["w"]).sum().backward()
    assert torch.equal(a.grad, g["grad_a"])


def test_c_abi_exports_every_declared_symbol():
    """include/ctk.h <-> libctk.so <-> ctypes table stay in lock-step (no compute: no GPU here)."""
    from vit_exp_b200 import _lib
    from vit_exp_b200.build import build
    lib_path = build()
    header = open(os.path.join(ROOT, "include", "ctk.h")).read()
    declared = set(re.findall(r"\b(ctk_[a-z0-9_]+)\s*\(", header))
    declared -= {"ctk_status", "ctk_epilogue"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(str(lib_path))
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().ctk_version() == 100


def test_no_cpu_fallback_without_gpu():
    """On a box without an sm_100 GPU the library refuses instead of computing elsewhere."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vit_exp_b200 import _lib
    lib = _lib.load()
    assert lib.ctk_device_ok() == -3                 # CTK_ERR_ARCH
    assert b"no CPU path" in lib.ctk_last_error() or b"sm_100" in lib.ctk_last_error()
    rc = lib.ctk_fill_f32(None, 0.0, 16, None)
    assert rc == -3


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "vit_exp_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_unused_parameter_names_are_exactly_the_gradient_free_ones():
    """CTCLIP.unused_parameter_names() (what bench.py hands to DDP as ignored parameters) = SURVEY appendix C: every name
    exists, covers the first-frame / pixel heads, cross-attention norms, empty null_kv, BERT pooler and *_extra
    projections, and nothing the backward pass produces a gradient for."""
    from transformers import BertConfig, BertModel
    from vit_exp_b200.ct_clip import CTCLIP
    from vit_exp_b200.transformer_maskgit import CTViT
    vit = CTViT(dim=64, codebook_size=32, image_size=(8, 8), patch_size=(4, 4), temporal_patch_size=2, spatial_depth=1,
                temporal_depth=1, dim_head=32, heads=2)
    bert = BertModel(BertConfig(vocab_size=50, hidden_size=128, num_hidden_layers=1, num_attention_heads=2, intermediate_size=256))
    clip = CTCLIP(image_encoder=vit, text_encoder=bert, dim_text=128, dim_image=64, dim_latent=16, config={})
    names = clip.unused_parameter_names()
    all_names = dict(clip.named_parameters())
    assert len(set(names)) == len(names) and all(n in all_names for n in names)
    for frag in ("to_patch_emb_first_frame.", "to_pixels.", "to_pixels_first_frame.", "context_norm.gamma", "null_kv", "pooler.dense.",
                 "to_text_latent_extra.weight", "to_visual_latent_extra.weight"):
        assert any(frag in n for n in names), frag
    used = {id(p) for p in vit._flat_params()}
    for n in names:
        assert id(all_names[n]) not in used
    for n in all_names:       # everything else is on the gradient path: encoder flat params, BERT minus pooler, the head
        if n not in names:
            assert n.startswith(("visual_transformer.", "text_transformer.")) or n in ("to_text_latent.weight", "to_visual_latent.weight", "temperature")
