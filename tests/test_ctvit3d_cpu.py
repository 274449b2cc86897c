"""Host logic of the drop-in CTViT3D (vit_exp_b200/ctvit3d.py) on CPU: state-dict compatibility with the reference
module, the fixed position table, and `_forward` / `_backward` on torch doubles of the libctk kernels against autograd
through the oracle that tests/test_oracle_cpu.py pins to the real reference CTViT3D."""
import os

import pytest
import torch

import emulated_ops as E
from oracle import ctclip_oracle as O
from vit_exp_b200 import ctvit3d as C3

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ctvit3d_golden.pt")


class _Ops:
    def __getattr__(self, name):
        full = {"gemm": E.gemm_full, "cast_bf16": E.cast_bf16_full, "transpose_cast_bf16": E.transpose_cast_bf16_full,
                "layernorm_fwd": E.layernorm_fwd_full, "layernorm_bwd": E.layernorm_bwd_full}
        return full[name] if name in full else getattr(E, name)


@pytest.fixture
def doubles(monkeypatch):
    monkeypatch.setattr(C3, "ops", _Ops())
    monkeypatch.setattr(C3, "OPERAND_DTYPE", torch.float32)
    E.OPERAND = torch.float32
    yield
    E.OPERAND = torch.bfloat16


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_state_dict_and_surface_match_reference():
    gold = torch.load(GOLD, weights_only=False)
    for case in gold["cases"]:
        c = case["cfg"]
        m = C3.CTViT3D(dim=c["dim"], image_size=c["image_size"], patch_size=c["patch_size"],
                       temporal_size=c["temporal_size"], temporal_patch_size=c["temporal_patch_size"],
                       transformer_blocks=c["transformer_blocks"], dim_head=32, heads=c["heads"])
        ref = case["state_dict"]
        mine = m.state_dict()
        assert {k for k in mine if not k.startswith("to_pixels")} == set(ref.keys())
        for k, v in ref.items():
            assert tuple(mine[k].shape) == tuple(v.shape), k
        assert (mine["pos_embed"] - ref["pos_embed"]).abs().max() < 1e-6          # built here, not loaded
        assert not m.pos_embed.requires_grad
        missing, unexpected = m.load_state_dict(ref, strict=False)
        assert not unexpected and all(k.startswith("to_pixels") for k in missing)
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 1, c["temporal_size"], c["image_size"], c["image_size"]))
    with pytest.raises(AssertionError):                                          # no CPU path
        m(torch.zeros(1, 1, c["temporal_size"], c["image_size"], c["image_size"]), return_encoded_tokens=True)


def test_forward_backward_match_oracle_on_reference_weights(doubles):
    """weights, input and output gradient of the golden case: tokens and EVERY parameter gradient against the oracle
    (itself pinned to the reference), plus the reference's own recorded output."""
    gold = torch.load(GOLD, weights_only=False)
    for case in gold["cases"]:
        c = case["cfg"]
        m = C3.CTViT3D(dim=c["dim"], image_size=c["image_size"], patch_size=c["patch_size"],
                       temporal_size=c["temporal_size"], temporal_patch_size=c["temporal_patch_size"],
                       transformer_blocks=c["transformer_blocks"], dim_head=32, heads=c["heads"])
        m.load_state_dict(case["state_dict"], strict=False)
        params = [p.detach() for p in m._flat_params()]
        out, ctx = C3._forward(m, case["video"], params, save=True)
        assert _rel(out, case["out"]) < 1e-5
        p = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and k != "pos_embed")
             for k, v in m.state_dict().items()}
        ref = O.ctvit3d_forward(case["video"], p, patch=c["patch_size"], tpatch=c["temporal_patch_size"],
                                blocks=c["transformer_blocks"], heads=c["heads"])
        (ref * case["dy"]).sum().backward()
        grads = C3._backward(m, params, ctx, case["dy"])
        by_id = {id(q): n for n, q in m.named_parameters()}
        assert len(grads) == len(params)
        for q, g in zip(m._flat_params(), grads):
            n = by_id[id(q)]
            assert tuple(g.shape) == tuple(q.shape), n
            assert _rel(g, p[n].grad) < 5e-4, (n, _rel(g, p[n].grad))
        for n, g in case["grads"].items():                                       # the reference's own gradients
            q = dict(m.named_parameters())[n]
            mine = grads[[id(x) for x in m._flat_params()].index(id(q))]
            assert _rel(mine, g) < 5e-4, n


def test_no_grad_path_and_broadcast_gradient(doubles):
    torch.manual_seed(0)
    m = C3.CTViT3D(dim=96, image_size=8, patch_size=4, temporal_size=4, temporal_patch_size=2, transformer_blocks=1,
                   dim_head=32, heads=2)
    video = torch.rand(2, 1, 4, 8, 8)
    params = [p.detach() for p in m._flat_params()]
    out, ctx = C3._forward(m, video, params, save=False)
    assert ctx is None and out.shape == (2, 2, 2, 2, 96)
    dpool = torch.randn(2, 96)
    dense = (dpool / 8).view(2, 1, 1, 1, 96).expand(2, 2, 2, 2, 96)               # what _ClipHead hands back
    _, c1 = C3._forward(m, video, params, save=True)
    g1 = C3._backward(m, params, c1, dense)
    _, c2 = C3._forward(m, video, params, save=True)
    g2 = C3._backward(m, params, c2, dense.contiguous())
    for a, b in zip(g1, g2):
        assert _rel(a, b) < 1e-4
