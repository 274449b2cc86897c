"""ctypes binding of libctk.so (the C ABI declared in include/ctk.h).

There is deliberately no fallback: if the shared library is missing or the device is not an
sm_100-class GPU, calls raise.  PyTorch is only used by callers for memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libctk.so"

CTK_OK = 0
(EPI_BF16, EPI_F32, EPI_RESID_F32, EPI_GEGLU, EPI_GEGLU_BWD, EPI_QKV, EPI_ATOMIC_F32, EPI_ARGMAX, EPI_GELU,
 EPI_GELU_BWD, EPI_LSE_PART, EPI_CLIP_GRAD, EPI_ARGMAX_PART) = range(13)

_vp = C.c_void_p
_ll = C.c_longlong
_i = C.c_int
_f = C.c_float
_sz = C.c_size_t


class GemmEpilogue(C.Structure):
    _fields_ = [
        ("C", _vp), ("ldc", _ll), ("bias", _vp), ("resid", _vp), ("ldr", _ll),
        ("aux0", _vp), ("ld_aux0", _ll), ("vec0", _vp), ("vec1", _vp), ("row_map", _vp),
        ("alpha", _f), ("i0", _i), ("i1", _i),
    ]


# name -> (restype, argtypes); mirrors include/ctk.h one to one
SIGNATURES = {
    "ctk_last_error": (C.c_char_p, []),
    "ctk_version": (_i, []),
    "ctk_device_ok": (_i, []),
    "ctk_launch_count": (C.c_ulonglong, []),
    "ctk_gemm_bf16": (_i, [_vp, _ll, _i, _vp, _ll, _i, _i, _i, _i, _i, C.POINTER(GemmEpilogue), _i, _vp]),
    "ctk_cast_bf16": (_i, [_vp, _vp, _ll, _ll, _ll, _vp, _vp]),
    "ctk_transpose_cast_bf16": (_i, [_vp, _vp, _ll, _ll, _ll, _vp]),
    "ctk_pack_ff_w1": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "ctk_patch_norm_fwd": (_i, [_vp, _vp, _ll, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "ctk_layernorm_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _f, _i, _i, _vp]),
    "ctk_layernorm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _ll, _i, _i, _i, _ll, _f, _vp]),
    "ctk_peg_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ctk_peg_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ctk_cpb_fwd": (_i, [_vp] * 9 + [_i, _i, _i, _i, _vp]),
    "ctk_cpb_bwd": (_i, [_vp] * 13 + [_i, _i, _i, _i, _vp]),
    "ctk_opt_chunk_elems": (_i, []),
    "ctk_copy_from_pinned": (_i, [_vp, _vp, _ll, _vp]),
    "ctk_multi_sqnorm": (_i, [_vp, _i, _ll, _vp, _vp]),
    "ctk_multi_adam": (_i, [_vp, _i, _ll, _vp, _f, _f, _f, _f, _f, _i, _ll, _f, _i, _vp]),
    "ctk_attn_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ctk_attn_bwd": (_i, [_vp] * 8 + [_i, _i, _i, _i, _i, _vp]),
    "ctk_qknorm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _ll, _i, _vp]),
    "ctk_l2norm_rows": (_i, [_vp, _vp, _vp, _ll, _i, _vp]),
    "ctk_vq_select": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _f, _vp]),
    "ctk_vq_gather": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _vp]),
    "ctk_vq_ema_update": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _f, _vp]),
    "ctk_mean_pool_fwd": (_i, [_vp, _vp, _i, _ll, _i, _vp]),
    "ctk_latent_fwd": (_i, [_vp, _ll, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "ctk_latent_bwd": (_i, [_vp, _vp, _vp, _vp, _ll, _vp, _vp, _vp, _ll, _i, _i, _i, _vp]),
    "ctk_clip_loss_ws_bytes": (_sz, [_i, _i]),
    "ctk_clip_loss_fwd_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _vp]),
    "ctk_pair_logits": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "ctk_fill_f32": (_i, [_vp, _f, _ll, _vp]),
    "ctk_patch_affine_bwd": (_i, [_vp] * 8 + [_i, _i, _vp]),
    "ctk_colsum": (_i, [_vp, _vp, _vp, _ll, _i, _vp]),
    "ctk_transpose_cast_bf16_slice": (_i, [_vp, _vp, _ll, _ll, _ll, _vp]),
    "ctk_bert_embed_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _vp]),
    "ctk_bert_embed_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _ll, _vp]),
    "ctk_mha_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _f, _f, _vp, C.c_ulonglong, _vp]),
    "ctk_mha_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _vp,
                         C.c_ulonglong, _vp]),
    "ctk_volume_prep": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp]),
}

_lib = None


class CtkError(RuntimeError):
    pass


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load libctk.so, building it in-tree with nvcc on first use. Never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if not build_if_missing or os.environ.get("CTK_NO_BUILD"):
            raise CtkError(f"{LIB_PATH} is missing: run `python -m vit_exp_b200.build` (no fallback path exists)")
        from .build import build
        build()
    lib = C.CDLL(str(LIB_PATH))
    partial = bool(os.environ.get("CTK_DEV_PARTIAL"))   # kernel bring-up only
    for name, (res, args) in SIGNATURES.items():
        if partial and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)          # AttributeError if the ABI and the header drifted apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != CTK_OK:
        msg = load().ctk_last_error()
        raise CtkError(f"{what or 'ctk call'} failed (status {rc}): {msg.decode() if msg else ''}")
