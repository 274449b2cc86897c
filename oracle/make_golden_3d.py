"""Generate tests/golden/ctvit3d_golden.pt by running the REAL reference CTViT3D (ctvit3d.py) in the build container:

    python -m oracle.make_golden_3d

Tiny configurations, seeded inputs; records the state dict (incl. the fixed sin/cos `pos_embed` the reference
builds), the encoded tokens and a sample of parameter gradients.  tests/test_oracle_cpu.py pins
oracle.ctclip_oracle.{flash_attention, sincos_pos_embed_3d, ctvit3d_forward} against it.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "ctvit3d_golden.pt")


def main():
    from oracle.ref_import import import_reference
    import_reference()
    from transformer_maskgit.ctvit3d import CTViT3D, get_3d_sincos_pos_embed
    cases = []
    for seed, (dim, img, ps, ts, tps, blocks, heads) in enumerate([(96, 8, 4, 6, 2, 2, 2), (48, 12, 4, 4, 2, 1, 1)]):
        torch.manual_seed(seed)
        m = CTViT3D(dim=dim, image_size=img, patch_size=ps, temporal_size=ts, temporal_patch_size=tps,
                    transformer_blocks=blocks, dim_head=32, heads=heads)
        g = torch.Generator().manual_seed(100 + seed)
        with torch.no_grad():
            for n, p in m.named_parameters():
                if n.endswith(("gamma", "q_scale", "k_scale")) or (".0.weight" in n and p.dim() == 1):
                    p.mul_(1 + 0.2 * torch.randn(p.shape, generator=g))
                elif p.dim() == 1 and p.requires_grad:
                    p.add_(0.1 * torch.randn(p.shape, generator=g))
        video = torch.rand(2, 1, ts, img, img, generator=g)
        out = m(video, return_encoded_tokens=True)
        dy = torch.randn(out.shape, generator=g)
        (out * dy).sum().backward()
        keep = ("enc_3D.layers.0.1.null_kv", "enc_3D.layers.0.1.k_scale", "enc_3D.layers.0.1.q_scale",
                "enc_3D.layers.0.1.to_kv.weight", "enc_3D.layers.0.3.1.weight", "to_patch_emb.2.weight",
                "enc_3D.norm_out.gamma", "enc_3D.layers.0.1.norm.gamma")
        grads = {n: p.grad.clone() for n, p in m.named_parameters() if n in keep}
        sd = {k: v.clone() for k, v in m.state_dict().items() if not k.startswith("to_pixels")}
        cases.append(dict(cfg=dict(dim=dim, image_size=img, patch_size=ps, temporal_size=ts, temporal_patch_size=tps,
                                   transformer_blocks=blocks, heads=heads), state_dict=sd, video=video, out=out.detach(),
                          dy=dy, grads=grads))
    # the position table alone at a shape where n_t != n_w (the reshape quirk matters) and at the production grid's aspect
    pos = {str(g): torch.tensor(get_3d_sincos_pos_embed(d, list(g))).float() for d, g in ((96, (3, 2, 4)), (48, (2, 3, 3)))}
    torch.save(dict(cases=cases, pos=pos, unused=sorted(n for n, p in m.named_parameters() if p.grad is None)), OUT)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
