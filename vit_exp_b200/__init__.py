"""vit_exp_b200 - B200-native (sm_100a) CT-CLIP training hot path.

Drop-in modules mirroring the reference packages:
  vit_exp_b200.transformer_maskgit.CTViT   (reference: transformer_maskgit/ctvit.py)
  vit_exp_b200.ct_clip.CTCLIP              (reference: CT_CLIP/ct_clip/ct_clip.py)
Host code is PyTorch; all arithmetic on the path runs in libctk.so (include/ctk.h).
"""
__version__ = "0.1.0"
