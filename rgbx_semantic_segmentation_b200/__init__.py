"""cmx-b200: B200-native hot path of CMX-style RGB-X semantic segmentation (reference:
ynalcakan/RGBX_Semantic_Segmentation).  Host-side mirror of the reference's `models.builder.EncoderDecoder`
and `utils.metric` on top of hand-written sm_100a kernels (csrc/, C ABI in include/cmx_b200.h)."""
__version__ = "0.1.0"
