"""ORACLE helper (test infrastructure): deterministic synthetic weights and inputs.

Weights are drawn per key from a generator seeded by (seed, crc32(key)) so that a
fixture can be regenerated on any machine with the same torch build without
committing 266 MB of parameters.  Distributions are chosen so that every op of
the path is exercised non-trivially (non-zero biases, non-unit norm scales,
non-trivial BN running statistics) — unlike the reference's init, which zeroes
all biases (models/encoders/dual_segformer.py:52-65).
"""
import zlib
from typing import Dict

import torch

from .cmx_ref import MitSpec, state_dict_schema


def synth_state_dict(spec: MitSpec, num_classes: int, seed: int = 0, embed_dim: int = 512,
                     ctx_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """ctx_gain scales the FFM cross-attention kv weights.  With unit-variance K/V the context logits
    (net_utils.py:209: K^T V * scale, a sum over all N tokens) have std ~ sqrt(N)/8 — measured 119 at N = 4800 and
    ~240 at N = 19200 (480x640, stage 1) — so the dim=-2 softmax is one-hot and amplifies the bf16 rounding of K/V
    into O(10 %) gradient noise; the oracle itself run under bf16 autocast shows the same.  Full-resolution parity
    tests therefore pass ctx_gain=0.1 (logit std ~2.4), which keeps the softmax in its sensitive, non-saturated range."""
    sd = {}
    for name, (shape, kind) in state_dict_schema(spec, num_classes, embed_dim).items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        if kind == "linear_w":
            t = torch.randn(shape, generator=g) * (1.0 / shape[1]) ** 0.5
        elif kind == "conv_w":
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5
        elif kind == "bias":
            t = torch.randn(shape, generator=g) * 0.05
        elif kind == "norm_w":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "norm_b":
            t = 0.05 * torch.randn(shape, generator=g)
        elif kind == "bn_mean":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_var":
            t = 0.5 + torch.rand(shape, generator=g)
        elif kind == "bn_count":
            t = torch.tensor(3, dtype=torch.long)
        else:
            raise KeyError(kind)
        if ctx_gain != 1.0 and ".cross_attn.kv" in name:
            t = t * ctx_gain
        sd[name] = t
    return sd


def synth_inputs(batch: int, height: int, width: int, num_classes: int, seed: int = 1,
                 ignore_frac: float = 0.03, ignore_index: int = 255):
    """BASELINE.json config inputs: N(0,1) RGB and X (already-normalised image statistics),
    int64 labels with a fraction of ignore pixels."""
    g = torch.Generator().manual_seed(seed)
    rgb = torch.randn(batch, 3, height, width, generator=g)
    x = torch.randn(batch, 3, height, width, generator=g)
    gt = torch.randint(0, num_classes, (batch, height, width), generator=g)
    ign = torch.rand(batch, height, width, generator=g) < ignore_frac
    gt[ign] = ignore_index
    return rgb, x, gt
