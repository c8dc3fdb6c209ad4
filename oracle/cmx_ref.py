"""ORACLE (test infrastructure, NOT product code).

CPU restatement, in plain functional PyTorch fp32, of the CMX dual-branch
MiT RGB-X segmentation hot path of ynalcakan/RGBX_Semantic_Segmentation.
It is written against a flat ``state_dict`` (the reference's own key schema)
so that the very same weights can be loaded into the reference model, into
this oracle and into the CUDA product model.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this file.  The product package
``rgbx_semantic_segmentation_b200`` never does.

Parity pin: ``tests/golden/make_golden.py`` imports the real reference from
``/root/reference`` (4 import shims), loads ``oracle.synth`` weights into it,
and stores its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this restatement against those fixtures (fp32 round-off tolerance).

Every function cites the reference file:line it restates (paths relative to
the reference repo root).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------
# Architecture table: models/encoders/dual_segformer.py:483-527 (mit_b0..b5)
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class MitSpec:
    embed_dims: Tuple[int, int, int, int]
    num_heads: Tuple[int, int, int, int] = (1, 2, 5, 8)
    mlp_ratios: Tuple[int, int, int, int] = (4, 4, 4, 4)
    depths: Tuple[int, int, int, int] = (3, 4, 6, 3)
    sr_ratios: Tuple[int, int, int, int] = (8, 4, 2, 1)
    drop_path_rate: float = 0.1


MIT_SPECS: Dict[str, MitSpec] = {
    "mit_b0": MitSpec((32, 64, 160, 256), depths=(2, 2, 2, 2)),
    "mit_b1": MitSpec((64, 128, 320, 512), depths=(2, 2, 2, 2)),
    "mit_b2": MitSpec((64, 128, 320, 512), depths=(3, 4, 6, 3)),
    "mit_b3": MitSpec((64, 128, 320, 512), depths=(3, 4, 18, 3)),
    "mit_b4": MitSpec((64, 128, 320, 512), depths=(3, 8, 27, 3)),
    "mit_b5": MitSpec((64, 128, 320, 512), depths=(3, 6, 40, 3)),
}


def drop_path_probs(spec: MitSpec) -> Tuple[List[List[float]], List[List[float]]]:
    """Per-block DropPath probabilities for the RGB and X branches.

    dual_segformer.py:256-309: ``dpr = linspace(0, rate, sum(depths))``; stages
    1/3/4 use ``dpr[cur+i]`` for both branches, stage 2 uses ``dpr[cur]`` for every
    RGB block and ``dpr[cur+1]`` for every X block.
    """
    dpr = [x.item() for x in torch.linspace(0, spec.drop_path_rate, sum(spec.depths))]
    rgb, ext = [], []
    cur = 0
    for s, d in enumerate(spec.depths):
        if s == 1:
            rgb.append([dpr[cur]] * d)
            ext.append([dpr[cur + 1]] * d)
        else:
            rgb.append([dpr[cur + i] for i in range(d)])
            ext.append([dpr[cur + i] for i in range(d)])
        cur += d
    return rgb, ext


# --------------------------------------------------------------------------
# Encoder pieces
# --------------------------------------------------------------------------
def overlap_patch_embed(sd: SD, p: str, x: Tensor, k: int, s: int) -> Tuple[Tensor, int, int]:
    """dual_segformer.py:217-225 — conv(k, stride s, pad k//2) → NLC → LN(eps 1e-5)."""
    x = F.conv2d(x, sd[p + ".proj.weight"], sd[p + ".proj.bias"], stride=s, padding=k // 2)
    _, C, H, W = x.shape
    x = x.flatten(2).transpose(1, 2)
    x = F.layer_norm(x, (C,), sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-5)
    return x, H, W


def attention(sd: SD, p: str, x: Tensor, H: int, W: int, heads: int, sr: int) -> Tensor:
    """dual_segformer.py:116-138 — spatial-reduction efficient self-attention."""
    B, N, C = x.shape
    d = C // heads
    scale = d ** -0.5
    q = F.linear(x, sd[p + ".q.weight"], sd[p + ".q.bias"]).reshape(B, N, heads, d).permute(0, 2, 1, 3)
    if sr > 1:
        x_ = x.permute(0, 2, 1).reshape(B, C, H, W)
        x_ = F.conv2d(x_, sd[p + ".sr.weight"], sd[p + ".sr.bias"], stride=sr).reshape(B, C, -1).permute(0, 2, 1)
        x_ = F.layer_norm(x_, (C,), sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-5)
    else:
        x_ = x
    kv = F.linear(x_, sd[p + ".kv.weight"], sd[p + ".kv.bias"]).reshape(B, -1, 2, heads, d).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    attn = (q @ k.transpose(-2, -1)) * scale
    attn = attn.softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(x, sd[p + ".proj.weight"], sd[p + ".proj.bias"])


def mix_ffn(sd: SD, p: str, x: Tensor, H: int, W: int) -> Tensor:
    """dual_segformer.py:27-33, 67-74 — fc1 → depthwise 3×3 (pad 1) → GELU(erf) → fc2."""
    B, N, _ = x.shape
    x = F.linear(x, sd[p + ".fc1.weight"], sd[p + ".fc1.bias"])
    Ch = x.shape[-1]
    x = x.permute(0, 2, 1).reshape(B, Ch, H, W)
    x = F.conv2d(x, sd[p + ".dwconv.dwconv.weight"], sd[p + ".dwconv.dwconv.bias"], padding=1, groups=Ch)
    x = x.flatten(2).transpose(1, 2)
    x = F.gelu(x)
    return F.linear(x, sd[p + ".fc2.weight"], sd[p + ".fc2.bias"])


def block(sd: SD, p: str, x: Tensor, H: int, W: int, heads: int, sr: int,
          dp_scale: Optional[Tuple[Tensor, Tensor]] = None) -> Tensor:
    """dual_segformer.py:176-180.  ``dp_scale`` = per-sample DropPath multipliers
    ``mask/keep_prob`` of shape [B] for the attention and MLP residuals (None = eval / p=0).
    LayerNorm eps 1e-6 (dual_segformer.py:487)."""
    C = x.shape[-1]
    a = attention(sd, p + ".attn", F.layer_norm(x, (C,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-6),
                  H, W, heads, sr)
    if dp_scale is not None:
        a = a * dp_scale[0].view(-1, 1, 1)
    x = x + a
    m = mix_ffn(sd, p + ".mlp", F.layer_norm(x, (C,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-6), H, W)
    if dp_scale is not None:
        m = m * dp_scale[1].view(-1, 1, 1)
    return x + m


# --------------------------------------------------------------------------
# Fusion modules (models/net_utils.py)
# --------------------------------------------------------------------------
def frm(sd: SD, p: str, x1: Tensor, x2: Tensor) -> Tuple[Tensor, Tensor]:
    """net_utils.py:22-30 (ChannelWeights), 79-83 (SpatialWeights), 147-152 (rectify, λ=0.5)."""
    B, C, H, W = x1.shape
    x = torch.cat((x1, x2), dim=1)
    avg = x.mean(dim=(2, 3))
    mx = F.adaptive_max_pool2d(x, 1).flatten(1)  # nn.AdaptiveMaxPool2d(1): gradient goes to the first maximum
    y = torch.cat((avg, mx), dim=1)
    y = F.relu(F.linear(y, sd[p + ".channel_weights.mlp.0.weight"], sd[p + ".channel_weights.mlp.0.bias"]))
    y = torch.sigmoid(F.linear(y, sd[p + ".channel_weights.mlp.2.weight"], sd[p + ".channel_weights.mlp.2.bias"]))
    cw = y.reshape(B, 2, C, 1, 1).permute(1, 0, 2, 3, 4)
    s = F.relu(F.conv2d(x, sd[p + ".spatial_weights.mlp.0.weight"], sd[p + ".spatial_weights.mlp.0.bias"]))
    s = torch.sigmoid(F.conv2d(s, sd[p + ".spatial_weights.mlp.2.weight"], sd[p + ".spatial_weights.mlp.2.bias"]))
    sw = s.reshape(B, 2, 1, H, W).permute(1, 0, 2, 3, 4)
    out1 = x1 + 0.5 * cw[1] * x2 + 0.5 * sw[1] * x2
    out2 = x2 + 0.5 * cw[0] * x1 + 0.5 * sw[0] * x1
    return out1, out2


def cross_attention(sd: SD, p: str, u1: Tensor, u2: Tensor, heads: int) -> Tuple[Tensor, Tensor]:
    """net_utils.py:199-214 — linear-complexity cross attention; softmax over dim -2."""
    B, N, C = u1.shape
    d = C // heads
    scale = d ** -0.5
    q1 = u1.reshape(B, N, heads, d).permute(0, 2, 1, 3)
    q2 = u2.reshape(B, N, heads, d).permute(0, 2, 1, 3)
    k1, v1 = F.linear(u1, sd[p + ".kv1.weight"]).reshape(B, N, 2, heads, d).permute(2, 0, 3, 1, 4)
    k2, v2 = F.linear(u2, sd[p + ".kv2.weight"]).reshape(B, N, 2, heads, d).permute(2, 0, 3, 1, 4)
    ctx1 = ((k1.transpose(-2, -1) @ v1) * scale).softmax(dim=-2)
    ctx2 = ((k2.transpose(-2, -1) @ v2) * scale).softmax(dim=-2)
    o1 = (q1 @ ctx2).permute(0, 2, 1, 3).reshape(B, N, C)
    o2 = (q2 @ ctx1).permute(0, 2, 1, 3).reshape(B, N, C)
    return o1, o2


def cross_path(sd: SD, p: str, x1: Tensor, x2: Tensor, heads: int) -> Tuple[Tensor, Tensor]:
    """net_utils.py:273-281."""
    C = x1.shape[-1]
    y1, u1 = F.relu(F.linear(x1, sd[p + ".channel_proj1.weight"], sd[p + ".channel_proj1.bias"])).chunk(2, dim=-1)
    y2, u2 = F.relu(F.linear(x2, sd[p + ".channel_proj2.weight"], sd[p + ".channel_proj2.bias"])).chunk(2, dim=-1)
    v1, v2 = cross_attention(sd, p + ".cross_attn", u1, u2, heads)
    y1 = torch.cat((y1, v1), dim=-1)
    y2 = torch.cat((y2, v2), dim=-1)
    o1 = F.layer_norm(x1 + F.linear(y1, sd[p + ".end_proj1.weight"], sd[p + ".end_proj1.bias"]), (C,),
                      sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-5)
    o2 = F.layer_norm(x2 + F.linear(y2, sd[p + ".end_proj2.weight"], sd[p + ".end_proj2.bias"]), (C,),
                      sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-5)
    return o1, o2


def _batch_norm(sd: SD, p: str, x: Tensor, training: bool, eps: float, momentum: float,
                new_stats: Optional[SD]) -> Tensor:
    """nn.BatchNorm2d semantics; in training mode the updated running stats are
    written to ``new_stats`` (functional: ``sd`` itself is left untouched)."""
    rm, rv = sd[p + ".running_mean"], sd[p + ".running_var"]
    if training:
        rm, rv = rm.clone(), rv.clone()
    y = F.batch_norm(x, rm, rv, sd[p + ".weight"], sd[p + ".bias"], training, momentum, eps)
    if training and new_stats is not None:
        new_stats[p + ".running_mean"] = rm
        new_stats[p + ".running_var"] = rv
        new_stats[p + ".num_batches_tracked"] = sd[p + ".num_batches_tracked"] + 1
    return y


def channel_embed(sd: SD, p: str, x: Tensor, H: int, W: int, training: bool, new_stats: Optional[SD]) -> Tensor:
    """net_utils.py:323-329; BN = plain BatchNorm2d eps 1e-5 momentum 0.1."""
    B, N, C2 = x.shape
    x = x.permute(0, 2, 1).reshape(B, C2, H, W)
    residual = F.conv2d(x, sd[p + ".residual.weight"])
    C = residual.shape[1]
    y = F.conv2d(x, sd[p + ".channel_embed.0.weight"], sd[p + ".channel_embed.0.bias"])
    y = F.conv2d(y, sd[p + ".channel_embed.1.weight"], sd[p + ".channel_embed.1.bias"], padding=1, groups=C)
    y = F.relu(y)
    y = F.conv2d(y, sd[p + ".channel_embed.3.weight"], sd[p + ".channel_embed.3.bias"])
    y = _batch_norm(sd, p + ".channel_embed.4", y, training, 1e-5, 0.1, new_stats)
    return _batch_norm(sd, p + ".norm", residual + y, training, 1e-5, 0.1, new_stats)


def ffm(sd: SD, p: str, x1: Tensor, x2: Tensor, heads: int, training: bool, new_stats: Optional[SD]) -> Tensor:
    """net_utils.py:376-384."""
    B, C, H, W = x1.shape
    t1 = x1.flatten(2).transpose(1, 2)
    t2 = x2.flatten(2).transpose(1, 2)
    t1, t2 = cross_path(sd, p + ".cross", t1, t2, heads)
    merge = torch.cat((t1, t2), dim=-1)
    return channel_embed(sd, p + ".channel_emb", merge, H, W, training, new_stats)


# --------------------------------------------------------------------------
# Backbone, decoder, loss
# --------------------------------------------------------------------------
def backbone(sd: SD, spec: MitSpec, rgb: Tensor, x: Tensor, training: bool = False,
             new_stats: Optional[SD] = None,
             dp_scales: Optional[Dict[str, Tuple[Tensor, Tensor]]] = None,
             p: str = "backbone", trace: Optional[Dict[str, Tensor]] = None) -> List[Tensor]:
    """dual_segformer.py:366-442.  ``dp_scales`` maps a block prefix (e.g.
    ``backbone.block1.2``) to its (attn, mlp) per-sample DropPath multipliers."""
    B = rgb.shape[0]
    outs = []
    x_rgb, x_e = rgb, x

    def tr(key, t, nchw=False):
        if trace is not None:
            t = t.detach()
            trace[key] = t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]) if nchw else t.reshape(-1, t.shape[-1])

    for s in range(4):
        k, st = (7, 4) if s == 0 else (3, 2)
        x_rgb, H, W = overlap_patch_embed(sd, f"{p}.patch_embed{s + 1}", x_rgb, k, st)
        x_e, _, _ = overlap_patch_embed(sd, f"{p}.extra_patch_embed{s + 1}", x_e, k, st)
        tr(f"{p}.patch_embed{s + 1}", x_rgb)
        tr(f"{p}.extra_patch_embed{s + 1}", x_e)
        for i in range(spec.depths[s]):
            bp = f"{p}.block{s + 1}.{i}"
            x_rgb = block(sd, bp, x_rgb, H, W, spec.num_heads[s], spec.sr_ratios[s],
                          None if dp_scales is None else dp_scales.get(bp))
            tr(bp, x_rgb)
        for i in range(spec.depths[s]):
            bp = f"{p}.extra_block{s + 1}.{i}"
            x_e = block(sd, bp, x_e, H, W, spec.num_heads[s], spec.sr_ratios[s],
                        None if dp_scales is None else dp_scales.get(bp))
            tr(bp, x_e)
        C = spec.embed_dims[s]
        x_rgb = F.layer_norm(x_rgb, (C,), sd[f"{p}.norm{s + 1}.weight"], sd[f"{p}.norm{s + 1}.bias"], 1e-6)
        x_e = F.layer_norm(x_e, (C,), sd[f"{p}.extra_norm{s + 1}.weight"], sd[f"{p}.extra_norm{s + 1}.bias"], 1e-6)
        x_rgb = x_rgb.reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous()
        x_e = x_e.reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous()
        tr(f"{p}.norm{s + 1}", x_rgb, True)
        tr(f"{p}.extra_norm{s + 1}", x_e, True)
        x_rgb, x_e = frm(sd, f"{p}.FRMs.{s}", x_rgb, x_e)
        tr(f"{p}.FRMs.{s}.out1", x_rgb, True)
        tr(f"{p}.FRMs.{s}.out2", x_e, True)
        outs.append(ffm(sd, f"{p}.FFMs.{s}", x_rgb, x_e, spec.num_heads[s], training, new_stats))
        tr(f"{p}.FFMs.{s}", outs[-1], True)
    return outs


def _decoder_pre_bn(sd: SD, feats: Sequence[Tensor], p: str) -> Tensor:
    """MLPDecoder.py:59-77: linear_c{1-4}, bilinear upsample to 1/4 resolution, concat [c4, c3, c2, c1], 1x1 linear_fuse conv"""
    c1, c2, c3, c4 = feats
    n = c4.shape[0]
    size = c1.shape[2:]

    def mlp(c: Tensor, name: str) -> Tensor:
        t = F.linear(c.flatten(2).transpose(1, 2), sd[f"{p}.{name}.proj.weight"], sd[f"{p}.{name}.proj.bias"])
        return t.permute(0, 2, 1).reshape(n, -1, c.shape[2], c.shape[3])

    _c4 = F.interpolate(mlp(c4, "linear_c4"), size=size, mode="bilinear", align_corners=False)
    _c3 = F.interpolate(mlp(c3, "linear_c3"), size=size, mode="bilinear", align_corners=False)
    _c2 = F.interpolate(mlp(c2, "linear_c2"), size=size, mode="bilinear", align_corners=False)
    _c1 = mlp(c1, "linear_c1")
    return F.conv2d(torch.cat([_c4, _c3, _c2, _c1], dim=1), sd[f"{p}.linear_fuse.0.weight"], sd[f"{p}.linear_fuse.0.bias"])


def _decoder_post_bn(sd: SD, y: Tensor, dropout_scale: Optional[Tensor], p: str) -> Tensor:
    """MLPDecoder.py:77-79: ReLU, Dropout2d multiplier, linear_pred"""
    y = F.relu(y)
    if dropout_scale is not None:
        y = y * dropout_scale[:, :, None, None]
    return F.conv2d(y, sd[f"{p}.linear_pred.weight"], sd[f"{p}.linear_pred.bias"])


def decoder_head(sd: SD, feats: Sequence[Tensor], training: bool = False, bn_eps: float = 1e-5,
                 bn_momentum: float = 0.1, new_stats: Optional[SD] = None,
                 dropout_scale: Optional[Tensor] = None, p: str = "decode_head") -> Tensor:
    """models/decoders/MLPDecoder.py:59-81.  ``dropout_scale`` = Dropout2d multiplier
    ``mask/0.9`` of shape [B, E] (None = eval / p=0)."""
    y = _decoder_pre_bn(sd, feats, p)
    y = _batch_norm(sd, f"{p}.linear_fuse.1", y, training, bn_eps, bn_momentum, new_stats)
    return _decoder_post_bn(sd, y, dropout_scale, p)


def encode_decode(sd: SD, spec: MitSpec, rgb: Tensor, x: Tensor, training: bool = False,
                  decoder_bn_eps: float = 1e-5, new_stats: Optional[SD] = None,
                  dp_scales=None, dropout_scale=None, trace=None) -> Tensor:
    """models/builder.py:212-238 — backbone → decoder → bilinear to input size."""
    feats = backbone(sd, spec, rgb, x, training, new_stats, dp_scales, trace=trace)
    out = decoder_head(sd, feats, training, decoder_bn_eps, 0.1, new_stats, dropout_scale)
    if trace is not None:
        trace["decode_head.logits"] = out.detach().permute(0, 2, 3, 1).reshape(-1, out.shape[1])
    return F.interpolate(out, size=rgb.shape[2:], mode="bilinear", align_corners=False)


def forward(sd: SD, spec: MitSpec, rgb: Tensor, x: Tensor, label: Optional[Tensor] = None,
            training: bool = False, decoder_bn_eps: float = 1e-5, ignore_index: int = 255,
            new_stats: Optional[SD] = None, dp_scales=None, dropout_scale=None, trace=None) -> Tensor:
    """models/builder.py:240-253 with criterion = CrossEntropyLoss(mean, ignore 255) (train.py:72-73)."""
    out = encode_decode(sd, spec, rgb, x, training, decoder_bn_eps, new_stats, dp_scales, dropout_scale, trace)
    if label is not None:
        return F.cross_entropy(out, label.long(), ignore_index=ignore_index, reduction="mean")
    return out


def forward_data_parallel(sd: SD, spec: MitSpec, shards: Sequence[Tuple[Tensor, Tensor, Tensor]],
                          sync_decoder_bn: bool = True, decoder_bn_eps: float = 1e-3, ignore_index: int = 255,
                          new_stats: Optional[SD] = None) -> Tuple[Tensor, List[Tensor]]:
    """What the reference's distributed training step computes (train.py:64-67, 145-146, 186-200), restated on ONE set of
    parameters: every rank runs the model on its own shard (rgb, x, label) - so the FFM BatchNorms, which are always plain
    nn.BatchNorm2d, use per-rank statistics -, the decoder norm is an nn.SyncBatchNorm whose statistics
    are taken over the shards of all ranks (equal per-rank counts), every rank's loss is the mean CE over ITS valid pixels,
    and DistributedDataParallel averages the per-rank gradients: d/dtheta of mean_r(loss_r).  Returns (mean_r loss_r,
    [loss_r]); call .backward() on the first to get the averaged gradients.  DropPath / Dropout2d are off.
    Running statistics in ``new_stats``: the FFM ones are rank 0's (DDP broadcasts rank 0's buffers before every forward)."""
    pre, p = [], "decode_head"
    for r, (rgb, x, _) in enumerate(shards):
        feats = backbone(sd, spec, rgb, x, True, new_stats if r == 0 else None)
        pre.append(_decoder_pre_bn(sd, feats, p))
    if sync_decoder_bn:
        y = _batch_norm(sd, f"{p}.linear_fuse.1", torch.cat(pre, 0), True, decoder_bn_eps, 0.1, new_stats)
        ys = list(y.split([t.shape[0] for t in pre], 0))
    else:
        ys = [_batch_norm(sd, f"{p}.linear_fuse.1", t, True, decoder_bn_eps, 0.1, new_stats if r == 0 else None)
              for r, t in enumerate(pre)]
    losses = []
    for (rgb, _, label), y in zip(shards, ys):
        out = F.interpolate(_decoder_post_bn(sd, y, None, p), size=rgb.shape[2:], mode="bilinear", align_corners=False)
        losses.append(F.cross_entropy(out, label.long(), ignore_index=ignore_index, reduction="mean"))
    return torch.stack(losses).mean(), losses


# --------------------------------------------------------------------------
# state_dict schema: names, shapes; 837 entries for mit_b2 / 9 classes
# --------------------------------------------------------------------------
def state_dict_schema(spec: MitSpec, num_classes: int, embed_dim: int = 512,
                      decoder_channels: Optional[Sequence[int]] = None) -> Dict[str, Tuple[Tuple[int, ...], str]]:
    """name -> (shape, kind) in the reference's registration order.
    kind ∈ {linear_w, conv_w, bias, norm_w, norm_b, bn_mean, bn_var, bn_count}."""
    E = spec.embed_dims
    dec_ch = list(decoder_channels) if decoder_channels is not None else list(E)
    out: Dict[str, Tuple[Tuple[int, ...], str]] = {}

    def lin(p, i, o, bias=True):
        out[p + ".weight"] = ((o, i), "linear_w")
        if bias:
            out[p + ".bias"] = ((o,), "bias")

    def conv(p, i, o, k, groups=1, bias=True):
        out[p + ".weight"] = ((o, i // groups, k, k), "conv_w")
        if bias:
            out[p + ".bias"] = ((o,), "bias")

    def ln(p, c):
        out[p + ".weight"] = ((c,), "norm_w")
        out[p + ".bias"] = ((c,), "norm_b")

    def bn(p, c):
        ln(p, c)
        out[p + ".running_mean"] = ((c,), "bn_mean")
        out[p + ".running_var"] = ((c,), "bn_var")
        out[p + ".num_batches_tracked"] = ((), "bn_count")

    def pe(p, cin, cout, k):
        conv(p + ".proj", cin, cout, k)
        ln(p + ".norm", cout)

    def blk(p, c, sr):
        ln(p + ".norm1", c)
        lin(p + ".attn.q", c, c)
        lin(p + ".attn.kv", c, 2 * c)
        lin(p + ".attn.proj", c, c)
        if sr > 1:
            conv(p + ".attn.sr", c, c, sr)
            ln(p + ".attn.norm", c)
        ln(p + ".norm2", c)
        lin(p + ".mlp.fc1", c, 4 * c)
        conv(p + ".mlp.dwconv.dwconv", 4 * c, 4 * c, 3, groups=4 * c)
        lin(p + ".mlp.fc2", 4 * c, c)

    b = "backbone"
    # registration order of RGBXTransformer.__init__ (dual_segformer.py:237-335)
    for pre in ("patch_embed", "extra_patch_embed"):
        for s in range(4):
            pe(f"{b}.{pre}{s + 1}", 3 if s == 0 else E[s - 1], E[s], 7 if s == 0 else 3)
    for s in range(4):
        for pre, npre in (("block", "norm"), ("extra_block", "extra_norm")):
            for i in range(spec.depths[s]):
                blk(f"{b}.{pre}{s + 1}.{i}", E[s], spec.sr_ratios[s])
            ln(f"{b}.{npre}{s + 1}", E[s])
    for s in range(4):
        p = f"{b}.FRMs.{s}"
        c = E[s]
        lin(p + ".channel_weights.mlp.0", 4 * c, 4 * c)
        lin(p + ".channel_weights.mlp.2", 4 * c, 2 * c)
        conv(p + ".spatial_weights.mlp.0", 2 * c, c, 1)
        conv(p + ".spatial_weights.mlp.2", c, 2, 1)
    for s in range(4):
        p = f"{b}.FFMs.{s}"
        c = E[s]
        lin(p + ".cross.channel_proj1", c, 2 * c)
        lin(p + ".cross.channel_proj2", c, 2 * c)
        lin(p + ".cross.cross_attn.kv1", c, 2 * c, bias=False)
        lin(p + ".cross.cross_attn.kv2", c, 2 * c, bias=False)
        lin(p + ".cross.end_proj1", 2 * c, c)
        lin(p + ".cross.end_proj2", 2 * c, c)
        ln(p + ".cross.norm1", c)
        ln(p + ".cross.norm2", c)
        conv(p + ".channel_emb.residual", 2 * c, c, 1, bias=False)
        conv(p + ".channel_emb.channel_embed.0", 2 * c, c, 1)
        conv(p + ".channel_emb.channel_embed.1", c, c, 3, groups=c)
        conv(p + ".channel_emb.channel_embed.3", c, c, 1)
        bn(p + ".channel_emb.channel_embed.4", c)
        bn(p + ".channel_emb.norm", c)
    d = "decode_head"
    for i in (4, 3, 2, 1):  # MLPDecoder.py:46-49 registration order
        lin(f"{d}.linear_c{i}.proj", dec_ch[i - 1], embed_dim)
    conv(f"{d}.linear_fuse.0", 4 * embed_dim, embed_dim, 1)
    bn(f"{d}.linear_fuse.1", embed_dim)
    conv(f"{d}.linear_pred", embed_dim, num_classes, 1)
    return out
