"""Dual-branch MiT (SegFormer) encoder containers — same module/parameter tree as the reference
models/encoders/dual_segformer.py:19-527 (RGBXTransformer, mit_b0..b5).  Parameters live in ordinary
nn.Linear / nn.Conv2d / nn.LayerNorm / nn.BatchNorm2d leaves; the forward pass is executed by
rgbx_semantic_segmentation_b200.engine (token-major bf16 kernels), not by these modules."""
from functools import partial

import torch
import torch.nn as nn

from ..net_utils import FeatureFusionModule, FeatureRectifyModule, _Container, mit_init_


class DropPath(_Container):
    """Marker module holding the stochastic-depth probability (timm.DropPath in the reference)."""
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob


class DWConv(_Container):
    def __init__(self, dim=768):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=3, stride=1, padding=1, bias=True, groups=dim)


class Mlp(_Container):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.dwconv = DWConv(hidden_features)
        self.fc2 = nn.Linear(hidden_features, in_features)


class Attention(_Container):
    def __init__(self, dim, num_heads=8, qkv_bias=False, sr_ratio=1):
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} should be divided by num_heads {num_heads}."
        self.dim, self.num_heads, self.sr_ratio = dim, num_heads, sr_ratio
        self.scale = (dim // num_heads) ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        if sr_ratio > 1:
            self.sr = nn.Conv2d(dim, dim, kernel_size=sr_ratio, stride=sr_ratio)
            self.norm = nn.LayerNorm(dim)


class Block(_Container):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, drop_path=0., norm_layer=nn.LayerNorm, sr_ratio=1):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, sr_ratio=sr_ratio)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio))


class OverlapPatchEmbed(_Container):
    def __init__(self, patch_size=7, stride=4, in_chans=3, embed_dim=768):
        super().__init__()
        self.patch_size, self.stride = patch_size, stride
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=stride, padding=patch_size // 2)
        self.norm = nn.LayerNorm(embed_dim)


class RGBXTransformer(_Container):
    def __init__(self, in_chans=3, embed_dims=(64, 128, 256, 512), num_heads=(1, 2, 4, 8), mlp_ratios=(4, 4, 4, 4),
                 qkv_bias=False, drop_path_rate=0., norm_layer=nn.LayerNorm, norm_fuse=nn.BatchNorm2d,
                 depths=(3, 4, 6, 3), sr_ratios=(8, 4, 2, 1)):
        super().__init__()
        self.depths, self.embed_dims, self.num_heads = list(depths), list(embed_dims), list(num_heads)
        self.sr_ratios, self.mlp_ratios = list(sr_ratios), list(mlp_ratios)
        for pre in ("patch_embed", "extra_patch_embed"):
            for s in range(4):
                setattr(self, f"{pre}{s + 1}", OverlapPatchEmbed(7 if s == 0 else 3, 4 if s == 0 else 2,
                                                                 in_chans if s == 0 else embed_dims[s - 1], embed_dims[s]))
        # stochastic-depth schedule incl. the reference's stage-2 quirk (dual_segformer.py:256-309)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        cur = 0
        for s in range(4):
            for pre, npre, off in (("block", "norm", 0), ("extra_block", "extra_norm", 1)):
                probs = [dpr[cur + (off if s == 1 else i)] for i in range(depths[s])]
                setattr(self, f"{pre}{s + 1}", nn.ModuleList([
                    Block(dim=embed_dims[s], num_heads=num_heads[s], mlp_ratio=mlp_ratios[s], qkv_bias=qkv_bias,
                          drop_path=probs[i], norm_layer=norm_layer, sr_ratio=sr_ratios[s]) for i in range(depths[s])]))
                setattr(self, f"{npre}{s + 1}", norm_layer(embed_dims[s]))
            cur += depths[s]
        self.FRMs = nn.ModuleList([FeatureRectifyModule(dim=embed_dims[s], reduction=1) for s in range(4)])
        self.FFMs = nn.ModuleList([FeatureFusionModule(dim=embed_dims[s], reduction=1, num_heads=num_heads[s],
                                                       norm_layer=norm_fuse) for s in range(4)])
        self.apply(mit_init_)

    def init_weights(self, pretrained=None):
        if isinstance(pretrained, str):
            load_dualpath_model(self, pretrained)
        else:
            raise TypeError('pretrained must be a str or None')


def load_dualpath_model(model, model_file):
    """Reference dual_segformer.py:449-480: duplicate every single-branch SegFormer key to its `extra_*` twin."""
    if isinstance(model_file, str):
        raw = torch.load(model_file, map_location=torch.device('cpu'))
        if 'model' in raw.keys():
            raw = raw['model']
    else:
        raw = model_file
    sd = {}
    for k, v in raw.items():
        if k.find('patch_embed') >= 0:
            sd[k] = v
            sd[k.replace('patch_embed', 'extra_patch_embed')] = v
        elif k.find('block') >= 0:
            sd[k] = v
            sd[k.replace('block', 'extra_block')] = v
        elif k.find('norm') >= 0:
            sd[k] = v
            sd[k.replace('norm', 'extra_norm')] = v
    model.load_state_dict(sd, strict=False)


def _mit(embed_dims, depths):
    class _M(RGBXTransformer):
        def __init__(self, fuse_cfg=None, **kwargs):  # kwargs (norm_fuse) swallowed like the reference (App. A-3)
            super().__init__(embed_dims=embed_dims, num_heads=[1, 2, 5, 8], mlp_ratios=[4, 4, 4, 4], qkv_bias=True,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), depths=depths, sr_ratios=[8, 4, 2, 1],
                             drop_path_rate=0.1)
    return _M


mit_b0 = _mit([32, 64, 160, 256], [2, 2, 2, 2])
mit_b1 = _mit([64, 128, 320, 512], [2, 2, 2, 2])
mit_b2 = _mit([64, 128, 320, 512], [3, 4, 6, 3])
mit_b3 = _mit([64, 128, 320, 512], [3, 4, 18, 3])
mit_b4 = _mit([64, 128, 320, 512], [3, 8, 27, 3])
mit_b5 = _mit([64, 128, 320, 512], [3, 6, 40, 3])
for _n, _c in list(globals().items()):
    if _n.startswith("mit_b"):
        _c.__name__ = _c.__qualname__ = _n
