"""Parameter containers of the FRM / FFM fusion modules, mirroring the reference module tree
(models/net_utils.py:10-30, 69-83, 124-152, 187-214, 260-281, 309-329, 354-384) so that state_dict keys,
`group_weight` (utils/init_func.py:33-57) and `init_weight` behave identically.  The containers hold
parameters only; the arithmetic runs in rgbx_semantic_segmentation_b200.engine on the CUDA kernels."""
import math

import torch.nn as nn


def _no_forward(self, *a, **k):
    raise RuntimeError("cmx_b200: parameter container — the forward pass runs through EncoderDecoder (CUDA engine)")


class _Container(nn.Module):
    forward = _no_forward


def mit_init_(m):
    """Reference `_init_weights` (dual_segformer.py:52-65, net_utils.py:362-374)."""
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)
    elif isinstance(m, nn.Conv2d):
        fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
        fan_out //= m.groups
        m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
        if m.bias is not None:
            m.bias.data.zero_()


class ChannelWeights(_Container):
    def __init__(self, dim, reduction=1):
        super().__init__()
        self.dim = dim
        self.mlp = nn.Sequential(nn.Linear(dim * 4, dim * 4 // reduction), nn.ReLU(inplace=True),
                                 nn.Linear(dim * 4 // reduction, dim * 2), nn.Sigmoid())


class SpatialWeights(_Container):
    def __init__(self, dim, reduction=1):
        super().__init__()
        self.dim = dim
        self.mlp = nn.Sequential(nn.Conv2d(dim * 2, dim // reduction, kernel_size=1), nn.ReLU(inplace=True),
                                 nn.Conv2d(dim // reduction, 2, kernel_size=1), nn.Sigmoid())


class FeatureRectifyModule(_Container):
    def __init__(self, dim, reduction=1, lambda_c=.5, lambda_s=.5):
        super().__init__()
        assert reduction == 1 and lambda_c == .5 and lambda_s == .5, "only the reference defaults are built"
        self.lambda_c, self.lambda_s = lambda_c, lambda_s
        self.channel_weights = ChannelWeights(dim, reduction)
        self.spatial_weights = SpatialWeights(dim, reduction)


class CrossAttention(_Container):
    def __init__(self, dim, num_heads=8):
        super().__init__()
        assert dim % num_heads == 0
        self.dim, self.num_heads = dim, num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.kv1 = nn.Linear(dim, dim * 2, bias=False)
        self.kv2 = nn.Linear(dim, dim * 2, bias=False)


class CrossPath(_Container):
    def __init__(self, dim, reduction=1, num_heads=None):
        super().__init__()
        self.channel_proj1 = nn.Linear(dim, dim // reduction * 2)
        self.channel_proj2 = nn.Linear(dim, dim // reduction * 2)
        self.cross_attn = CrossAttention(dim // reduction, num_heads=num_heads)
        self.end_proj1 = nn.Linear(dim // reduction * 2, dim)
        self.end_proj2 = nn.Linear(dim // reduction * 2, dim)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)


class ChannelEmbed(_Container):
    def __init__(self, in_channels, out_channels, reduction=1, norm_layer=nn.BatchNorm2d):
        super().__init__()
        self.out_channels = out_channels
        mid = out_channels // reduction
        self.residual = nn.Conv2d(in_channels, out_channels, kernel_size=1, bias=False)
        self.channel_embed = nn.Sequential(
            nn.Conv2d(in_channels, mid, kernel_size=1, bias=True),
            nn.Conv2d(mid, mid, kernel_size=3, stride=1, padding=1, bias=True, groups=mid),
            nn.ReLU(inplace=True),
            nn.Conv2d(mid, out_channels, kernel_size=1, bias=True),
            norm_layer(out_channels))
        self.norm = norm_layer(out_channels)


class FeatureFusionModule(_Container):
    def __init__(self, dim, reduction=1, num_heads=None, norm_layer=nn.BatchNorm2d):
        super().__init__()
        self.cross = CrossPath(dim=dim, reduction=reduction, num_heads=num_heads)
        self.channel_emb = ChannelEmbed(in_channels=dim * 2, out_channels=dim, reduction=reduction, norm_layer=norm_layer)
        self.apply(mit_init_)
