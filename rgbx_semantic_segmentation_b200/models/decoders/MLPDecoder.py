"""All-MLP decoder head containers — same tree as the reference models/decoders/MLPDecoder.py:8-81."""
import torch.nn as nn

from ..net_utils import _Container


class MLP(_Container):
    def __init__(self, input_dim=2048, embed_dim=768):
        super().__init__()
        self.proj = nn.Linear(input_dim, embed_dim)


class DecoderHead(_Container):
    def __init__(self, in_channels=(64, 128, 320, 512), num_classes=40, dropout_ratio=0.1, norm_layer=nn.BatchNorm2d,
                 embed_dim=768, align_corners=False):
        super().__init__()
        assert not align_corners, "reference default only"
        self.num_classes, self.dropout_ratio, self.align_corners = num_classes, dropout_ratio, align_corners
        self.in_channels = list(in_channels)
        self.dropout = nn.Dropout2d(dropout_ratio) if dropout_ratio > 0 else None
        c1, c2, c3, c4 = self.in_channels
        self.linear_c4 = MLP(c4, embed_dim)
        self.linear_c3 = MLP(c3, embed_dim)
        self.linear_c2 = MLP(c2, embed_dim)
        self.linear_c1 = MLP(c1, embed_dim)
        self.linear_fuse = nn.Sequential(nn.Conv2d(embed_dim * 4, embed_dim, kernel_size=1), norm_layer(embed_dim),
                                         nn.ReLU(inplace=True))
        self.linear_pred = nn.Conv2d(embed_dim, num_classes, kernel_size=1)
