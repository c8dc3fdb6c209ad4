"""Loader for the UNMODIFIED reference (ynalcakan/RGBX_Semantic_Segmentation) used by `bench.py --impl reference`, the
`cpu_baseline` leg and the fixture generators - never by the product package.

The reference has no setup.py / pyproject.toml (pip cannot install it), so `install()` copies its source tree verbatim to
the git-ignored `baseline/_ref/` (which gpurun ships to the GPU box, where /root/reference does not exist).  Importing it
needs four import shims for packages that are absent from this image (tests/golden/_shims: timm.models.layers, easydict,
tensorboardX; plus the collections.Iterable alias for utils/transforms.py:13) - none of them touches the hot path's
arithmetic (DropPath / to_2tuple / trunc_normal_ are restated from their public definitions)."""
import collections
import collections.abc
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref")
SHIMS = os.path.join(ROOT, "tests", "golden", "_shims")
SRC = os.environ.get("CMX_REFERENCE", "/root/reference")


def install(force=False):
    """copy the reference tree to baseline/_ref (no-op when the source is absent, e.g. on the GPU box)"""
    if not os.path.isdir(SRC):
        return os.path.isdir(REF_DIR)
    if os.path.isdir(REF_DIR) and not force:
        return True
    if os.path.isdir(REF_DIR):
        shutil.rmtree(REF_DIR)
    shutil.copytree(SRC, REF_DIR, ignore=shutil.ignore_patterns(".git", "*.ipynb", "*.jpg", "__pycache__"))
    return True


def available():
    return os.path.isfile(os.path.join(REF_DIR, "models", "builder.py")) or os.path.isdir(SRC)


def load():
    """-> (config, EncoderDecoder) of the reference; raises ImportError when it is not installed"""
    root = REF_DIR if os.path.isfile(os.path.join(REF_DIR, "models", "builder.py")) else SRC
    if not os.path.isfile(os.path.join(root, "models", "builder.py")):
        raise ImportError("reference not installed (baseline/_ref missing and %s absent)" % SRC)
    collections.Iterable = collections.abc.Iterable
    for p in (SHIMS, root):
        if p not in sys.path:
            sys.path.insert(0, p)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())      # config.py computes cwd-relative log directories at import time
    try:
        from config import config
        from models.builder import EncoderDecoder
    finally:
        os.chdir(cwd)
    config.pretrained_model = None
    return config, EncoderDecoder


def build_model(backbone="mit_b2", num_classes=9, criterion=None, norm_layer=None, embed_dim=512):
    """the reference EncoderDecoder for the CMX MiT + MLPDecoder path.  mit_b4 / mit_b5: the reference builder passes the wrong
    decoder channels for them and crashes, so its own backbone and DecoderHead classes are assembled with
    the channels the backbone really emits - same modules, same forward code."""
    import torch.nn as nn
    config, EncoderDecoder = load()
    norm_layer = norm_layer or nn.BatchNorm2d
    config.num_classes, config.decoder_embed_dim, config.pretrained_model = num_classes, embed_dim, None
    if backbone in ("mit_b4", "mit_b5"):
        config.backbone = "mit_b2"
        m = EncoderDecoder(cfg=config, criterion=criterion, norm_layer=norm_layer)
        from models.decoders.MLPDecoder import DecoderHead
        from models.encoders import dual_segformer
        m.backbone = getattr(dual_segformer, backbone)()
        m.decode_head = DecoderHead(in_channels=[64, 128, 320, 512], num_classes=num_classes, norm_layer=norm_layer, embed_dim=embed_dim)
        if criterion:
            m.init_weights(config, pretrained=None)
        config.backbone = backbone
        return m
    config.backbone = backbone
    return EncoderDecoder(cfg=config, criterion=criterion, norm_layer=norm_layer)
