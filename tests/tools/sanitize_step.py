"""One small training step (mit_b0, 64x64, batch 2) for compute-sanitizer (one tool per gpurun call)."""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.synth import synth_inputs  # noqa: E402
from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder  # noqa: E402


class Cfg:
    backbone = "mit_b0"; decoder = "MLPDecoder"; decoder_embed_dim = 256; num_classes = 5
    pretrained_model = None; bn_eps = 1e-3; bn_momentum = 0.1


torch.manual_seed(0)
rgb, x, gt = (t.cuda() for t in synth_inputs(2, 64, 64, 5, seed=3))
m = EncoderDecoder(Cfg, nn.CrossEntropyLoss(ignore_index=255), nn.BatchNorm2d).cuda().train()
m.use_cuda_graph = False
loss = m(rgb, x, gt)
loss.backward()
torch.cuda.synchronize()
print("loss", loss.item())
