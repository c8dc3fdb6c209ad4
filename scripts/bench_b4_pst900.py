"""BASELINE.json configs[3] on one GPU: CMX MiT-B4, PST900 native shape (720x1280, 5 classes), bf16 training step
(fwd+bwd+AdamW) and eval forward.  Nkv = 880/920 > 320, so self-attention takes the unfused path (batched tcgen05
GEMMs + row-softmax kernels).  Usage: python scripts/bench_b4_pst900.py [batch]"""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder  # noqa: E402
from rgbx_semantic_segmentation_b200.optim import FlatAdamW  # noqa: E402


class Cfg(bench.Cfg):
    backbone = "mit_b4"
    num_classes = 5


B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
H, W = 720, 1280
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = EncoderDecoder(Cfg, nn.CrossEntropyLoss(reduction="mean", ignore_index=255), nn.BatchNorm2d).to(dev).train()
opt = FlatAdamW(bench.group_weight(m, 6e-5), lr=6e-5, weight_decay=0.01)
g = torch.Generator().manual_seed(1)
rgb = torch.randn(B, 3, H, W, generator=g).to(dev)
x = torch.randn(B, 3, H, W, generator=g).to(dev)
gt = torch.randint(0, 5, (B, H, W), generator=g).to(dev)


def step():
    loss = m(rgb, x, gt)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss


for _ in range(4):
    loss = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print("MiT-B4 720x1280 batch %d: train %.1f ms/step = %.2f img/s, loss %.4f, peak memory %.1f GB" % (
    B, ms, B / (ms * 1e-3), loss.item(), torch.cuda.max_memory_allocated() / 2 ** 30), flush=True)
m.eval()
with torch.no_grad():
    for _ in range(4):
        m(rgb[:1], x[:1])
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        m(rgb[:1], x[:1])
    e1.record()
    torch.cuda.synchronize()
print("MiT-B4 720x1280 eval forward batch 1: %.2f ms" % (e0.elapsed_time(e1) / n), flush=True)
