"""Soak: N graph-replayed training steps at the bench shape with DropPath / Dropout2d active, FlatAdamW, fixed batch:
the loss must fall monotonically-ish, stay finite, and memory must not grow.  python scripts/soak.py [steps]"""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder  # noqa: E402
from rgbx_semantic_segmentation_b200.optim import FlatAdamW  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = EncoderDecoder(bench.Cfg, nn.CrossEntropyLoss(reduction="mean", ignore_index=255), nn.BatchNorm2d).to(dev).train()
opt = FlatAdamW(bench.group_weight(m, 6e-5), lr=6e-5, betas=(0.9, 0.999), weight_decay=0.01)
rgb, x, gt = bench.synth_batch(bench.PER_GPU_BATCH, 1, device=dev)
# learnable labels: a function of the input at the model's resolution (signs of 16x16 block means of three channels)
import torch.nn.functional as F  # noqa: E402
pool = lambda t: F.interpolate((F.avg_pool2d(t[:, None], 16) > 0).float(), scale_factor=16, mode="nearest")[:, 0].long()  # noqa: E731
gt = (pool(rgb[:, 0]) + 2 * pool(x[:, 0]) + 4 * pool(rgb[:, 1])).clamp_(0, bench.NCLS - 1)
losses, mem = [], []
for i in range(steps):
    loss = m(rgb, x, gt)
    opt.zero_grad()
    loss.backward()
    opt.step()
    if i % 25 == 0 or i == steps - 1:
        losses.append(loss.item())
        mem.append(torch.cuda.memory_allocated() / 2 ** 30)
        print("step %4d  loss %.4f  allocated %.2f GB" % (i, losses[-1], mem[-1]), flush=True)
assert all(l == l and abs(l) < 1e4 for l in losses), "non-finite loss"
assert losses[-1] < 0.5 * losses[0], "loss did not fall: %s" % losses
assert mem[-1] <= mem[1] + 0.05, "memory grows: %s" % mem
print("soak ok: loss %.3f -> %.3f over %d steps" % (losses[0], losses[-1], steps))
