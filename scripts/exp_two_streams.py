"""Experiment: does stream-level concurrency help?  Two independent models, batch 4 each, graph-replayed training steps on two
streams, against one model at batch 8 (analysis only; not a bench number - the two half-batch models do not share BatchNorm
statistics, so this is an upper bound for a batch-split pipeline)."""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder  # noqa: E402
from rgbx_semantic_segmentation_b200.optim import FlatAdamW  # noqa: E402

dev = torch.device("cuda", 0)


def make(batch):
    torch.manual_seed(0)
    m = EncoderDecoder(bench.Cfg, nn.CrossEntropyLoss(reduction="mean", ignore_index=255), nn.BatchNorm2d).to(dev).train()
    opt = FlatAdamW(bench.group_weight(m, 6e-5), lr=6e-5, betas=(0.9, 0.999), weight_decay=0.01)
    data = bench.synth_batch(batch, 1, device=dev)

    def step():
        loss = m(*data)
        opt.zero_grad()
        loss.backward()
        opt.step()
    return step


def timeit(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


one = make(8)
print("1 x batch 8: %.2f ms/step" % timeit(one))
del one
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
a, b = make(4), make(4)
for _ in range(4):     # warm-up + graph capture, one at a time
    with torch.cuda.stream(sa):
        a()
    torch.cuda.synchronize()
    with torch.cuda.stream(sb):
        b()
    torch.cuda.synchronize()


def both():
    with torch.cuda.stream(sa):
        a()
    with torch.cuda.stream(sb):
        b()


cur = torch.cuda.current_stream()
for _ in range(3):
    both()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
sa.wait_stream(cur); sb.wait_stream(cur)
for _ in range(20):
    both()
cur.wait_stream(sa); cur.wait_stream(sb)
e1.record()
torch.cuda.synchronize()
print("2 x batch 4 on two streams: %.2f ms per pair of steps (8 images)" % (e0.elapsed_time(e1) / 20))
with torch.cuda.stream(sa):
    print("1 x batch 4 alone: %.2f ms/step" % timeit(a))
