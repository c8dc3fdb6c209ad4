"""One eager (no CUDA graph) training step of the bench workload for ncu: 2 untimed warm-up steps, then `--steps` steps.
Prints the number of library launches per step so that `ncu -s` can skip the warm-up."""
import argparse
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rgbx_semantic_segmentation_b200 import ops  # noqa: E402
from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--batch", type=int, default=bench.PER_GPU_BATCH)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = EncoderDecoder(bench.Cfg, nn.CrossEntropyLoss(reduction="mean", ignore_index=255), nn.BatchNorm2d).to(dev).train()
m.use_cuda_graph = False
rgb, x, gt = bench.synth_batch(a.batch, 1, device=dev)
for i in range(2 + a.steps):
    n0 = ops.launch_count()
    m.zero_grad(set_to_none=True)   # like optimizer.zero_grad(): without it autograd ADDS into every .grad (one torch kernel per parameter)
    loss = m(rgb, x, gt)
    loss.backward()
    torch.cuda.synchronize()
    print("step %d: loss %.4f, %d launches" % (i, loss.item(), ops.launch_count() - n0), flush=True)
