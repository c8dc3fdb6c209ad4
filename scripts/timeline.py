"""Kernel timeline of graph-replayed training steps via torch.profiler (CUPTI): analysis only, never a bench number.
Writes gpurun_out/timeline.csv: name,stream,start_us,dur_us for every kernel of the profiled steps."""
import os
import sys

import torch
import torch.nn as nn
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = EncoderDecoder(bench.Cfg, nn.CrossEntropyLoss(reduction="mean", ignore_index=255), nn.BatchNorm2d).to(dev).train()
opt = torch.optim.AdamW(m.parameters(), lr=6e-5, fused=True)
rgb, x, gt = bench.synth_batch(bench.PER_GPU_BATCH, 1, device=dev)


def step():
    loss = m(rgb, x, gt)
    opt.zero_grad()
    loss.backward()
    opt.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
print("activities:", torch.profiler.supported_activities())
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
out = os.path.join("gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "timeline.csv")
n = 0
with open(out, "w") as f:
    f.write("name,stream,start_us,dur_us\n")
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            tr = e.time_range
            f.write('"%s",%s,%.3f,%.3f\n' % (e.name.replace('"', "'")[:120], getattr(e, "device_index", 0), tr.start, tr.end - tr.start))
            n += 1
print("wrote", n, "kernel records to", out)
try:
    prof.export_chrome_trace("gpurun_out/timeline_trace.json")
except Exception as ex:  # noqa: BLE001
    print("chrome trace export failed:", ex)
