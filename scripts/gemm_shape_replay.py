"""Per-shape timing of every kernel class of one training step (bench workload, one GPU): the launches of one eager step are
recorded with their problem sizes (CMX_PROFILE_SHAPES=1), then every (class, shape) group is captured into ONE CUDA graph (5
repetitions of the group's launches, one stream) and the graph is replayed between one pair of CUDA events - launch durations as
inside the captured step: no per-launch event records and no host enqueue time (a ctypes call + tensor-map encode costs the host
5-10 us, more than the small kernels run).
Output: one line per group, sorted by total time: launches, avg us, algorithmic MB, GB/s, TFLOP/s, fraction of the HBM peak."""
import json
import os
import sys

os.environ["CMX_PROFILE_SHAPES"] = "1"
os.environ["CMX_FFM_STREAM"] = "0"   # everything on one stream: the recorded stream handle is the capture stream
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rgbx_semantic_segmentation_b200 import ops  # noqa: E402
from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
peak = 6554.2
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
m = EncoderDecoder(bench.Cfg, nn.CrossEntropyLoss(reduction="mean", ignore_index=255), nn.BatchNorm2d).to(dev).train()
m.use_cuda_graph = False
rgb, x, gt = bench.synth_batch(bench.PER_GPU_BATCH, 1, device=dev)
eng = m._eng()
eng.wgrad_stream = False
s1 = torch.cuda.Stream()
torch.cuda.set_stream(s1)
for _ in range(2):
    m.zero_grad(set_to_none=True)
    m(rgb, x, gt).backward()
torch.cuda.synchronize()
eng.wgrad_stream = False
eng.keepalive = []
ops.RECORD, ops.PROFILE = [], []
m.zero_grad(set_to_none=True)
m(rgb, x, gt).backward()
torch.cuda.synchronize()
rec, prof = ops.RECORD, ops.PROFILE
ops.RECORD = ops.PROFILE = None
assert len(rec) == len(prof)
groups = {}
for (tag, fn, cargs), (_, _, _, fl, nb) in zip(rec, prof):
    g = groups.setdefault(tag, dict(calls=[], fl=0, nb=0))
    g["calls"].append((fn, cargs)); g["fl"] += fl; g["nb"] += nb
rows = []
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for tag, g in groups.items():
    cl = g["calls"]
    for fn, a in cl:
        fn(*a)
    torch.cuda.synchronize()
    reps = 5
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s1):
        for _ in range(reps):
            for fn, a in cl:
                fn(*a)
    gr.replay()
    flush.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gr.replay()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / (reps * len(cl))
    del gr
    nb, fl = g["nb"] / len(cl), g["fl"] / len(cl)
    rows.append((us * len(cl), tag, len(cl), us, nb, fl))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("# %d library calls, %.2f ms summed graph-replayed kernel time per step; HBM peak %.0f GB/s" % (sum(r[2] for r in rows), tot * 1e-3, peak))
print("%-58s %4s %9s %9s %8s %8s %6s %6s" % ("class_shape", "n", "total_us", "avg_us", "MB", "GB/s", "TF/s", "frac"))
for t, tag, n, us, nb, fl in rows:
    gbps = nb / us * 1e-3 if us > 0 else 0
    print("%-58s %4d %9.1f %9.2f %8.2f %8.0f %6.1f %6.3f" % (tag[:58], n, t, us, nb * 1e-6, gbps, fl / us * 1e-6, gbps / peak))
