"""Uninitialised-read detector: run the same training step twice; before the second run fill the caching
allocator's free memory with a poison pattern.  Any kernel that reads memory it (or a predecessor) did not write
shows up as a gradient difference / NaN, reported per parameter."""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cmx_ref  # noqa: E402
from oracle.synth import synth_inputs, synth_state_dict  # noqa: E402
from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder  # noqa: E402


class Cfg:
    backbone = "mit_b0"; decoder = "MLPDecoder"; decoder_embed_dim = 512; num_classes = 5
    pretrained_model = None; bn_eps = 1e-3; bn_momentum = 0.1


def poison(value):
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    n = min(int(free * 0.5), 8 << 30) // 4
    t = torch.full((n,), value, device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    del t            # stays in the caching allocator; the next step's torch.empty() calls carve it up


def step(m, sd, rgb, x, gt):
    m.load_state_dict(sd, strict=True)
    for p in m.parameters():
        p.grad = None
    loss = m(rgb, x, gt)
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), {n: p.grad.clone() for n, p in m.named_parameters()}


spec = cmx_ref.MIT_SPECS["mit_b0"]
sd = synth_state_dict(spec, 5, seed=0)
rgb, x, gt = (t.cuda() for t in synth_inputs(2, 64, 64, 5, seed=3))
m = EncoderDecoder(Cfg, nn.CrossEntropyLoss(ignore_index=255), nn.BatchNorm2d).cuda().train()
m.use_cuda_graph = False
m._eng().stochastic = False
l0, g0 = step(m, sd, rgb, x, gt)
l1, g1 = step(m, sd, rgb, x, gt)
print("eager vs eager: loss %.7f %.7f" % (l0, l1))
for label, val in (("NaN", float("nan")), ("1e30", 1e30), ("zero", 0.0)):
    poison(val)
    l2, g2 = step(m, sd, rgb, x, gt)
    bad = []
    for n in g0:
        d = (g2[n] - g0[n]).norm().item() / (g0[n].norm().item() + 1e-20)
        if not (d < 1e-4):
            bad.append((d, n))
    print("poison %-5s: loss %.7f, %d/%d parameter grads changed" % (label, l2, len(bad), len(g0)))
    badn = {n for _, n in bad}
    print("     UNCHANGED:", [n for n in g0 if n not in badn][:30])
    for d, n in sorted(bad)[-6:]:
        print("     %-60s rel diff %.3g" % (n, d))

# ---- allocation-site report: NaN-fill every engine buffer at allocation, list the ones that still hold NaNs
eng = m._eng()
eng.poison = []
l3, g3 = step(m, sd, rgb, x, gt)
print("NaN-at-allocation run: loss", l3)
seen = {}
for t, site in eng.poison:
    if t.dtype.is_floating_point:
        n = int(torch.isnan(t).sum())
        if n:
            k = (site, tuple(t.shape), str(t.dtype))
            seen.setdefault(k, [0, 0])
            seen[k][0] += 1
            seen[k][1] += n
for (site, shape, dt), (cnt, n) in seen.items():
    print("   %-28s %-22s %-15s buffers %3d  NaN elements %d" % (site, shape, dt, cnt, n))
nan_grads = [n for n, g in g3.items() if torch.isnan(g).any()]
print("   params with NaN grads: %d" % len(nan_grads), nan_grads[:10])
