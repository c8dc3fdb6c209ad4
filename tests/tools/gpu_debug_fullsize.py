"""GPU debugging aid: gradient parity of the whole training step against the fp32 oracle at growing sizes.
Prints every parameter whose gradient cosine falls below 0.98.  Usage: python tests/tools/gpu_debug_fullsize.py [HxW ...]"""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cmx_ref  # noqa: E402
from oracle.synth import synth_inputs, synth_state_dict  # noqa: E402
from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder  # noqa: E402


class Cfg:
    decoder = "MLPDecoder"
    decoder_embed_dim = 512
    pretrained_model = None
    bn_eps = 1e-3
    bn_momentum = 0.1
    backbone = "mit_b2"
    num_classes = 9


def run(B, H, W, reps=2):
    spec = cmx_ref.MIT_SPECS["mit_b2"]
    sd = synth_state_dict(spec, 9, seed=0, ctx_gain=float(os.environ.get("CTX_GAIN", "0.1")))
    rgb, x, gt = synth_inputs(B, H, W, 9, seed=1)
    params = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var")))
              for k, v in sd.items()}
    torch.set_num_threads(os.cpu_count() or 1)
    ref = cmx_ref.forward(params, spec, rgb, x, gt, training=True, decoder_bn_eps=1e-3)
    ref.backward()
    gmax = max(p.grad.norm().item() for p in params.values() if p.requires_grad and p.grad is not None)
    prev = None
    for rep in range(reps):
        m = EncoderDecoder(Cfg(), nn.CrossEntropyLoss(reduction="mean", ignore_index=255), nn.BatchNorm2d)
        m.load_state_dict(sd, strict=True)
        m.cuda().train()
        m.use_cuda_graph = False
        m._eng().stochastic = False
        loss = m(rgb.cuda(), x.cuda(), gt.cuda())
        loss.backward()
        print("== B=%d %dx%d rep %d: loss %.6f (oracle %.6f)" % (B, H, W, rep, loss.item(), ref.item()))
        cur = {}
        for n, p in m.named_parameters():
            g, gr = p.grad.double().cpu().flatten(), params[n].grad.double().flatten()
            cur[n] = g
            if gr.norm().item() < 1e-6 * gmax:
                continue
            cos = (g @ gr / (g.norm() * gr.norm())).item()
            extra = ""
            if prev is not None:
                extra = " | run-to-run rel %.3e" % ((g - prev[n]).norm() / (g.norm() + 1e-30)).item()
            if cos < 0.98 or abs(g.norm().item() / gr.norm().item() - 1) > 0.1:
                print("   %-58s cos %.4f  |g|/|ref| %.3f  |ref|/gmax %.2e%s" % (n, cos, g.norm().item() / gr.norm().item(),
                                                                             gr.norm().item() / gmax, extra))
        prev = cur


if __name__ == "__main__":
    sizes = sys.argv[1:] or ["64x96", "128x160", "240x320", "480x640"]
    for s in sizes:
        h, w = s.split("x")
        run(2, int(h), int(w))
