"""GPU debugging aid: layer-by-layer comparison of the CUDA engine against the fp32 oracle (same weights,
same inputs), then loss / gradient comparison.  Usage: python tests/tools/gpu_debug_model.py [case ...]"""
import os
import sys
import time

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cmx_ref  # noqa: E402
from oracle.synth import synth_inputs, synth_state_dict  # noqa: E402
from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder  # noqa: E402

CASES = {"b2_small": ("mit_b2", 9, 2, 64, 96), "b0_odd": ("mit_b0", 5, 1, 96, 160), "b4_small": ("mit_b4", 5, 1, 64, 64),
         "b2_mfnet": ("mit_b2", 9, 1, 480, 640), "b2_pst": ("mit_b2", 5, 1, 360, 640)}


class Cfg:
    decoder = "MLPDecoder"
    decoder_embed_dim = 512
    pretrained_model = None
    bn_eps = 1e-3
    bn_momentum = 0.1


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def run(name):
    backbone, ncls, B, H, W = CASES[name]
    spec = cmx_ref.MIT_SPECS[backbone]
    cfg = Cfg()
    cfg.backbone, cfg.num_classes = backbone, ncls
    sd = synth_state_dict(spec, ncls, seed=0)
    rgb, x, gt = synth_inputs(B, H, W, ncls, seed=1)
    print(f"===== {name}: {backbone} B={B} {H}x{W} ncls={ncls}")
    # ---------------- eval logits, traced
    m = EncoderDecoder(cfg, None, nn.BatchNorm2d)
    m.load_state_dict(sd, strict=True)
    m.cuda().eval()
    m.use_cuda_graph = False
    m._eng().trace = {}
    t0 = time.time()
    out = m(rgb.cuda(), x.cuda())
    torch.cuda.synchronize()
    print("  eval forward ok in %.2fs, logits %s" % (time.time() - t0, tuple(out.shape)))
    tr_ref = {}
    with torch.no_grad():
        ref = cmx_ref.forward(sd, spec, rgb, x, training=False, decoder_bn_eps=1e-5, trace=tr_ref)
    tr = m._eng().trace
    for k in tr_ref:
        if k in tr:
            e = rel(tr[k], tr_ref[k])
            flag = "  <<<<<<" if e > 3e-2 else ""
            print("   %-40s rel-L2 %.3e%s" % (k, e, flag))
    m._eng().trace = None
    d = (out.cpu() - ref)
    print("  LOGITS: rel-L2 %.3e  max-abs %.4f  mean-abs %.4f  (ref std %.3f, absmax %.2f)  argmax agree %.4f" % (
        rel(out, ref), d.abs().max().item(), d.abs().mean().item(), ref.std().item(), ref.abs().max().item(),
        (out.cpu().argmax(1) == ref.argmax(1)).float().mean().item()))
    # ---------------- train loss + grads (stochastic ops off)
    m = EncoderDecoder(cfg, nn.CrossEntropyLoss(reduction="mean", ignore_index=255), nn.BatchNorm2d)
    m.load_state_dict(sd, strict=True)
    m.cuda().train()
    m.use_cuda_graph = False
    m._eng().stochastic = False
    loss = m(rgb.cuda(), x.cuda(), gt.cuda())
    loss.backward()
    torch.cuda.synchronize()
    params = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var")))
              for k, v in sd.items()}
    new_stats = {}
    lref = cmx_ref.forward(params, spec, rgb, x, gt, training=True, decoder_bn_eps=1e-3, new_stats=new_stats)
    lref.backward()
    print("  LOSS ours %.6f  oracle %.6f" % (loss.item(), lref.item()))
    rows = []
    for n, p in m.named_parameters():
        g, gr = p.grad.double().cpu().flatten(), params[n].grad.double().flatten()
        cos = (g @ gr / (g.norm() * gr.norm() + 1e-30)).item()
        rows.append((cos, n, g.norm().item(), gr.norm().item()))
    rows.sort()
    print("  GRADS: worst 12 by cosine (cos, name, |ours|, |oracle|)")
    for r in rows[:12]:
        print("    %.5f  %-58s %.4e %.4e" % r)
    import statistics
    print("  GRADS: median cos %.5f ; #cos<0.99: %d / %d ; max norm ratio dev %.3f" % (
        statistics.median(r[0] for r in rows), sum(r[0] < 0.99 for r in rows), len(rows),
        max(abs(r[2] / (r[3] + 1e-30) - 1) for r in rows if r[3] > 1e-8)))
    bufs = dict(m.named_buffers())
    for k in ("backbone.FFMs.0.channel_emb.channel_embed.4.running_mean", "decode_head.linear_fuse.1.running_var"):
        print("   buffer %-60s rel %.3e" % (k, rel(bufs[k], new_stats[k])))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    for c in (sys.argv[1:] or ["b2_small", "b0_odd"]):
        run(c)
