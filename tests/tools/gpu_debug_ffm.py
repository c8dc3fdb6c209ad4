"""debug: where does the ~9 % relative error of the gradient that reaches the FRM outputs come from?  Compares, per stage,
the decoder's output gradient (d fused feature) and the FFM input gradients of the engine with the fp32 oracle's."""
import os, sys
import torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cmx_ref
from oracle.synth import synth_inputs, synth_state_dict
from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder
from rgbx_semantic_segmentation_b200 import engine as E
class Cfg:
    decoder = "MLPDecoder"; decoder_embed_dim = 512; pretrained_model = None; bn_eps = 1e-3; bn_momentum = 0.1
    backbone = "mit_b2"; num_classes = 9
spec = cmx_ref.MIT_SPECS["mit_b2"]
sd = synth_state_dict(spec, 9, seed=0)
rgb, x, gt = synth_inputs(2, 64, 96, 9, seed=1)
orc = {}
_ffm, _frm = cmx_ref.ffm, cmx_ref.frm
def ffm_hook(sd_, p_, x1, x2, heads, training, new_stats):
    x1.retain_grad(); x2.retain_grad()
    out = _ffm(sd_, p_, x1, x2, heads, training, new_stats)
    out.retain_grad()
    orc[p_] = (x1, x2, out)
    return out
cmx_ref.ffm = ffm_hook
params = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var"))) for k, v in sd.items()}
cmx_ref.forward(params, spec, rgb, x, gt, training=True, decoder_bn_eps=1e-3).backward()
tm = lambda t: t.detach().permute(0, 2, 3, 1).reshape(-1, t.shape[1]).double().cuda()   # noqa: E731
rel = lambda a_, b_: ((a_.double() - b_).norm() / b_.norm()).item()   # noqa: E731
cosf = lambda a_, b_: (a_.double().flatten() @ b_.flatten() / (a_.double().norm() * b_.norm())).item()   # noqa: E731
orig = E.Engine.ffm_bwd
def patched(self, c, dout, B):
    o1, o2, oo = orc[c.p]
    dr = orig(self, c, dout, B)
    torch.cuda.synchronize()
    g1, g2 = tm(o1.grad), tm(o2.grad)
    # for stages < 3 the oracle's x.grad also holds the next stage's contribution; only stage 3 isolates the FFM
    print("stage %d: d(fused) rel-L2 %.3e cos %.5f | FFM dx1 rel %.3e cos %.5f, dx2 rel %.3e cos %.5f%s" % (
        c.s, rel(dout, tm(oo.grad)), cosf(dout, tm(oo.grad)), rel(dr[0], g1), cosf(dr[0], g1), rel(dr[1], g2), cosf(dr[1], g2),
        "" if c.s == 3 else "   (oracle side includes the next stage's gradient)"))
    return dr
E.Engine.ffm_bwd = patched
m = EncoderDecoder(Cfg, nn.CrossEntropyLoss(reduction="mean", ignore_index=255), nn.BatchNorm2d)
m.load_state_dict(sd, strict=True); m.cuda().train(); m._eng().stochastic = False
m.use_cuda_graph = False
m(rgb.cuda(), x.cuda(), gt.cuda()).backward()
