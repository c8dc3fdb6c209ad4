"""debug: localise the error of the FRM spatial-gate bias gradient (b2_small case): kernel vs fp64 torch on the SAME
saved tensors vs the fp32 oracle."""
import os, sys
import numpy as np, torch, torch.nn as nn
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
# record the oracle's FRM inputs / output gradients (token-major) for the cross experiments below
_orc = {}
_frm = cmx_ref.frm
def _frm_hook(sd_, p_, x1, x2):
    o1, o2 = _frm(sd_, p_, x1, x2)
    o1.retain_grad(); o2.retain_grad()
    _orc[p_] = (x1, x2, o1, o2)
    return o1, o2
cmx_ref.frm = _frm_hook
params = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var"))) for k, v in sd.items()}
cmx_ref.forward(params, spec, rgb, x, gt, training=True, decoder_bn_eps=1e-3).backward()
orig = E.Engine.frm_bwd
def patched(self, c, dr1, dr2, B, HW):
    C = self.dims[c.s]
    a = c.cat12.double(); a1, a2 = a[:, :C], a[:, C:]
    sw = c.sw.double()
    ds0 = 0.5 * (dr2.double() * a1).sum(1) * sw[:, 0] * (1 - sw[:, 0])
    ds1 = 0.5 * (dr1.double() * a2).sum(1) * sw[:, 1] * (1 - sw[:, 1])
    ref_same = torch.stack([ds0.sum(), ds1.sum()])
    g = self.G(c.p + ".spatial_weights.mlp.2.bias")
    before = g.clone()
    out = orig(self, c, dr1, dr2, B, HW)
    torch.cuda.synchronize()
    kern = (g - before).double()
    orc = params[c.p + ".spatial_weights.mlp.2.bias"].grad.double()
    tm = lambda t: t.detach().permute(0, 2, 3, 1).reshape(-1, t.shape[1]).double().cuda()   # noqa: E731
    ox1, ox2, oo1, oo2 = _orc[c.p]
    oa1, oa2, og1, og2 = tm(ox1), tm(ox2), tm(oo1.grad), tm(oo2.grad)
    def dsum(g1, g2, b1, b2):
        return torch.stack([(0.5 * (g2 * b1).sum(1) * sw[:, 0] * (1 - sw[:, 0])).sum(), (0.5 * (g1 * b2).sum(1) * sw[:, 1] * (1 - sw[:, 1])).sum()])
    rel = lambda a_, b_: ((a_ - b_).norm() / b_.norm()).item()   # noqa: E731
    print("   upstream gradient dr vs oracle: rel-L2 %.3e / %.3e ; inputs a vs oracle: %.3e / %.3e" % (
        rel(dr1.double(), og1), rel(dr2.double(), og2), rel(a1, oa1), rel(a2, oa2)))
    print("   ds sums with (our dr, oracle a) %s | (oracle dr, our a) %s | (oracle dr, oracle a, our sw) %s" % (
        dsum(dr1.double(), dr2.double(), oa1, oa2).cpu().numpy().round(6), dsum(og1, og2, a1, a2).cpu().numpy().round(6),
        dsum(og1, og2, oa1, oa2).cpu().numpy().round(6)))
    print("stage %d: kernel %s | fp64 on same tensors %s | oracle %s | sum|ds| %.3e" % (
        c.s, kern.cpu().numpy().round(6), ref_same.cpu().numpy().round(6), orc.numpy().round(6), (ds0.abs().sum() + ds1.abs().sum()).item()))
    return out
E.Engine.frm_bwd = patched
m = EncoderDecoder(Cfg, nn.CrossEntropyLoss(reduction="mean", ignore_index=255), nn.BatchNorm2d)
m.load_state_dict(sd, strict=True); m.cuda().train(); m._eng().stochastic = False
m.use_cuda_graph = False
m(rgb.cuda(), x.cuda(), gt.cuda()).backward()
