"""debug: gradients inside the decoder head (low-res logits, post-BN activation, pre-BN fuse tensor) vs the fp32 oracle"""
import os, sys
import torch, torch.nn as nn, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cmx_ref
from oracle.synth import synth_inputs, synth_state_dict
from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder
class Cfg:
    decoder = "MLPDecoder"; decoder_embed_dim = 512; pretrained_model = None; bn_eps = 1e-3; bn_momentum = 0.1
    backbone = "mit_b2"; num_classes = 9
spec = cmx_ref.MIT_SPECS["mit_b2"]
sd = synth_state_dict(spec, 9, seed=0)
rgb, x, gt = synth_inputs(2, 64, 96, 9, seed=1)
keep = {}
def head(sd_, feats, training=False, bn_eps=1e-5, bn_momentum=0.1, new_stats=None, dropout_scale=None, p="decode_head"):
    c1, c2, c3, c4 = feats
    n, size = c4.shape[0], c1.shape[2:]
    def mlp(c, name):
        t = F.linear(c.flatten(2).transpose(1, 2), sd_[f"{p}.{name}.proj.weight"], sd_[f"{p}.{name}.proj.bias"])
        return t.permute(0, 2, 1).reshape(n, -1, c.shape[2], c.shape[3])
    cat = torch.cat([F.interpolate(mlp(c4, "linear_c4"), size=size, mode="bilinear", align_corners=False),
                     F.interpolate(mlp(c3, "linear_c3"), size=size, mode="bilinear", align_corners=False),
                     F.interpolate(mlp(c2, "linear_c2"), size=size, mode="bilinear", align_corners=False), mlp(c1, "linear_c1")], dim=1)
    y0 = F.conv2d(cat, sd_[f"{p}.linear_fuse.0.weight"], sd_[f"{p}.linear_fuse.0.bias"])
    y1 = F.relu(cmx_ref._batch_norm(sd_, f"{p}.linear_fuse.1", y0, training, bn_eps, bn_momentum, new_stats))
    out = F.conv2d(y1, sd_[f"{p}.linear_pred.weight"], sd_[f"{p}.linear_pred.bias"])
    for t in (y0, y1, out):
        t.retain_grad()
    keep.update(fuse=y0, post_bn=y1, logits=out)
    return out
cmx_ref.decoder_head = head
params = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var"))) for k, v in sd.items()}
cmx_ref.forward(params, spec, rgb, x, gt, training=True, decoder_bn_eps=1e-3).backward()
m = EncoderDecoder(Cfg, nn.CrossEntropyLoss(reduction="mean", ignore_index=255), nn.BatchNorm2d)
m.load_state_dict(sd, strict=True); m.cuda().train(); m._eng().stochastic = False
m.use_cuda_graph = False
m._eng().trace = {}
m(rgb.cuda(), x.cuda(), gt.cuda()).backward()
tr = m._eng().trace
tm = lambda t: t.detach().permute(0, 2, 3, 1).reshape(-1, t.shape[1]).double()   # noqa: E731
for k in ("logits", "post_bn", "fuse"):
    mine = tr["grad.decode_head." + k].double().cpu()
    ref = tm(keep[k].grad)
    mine = mine[:, :ref.shape[1]]
    print("%-8s grad: rel-L2 %.3e  cos %.6f  |mine|/|ref| %.4f" % (k, ((mine - ref).norm() / ref.norm()).item(),
          (mine.flatten() @ ref.flatten() / (mine.norm() * ref.norm())).item(), (mine.norm() / ref.norm()).item()))
fwd = tr.get("decode_head.logits")
if fwd is not None:
    ref = tm(keep["logits"])
    print("logits fwd: rel-L2 %.3e" % ((fwd.double().cpu()[:, :ref.shape[1]] - ref).norm() / ref.norm()).item())
