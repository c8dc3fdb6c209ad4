"""debug: gradient norm ratios of the smallest tensors of the b0_odd golden case (several runs)"""
import os, sys
import numpy as np, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cmx_ref
from oracle.synth import synth_inputs, synth_state_dict
from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder
z = np.load("tests/golden/b0_odd.npz")
backbone, ncls, B, H, W, sub, stoch = z["meta"]
class Cfg:
    decoder = "MLPDecoder"; decoder_embed_dim = 512; pretrained_model = None; bn_eps = 1e-3; bn_momentum = 0.1
    backbone = "mit_b0"; num_classes = int(ncls)
spec = cmx_ref.MIT_SPECS["mit_b0"]
sd = synth_state_dict(spec, int(ncls), seed=0)
rgb, x, gt = synth_inputs(int(B), int(H), int(W), int(ncls), seed=1)
names = [str(n) for n in z["grad_names"]]; norms = dict(zip(names, z["grad_norms"]))
for rep in range(3):
    m = EncoderDecoder(Cfg, nn.CrossEntropyLoss(reduction="mean", ignore_index=255), nn.BatchNorm2d)
    m.load_state_dict(sd, strict=True); m.cuda().train(); m._eng().stochastic = False
    m(rgb.cuda(), x.cuda(), gt.cuda()).backward()
    pd = dict(m.named_parameters())
    out = []
    for n in names:
        if "FRMs.3" in n or "linear_c" in n or "linear_fuse.0" in n:
            out.append("%s %.3f" % (n.split("backbone.")[-1], pd[n].grad.double().norm().item() / max(norms[n], 1e-30)))
    print(" | ".join(out))
