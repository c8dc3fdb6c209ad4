#!/usr/bin/env python
"""Golden values of the reference's FocalLoss / CE_Focal criteria (utils/loss_opr.py:157-196, builder.py:246-247) on seeded
logits: loss values and d loss / d logits from the REAL reference class.   python tests/golden/make_golden_focal.py"""
import collections
import collections.abc
import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CMX_REFERENCE", "/root/reference")
collections.Iterable = collections.abc.Iterable
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, REF)
os.chdir(tempfile.mkdtemp())
from utils.loss_opr import FocalLoss  # noqa: E402  (reference)

out = {}
for name, (ncls, gamma, alpha, seed) in {"g2": (5, 2.0, 0.25, 0), "g4": (9, 4.0, 0.25, 1), "g1": (4, 1.0, 0.6, 2)}.items():
    g = torch.Generator().manual_seed(seed)
    logits = (2.5 * torch.randn(2, ncls, 12, 16, generator=g)).requires_grad_(True)
    target = torch.randint(0, ncls, (2, 12, 16), generator=g)
    target[torch.rand(2, 12, 16, generator=g) < 0.15] = 255
    fl = FocalLoss(ignore_label=255, gamma=gamma, alpha=alpha, reduction='mean')
    ce = nn.CrossEntropyLoss(reduction='mean', ignore_index=255)
    lf = fl(logits, target)
    gf, = torch.autograd.grad(lf, logits)
    lc = ce(logits, target) + 0.2 * fl(logits, target)       # builder.py:246-247
    gc, = torch.autograd.grad(lc, logits)
    out.update({name + "_meta": np.array([ncls, gamma, alpha, seed], np.float64), name + "_logits": logits.detach().numpy(),
                name + "_target": target.numpy(), name + "_focal": np.float64(lf.item()), name + "_focal_grad": gf.numpy(),
                name + "_cefocal": np.float64(lc.item()), name + "_cefocal_grad": gc.numpy()})
    print(name, lf.item(), lc.item())
np.savez_compressed(os.path.join(HERE, "focal.npz"), **out)
