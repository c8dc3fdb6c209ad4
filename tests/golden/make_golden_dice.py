#!/usr/bin/env python
"""Golden values of the reference's DiceCELoss (utils/loss_opr.py:103-156; train.py:79-80) on seeded logits: loss and
d loss / d logits from the REAL reference class.   python tests/golden/make_golden_dice.py"""
import collections
import collections.abc
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CMX_REFERENCE", "/root/reference")
collections.Iterable = collections.abc.Iterable
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, REF)
os.chdir(tempfile.mkdtemp())
from utils.loss_opr import DiceCELoss  # noqa: E402  (reference)

out = {}
for name, (ncls, alpha, B, h, w, seed) in {"d5": (5, 0.5, 2, 12, 16, 0), "d9": (9, 0.5, 3, 9, 11, 1), "d40": (40, 0.3, 1, 16, 8, 2)}.items():
    g = torch.Generator().manual_seed(seed)
    logits = (2.5 * torch.randn(B, ncls, h, w, generator=g)).requires_grad_(True)
    target = torch.randint(0, ncls, (B, h, w), generator=g)
    target[torch.rand(B, h, w, generator=g) < 0.15] = 255
    crit = DiceCELoss(alpha=alpha, ignore_index=255, reduction='mean')
    loss = crit(logits, target)
    grad, = torch.autograd.grad(loss, logits)
    out.update({name + "_meta": np.array([ncls, alpha, seed], np.float64), name + "_logits": logits.detach().numpy(),
                name + "_target": target.numpy(), name + "_loss": np.float64(loss.item()), name + "_grad": grad.numpy()})
    print(name, loss.item())
np.savez_compressed(os.path.join(HERE, "dice.npz"), **out)
