"""utils/loss_opr.FocalLoss (host mirror) vs golden values produced by the reference's own class
(tests/golden/make_golden_focal.py): loss and gradient for FocalLoss and the CE_Focal combination."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from rgbx_semantic_segmentation_b200.utils.loss_opr import FocalLoss


@pytest.mark.parametrize("name", ["g2", "g4", "g1"])
def test_focal_mirror_matches_reference(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "focal.npz"))
    ncls, gamma, alpha, _ = z[name + "_meta"]
    logits = torch.from_numpy(z[name + "_logits"]).requires_grad_(True)
    target = torch.from_numpy(z[name + "_target"])
    fl = FocalLoss(ignore_label=255, gamma=float(gamma), alpha=float(alpha))
    lf = fl(logits, target)
    gf, = torch.autograd.grad(lf, logits)
    assert abs(lf.item() - float(z[name + "_focal"])) < 1e-5
    assert np.allclose(gf.numpy(), z[name + "_focal_grad"], rtol=1e-4, atol=1e-7)
    lc = nn.CrossEntropyLoss(reduction='mean', ignore_index=255)(logits, target) + 0.2 * fl(logits, target)
    gc, = torch.autograd.grad(lc, logits)
    assert abs(lc.item() - float(z[name + "_cefocal"])) < 1e-5
    assert np.allclose(gc.numpy(), z[name + "_cefocal_grad"], rtol=1e-4, atol=1e-7)
