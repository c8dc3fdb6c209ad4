"""Shared by make_golden_eval.py (reference side) and tests/test_sliding_eval_cpu.py (our side): the deterministic stub
network, the cases and their seeded inputs."""
import numpy as np
import torch

CASES = {
    # crop (h, w), stride rate, scales, flip, image (h, w), X channels (3 = HxWx3, 1 = HxW grey handled as 2-D)
    "square_crop_flip": dict(crop=(32, 32), stride_rate=2 / 3, scales=[0.75, 1.0, 1.5], flip=True, hw=(50, 70), xch=3, ncls=5, seed=3),
    "whole_image_noflip": dict(crop=(48, 64), stride_rate=2 / 3, scales=[0.75, 1.0], flip=False, hw=(48, 64), xch=3, ncls=9, seed=4),
    "grey_x_square": dict(crop=(24, 24), stride_rate=0.5, scales=[1.0, 1.25], flip=True, hw=(40, 33), xch=1, ncls=4, seed=5),
}


def make_inputs(case):
    rng = np.random.default_rng(case["seed"])
    h, w = case["hw"]
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    mx = rng.integers(0, 256, (h, w, 3) if case["xch"] == 3 else (h, w), dtype=np.uint8)
    return img, mx


class StubNet:
    """score[t, c] = sum_k wa[c,k] a[t,k] + sum_k wb[c,k] b[t,k] + 0.05 * cumsum_x(a[t,0]) * (c+1)/ncls - elementwise ops,
    a 3-term channel sum and a per-row cumsum: per-sample independent and left/right asymmetric (catches flip errors)"""

    def __init__(self, ncls):
        self.ncls = ncls
        g = torch.Generator().manual_seed(1234)
        self.wa = torch.randn(ncls, 3, generator=g)
        self.wb = torch.randn(ncls, 3, generator=g)

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def __call__(self, a, b):
        a, b = a.float(), b.float()
        if b.shape[1] == 1:
            b = b.expand(-1, 3, -1, -1)
        wa, wb = self.wa.to(a.device), self.wb.to(a.device)
        s = (a[:, None] * wa[None, :, :, None, None]).sum(2) + (b[:, None] * wb[None, :, :, None, None]).sum(2)
        ramp = torch.cumsum(a[:, 0], dim=-1)[:, None] * ((torch.arange(self.ncls, device=a.device).float() + 1) / self.ncls)[None, :, None, None]
        return 0.3 * s + 0.05 * ramp
