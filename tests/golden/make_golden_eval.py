#!/usr/bin/env python
"""Golden fixture for the batched sliding-window driver, produced by the REAL reference
`Evaluator.sliding_eval_rgbX` (engine/evaluator.py:306-395) with a deterministic stub network on the CPU
(`Tensor.cuda` patched to a no-op; the stub is elementwise + cumsum, hence bit-exact for any batch size).

    python tests/golden/make_golden_eval.py     # build container only (needs /root/reference)
"""
import collections
import collections.abc
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CMX_REFERENCE", "/root/reference")
collections.Iterable = collections.abc.Iterable
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, REF)
os.chdir(tempfile.mkdtemp())

from engine.evaluator import Evaluator  # noqa: E402  (reference)
sys.path.insert(0, HERE)
from eval_stub import StubNet, CASES, make_inputs  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self   # the reference hard-codes .cuda(device)


class _NoDev:
    def __init__(self, *a): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False


torch.cuda.device = _NoDev
torch.Tensor.get_device = lambda self: -1

out = {}
for name, case in CASES.items():
    img, mx = make_inputs(case)
    ev = object.__new__(Evaluator)
    ev.class_num, ev.multi_scales, ev.is_flip = case["ncls"], case["scales"], case["flip"]
    ev.norm_mean, ev.norm_std = np.array([0.485, 0.456, 0.406]), np.array([0.229, 0.224, 0.225])
    ev.val_func = StubNet(case["ncls"])
    pred = ev.sliding_eval_rgbX(img, mx, case["crop"], case["stride_rate"], None)
    out[name] = pred.astype(np.int64)
    print(name, pred.shape, np.bincount(pred.ravel(), minlength=case["ncls"]))
np.savez_compressed(os.path.join(HERE, "sliding_eval.npz"), **out)
