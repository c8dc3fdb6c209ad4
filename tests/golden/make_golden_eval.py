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


# ---- the same driver through the REAL reference model (fp32, CPU): prediction map + the summed score map, so that a bf16
# implementation can be held to "equal wherever the reference's top-2 scores are separated"
import torch.nn as nn  # noqa: E402
from oracle import cmx_ref  # noqa: E402
from oracle.synth import synth_state_dict  # noqa: E402
sys.path.insert(0, ROOT)
from baseline import ref_loader  # noqa: E402

REAL = dict(crop=(64, 64), stride_rate=2 / 3, scales=[0.75, 1.0, 1.5], flip=True, hw=(96, 128), ncls=5, seed=7)
rng = np.random.default_rng(REAL["seed"])
img = rng.integers(0, 256, (*REAL["hw"], 3), dtype=np.uint8)
# smooth-ish thermal image replicated to 3 channels like RGBXDataset.py:57-59
import cv2  # noqa: E402
grey = cv2.resize(rng.integers(0, 256, (12, 16), dtype=np.uint8), (REAL["hw"][1], REAL["hw"][0]), interpolation=cv2.INTER_LINEAR)
mx = cv2.merge([grey, grey, grey])
model = ref_loader.build_model("mit_b0", REAL["ncls"], None, nn.BatchNorm2d, embed_dim=256)
model.load_state_dict(synth_state_dict(cmx_ref.MIT_SPECS["mit_b0"], REAL["ncls"], seed=0, embed_dim=256), strict=True)
model.eval()
ev = object.__new__(Evaluator)
ev.class_num, ev.multi_scales, ev.is_flip = REAL["ncls"], REAL["scales"], REAL["flip"]
ev.norm_mean, ev.norm_std = np.array([0.485, 0.456, 0.406]), np.array([0.229, 0.224, 0.225])


class _OnCpu:   # the reference calls val_func.to(input.get_device()) / .eval(): keep the unmodified model on the CPU
    def to(self, *a, **k): return self
    def eval(self): return self
    def __call__(self, a, b): return model(a, b)


ev.val_func = _OnCpu()
total = np.zeros((*REAL["hw"], REAL["ncls"]))
orig = ev.scale_process_rgbX


def capture(*a, **k):
    r = orig(*a, **k)
    total[...] += r
    return r


ev.scale_process_rgbX = capture
pred = ev.sliding_eval_rgbX(img, mx, REAL["crop"], REAL["stride_rate"], None)
assert np.array_equal(pred, total.argmax(2))
np.savez_compressed(os.path.join(HERE, "sliding_eval_real_model.npz"), img=img, mx=mx, pred=pred.astype(np.int64),
                    score=total.astype(np.float32), crop=np.array(REAL["crop"]), stride_rate=np.float64(REAL["stride_rate"]),
                    scales=np.array(REAL["scales"]), flip=np.int64(REAL["flip"]), ncls=np.int64(REAL["ncls"]))
top2 = np.sort(total, axis=2)[:, :, -2:]
print("real model:", pred.shape, np.bincount(pred.ravel(), minlength=REAL["ncls"]), "median rel top-2 gap",
      np.median((top2[:, :, 1] - top2[:, :, 0]) / top2[:, :, 1]))
