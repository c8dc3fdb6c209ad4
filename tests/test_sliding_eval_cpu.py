"""Batched sliding-window driver (utils/sliding_eval.py) vs the REAL reference `Evaluator.sliding_eval_rgbX`:
tests/golden/sliding_eval.npz was produced by the reference code itself with the same stub network
(tests/golden/make_golden_eval.py).  Host logic only - runs on the CPU with device="cpu"."""
import os
import sys

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from eval_stub import CASES, StubNet, make_inputs  # noqa: E402

from rgbx_semantic_segmentation_b200.utils.sliding_eval import SlidingEvalContext, sliding_eval_rgbX_batched, sliding_eval_rgbX_gpu  # noqa: E402


def _ctx(case):
    return SlidingEvalContext(StubNet(case["ncls"]), case["ncls"], case["scales"], case["flip"])


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("max_batch", [1, 3, 64])
def test_batched_driver_reproduces_reference_predictions(golden_dir, name, max_batch):
    case = CASES[name]
    gold = np.load(os.path.join(golden_dir, "sliding_eval.npz"))[name]
    img, mx = make_inputs(case)
    pred = sliding_eval_rgbX_batched(_ctx(case), img, mx, case["crop"], case["stride_rate"], device="cpu", max_batch=max_batch)
    assert pred.shape == gold.shape and pred.dtype == gold.dtype
    assert np.array_equal(pred, gold), "%d of %d pixels differ" % ((pred != gold).sum(), gold.size)


@pytest.mark.parametrize("name", sorted(CASES))
def test_device_resident_driver_matches_reference_predictions(golden_dir, name):
    """device preprocessing is bit-identical (float64 normalisation); only the score-map resize differs from cv2 by ~5e-7,
    so at most a handful of tied pixels may flip"""
    case = CASES[name]
    gold = np.load(os.path.join(golden_dir, "sliding_eval.npz"))[name]
    img, mx = make_inputs(case)
    pred = sliding_eval_rgbX_gpu(_ctx(case), img, mx, case["crop"], case["stride_rate"], device="cpu", max_batch=4)
    assert pred.shape == gold.shape and pred.dtype == gold.dtype
    assert (pred != gold).mean() <= 1e-3, "%d of %d pixels differ" % ((pred != gold).sum(), gold.size)


def test_device_normalisation_is_bit_identical_to_the_reference_formula():
    import torch
    from rgbx_semantic_segmentation_b200.utils.sliding_eval import _normalize_dev
    rng = np.random.default_rng(0)
    u8 = rng.integers(0, 256, (37, 41, 3), dtype=np.uint8)
    mean, std = np.array([0.485, 0.456, 0.406]), np.array([0.229, 0.224, 0.225])
    ref = ((u8.astype(np.float64) / 255.0 - mean) / std).astype(np.float32)      # utils/transforms.py:182-187 + evaluator.py:375
    ours = _normalize_dev(torch.from_numpy(u8), torch.from_numpy(mean), torch.from_numpy(std)).numpy()
    assert np.array_equal(ref, ours)
