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

from rgbx_semantic_segmentation_b200.utils.sliding_eval import SlidingEvalContext, sliding_eval_rgbX_batched  # noqa: E402


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


def test_device_driver_has_no_cpu_path():
    """the device-resident driver runs on the library's kernels only (tests/test_eval_gpu.py); a CPU device raises loudly"""
    from rgbx_semantic_segmentation_b200.utils.sliding_eval import sliding_eval_rgbX_gpu
    case = CASES["whole_image_noflip"]
    img, mx = make_inputs(case)
    with pytest.raises(RuntimeError):
        sliding_eval_rgbX_gpu(_ctx(case), img, mx, case["crop"], case["stride_rate"], device="cpu")
