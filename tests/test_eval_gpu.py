"""Device-resident sliding-window / multi-scale evaluation driver (utils/sliding_eval.py::sliding_eval_rgbX_gpu)
on the library's own kernels, against fixtures written by the REAL reference `Evaluator.sliding_eval_rgbX`
(engine/evaluator.py:306-431; tests/golden/make_golden_eval.py):
  * the network crops built by cmx_eval_pack_crop are BIT-identical to the reference's process_image_rgbX output;
  * with the deterministic stub network the prediction maps equal the reference's except for ties (see below);
  * through the real model (bf16 kernels vs the reference's fp32 model) the prediction map equals the reference's wherever
    the reference's top-2 summed scores are separated by more than the score tolerance.
Tie rule: the score-map resize evaluates cv2.INTER_LINEAR's positions and weights in fp32 on the device; cv2 may contract
multiply-adds differently (~1 ulp), so a pixel may differ only if its two best summed scores agree to 1e-5 relative."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from eval_stub import CASES, StubNet, make_inputs  # noqa: E402

from oracle import cmx_ref  # noqa: E402
from oracle.synth import synth_state_dict  # noqa: E402

if torch.cuda.is_available():
    from rgbx_semantic_segmentation_b200 import ops
    from rgbx_semantic_segmentation_b200.utils.sliding_eval import (SlidingEvalContext, _pad_margin, sliding_eval_rgbX_batched,
                                                                    sliding_eval_rgbX_gpu)


class CudaStub(StubNet):
    def __init__(self, ncls):
        super().__init__(ncls)
        self.wa, self.wb = self.wa.cuda(), self.wb.cuda()


@pytest.mark.parametrize("grey", [False, True])
@pytest.mark.parametrize("flip", [0, 1])
def test_pack_crop_is_bit_identical_to_process_image_rgbX(grey, flip):
    """whole-image mode (normalise, then zero-pad to the crop) and tile mode (zero-pad the RAW canvas, cut, normalise, pad)"""
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (37, 41, 3), dtype=np.uint8)
    mx = rng.integers(0, 256, (37, 41) if grey else (37, 41, 3), dtype=np.uint8)
    ctx = SlidingEvalContext(None, 5)
    mean, std = [float(v) for v in ctx.norm_mean], [float(v) for v in ctx.norm_std]
    for src, m, s_ in ((img, mean, std), (mx, [0.0] * 3 if grey else mean, [1.0] * 3 if grey else std)):
        d = torch.from_numpy(src).cuda()
        ch = 1 if src.ndim == 2 else 3
        # whole image into a 48 x 64 crop
        a, b, margin = ctx.process_image_rgbX(img, mx, (48, 64))
        ref = a if src is img else b
        out = torch.empty(ch, 48, 64, device="cuda")
        ops.eval_pack_crop(d, (0, 0), (0, 0, 37, 41), (int(margin[0]), int(margin[2])), m, s_, flip, out)
        want = np.ascontiguousarray(ref[:, :, ::-1] if flip else ref).astype(np.float32)
        assert np.array_equal(out.cpu().numpy(), want)
        # a 20 x 24 tile at canvas offset (9, 30) of the image zero-padded by (5, 7) rows / (3, 2) columns, into a 32 x 32 crop
        pad = cv2.copyMakeBorder(img, 5, 7, 3, 2, cv2.BORDER_CONSTANT, value=0)
        padx = cv2.copyMakeBorder(mx, 5, 7, 3, 2, cv2.BORDER_CONSTANT, value=0)
        ta, tb, tm = ctx.process_image_rgbX(pad[9:29, 22:46], padx[9:29, 22:46], (32, 32))
        ref = ta if src is img else tb
        out = torch.empty(ch, 32, 32, device="cuda")
        ops.eval_pack_crop(d, (5, 3), (9, 22, 20, 24), (int(tm[0]), int(tm[2])), m, s_, flip, out)
        want = np.ascontiguousarray(ref[:, :, ::-1] if flip else ref).astype(np.float32)
        assert np.array_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize("name", sorted(CASES))
def test_device_driver_vs_reference_evaluator_golden_stub_network(golden_dir, name):
    case = CASES[name]
    gold = np.load(os.path.join(golden_dir, "sliding_eval.npz"))[name]
    img, mx = make_inputs(case)
    ctx = SlidingEvalContext(CudaStub(case["ncls"]), case["ncls"], case["scales"], case["flip"])
    pred = sliding_eval_rgbX_gpu(ctx, img, mx, case["crop"], case["stride_rate"], device="cuda", max_batch=4)
    assert pred.shape == gold.shape and pred.dtype == gold.dtype
    assert (pred != gold).mean() <= 1e-3, "%d of %d pixels differ" % ((pred != gold).sum(), gold.size)
    # fused confusion matrix == utils/metric.py on the device prediction
    from oracle import metric_ref
    gt = np.random.default_rng(1).integers(0, case["ncls"], gold.shape).astype(np.uint8)
    gt[::7] = 255
    p2, hist, labeled, correct = sliding_eval_rgbX_gpu(ctx, img, mx, case["crop"], case["stride_rate"], device="cuda", gt=gt)
    assert np.array_equal(p2, pred)
    hr, lr, cr = metric_ref.hist_info(case["ncls"], pred, gt)
    assert np.array_equal(hist, hr) and (labeled, correct) == (lr, cr)
    # dataset streaming: device accumulators, no read-back per image; two images == twice the single-image sums
    acc_h = torch.zeros(case["ncls"], case["ncls"], dtype=torch.int64, device="cuda")
    acc_s = torch.zeros(2, dtype=torch.int64, device="cuda")
    for _ in range(2):
        p3 = sliding_eval_rgbX_gpu(ctx, img, mx, case["crop"], case["stride_rate"], device="cuda", gt=gt, accum=(acc_h, acc_s))
    assert p3.is_cuda and np.array_equal(p3.cpu().numpy().astype(np.int64), pred)
    assert np.array_equal(acc_h.cpu().numpy(), 2 * hr) and acc_s.tolist() == [2 * lr, 2 * cr]


def test_device_driver_through_the_real_model_vs_reference_evaluator_golden(golden_dir):
    """fixture: the reference Evaluator driving the reference model in fp32 (prediction map + summed score map)"""
    from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder
    z = np.load(os.path.join(golden_dir, "sliding_eval_real_model.npz"))
    ncls = int(z["ncls"])

    class Cfg:
        backbone, decoder, decoder_embed_dim, num_classes, pretrained_model = "mit_b0", "MLPDecoder", 256, ncls, None
    m = EncoderDecoder(Cfg, None, nn.BatchNorm2d)
    m.load_state_dict(synth_state_dict(cmx_ref.MIT_SPECS["mit_b0"], ncls, seed=0, embed_dim=256), strict=True)
    m = m.cuda().eval()
    ctx = SlidingEvalContext(m, ncls, [float(s) for s in z["scales"]], bool(int(z["flip"])))
    crop = tuple(int(v) for v in z["crop"])
    gold, score = z["pred"], z["score"].astype(np.float64)
    top2 = np.sort(score, axis=2)[:, :, -2:]
    gap = (top2[:, :, 1] - top2[:, :, 0]) / top2[:, :, 1]
    for name, fn in (("device", sliding_eval_rgbX_gpu), ("host-batched", sliding_eval_rgbX_batched)):
        pred = fn(ctx, z["img"], z["mx"], crop, float(z["stride_rate"]), "cuda", max_batch=8)
        assert pred.shape == gold.shape and pred.dtype == np.int64
        diff = pred != gold
        # bf16 logits deviate ~1.5e-2 rel-L2 from fp32; summed exp-scores by a few percent: pixels whose best two reference
        # scores are more than 10 % apart must agree
        assert not diff[gap > 0.10].any(), "%s: %d separated pixels differ" % (name, int(diff[gap > 0.10].sum()))
        assert diff.mean() < 0.02, "%s: %.2f %% of the pixels differ" % (name, 100 * diff.mean())
