"""Input pipeline fused into the stage-1 patch-embed load: raw uint8 images on the device instead of the
host-normalised fp32 CHW tensors the reference's loader produces (dataloader/dataloader.py:85-112, RGBXDataset.py:57-59,
utils/transforms.py:182-187).
  * 3-channel images: the im2col rows are BIT-identical to normalise-on-host + cmx_im2col_nchw;
  * grey X: the replicated channels are folded into a 2-column-per-tap operand and 7x7x2 weights - the convolution equals the
    3-channel one up to bf16 operand rounding; the folded weight gradient is chained back to the three channel slices;
  * whole model: loss / logits / gradients from uint8 inputs equal those from the reference-style fp32 inputs."""
import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

from oracle import cmx_ref  # noqa: E402
from oracle.synth import synth_state_dict  # noqa: E402

if torch.cuda.is_available():
    from rgbx_semantic_segmentation_b200 import ops
    from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder

MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def host_normalize(u8_hwc):
    """utils/transforms.py:182-187 + the loader's HWC -> CHW and float32 cast"""
    x = u8_hwc.astype(np.float64) / 255.0
    x = (x - np.array(MEAN)) / np.array(STD)
    return np.ascontiguousarray(np.moveaxis(x, -1, -3)).astype(np.float32)


def test_im2col_u8_bit_identical_to_host_pipeline_and_grey_fold():
    rng = np.random.default_rng(0)
    B, H, W = 2, 37, 50
    img = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    Ho, Wo = (H + 6 - 7) // 4 + 1, (W + 6 - 7) // 4 + 1
    ref = torch.empty(B * Ho * Wo, 152, device="cuda", dtype=torch.bfloat16)
    ops.im2col_nchw(torch.from_numpy(host_normalize(img)).cuda(), ref, 7, 4, 3, Ho, Wo)
    got = torch.empty_like(ref)
    ops.im2col_u8(torch.from_numpy(img).cuda(), got, 7, 4, 3, Ho, Wo, MEAN, STD)
    assert torch.equal(got, ref)
    # grey: conv(W, replicate+normalise(v)) == colx @ [Wa | Wb]
    grey = rng.integers(0, 256, (B, H, W), dtype=np.uint8)
    Wt = torch.randn(16, 3, 7, 7, generator=torch.Generator().manual_seed(1))
    x3 = torch.from_numpy(host_normalize(np.stack([grey] * 3, -1)))
    want = torch.nn.functional.conv2d(x3, Wt, stride=4, padding=3).permute(0, 2, 3, 1).reshape(-1, 16)
    colx = torch.empty(B * Ho * Wo, 104, device="cuda", dtype=torch.bfloat16)
    ops.im2col_u8(torch.from_numpy(grey).cuda(), colx, 7, 4, 3, Ho, Wo, MEAN, STD)
    inv = torch.tensor([1 / s for s in STD]).view(1, 3, 1, 1)
    nms = torch.tensor([-m / s for m, s in zip(MEAN, STD)]).view(1, 3, 1, 1)
    fold = torch.stack([(Wt * inv).sum(1), (Wt * nms).sum(1)], -1).reshape(16, 98)
    got = colx[:, :98].float().cpu() @ fold.t()
    assert float((got - want).abs().max()) < 2e-2 * float(want.abs().max())
    assert float(colx[:, 98:].abs().max()) == 0.0


@pytest.mark.parametrize("grey", [True, False])
def test_model_from_raw_uint8_inputs_equals_reference_style_inputs(grey):
    class Cfg:
        backbone, decoder, decoder_embed_dim, num_classes, pretrained_model, bn_eps, bn_momentum = "mit_b0", "MLPDecoder", 256, 9, None, 1e-3, 0.1
    sd = synth_state_dict(cmx_ref.MIT_SPECS["mit_b0"], 9, seed=0, embed_dim=256)
    rng = np.random.default_rng(3)
    B, H, W = 2, 64, 96
    img = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    xg = rng.integers(0, 256, (B, H, W) if grey else (B, H, W, 3), dtype=np.uint8)
    x3 = np.stack([xg] * 3, -1) if grey else xg                      # RGBXDataset.py:57-59
    gt = torch.from_numpy(rng.integers(0, 9, (B, H, W))).cuda()
    res = []
    for mode in ("fp32", "u8"):
        m = EncoderDecoder(Cfg, nn.CrossEntropyLoss(reduction="mean", ignore_index=255), nn.BatchNorm2d)
        m.load_state_dict(sd, strict=True)
        m = m.cuda().train()
        m._eng().stochastic = False
        if mode == "fp32":
            a, b = torch.from_numpy(host_normalize(img)).cuda(), torch.from_numpy(host_normalize(x3)).cuda()
        else:
            a, b = torch.from_numpy(img).cuda(), torch.from_numpy(xg).cuda()
        for _ in range(3):      # eager, capture, replay
            m.load_state_dict({k: v.cuda() for k, v in sd.items()}, strict=True)
            m.zero_grad()
            loss = m(a, b, gt)
            loss.backward()
        grads = {n: p.grad.clone() for n, p in m.named_parameters()}
        m.eval()
        with torch.no_grad():
            logits = m(a, b)
        res.append((loss.item(), grads, logits))
    (l0, g0, y0), (l1, g1, y1) = res
    assert abs(l0 - l1) < 2e-3 * abs(l0), (l0, l1)
    assert float((y0 - y1).norm() / y0.norm()) < 2e-2
    for n in ("backbone.patch_embed1.proj.weight", "backbone.extra_patch_embed1.proj.weight", "backbone.extra_patch_embed1.proj.bias",
              "backbone.block1.0.attn.q.weight", "decode_head.linear_pred.weight"):
        a_, b_ = g0[n].flatten().double(), g1[n].flatten().double()
        cos = float(a_ @ b_ / (a_.norm() * b_.norm()))
        assert cos > 0.99 and abs(float(b_.norm() / a_.norm()) - 1) < 0.05, (n, cos, float(b_.norm() / a_.norm()))
