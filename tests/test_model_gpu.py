"""GPU end-to-end parity of the CUDA EncoderDecoder against (a) the committed golden fixtures produced by the REAL
reference and (b) the fp32 oracle on identical weights and inputs.

Tolerances (bf16 operands, fp32 accumulate/residual/statistics).  Yardstick measured on the reference itself
(SURVEY.md §8c): reference under torch.autocast(bf16) vs its own fp32 run gives logits rel-L2 1.5e-2, max-abs
0.10*std, argmax agreement 98.8 %.  We require at least that:
    logits   rel-L2 <= 2.5e-2 ; max-abs <= 0.25 * std(ref) (an extreme-value statistic over up to 2.8 M logits) ; argmax equal wherever the fp32 top-2 gap > 2*max-abs
    loss     |d| <= 5e-3 * |ref|
    grads    every parameter whose reference gradient is not structurally zero: cosine >= 0.90, median >= 0.99,
             norm ratio within 25 %
"""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

from oracle import cmx_ref  # noqa: E402
from oracle.synth import synth_inputs, synth_state_dict  # noqa: E402

if torch.cuda.is_available():
    from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder


class Cfg:
    decoder = "MLPDecoder"
    decoder_embed_dim = 512
    pretrained_model = None
    bn_eps = 1e-3
    bn_momentum = 0.1


def make(backbone, ncls, train_built, sd):
    cfg = Cfg()
    cfg.backbone, cfg.num_classes = backbone, ncls
    crit = nn.CrossEntropyLoss(reduction="mean", ignore_index=255) if train_built else None
    m = EncoderDecoder(cfg, crit, nn.BatchNorm2d)
    m.load_state_dict(sd, strict=True)
    return m.cuda()


def load_case(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    backbone, ncls, B, H, W, sub, stoch = z["meta"]
    return z, str(backbone), int(ncls), int(B), int(H), int(W), int(sub), bool(int(stoch))


def check_logits(out, ref, what):
    d = (out - ref)
    rel = (d.norm() / ref.norm()).item()
    mx = d.abs().max().item()
    std = ref.std().item()
    assert rel <= 2.5e-2, "%s: logits rel-L2 %.4g" % (what, rel)
    assert mx <= 0.25 * std, "%s: logits max-abs %.4g vs std %.4g" % (what, mx, std)
    top2 = ref.topk(2, dim=1).values
    sep = (top2[:, 0] - top2[:, 1]) > 2 * mx
    assert sep.float().mean().item() > 0.5
    assert torch.equal(out.argmax(1)[sep], ref.argmax(1)[sep]), "%s: argmax differs on separated pixels" % what
    return rel, mx


@pytest.mark.parametrize("name", ["b2_small", "b0_odd", "b4_small", "b2_mfnet"])
def test_eval_logits_vs_reference_golden(golden_dir, name):
    z, backbone, ncls, B, H, W, sub, _ = load_case(golden_dir, name)
    spec = cmx_ref.MIT_SPECS[backbone]
    sd = synth_state_dict(spec, ncls, seed=0)
    rgb, x, _ = synth_inputs(B, H, W, ncls, seed=1)
    m = make(backbone, ncls, False, sd).eval()
    out = None
    for _ in range(3):  # eager, CUDA-graph capture, CUDA-graph replay must all agree (up to the bf16 noise floor:
        # the split-K fp32 atomics of the FFM context GEMM make the summation order run-dependent)
        o = m(rgb.cuda(), x.cuda())
        assert o.shape == (B, ncls, H, W) and o.dtype == torch.float32
        if out is not None:
            assert ((o - out).norm() / out.norm()).item() < 1e-2, "graph replay differs from eager execution"
        out = o
    ref = torch.from_numpy(z["eval_logits"])
    check_logits(out.cpu()[:, :, ::sub, ::sub], ref, name)
    # evaluator numerics: score = exp(logits[0])  (engine/evaluator.py:393)
    s = torch.exp(out[0]).double().sum().item()
    assert abs(s - float(z["eval_exp_score0_sum"])) < 2e-2 * float(z["eval_exp_score0_sum"])


# worst per-parameter gradient cosine vs the fp32 oracle (measured values: profiles/r2_gradient_cosines.txt).  Named exceptions,
# held to 0.90: bias gradients that are sums over all pixels / tokens of terms that cancel 10-25x - the 2-element FRM spatial-gate
# bias and its hidden-layer bias (the kernel reproduces an fp64 evaluation on the same saved tensors, the
# deviation is ReLU-mask-flip noise of the bf16 forward) and the attention q bias (sum over tokens of dq, which sums to ~0
# per softmax row by construction)
MIN_COS = 0.95
MIN_COS_EXCEPTIONS = (("spatial_weights.mlp.2.bias", "spatial_weights.mlp.0.bias", "attn.q.bias"), 0.90)


def grads_vs(m, ref_grads, what):
    rows = []
    gmax = max(g.norm().item() for g in ref_grads.values())
    for n, p in m.named_parameters():
        assert p.grad is not None and p.grad.dtype == torch.float32, n
        g, gr = p.grad.double().cpu().flatten(), ref_grads[n].double().flatten()
        if gr.norm().item() < 1e-6 * gmax:      # structurally zero (bias feeding a BatchNorm etc.)
            assert g.norm().item() < 1e-3 * gmax, (n, g.norm().item())
            continue
        cos = (g @ gr / (g.norm() * gr.norm())).item()
        rows.append((cos, g.norm().item() / gr.norm().item(), n))
    cosv = sorted(r[0] for r in rows)
    print("%s: worst gradient cosines %s" % (what, [(round(r[0], 4), r[2]) for r in sorted(rows)[:4]]))
    low = [r for r in rows if r[0] < (MIN_COS_EXCEPTIONS[1] if r[2].endswith(MIN_COS_EXCEPTIONS[0]) else MIN_COS)]
    assert not low, "%s: gradient cosines below the gate: %s" % (what, sorted(low)[:5])
    assert cosv[len(cosv) // 2] >= 0.99, "%s: median grad cosine %.4f" % (what, cosv[len(cosv) // 2])
    bad = [r for r in rows if abs(r[1] - 1) > 0.25 and abs(r[1] - 1) * ref_grads[r[2]].double().norm().item() > 5e-4 * gmax]
    assert not bad, "%s: grad norm ratio off: %s" % (what, bad[:3])


@pytest.mark.parametrize("name", ["b2_small", "b0_odd", "b2_small_stochastic", "b4_pst900"])
def test_train_loss_and_grads_vs_reference(golden_dir, name):
    """b4_pst900 = BASELINE.json configs[3] at full size (MiT-B4, 720x1280, 5 classes, batch 1; Nkv = 880 / 920, non-integer
    23x40 -> 180x320 decoder ratio): loss, all 1 500+ gradient norms, 12 full gradients and the BatchNorm buffers against the
    fixture written by the REAL reference (tests/golden/make_golden.py b4_pst900)."""
    z, backbone, ncls, B, H, W, sub, stoch = load_case(golden_dir, name)
    spec = cmx_ref.MIT_SPECS[backbone]
    sd = synth_state_dict(spec, ncls, seed=0, ctx_gain=0.1 if name == "b4_pst900" else 1.0)   # see make_golden.py
    rgb, x, gt = synth_inputs(B, H, W, ncls, seed=1)
    m = make(backbone, ncls, True, sd).train()
    eng = m._eng()
    dp, dscale = None, None
    if stoch:
        dp = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("dp::")}
        dscale = torch.from_numpy(z["dropout_scale"])
        eng.forced_dp, eng.forced_dropout = dp, dscale
    else:
        eng.stochastic = False
    loss = m(rgb.cuda(), x.cuda(), gt.cuda())
    loss.backward()
    ref_loss = float(z["train_loss"])
    assert abs(loss.item() - ref_loss) <= 5e-3 * abs(ref_loss), (loss.item(), ref_loss)
    # golden: gradient norms of ALL parameters + a few full tensors straight from the reference
    names = [str(n) for n in z["grad_names"]]
    norms = dict(zip(names, z["grad_norms"]))
    gmax = max(norms.values())
    pd = dict(m.named_parameters())
    bad = []
    for n in names:
        if norms[n] > 1e-4 * gmax:
            r = pd[n].grad.double().norm().item() / norms[n]
            # 25 % on the norm, or - for tensors that are themselves ~1e-3 of the largest gradient - an absolute slack of
            # 5e-4 of the largest gradient norm.  The case that needs it: the 2-element FRM spatial-gate biases, whose
            # gradient is a sum over pixels of terms that cancel 13-24x (sum |ds| = 2e-2 vs |sum ds| = 1.5e-3 at stage 2,
            # tests/tools/gpu_debug_gate.py); the kernel reproduces an fp64 evaluation on the same saved tensors exactly, the
            # ~1.5 % (of sum |ds|) deviation is ReLU-mask-flip noise of the bf16 forward
            if not (abs(r - 1) < 0.25 or abs(r - 1) * norms[n] < 5e-4 * gmax):
                bad.append((round(r, 4), n))
    print("%s: %d gradient norms outside tolerance%s" % (name, len(bad), "" if not bad else ": " + str(sorted(bad)[:12]) + " ... " + str(sorted(bad)[-12:])))
    assert not bad, (name, len(bad), bad[:8])
    for k in z.files:
        if k.startswith("grad::"):
            gr = torch.from_numpy(z[k]).double().flatten()
            g = pd[k[6:]].grad.double().cpu().flatten()
            assert (g @ gr / (g.norm() * gr.norm())).item() > 0.97, k
        if k.startswith("post::") and not k.endswith("num_batches_tracked"):
            b = dict(m.named_buffers())[k[6:]].cpu()
            assert torch.allclose(b, torch.from_numpy(z[k]), rtol=2e-2, atol=2e-3), k
    assert int(dict(m.named_buffers())["decode_head.linear_fuse.1.num_batches_tracked"]) == \
        int(z["post::decode_head.linear_fuse.1.num_batches_tracked"])
    if name == "b4_pst900":
        return   # the fp32 oracle needs ~20 GB and minutes of CPU time at this size; the reference fixture above is the check
    # full gradient comparison against the oracle (CPU, seconds at this size)
    params = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var")))
              for k, v in sd.items()}
    dpo = None if dp is None else {k: (v[0], v[1]) for k, v in dp.items()}
    cmx_ref.forward(params, spec, rgb, x, gt, training=True, decoder_bn_eps=1e-3, dp_scales=dpo, dropout_scale=dscale).backward()
    grads_vs(m, {n: params[n].grad for n, _ in m.named_parameters()}, name)


def test_training_step_cuda_graph_matches_eager_and_learns():
    """4 calls = eager, eager, capture+replay, replay.  The loss is bit-stable; gradients are reproducible only up
    to the bf16 noise floor because fp32 atomic accumulation order (split-K wgrad, LN/BN parameter grads) varies
    from run to run and the perturbation is re-rounded to bf16 downstream — eager-vs-eager shows the same spread
    as eager-vs-graph.  A few AdamW steps must reduce the loss."""
    spec = cmx_ref.MIT_SPECS["mit_b0"]
    sd = synth_state_dict(spec, 5, seed=0)
    rgb, x, gt = (t.cuda() for t in synth_inputs(2, 64, 64, 5, seed=3))
    m = make("mit_b0", 5, True, sd).train()
    m._eng().stochastic = False
    ref = None
    m.use_cuda_graph = False
    for i in range(5):
        if i == 2:
            m.use_cuda_graph = True
        m.load_state_dict(sd, strict=True)     # also resets BN running statistics
        for p in m.parameters():
            p.grad = None
        loss = m(rgb, x, gt)
        loss.backward()
        g = torch.cat([p.grad.flatten() for p in m.parameters()])
        if ref is None:
            ref = (loss.item(), g.clone())
        else:
            assert abs(loss.item() - ref[0]) < 1e-6
            rel = ((g - ref[1]).norm() / ref[1].norm()).item()
            print("call %d vs eager: grad rel diff %.3g" % (i, rel))
            assert rel < 2e-2, "call %d: gradients differ from the first eager run (rel %.3g)" % (i, rel)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4)
    m._eng().stochastic = True     # DropPath / Dropout2d drawn inside the captured graph (Philox offset advances per replay)
    losses = []
    for _ in range(20):
        loss = m(rgb, x, gt)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert min(losses[-3:]) < 0.95 * losses[0], losses


def test_no_grad_loss_and_amp_scaler_contract():
    spec = cmx_ref.MIT_SPECS["mit_b0"]
    sd = synth_state_dict(spec, 5, seed=0)
    rgb, x, gt = (t.cuda() for t in synth_inputs(1, 64, 64, 5, seed=3))
    m = make("mit_b0", 5, True, sd).train()
    m._eng().stochastic = False
    with torch.no_grad():
        l0 = m(rgb, x, gt)
    assert l0.dim() == 0 and not l0.requires_grad
    m.load_state_dict(sd, strict=True)
    l1 = m(rgb, x, gt)
    assert abs(l0.item() - l1.item()) < 1e-5
    (l1 * 1024.0).backward()                     # GradScaler-style scaled loss (train.py:195-198)
    g1 = m.backbone.block1[0].attn.q.weight.grad.clone()
    m.load_state_dict(sd, strict=True)
    m.zero_grad()
    m(rgb, x, gt).backward()
    g0 = m.backbone.block1[0].attn.q.weight.grad
    # two separate runs: equal up to the bf16 noise floor (fp32 atomic ordering), see the CUDA-graph test
    cos = torch.nn.functional.cosine_similarity(g1.flatten(), g0.flatten(), dim=0).item()
    assert cos > 0.999 and abs(g1.norm().item() / (1024.0 * g0.norm().item()) - 1) < 2e-2, (cos, g1.norm(), g0.norm())
    # all-ignored labels -> NaN like torch (0/0)
    assert torch.isnan(m(rgb, x, torch.full_like(gt, 255)))


def test_metric_dropin_bit_exact(golden_dir):
    from rgbx_semantic_segmentation_b200.utils import metric
    z = np.load(os.path.join(golden_dir, "metric.npz"))
    i = 0
    while f"c{i}_n" in z.files:
        n = int(z[f"c{i}_n"])
        hist, labeled, correct = metric.hist_info(n, z[f"c{i}_pred"].astype(np.int64), z[f"c{i}_gt"])
        assert hist.dtype == np.int64 and np.array_equal(hist, z[f"c{i}_hist"])
        assert (labeled, correct) == (int(z[f"c{i}_labeled"]), int(z[f"c{i}_correct"]))
        sc = metric.compute_score(hist, correct, labeled)
        assert np.array_equal(np.asarray(sc[0]), z[f"c{i}_iou"], equal_nan=True)
        assert np.array_equal(np.asarray(sc[1:], dtype=np.float64), z[f"c{i}_scores"], equal_nan=True)
        i += 1
    assert i == 4


def test_full_size_training_step_vs_oracle():
    """BASELINE.json configs[1] shape (MiT-B2, 480x640, 9 classes) at batch 2: loss and gradients vs the fp32 oracle."""
    spec = cmx_ref.MIT_SPECS["mit_b2"]
    sd = synth_state_dict(spec, 9, seed=0, ctx_gain=0.1)  # see synth_state_dict: keeps the FFM context softmax unsaturated
    rgb, x, gt = synth_inputs(2, 480, 640, 9, seed=1)
    m = make("mit_b2", 9, True, sd).train()
    m._eng().stochastic = False
    loss = m(rgb.cuda(), x.cuda(), gt.cuda())
    loss.backward()
    torch.set_num_threads(os.cpu_count() or 1)
    params = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var")))
              for k, v in sd.items()}
    ref = cmx_ref.forward(params, spec, rgb, x, gt, training=True, decoder_bn_eps=1e-3)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 5e-3 * abs(ref.item()), (loss.item(), ref.item())
    grads_vs(m, {n: params[n].grad for n, _ in m.named_parameters()}, "b2 480x640 batch 2")


def test_mit_b4_pst900_shape_eval_vs_oracle():
    """BASELINE.json configs[3] shape: MiT-B4, 720x1280, 5 classes (Nkv = 880/920 > 320 -> unfused attention path,
    non-integer 23x40 -> 180x320 bilinear ratio in the decoder)."""
    spec = cmx_ref.MIT_SPECS["mit_b4"]
    sd = synth_state_dict(spec, 5, seed=0)
    rgb, x, _ = synth_inputs(1, 720, 1280, 5, seed=1)
    m = make("mit_b4", 5, False, sd).eval()
    out = m(rgb.cuda(), x.cuda()).cpu()
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ref = cmx_ref.forward(sd, spec, rgb, x, training=False, decoder_bn_eps=1e-5)
    check_logits(out[:, :, ::4, ::4], ref[:, :, ::4, ::4], "b4 720x1280")


def test_flat_adamw_matches_torch_adamw():
    """optim.FlatAdamW (one launch over the flat parameter buffer) vs torch.optim.AdamW on identical gradients:
    two parameter groups with different lr / weight decay (utils/init_func.py:33-57 grouping), 3 steps, then a
    state_dict round trip into a fresh optimizer and one more step."""
    from rgbx_semantic_segmentation_b200.optim import FlatAdamW
    spec = cmx_ref.MIT_SPECS["mit_b0"]
    sd = synth_state_dict(spec, 9, seed=0)
    rgb, x, gt = synth_inputs(2, 64, 96, 9, seed=1)
    m = make("mit_b0", 9, True, sd).train()
    m._eng().stochastic = False
    m(rgb.cuda(), x.cuda(), gt.cuda()).backward()
    decay = [p for n, p in m.named_parameters() if p.dim() > 1]
    no_decay = [p for n, p in m.named_parameters() if p.dim() <= 1]
    groups = lambda a, b: [dict(params=a, lr=3e-3, weight_decay=0.05), dict(params=b, lr=1e-3, weight_decay=0.0)]  # noqa: E731
    # reference copies (independent tensors, same gradients)
    ref_p = {p: p.detach().clone().requires_grad_(True) for p in decay + no_decay}
    for p, q in ref_p.items():
        q.grad = p.grad.detach().clone()
    ref = torch.optim.AdamW(groups([ref_p[p] for p in decay], [ref_p[p] for p in no_decay]), betas=(0.9, 0.99), eps=1e-8)
    opt = FlatAdamW(groups(decay, no_decay), betas=(0.9, 0.99), eps=1e-8)

    def check(tag):
        worst = max(((p.detach() - q.detach()).abs().max() / (q.detach().abs().max() + 1e-12)).item() for p, q in ref_p.items())
        assert worst < 2e-6, "%s: max relative parameter deviation %.3g" % (tag, worst)

    for i in range(3):
        ref.step()
        opt.step()
        check("step %d" % (i + 1))
    ssd = opt.state_dict()
    assert set(ssd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and float(ssd["state"][0]["step"]) == 3.0
    opt2 = FlatAdamW(groups(decay, no_decay), betas=(0.9, 0.99), eps=1e-8)
    opt2.load_state_dict(ssd)
    opt2.param_groups[0]["lr"] = ref.param_groups[0]["lr"] = 1e-3   # LR schedulers rewrite param_groups (train.py:160-163)
    ref.step()
    opt2.step()
    check("after state_dict round trip")
    # the engine sees the update: parameters are views of its flat buffer
    n0 = next(iter(dict(m.named_parameters())))
    assert torch.equal(m._eng().P(n0), dict(m.named_parameters())[n0].detach())
    # misuse fails loudly
    dict(m.named_parameters())[n0].grad = torch.zeros_like(dict(m.named_parameters())[n0])
    with pytest.raises(RuntimeError):
        opt2.step()


def test_sliding_window_driver_batched_equals_per_crop():
    """utils/sliding_eval.py on the real model: one batched forward over all crops of all scales gives the same
    prediction map as the reference's per-crop schedule (max_batch=1) wherever the per-crop scores are separated."""
    pytest.importorskip("cv2")
    from rgbx_semantic_segmentation_b200.utils.sliding_eval import SlidingEvalContext, sliding_eval_rgbX_batched
    spec = cmx_ref.MIT_SPECS["mit_b0"]
    sd = synth_state_dict(spec, 5, seed=0)
    m = make("mit_b0", 5, False, sd).eval()
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    mx = rng.integers(0, 256, (96, 128), dtype=np.uint8)          # grey X: replicated to 3 channels like RGBXDataset.py:57-59
    class Net:                                                       # noqa: E306
        def eval(self): return self
        def __call__(self, a, b): return m(a, b.expand(-1, 3, -1, -1).contiguous())
    ctx = SlidingEvalContext(Net(), 5, [0.75, 1.0, 1.5], True)
    p1 = sliding_eval_rgbX_batched(ctx, img, mx, (64, 64), 2 / 3, "cuda", max_batch=1)
    p8 = sliding_eval_rgbX_batched(ctx, img, mx, (64, 64), 2 / 3, "cuda", max_batch=8)
    assert p1.shape == (96, 128) and p1.dtype == np.int64
    assert (p1 != p8).mean() < 0.02, "batched and per-crop predictions differ on %.2f %% of the pixels" % (100 * (p1 != p8).mean())


def test_ragged_input_size_train_and_eval_vs_oracle():
    """input not a multiple of 32 (72x104 -> stages 18x26, 9x13, 5x7, 3x4: odd feature maps, non-integer bilinear ratios,
    SR convolutions that drop the last rows/columns like nn.Conv2d(k=s=R) does), batch 1, plus an all-ignored label row
    block; loss, all gradients and eval logits against the fp32 oracle."""
    spec = cmx_ref.MIT_SPECS["mit_b0"]
    sd = synth_state_dict(spec, 9, seed=0)
    rgb, x, gt = synth_inputs(1, 72, 104, 9, seed=3)
    gt[:, :7] = 255
    m = make("mit_b0", 9, True, sd).train()
    m._eng().stochastic = False
    loss = m(rgb.cuda(), x.cuda(), gt.cuda())
    loss.backward()
    params = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var")))
              for k, v in sd.items()}
    ref = cmx_ref.forward(params, spec, rgb, x, gt, training=True, decoder_bn_eps=1e-3)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 5e-3 * abs(ref.item()), (loss.item(), ref.item())
    grads_vs(m, {n: params[n].grad for n, _ in m.named_parameters()}, "b0 72x104")
    me = make("mit_b0", 9, False, sd).eval()
    out = me(rgb.cuda(), x.cuda()).cpu()
    with torch.no_grad():
        refl = cmx_ref.forward(sd, spec, rgb, x, training=False, decoder_bn_eps=1e-5)
    assert out.shape == refl.shape == (1, 9, 72, 104)
    check_logits(out, refl, "b0 72x104 eval")


def test_all_pixels_ignored_gives_nan_loss_like_torch():
    """nn.CrossEntropyLoss(mean, ignore_index) over zero valid pixels is 0/0 = NaN in the reference (train.py:72-73)"""
    spec = cmx_ref.MIT_SPECS["mit_b0"]
    sd = synth_state_dict(spec, 5, seed=0)
    rgb, x, gt = synth_inputs(1, 64, 64, 5, seed=2)
    gt[:] = 255
    m = make("mit_b0", 5, True, sd).train()
    with torch.no_grad():
        loss = m(rgb.cuda(), x.cuda(), gt.cuda())
    assert torch.isnan(loss).item()


def test_flat_data_parallel_two_graph_step_single_process_group():
    """FlatDataParallel path on one GPU (single-rank NCCL group): the step is captured as TWO CUDA graphs around the
    point where the first gradient slice is final, the slice all-reduces are issued asynchronously and awaited in
    backward(); loss and gradients must match the plain single-graph model, and the overlap plumbing must not leak
    pending work between steps."""
    import torch.distributed as dist
    from rgbx_semantic_segmentation_b200.parallel import FlatDataParallel
    spec = cmx_ref.MIT_SPECS["mit_b0"]
    sd = synth_state_dict(spec, 9, seed=0)
    rgb, x, gt = synth_inputs(2, 64, 96, 9, seed=1)
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29531", rank=0, world_size=1)
    try:
        plain = make("mit_b0", 9, True, sd).train()
        plain._eng().stochastic = False
        m = make("mit_b0", 9, True, sd).train()
        m._eng().stochastic = False
        net = FlatDataParallel(m)
        eng = m._eng()
        ref_loss = plain(rgb.cuda(), x.cuda(), gt.cuda())
        ref_loss.backward()
        losses = []
        for it in range(4):     # eager, eager->capture, replay, replay
            for p in m.parameters():
                p.grad = None
            loss = net(rgb.cuda(), x.cuda(), gt.cuda())
            loss.backward()
            losses.append(loss.item())
            assert not m._flat_pending.works, "all-reduce handles must be consumed by backward()"
        key = [k for k in m._graphs if k[0] == "train"][0]
        assert len(m._graphs[key]["graphs"]) == 2, "FlatDataParallel must capture the step as two graphs"
        assert 0 < eng.split_off < eng.total and eng.n_early > 0
        assert all(abs(v - ref_loss.item()) < 2e-3 * abs(ref_loss.item()) for v in losses), (losses, ref_loss.item())
        pg = dict(plain.named_parameters())
        worst = 1.0
        for n, p in m.named_parameters():
            a, b = p.grad.double().flatten(), pg[n].grad.double().flatten()
            if b.norm() > 1e-4:
                worst = min(worst, (a @ b / (a.norm() * b.norm())).item())
        assert worst > 0.98, worst
    finally:
        if created:
            dist.destroy_process_group()


def test_focal_and_ce_focal_criteria_train_step_vs_oracle():
    """criterion = FocalLoss / (CrossEntropyLoss, FocalLoss) / DiceCELoss (train.py:70-93, builder.py:246-247): loss and gradients of the whole
    model against the fp32 oracle with the same criterion applied to its logits"""
    from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder
    from rgbx_semantic_segmentation_b200.utils.loss_opr import DiceCELoss, FocalLoss
    spec = cmx_ref.MIT_SPECS["mit_b0"]
    sd = synth_state_dict(spec, 9, seed=0)
    rgb, x, gt = synth_inputs(2, 64, 96, 9, seed=1)
    ce = nn.CrossEntropyLoss(reduction="mean", ignore_index=255)
    for crit in (FocalLoss(ignore_label=255, gamma=4.0, alpha=0.25), (ce, FocalLoss(ignore_label=255, gamma=2.0, alpha=0.25)),
                 DiceCELoss(alpha=0.5, ignore_index=255)):
        cfg = Cfg()
        cfg.backbone, cfg.num_classes = "mit_b0", 9
        m = EncoderDecoder(cfg, crit, nn.BatchNorm2d)
        m.load_state_dict(sd, strict=True)
        m = m.cuda().train()
        m._eng().stochastic = False
        loss = m(rgb.cuda(), x.cuda(), gt.cuda())
        loss.backward()
        params = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var")))
                  for k, v in sd.items()}
        logits = cmx_ref.forward(params, spec, rgb, x, None, training=True, decoder_bn_eps=1e-3)
        ref = crit(logits, gt) if not isinstance(crit, tuple) else crit[0](logits, gt) + 0.2 * crit[1](logits, gt)
        ref.backward()
        assert abs(loss.item() - ref.item()) <= 5e-3 * abs(ref.item()), (loss.item(), ref.item())
        grads_vs(m, {n: params[n].grad for n, _ in m.named_parameters()}, "criterion %s" % (type(crit).__name__,))
        with torch.no_grad():
            assert abs(m(rgb.cuda(), x.cuda(), gt.cuda()).item() - ref.item()) <= 5e-3 * abs(ref.item())
    with pytest.raises(NotImplementedError):
        bad = EncoderDecoder(cfg, nn.MSELoss(), nn.BatchNorm2d).cuda().train()
        bad(rgb.cuda(), x.cuda(), gt.cuda())
