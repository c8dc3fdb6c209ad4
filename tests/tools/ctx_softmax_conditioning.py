#!/usr/bin/env python
"""Why the full-size gradient parity tests scale the FFM context kv weights by 0.1 (oracle/synth.py `ctx_gain`).

Pure fp32 oracle on the CPU, MiT-B2 480x640 batch 1, run twice: on the synthetic inputs and on the same inputs rounded to
bf16 (a relative perturbation of 2^-9 of the INPUT IMAGE only - every bf16 pipeline perturbs every activation by that much).
Per-parameter gradient cosine between the two runs, for ctx_gain = 1.0 and 0.1.  With unit-variance synthetic kv weights the
FFM context logits K^T V * scale (net_utils.py:209, a sum over N = 19 200 tokens) have a standard deviation in the hundreds,
the softmax over dim -2 is one-hot, and the gradient is discontinuous in the activations: the fp32 arithmetic itself cannot
reproduce its own gradients under a 0.2 % input perturbation.  The reference's real initialisation (trunc_normal std 0.02,
dual_segformer.py:52-65) gives logit std < 1.
    python tests/tools/ctx_softmax_conditioning.py > profiles/r2_ctx_softmax_conditioning.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cmx_ref  # noqa: E402
from oracle.synth import synth_inputs, synth_state_dict  # noqa: E402

torch.set_num_threads(os.cpu_count() or 1)
spec = cmx_ref.MIT_SPECS["mit_b2"]
rgb, x, gt = synth_inputs(1, 480, 640, 9, seed=1)


def grads(sd, a, b):
    params = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var"))) for k, v in sd.items()}
    trace = {}
    loss = cmx_ref.forward(params, spec, a, b, gt, training=True, decoder_bn_eps=1e-3, trace=trace)
    loss.backward()
    return loss.item(), {k: p.grad for k, p in params.items() if p.requires_grad}


for gain in (1.0, 0.1):
    sd = synth_state_dict(spec, 9, seed=0, ctx_gain=gain)
    l0, g0 = grads(sd, rgb, x)
    l1, g1 = grads(sd, rgb.bfloat16().float(), x.bfloat16().float())
    gmax = max(v.norm().item() for v in g0.values())
    rows = []
    for k in g0:
        a, b = g0[k].double().flatten(), g1[k].double().flatten()
        if a.norm().item() < 1e-6 * gmax:
            continue
        rows.append(((a @ b / (a.norm() * b.norm())).item(), b.norm().item() / a.norm().item(), k))
    rows.sort()
    cos = [r[0] for r in rows]
    print("ctx_gain %.1f: loss %.6f vs %.6f (inputs rounded to bf16); gradient cosine over %d parameters: min %.4f, 1%% quantile %.4f, "
          "median %.6f; parameters with cosine < 0.95: %d" % (gain, l0, l1, len(cos), cos[0], cos[len(cos) // 100], cos[len(cos) // 2],
                                                              sum(c < 0.95 for c in cos)))
    for r in rows[:8]:
        print("    cos %.4f  norm ratio %.3f  %s" % r)
