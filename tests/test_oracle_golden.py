"""CPU: pins oracle/ against fixtures produced by the REAL reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import cmx_ref, metric_ref
from oracle.synth import synth_inputs, synth_state_dict

CASES = ["b2_small", "b2_small_stochastic", "b0_odd", "b4_small", "b2_mfnet"]


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    backbone, ncls, B, H, W, sub, stoch = z["meta"]
    return z, backbone, int(ncls), int(B), int(H), int(W), int(sub), bool(int(stoch))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(golden_dir, name):
    torch.set_num_threads(os.cpu_count() or 1)
    z, backbone, ncls, B, H, W, sub, stoch = _load(golden_dir, name)
    spec = cmx_ref.MIT_SPECS[backbone]
    sd = synth_state_dict(spec, ncls, seed=0)
    rgb, x, gt = synth_inputs(B, H, W, ncls, seed=1)

    # eval-built logits (decoder BN eps 1e-5)
    with torch.no_grad():
        logits = cmx_ref.forward(sd, spec, rgb, x, training=False, decoder_bn_eps=1e-5)
    ref = torch.from_numpy(z["eval_logits"])
    got = logits[:, :, ::sub, ::sub]
    assert got.shape == ref.shape
    # fp32 round-off only: same ops, possibly different reduction order
    assert (got - ref).abs().max().item() < 2e-4 * max(1.0, ref.abs().max().item())
    assert abs(torch.exp(logits[0]).double().sum().item() - float(z["eval_exp_score0_sum"])) < 1e-4 * float(z["eval_exp_score0_sum"])

    # train-built loss + grads (decoder BN eps 1e-3, batch statistics)
    params = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var")))
              for k, v in sd.items()}
    dp, dscale = None, None
    if stoch:
        dp = {k[4:]: (torch.from_numpy(z[k][0]), torch.from_numpy(z[k][1])) for k in z.files if k.startswith("dp::")}
        dscale = torch.from_numpy(z["dropout_scale"])
    new_stats = {}
    loss = cmx_ref.forward(params, spec, rgb, x, gt, training=True, decoder_bn_eps=1e-3,
                           new_stats=new_stats, dp_scales=dp, dropout_scale=dscale)
    loss.backward()
    assert abs(loss.item() - float(z["train_loss"])) < 2e-5 * max(1.0, abs(float(z["train_loss"])))
    names = [str(n) for n in z["grad_names"]]
    norms = z["grad_norms"]
    for n, g in zip(names, norms):
        mine = params[n].grad.double().norm().item()
        assert abs(mine - g) <= 2e-3 * max(g, 1e-6) + 1e-7, (n, mine, g)
    for k in z.files:
        if k.startswith("grad::"):
            ref_g = torch.from_numpy(z[k])
            mine = params[k[6:]].grad
            assert (mine - ref_g).abs().max().item() <= 2e-3 * ref_g.abs().max().item() + 1e-7, k
        if k.startswith("post::"):
            ref_s = torch.from_numpy(z[k])
            mine = new_stats[k[6:]]
            assert torch.allclose(mine.double(), ref_s.double(), rtol=1e-4, atol=1e-6), k


def test_schema_counts():
    # 837 state_dict keys, 810 parameters, 66 565 521 params for MiT-B2 / 9 classes
    sch = cmx_ref.state_dict_schema(cmx_ref.MIT_SPECS["mit_b2"], 9)
    assert len(sch) == 837
    par = {k: v for k, v in sch.items() if v[1] not in ("bn_mean", "bn_var", "bn_count")}
    assert len(par) == 810
    assert sum(int(np.prod(s)) for s, _ in par.values()) == 66565521
    sch4 = cmx_ref.state_dict_schema(cmx_ref.MIT_SPECS["mit_b4"], 5)
    par4 = {k: v for k, v in sch4.items() if v[1] not in ("bn_mean", "bn_var", "bn_count")}
    assert sum(int(np.prod(s)) for s, _ in par4.values()) == 139856269


def test_metric_oracle_matches_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "metric.npz"))
    i = 0
    while f"c{i}_n" in z.files:
        n = int(z[f"c{i}_n"])
        hist, labeled, correct = metric_ref.hist_info(n, z[f"c{i}_pred"].astype(np.int64), z[f"c{i}_gt"])
        assert hist.dtype == np.int64
        assert np.array_equal(hist, z[f"c{i}_hist"])
        assert labeled == int(z[f"c{i}_labeled"]) and correct == int(z[f"c{i}_correct"])
        sc = metric_ref.compute_score(hist, correct, labeled)
        assert np.array_equal(np.asarray(sc[0]), z[f"c{i}_iou"], equal_nan=True)
        assert np.array_equal(np.asarray(sc[1:], dtype=np.float64), z[f"c{i}_scores"], equal_nan=True)
        i += 1
    assert i == 4
