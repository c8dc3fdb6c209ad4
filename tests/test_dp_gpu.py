"""Data-parallel NUMERICS of the CUDA path: what N ranks compute together must equal what the
reference's distributed step computes - DistributedDataParallel gradient averaging (train.py:145-146) with the decoder norm
an nn.SyncBatchNorm (train.py:64-67) and per-rank FFM BatchNorm statistics.

The checker is oracle.cmx_ref.forward_data_parallel (one set of fp32 parameters, shards run side by side, decoder BN over
all shards, gradient of the mean of the per-rank losses).

* test_two_rank_lockstep_*: runs on ONE GPU - two model replicas are driven through the engine's step generator in
  lock-step and the all-reduce events are executed by summing the two replicas' tensors in-process.
* test_two_gpu_*: real NCCL, 2 processes x 2 GPUs (skipped on a 1-GPU box; `gpurun --gpus 2`), FlatDataParallel and stock
  DistributedDataParallel, eager + graph-segment capture + replay; tests/dp_worker.py is the per-rank program."""
import os
import subprocess
import sys

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

from oracle import cmx_ref  # noqa: E402
from oracle.synth import synth_inputs, synth_state_dict  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Cfg:
    backbone = "mit_b0"
    decoder = "MLPDecoder"
    decoder_embed_dim = 256
    num_classes = 9
    pretrained_model = None
    bn_eps = 1e-3
    bn_momentum = 0.1


def make_shard(r):
    """per-rank shard; odd ranks see blocky / smooth images instead of white noise so that per-rank and global batch statistics
    of the decoder norm really differ (a mere rescaling of the inputs is removed by the LayerNorms)"""
    rgb, x, gt = synth_inputs(2, 64, 96, 9, seed=11 + r)
    if r % 2 == 1:
        g = torch.Generator().manual_seed(5 + r)
        rgb = torch.nn.functional.interpolate(torch.randn(2, 3, 4, 6, generator=g), size=(64, 96), mode="nearest") * 2.0 + 1.0
        x = torch.nn.functional.interpolate(torch.randn(2, 3, 8, 12, generator=g), size=(64, 96), mode="bilinear") * 0.3
    return rgb, x, gt


def oracle_dp(sd, spec, shards, sync):
    params = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var")))
              for k, v in sd.items()}
    new_stats = {}
    loss, losses = cmx_ref.forward_data_parallel(params, spec, shards, sync_decoder_bn=sync, decoder_bn_eps=1e-3, new_stats=new_stats)
    loss.backward()
    return [l.item() for l in losses], {k: p.grad for k, p in params.items() if p.requires_grad}, new_stats


def compare_grads(named_grads, ref, what, min_cos=0.97):
    gmax = max(g.norm().item() for g in ref.values())
    worst = (2.0, None)
    for n, g in named_grads.items():
        g, gr = g.double().cpu().flatten(), ref[n].double().flatten()
        if gr.norm().item() < 1e-6 * gmax:
            assert g.norm().item() < 1e-3 * gmax, (what, n)
            continue
        cos = (g @ gr / (g.norm() * gr.norm())).item()
        ratio = g.norm().item() / gr.norm().item()
        worst = min(worst, (cos, n))
        assert abs(ratio - 1) < 0.25 or abs(ratio - 1) * gr.norm().item() < 5e-4 * gmax, (what, n, ratio)
    assert worst[0] >= min_cos, "%s: worst gradient cosine %s" % (what, worst)
    return worst


@pytest.mark.parametrize("sync", [True, False])
def test_two_rank_lockstep_emulation_vs_oracle(sync):
    from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder
    spec = cmx_ref.MIT_SPECS["mit_b0"]
    sd = synth_state_dict(spec, 9, seed=0, embed_dim=256)
    shards = [make_shard(r) for r in range(2)]
    norm = nn.SyncBatchNorm if sync else nn.BatchNorm2d
    models = []
    for r in range(2):
        m = EncoderDecoder(Cfg, nn.CrossEntropyLoss(reduction="mean", ignore_index=255), norm)
        m.load_state_dict(sd, strict=True)
        m = m.cuda().train()
        eng = m._eng()
        eng.stochastic = False
        eng.sync_emulate_world = 2 if sync else 0
        models.append(m)
    gens = [m._eng().forward_loss_steps(*(t.cuda() for t in shards[r]), 255, True) for r, m in enumerate(models)]
    losses, n_sync = [None, None], 0
    while any(g is not None for g in gens):
        evs = []
        for r, g in enumerate(gens):
            try:
                evs.append(next(g))
            except StopIteration as done:
                losses[r] = done.value.item()
                gens[r] = None
                evs.append(None)
        assert (evs[0] is None) == (evs[1] is None) and type(evs[0]) is type(evs[1]), "the ranks must yield the same events"
        if isinstance(evs[0], tuple) and evs[0][0] == "allreduce_sum":
            n_sync += 1
            total = evs[0][1] + evs[1][1]
            evs[0][1].copy_(total)
            evs[1][1].copy_(total)
    assert n_sync == (2 if sync else 0), "SyncBatchNorm: one all-reduce in forward (sum, sumsq) and one in backward"
    ref_losses, ref_grads, ref_stats = oracle_dp(sd, spec, shards, sync)
    for r in range(2):
        assert abs(losses[r] - ref_losses[r]) <= 5e-3 * abs(ref_losses[r]), (r, losses[r], ref_losses[r])
    engs = [m._eng() for m in models]
    avg = {n: 0.5 * (engs[0].G(n) + engs[1].G(n)) for n in engs[0].names}
    compare_grads(avg, ref_grads, "2-rank lock-step, sync=%s" % sync)
    # decoder norm running statistics: over the GLOBAL batch when synchronised (identical on both ranks)
    for r in range(2 if sync else 1):
        b = dict(models[r].named_buffers())
        for k in ("running_mean", "running_var"):
            assert torch.allclose(b["decode_head.linear_fuse.1." + k].cpu(), ref_stats["decode_head.linear_fuse.1." + k],
                                  rtol=2e-2, atol=2e-3), (r, k)
    if sync:
        # and the synchronised statistics really differ from per-rank ones (the test would otherwise prove nothing)
        _, _, local_stats = oracle_dp(sd, spec, shards, False)
        # (the batch MEAN of the decoder norm's input does not depend on the data: every fused feature leaves a BatchNorm)
        k = "decode_head.linear_fuse.1.running_var"
        assert not torch.allclose(local_stats[k], ref_stats[k], rtol=2e-2, atol=2e-3)


def _run_workers(mode, nproc=2, port=29541):
    if torch.cuda.device_count() < nproc:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (nproc, nproc))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dp_worker.py"), "--mode", mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, "dp_worker failed:\n%s\n%s" % (r.stdout[-3000:], r.stderr[-3000:])
    assert "DP_WORKER_OK" in r.stdout, r.stdout[-2000:]
    return r.stdout


def test_two_gpu_flat_data_parallel_syncbn_vs_oracle():
    _run_workers("flat", port=29541)


def test_two_gpu_torch_ddp_syncbn_vs_oracle():
    _run_workers("ddp", port=29542)
