"""Per-rank program of the multi-GPU data-parallel parity tests (tests/test_dp_gpu.py); launch with torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        tests/dp_worker.py --mode flat|ddp|reference

flat / ddp : the CUDA EncoderDecoder with norm_layer=nn.SyncBatchNorm (what train.py:64-67 passes when distributed) wrapped
             in FlatDataParallel / torch DistributedDataParallel; every rank trains on its own shard; 4 calls (eager, eager,
             graph-segment capture + replay, replay); the rank-averaged gradients, per-rank losses and the synchronised
             running statistics are compared with oracle.cmx_ref.forward_data_parallel on rank 0.
reference  : the UNMODIFIED reference model (baseline/_ref) under DistributedDataParallel + nn.SyncBatchNorm in fp32 on the
             GPUs against the same oracle function - this pins the oracle's restatement of the distributed step to what
             torch's DDP + SyncBatchNorm really compute (skipped with a message when the reference is not installed)."""
import argparse
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import cmx_ref  # noqa: E402
from oracle.synth import synth_inputs, synth_state_dict  # noqa: E402


class Cfg:
    backbone = "mit_b0"
    decoder = "MLPDecoder"
    decoder_embed_dim = 256
    num_classes = 9
    pretrained_model = None
    bn_eps = 1e-3
    bn_momentum = 0.1


def oracle_dp(sd, spec, shards):
    params = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var")))
              for k, v in sd.items()}
    stats = {}
    loss, losses = cmx_ref.forward_data_parallel(params, spec, shards, sync_decoder_bn=True, decoder_bn_eps=1e-3, new_stats=stats)
    loss.backward()
    return [l.item() for l in losses], {k: p.grad for k, p in params.items() if p.requires_grad}, stats


def check(named_grads, ref, what, min_cos, max_ratio):
    gmax = max(g.norm().item() for g in ref.values())
    worst = (2.0, None)
    for n, g in named_grads.items():
        g, gr = g.double().cpu().flatten(), ref[n].double().flatten()
        if gr.norm().item() < 1e-6 * gmax:
            assert g.norm().item() < 1e-3 * gmax, (what, n, g.norm().item())
            continue
        cos = (g @ gr / (g.norm() * gr.norm())).item()
        ratio = g.norm().item() / gr.norm().item()
        worst = min(worst, (cos, n))
        assert abs(ratio - 1) < max_ratio or abs(ratio - 1) * gr.norm().item() < 5e-4 * gmax, (what, n, ratio)
    assert worst[0] >= min_cos, "%s: worst gradient cosine %s" % (what, worst)
    return worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="flat", choices=["flat", "ddp", "reference"])
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    spec = cmx_ref.MIT_SPECS["mit_b0"]
    sd = synth_state_dict(spec, 9, seed=0, embed_dim=256)
    from test_dp_gpu import make_shard   # odd ranks see blocky / smooth images: per-rank and global batch statistics really differ
    shards = [make_shard(r) for r in range(world)]
    rgb, x, gt = (t.to(dev) for t in shards[rank])
    crit = nn.CrossEntropyLoss(reduction="mean", ignore_index=255)
    if rank == 0:
        torch.set_num_threads(os.cpu_count() or 1)
        ref_losses, ref_grads, ref_stats = oracle_dp(sd, spec, shards)
    if args.mode == "reference":
        sys.path.insert(0, os.path.join(ROOT))
        from baseline import ref_loader
        if not ref_loader.available():
            if rank == 0:
                print("reference not installed: skipped\nDP_WORKER_OK")
            dist.destroy_process_group()
            return
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        config, RefED = ref_loader.load()
        config.backbone, config.num_classes, config.decoder_embed_dim = "mit_b0", 9, 256
        m = RefED(cfg=config, criterion=crit, norm_layer=nn.SyncBatchNorm)
        m.load_state_dict(sd, strict=True)
        for mod in m.modules():
            if type(mod).__name__ == "DropPath":
                mod.drop_prob = 0.0
        m.decode_head.dropout.p = 0.0
        m = m.to(dev).train()
        net = torch.nn.parallel.DistributedDataParallel(m, device_ids=[local])
        loss = net(rgb, x, gt)
        loss.backward()
        losses = [torch.zeros((), device=dev) for _ in range(world)]
        dist.all_gather(losses, loss.detach())
        if rank == 0:
            for r in range(world):
                assert abs(losses[r].item() - ref_losses[r]) <= 1e-4 * abs(ref_losses[r]), (r, losses[r].item(), ref_losses[r])
            w = check({n: p.grad for n, p in m.named_parameters()}, ref_grads, "reference DDP+SyncBN vs oracle", 0.9999, 2e-3)
            b = dict(m.named_buffers())
            for k in ("running_mean", "running_var"):
                assert torch.allclose(b["decode_head.linear_fuse.1." + k].cpu(), ref_stats["decode_head.linear_fuse.1." + k],
                                      rtol=1e-4, atol=1e-5), k
            print("reference DDP + SyncBatchNorm == oracle.forward_data_parallel: losses %s, worst gradient cosine %.6f (%s)"
                  % ([round(l.item(), 6) for l in losses], w[0], w[1]))
            print("DP_WORKER_OK")
        dist.destroy_process_group()
        return

    from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder
    from rgbx_semantic_segmentation_b200.parallel import FlatDataParallel
    m = EncoderDecoder(Cfg, crit, nn.SyncBatchNorm)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).train()
    m._eng().stochastic = False
    net = FlatDataParallel(m) if args.mode == "flat" else torch.nn.parallel.DistributedDataParallel(m, device_ids=[local])
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    for call in range(4):
        m.load_state_dict(sd_dev, strict=True)   # also resets the running statistics
        for p in m.parameters():
            p.grad = None
        loss = net(rgb, x, gt)
        loss.backward()
        losses = [torch.zeros((), device=dev) for _ in range(world)]
        dist.all_gather(losses, loss.detach())
        flat = torch.cat([p.grad.flatten() for p in m.parameters()])
        spread = flat.clone()
        dist.broadcast(spread, src=0)
        same = torch.tensor([float((spread - flat).abs().max() <= 1e-6 * flat.abs().max())], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        assert same.item() == 1.0, "call %d: the ranks hold different averaged gradients" % call
        if rank == 0:
            for r in range(world):
                assert abs(losses[r].item() - ref_losses[r]) <= 5e-3 * abs(ref_losses[r]), (call, r, losses[r].item(), ref_losses[r])
            w = check({n: p.grad for n, p in m.named_parameters()}, ref_grads, "%s call %d" % (args.mode, call), 0.97, 0.25)
            b = dict(m.named_buffers())
            for k in ("running_mean", "running_var"):
                assert torch.allclose(b["decode_head.linear_fuse.1." + k].cpu(), ref_stats["decode_head.linear_fuse.1." + k],
                                      rtol=2e-2, atol=2e-3), (call, k)
            print("%s call %d: losses %s (oracle %s), worst gradient cosine %.4f (%s)"
                  % (args.mode, call, [round(l.item(), 5) for l in losses], [round(l, 5) for l in ref_losses], w[0], w[1]))
    if args.mode == "flat":
        key = [k for k in m._graphs if k[0] == "train"][0]
        n_graphs = len(m._graphs[key]["graphs"])
        assert n_graphs == 4, "expected 4 graph segments (SyncBN fwd, SyncBN bwd, early-gradient all-reduce), got %d" % n_graphs
    if rank == 0:
        print("DP_WORKER_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
