#!/usr/bin/env python
"""Benchmark of the CMX hot path (BASELINE.json metric: train img/s, MiT-B2 RGB-T 480x640, 9 classes).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores (oracle port)

A step = forward + backward + AdamW update on one batch of synthetic input (batch 8 per GPU, random-init
weights; there is no network for datasets/checkpoints).  `value` is timed with inputs resident in HBM;
`e2e` is the same step driven through the public EncoderDecoder API from pinned HOST buffers, with the
host->device copies of rgb/modal_x/label and the device->host read of the loss inside the timed region.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[1] (N = 1) / configs[2] (N > 1): the configuration the headline metric is quoted on
    "b2_mfnet": dict(backbone="mit_b2", H=480, W=640, ncls=9, batch=8,
                     metric="train img/s, CMX MiT-B2 RGB-T 480x640 (fwd+bwd+AdamW)",
                     workload="CMX MiT-B2 RGB-T MFNet shape (480x640, 9 classes) bf16 training, batch 8 per GPU "
                              "(BASELINE.json configs[1]; configs[2] for N>1)"),
    # BASELINE.json configs[3]: MiT-B4 at the PST900 native shape (the reference builder's b4 channel bug fixed, App. A-1)
    "b4_pst900": dict(backbone="mit_b4", H=720, W=1280, ncls=5, batch=4,
                      metric="train img/s, CMX MiT-B4 RGB-T 720x1280 (fwd+bwd+AdamW)",
                      workload="CMX MiT-B4 RGB-T PST900 shape (720x1280, 5 classes) bf16 training, batch 4 per GPU "
                               "(BASELINE.json configs[3])"),
}
H, W, NCLS, PER_GPU_BATCH = 480, 640, 9, 8
BACKBONE = "mit_b2"
METRIC, UNIT = CONFIGS["b2_mfnet"]["metric"], "img/s"
WORKLOAD = CONFIGS["b2_mfnet"]["workload"]
RIDGE_FLOP_PER_BYTE = 213.0  # 1396.8 TF / 6.554 TB/s (MEASURED_PEAKS.json)


class Cfg:
    backbone = "mit_b2"
    decoder = "MLPDecoder"
    decoder_embed_dim = 512
    num_classes = NCLS
    pretrained_model = None
    bn_eps = 1e-3
    bn_momentum = 0.1
    feature_rectify_module = "FRM"
    feature_fusion_module = "FFM"


def select_config(name):
    global H, W, NCLS, PER_GPU_BATCH, BACKBONE, METRIC, WORKLOAD
    c = CONFIGS[name]
    H, W, NCLS, PER_GPU_BATCH, BACKBONE = c["H"], c["W"], c["ncls"], c["batch"], c["backbone"]
    METRIC, WORKLOAD = c["metric"], c["workload"]
    Cfg.backbone, Cfg.num_classes = BACKBONE, NCLS


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        try:
            d = json.load(open(f))
            p.update({k: d[k] for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in d})
            p["src"] = "measured"
        except Exception:
            pass
    return p


def group_weight(module, lr):
    """utils/init_func.py:33-57 — weights of Linear/Conv decay, biases and norm parameters do not."""
    decay, no_decay = [], []
    for m in module.modules():
        if isinstance(m, (nn.Linear, nn.Conv2d)):
            decay.append(m.weight)
            if m.bias is not None:
                no_decay.append(m.bias)
        elif isinstance(m, (nn.BatchNorm2d, nn.LayerNorm, nn.SyncBatchNorm)):
            no_decay += [m.weight, m.bias]
    return [dict(params=decay, lr=lr), dict(params=no_decay, weight_decay=0.0, lr=lr)]


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            text, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in text.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def synth_batch(batch, seed, device=None, pin=False):
    g = torch.Generator().manual_seed(seed)
    rgb = torch.randn(batch, 3, H, W, generator=g)
    x = torch.randn(batch, 3, H, W, generator=g)
    gt = torch.randint(0, NCLS, (batch, H, W), generator=g)
    gt[torch.rand(batch, H, W, generator=g) < 0.03] = 255
    if pin:
        return rgb.pin_memory(), x.pin_memory(), gt.pin_memory()
    return rgb.to(device), x.to(device), gt.to(device)


def workload_config(world):
    """`config` of the JSON line - the WORKLOAD both arms are quoted on (`--impl reference` prints the same dict: it times a
    bounded sample of this workload, described in its `cpu_baseline.sample`); what is specific to this repo's run (optimizer,
    norm layer, gradient all-reduce, CUDA graphs) goes into the separate `impl_config` key"""
    return {"workload": WORKLOAD, "global_batch": PER_GPU_BATCH * world, "per_gpu_batch": PER_GPU_BATCH, "parallelism": "dp%d" % world,
            "l2": "per-step working set (>5 GB of activations at batch 8) exceeds the 126 MB L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------------------
def cpu_reference_rate(steps, warmup, threads=None):
    """The reference's own CPU implementation of the path on the host cores, batch-1 fwd+bwd+AdamW: the UNMODIFIED reference
    model (baseline/_ref, `kind: "reference"`) through its public API `loss = model(rgb, modal_x, label)` (train.py:186);
    when it is not installed, the fp32 oracle port of the same algorithm (`kind: "port"`)."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    rgb, x, gt = synth_batch(1, 1, device="cpu")
    kind = "port"
    try:
        from baseline import ref_loader
        if ref_loader.available():
            kind = "reference"
    except ImportError:
        pass
    if kind == "reference":
        torch.manual_seed(0)
        model = ref_loader.build_model(BACKBONE, NCLS, nn.CrossEntropyLoss(reduction="mean", ignore_index=255), nn.BatchNorm2d).train()
        opt = torch.optim.AdamW(group_weight(model, 6e-5), lr=6e-5, betas=(0.9, 0.999), weight_decay=0.01)

        def train_step():
            loss = model(rgb, x, gt)
            opt.zero_grad()
            loss.backward()
            opt.step()

        def infer():
            model.eval()
            with torch.no_grad():
                model(rgb, x)
            model.train()
    else:
        from oracle import cmx_ref
        from oracle.synth import synth_state_dict
        spec = cmx_ref.MIT_SPECS[BACKBONE]
        sd = synth_state_dict(spec, NCLS, seed=0)
        params = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var")))
                  for k, v in sd.items()}
        opt = torch.optim.AdamW([p for p in params.values() if p.requires_grad], lr=6e-5, weight_decay=0.01)

        def train_step():
            loss = cmx_ref.forward(params, spec, rgb, x, gt, training=True, decoder_bn_eps=1e-3)
            opt.zero_grad()
            loss.backward()
            opt.step()

        def infer():
            with torch.no_grad():
                cmx_ref.forward(sd, spec, rgb, x, training=False)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        train_step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    t0 = time.perf_counter()
    infer()
    t_inf = time.perf_counter() - t0
    what = ("the unmodified reference EncoderDecoder (baseline/_ref)" if kind == "reference" else "fp32 oracle port of the reference")
    return {"value": len(times) / total, "unit": UNIT, "cores": threads, "kind": kind, "inference_img_s": 1.0 / t_inf,
            "sample": "%d steps of batch 1 (fwd+bwd+AdamW, %s, fp32, %d threads) after %d warm-up"
                      % (len(times), what, threads, warmup), "ms_per_step": 1e3 * total / len(times)}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    r = cpu_reference_rate(steps, max(1, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": max(1, args.warmup), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(max(1, args.gpus)),
            "sample": "the same training step (fwd+bwd+AdamW) on one image of the batch per step: the reference's fp32 "
                      "algorithm on the host CPU, all host threads (img/s does not depend on the batch size there)",
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------------------
def main_ours(args):
    import torch.distributed as dist
    from rgbx_semantic_segmentation_b200 import ops
    from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl ours) needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    # train.py:64-67: the norm layer is nn.SyncBatchNorm whenever the run is distributed (it reaches the decoder norm only,
    # SURVEY App. A-3): its statistics are all-reduced inside the step
    norm_layer = nn.SyncBatchNorm if world > 1 else nn.BatchNorm2d
    model = EncoderDecoder(Cfg, nn.CrossEntropyLoss(reduction="mean", ignore_index=255), norm_layer).to(dev).train()
    net = model
    if world > 1:
        if os.environ.get("CMX_BENCH_TORCH_DDP", "0") == "1":   # the reference's wrapper also works (slower: per-param copies)
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
        else:
            from rgbx_semantic_segmentation_b200.parallel import FlatDataParallel
            net = FlatDataParallel(model)
    if args.optimizer == "flat":
        from rgbx_semantic_segmentation_b200.optim import FlatAdamW   # same update as torch.optim.AdamW, one launch
        opt = FlatAdamW(group_weight(model, 6e-5), lr=6e-5, betas=(0.9, 0.999), weight_decay=0.01)
    else:
        opt = torch.optim.AdamW(group_weight(model, 6e-5), lr=6e-5, betas=(0.9, 0.999), weight_decay=0.01, fused=True)
    B = PER_GPU_BATCH
    rgb, x, gt = synth_batch(B, 1 + rank, device=dev)

    def step(a, b, c):
        loss = net(a, b, c)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(3, args.warmup)
    for _ in range(warm):
        step(rgb, x, gt)
    barrier()
    # ---------------- device-resident timing
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step(rgb, x, gt)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    clk = clocks.stop() if clocks else None
    # ---------------- end-to-end timing through the public API from pinned host buffers
    hr, hx, hg = synth_batch(B, 101 + rank, pin=True)
    from rgbx_semantic_segmentation_b200.utils.prefetch import CudaPrefetcher

    def host_batches():
        while True:
            yield hr, hx, hg

    # batch i+1 is uploaded on a side stream (double buffer) while step i runs; every step copies its own inputs once
    feed = CudaPrefetcher(host_batches(), dev)

    def e2e_step():
        (dr, dx_, dg), k = next(feed)
        loss = step(dr, dx_, dg)
        feed.done_with(k)
        return float(loss.item())  # .item() = device->host read of the loss (train.py:170 reads it every iteration)

    for _ in range(2):
        e2e_step()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        last = e2e_step()
    e3.record()
    barrier()
    ms2 = torch.tensor([e2.elapsed_time(e3)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    ms2 = float(ms2)
    # ---------------- per-kernel profile of one eager step (CUDA events around every launch, same stream)
    roof, launches_per_step, top = None, None, []
    # Isolated kernel durations: one stream (no branch / weight-gradient concurrency, so kernels do not share the GPU) and
    # a GPU-side delay in front so that the host has enqueued the whole step before the first kernel starts - otherwise
    # every event pair would also time the ~10 us of host work (tensor-map encoding, ctypes) between two launches.
    model.use_cuda_graph = False   # every rank takes the eager steps (DDP collectives need all ranks)
    eng = model._eng()
    saved_streams = eng.wgrad_stream
    step(rgb, x, gt)
    torch.cuda.synchronize()
    eng.wgrad_stream = False
    step(rgb, x, gt)
    torch.cuda.synchronize()
    n0 = ops.launch_count()
    ops.PROFILE = [] if rank == 0 else None
    torch.cuda._sleep(int(0.12 * 1.9e9))   # ~120 ms: longer than the host needs to enqueue the eager step
    step(rgb, x, gt)
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    launches_per_step = ops.launch_count() - n0
    # calibration: the same event bracket around a trivial launch (8 elements) - the fixed cost every per-launch duration above
    # carries (event records between kernels), reported beside the class figures; no figure is corrected with it
    floor_us = None
    if rank == 0:
        x8, y8 = torch.zeros(8, device=dev), torch.zeros(8, device=dev, dtype=torch.bfloat16)
        ops.PROFILE = []
        torch.cuda._sleep(int(0.01 * 1.9e9))
        for _ in range(65):
            ops.cast_f32_bf16(x8, y8)
        torch.cuda.synchronize()
        cal, ops.PROFILE = ops.PROFILE, None
        floor_us = 1e3 * statistics.median(a_.elapsed_time(b_) for _, a_, b_, _, _ in cal[1:])
    # ---------------- back-to-back replay of each hot class: the SAME launches (arguments recorded from one more eager step whose
    # buffers are kept alive) issued consecutively between ONE pair of events, 3 repetitions - the average launch duration without
    # the event records between kernels (6-7 us each, `event_floor_us`), i.e. as the kernels run inside the captured step
    replay_us = {}
    if rank == 0:
        agg0 = {}
        for tag, a_, b_, _, _ in prof:
            agg0[tag] = agg0.get(tag, 0.0) + a_.elapsed_time(b_)
        hot = [k for k, _ in sorted(agg0.items(), key=lambda kv: -kv[1])[:12] if k.startswith(("gemm_", "cmx_dwconv", "cmx_layernorm", "cmx_attn", "cmx_colsum", "cmx_col2im", "cmx_im2col"))]
    eng.keepalive = []
    ops.RECORD = [] if rank == 0 else None
    step(rgb, x, gt)
    torch.cuda.synchronize()
    rec, ops.RECORD = ops.RECORD, None
    if rank == 0:
        calls = {}
        for tag, fn, cargs in rec:
            calls.setdefault(tag, []).append((fn, cargs))
        for tag in hot:
            cl = calls.get(tag, [])
            if not cl:
                continue
            for fn, cargs in cl:
                fn(*cargs)
            torch.cuda.synchronize()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            for _ in range(3):
                for fn, cargs in cl:
                    fn(*cargs)
            r1.record()
            torch.cuda.synchronize()
            replay_us[tag] = 1e3 * r0.elapsed_time(r1) / (3 * len(cl))
        del calls, rec
    eng.keepalive = None
    eng.wgrad_stream = saved_streams
    model.use_cuda_graph = os.environ.get("CMX_CUDA_GRAPH", "1") != "0"
    if rank == 0:
        agg = {}
        for tag, a, b, fl, nb in prof:
            t = a.elapsed_time(b)
            d = agg.setdefault(tag, [0.0, 0, 0, 0, 0.0])
            d[0] += t; d[1] += 1; d[2] += fl; d[3] += nb; d[4] += t
        # class time = replayed back-to-back duration where available (hot classes), else the sum of the per-launch event pairs
        for tag, d in agg.items():
            if tag in replay_us:
                d[0] = replay_us[tag] * d[1] * 1e-3
        agg = {k: tuple(v) for k, v in agg.items()}
        total = sum(v[0] for v in agg.values())
        top = sorted(agg.items(), key=lambda kv: -kv[1][0])
        pk = peaks()
        name, (t, n, fl, nb, t_pairs) = top[0]
        tensor_bound = nb > 0 and fl / max(nb, 1) > RIDGE_FLOP_PER_BYTE
        if tensor_bound:
            ach = fl / (t * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach / pk["bf16_tflops_sustained"], "traffic": None}
        else:
            ach = nb / (t * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                    "traffic": None}
        # whole-step roofline: every launch at max(bytes / HBM peak, flops / sustained bf16 peak), against the measured step
        t_roof = sum(max(nb_ / (pk["hbm_gbs"] * 1e9), fl_ / (pk["bf16_tflops_sustained"] * 1e12)) for _, _, _, fl_, nb_ in prof) * 1e3
        roof["step"] = {"algorithmic_gbytes": sum(p_[4] for p_ in prof) / 1e9, "algorithmic_tflop": sum(p_[3] for p_ in prof) / 1e12,
                        "t_roofline_ms": t_roof, "t_measured_ms": ms / args.steps, "frac": t_roof / (ms / args.steps),
                        "note": "sum over the launches of one step (launches without a byte count contribute 0); "
                                "measured = graph-replayed training step incl. optimizer"}
        # DRAM bytes actually moved per launch (dram__bytes_read + write) from the committed ncu capture of this step
        # (scripts/ncu_launch_summary.py -> profiles/ncu_traffic_by_class.json); None when no capture is committed
        traffic, traffic_src = {}, None
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic_by_class.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            traffic = {k: v.get("dram_bytes_per_launch") for k, v in tj.get("by_class", {}).items()}
            traffic_src = tj.get("source")
        roof["traffic"] = traffic.get(name)
        roof["traffic_src"] = traffic_src
        # the other classes that matter, same definitions (achieved = algorithmic bytes or flops / isolated duration)
        klist = []
        for kname, (kt, kn, kfl, knb, kt_pairs) in top[:12]:
            if not knb:
                continue
            tb = kfl / max(knb, 1) > RIDGE_FLOP_PER_BYTE
            kach = kfl / (kt * 1e-3) / 1e12 if tb else knb / (kt * 1e-3) / 1e9
            kpk = pk["bf16_tflops_sustained"] if tb else pk["hbm_gbs"]
            klist.append({"kernel": kname, "launches": kn, "share": round(kt / total, 4), "bound": "tensor" if tb else "hbm",
                          "achieved": round(kach, 1), "unit": "TFLOP/s" if tb else "GB/s", "frac": round(kach / kpk, 3),
                          "avg_us": round(1e3 * kt / kn, 2), "avg_us_event_pairs": round(1e3 * kt_pairs / kn, 2),
                          "algorithmic_bytes_per_launch": knb / kn, "traffic": traffic.get(kname)})
        roof["kernels"] = klist
        roof.update(kernel=name, launches_per_step=n, avg_us=1e3 * t / n, share_of_step=t / total, peak_src=pk["src"],
                    event_floor_us=floor_us, avg_us_event_pairs=1e3 * t_pairs / n,
                    achieved_event_pairs=(fl / (t_pairs * 1e-3) / 1e12) if tensor_bound else (nb / (t_pairs * 1e-3) / 1e9),
                    timing="hot classes: the launches of one eager step (recorded arguments, buffers kept alive) replayed back to back "
                           "on one stream between ONE pair of CUDA events, 3 repetitions; other classes and *_event_pairs: CUDA events "
                           "around every launch of one eager single-stream step with the host pre-enqueued (each pair carries "
                           "event_floor_us).  Sum over classes %.1f ms; the graph-replayed multi-stream step overlaps them" % total,
                    algorithmic_bytes_per_launch=nb / n, algorithmic_flops_per_launch=fl / n)
        if args.profile_out:
            with open(args.profile_out, "w") as f:
                f.write("kernel,launches,total_ms,share,avg_us,GB/s,TFLOP/s\n")
                for k, (t, n, fl, nb, _tp) in top:
                    f.write("%s,%d,%.3f,%.4f,%.1f,%.1f,%.2f\n" % (k, n, t, t / total, 1e3 * t / n, nb / (t * 1e-3) / 1e9 if t else 0,
                                                                 fl / (t * 1e-3) / 1e12 if t else 0))
    # ---------------- stock torch DistributedDataParallel around the same model (what the unchanged train.py:145 constructs)
    ddp_stock = None
    if world > 1 and os.environ.get("CMX_BENCH_TORCH_DDP", "0") != "1" and not args.no_ddp_compare:
        model._flat_dp = None
        model.__dict__.pop("_flat_pending", None)
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
        for _ in range(4):
            step(rgb, x, gt)
        barrier()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nst = max(5, args.steps // 2)
        d0.record()
        for _ in range(nst):
            step(rgb, x, gt)
        d1.record()
        barrier()
        msd = torch.tensor([d0.elapsed_time(d1)], device=dev)
        dist.all_reduce(msd, op=dist.ReduceOp.MAX)
        ddp_stock = {"img_s": B * world * nst / (float(msd) * 1e-3), "ms_per_step": float(msd) / nst, "steps": nst,
                     "note": "torch.nn.parallel.DistributedDataParallel(model) instead of FlatDataParallel: same kernels, DDP's own "
                             "bucket copies and all-reduces after the fused step has delivered all gradients at once"}
        net = model
    # ---------------- inference (eval mode, no_grad) on EVERY rank: the evaluator splits the images over the devices
    # (engine/evaluator.py:117-137), no collective; batch-8 throughput and batch-1 latency (evaluator.py:381-391 call shape)
    infer = {}
    model.eval()
    with torch.no_grad():
        for bs in (PER_GPU_BATCH, 1):
            a_, b_ = rgb[:bs].contiguous(), x[:bs].contiguous()
            for _ in range(10):
                model(a_, b_)
            barrier()
            i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_it = 30
            i0.record()
            for _ in range(n_it):
                out = model(a_, b_)
            i1.record()
            barrier()
            t_ms = torch.tensor([i0.elapsed_time(i1) / n_it], device=dev)
            if world > 1:
                dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
            t_ms = float(t_ms)
            infer["batch%d" % bs] = {"img_s": world * bs / (t_ms * 1e-3), "ms_per_forward": t_ms, "n_gpus": world}
    if rank == 0:
        # BASELINE.json configs[4]: sliding-window multi-scale evaluation of one 480x640 image (scales 0.75/1/1.25, crop
        # 480x640, stride 2/3, no flip = 6 crops) from HOST uint8 arrays to the host prediction map + confusion matrix:
        # the reference's per-crop loop (6 batch-1 forwards) vs the batched driver (one batch-6 forward), same model
        try:
            import numpy as np
            from rgbx_semantic_segmentation_b200.utils.metric import hist_info
            from rgbx_semantic_segmentation_b200.utils.sliding_eval import SlidingEvalContext, sliding_eval_rgbX_batched
            rng = np.random.default_rng(2)
            im = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
            mx = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
            gtm = rng.integers(0, NCLS, (H, W)).astype(np.uint8)
            ctx = SlidingEvalContext(model, NCLS, [0.75, 1.0, 1.25], False)
            from rgbx_semantic_segmentation_b200.utils.sliding_eval import sliding_eval_rgbX_gpu
            runs = (("per_crop", lambda: sliding_eval_rgbX_batched(ctx, im, mx, (H, W), 2 / 3, dev, max_batch=1),
                     "reference schedule: host float64 normalisation per crop, 6 batch-1 forwards, host score resize/argmax"),
                    ("batched", lambda: sliding_eval_rgbX_batched(ctx, im, mx, (H, W), 2 / 3, dev, max_batch=8),
                     "same host pre/post-processing, one batch-6 forward"),
                    ("device", lambda: sliding_eval_rgbX_gpu(ctx, im, mx, (H, W), 2 / 3, dev, max_batch=8),
                     "device-resident: uint8 upload, normalise/pad/tile, one batch-6 forward, exp, resize, sum, argmax on the GPU"))
            # dataset streaming (what an evaluation run over a dataset does): confusion matrix accumulated on the device, no
            # read-back per image - the host work of image i+1 overlaps the kernels of image i; one synchronisation at the end
            acc_h = torch.zeros(NCLS, NCLS, dtype=torch.int64, device=dev)
            acc_s = torch.zeros(2, dtype=torch.int64, device=dev)
            for _ in range(2):
                sliding_eval_rgbX_gpu(ctx, im, mx, (H, W), 2 / 3, dev, max_batch=8, gt=gtm, accum=(acc_h, acc_s))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n_stream = 20
            for _ in range(n_stream):
                sliding_eval_rgbX_gpu(ctx, im, mx, (H, W), 2 / 3, dev, max_batch=8, gt=gtm, accum=(acc_h, acc_s))
            torch.cuda.synchronize()
            acc_h.cpu()
            dt = (time.perf_counter() - t0) / n_stream
            infer["sliding_eval_device_stream"] = {"img_s": 1.0 / dt, "ms_per_image": dt * 1e3, "crops_per_image": 6, "images": n_stream,
                                                   "note": "device driver over a stream of images: confusion matrix accumulated on the device, "
                                                           "one synchronisation + read-back after the last image; host wall clock per image"}
            for tag, fn, note in runs:
                for _ in range(2):
                    pred = fn()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                n_img = 5
                for _ in range(n_img):
                    pred = fn()
                    hist_info(NCLS, pred, gtm)
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / n_img
                infer["sliding_eval_" + tag] = {"img_s": 1.0 / dt, "ms_per_image": dt * 1e3, "crops_per_image": 6,
                                                 "note": note + "; host wall clock from uint8 arrays to prediction map + confusion matrix"}
        except ImportError as ex:   # cv2 missing
            infer["sliding_eval"] = {"unavailable": str(ex)}
        model.train()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_rate(10, 2)   # ~10-15 s of host work on 16 cores
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "inference_img_s")}
    if rank == 0:
        gb = B * world
        line = {"metric": METRIC, "value": gb * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": workload_config(world),
                "impl_config": {"optimizer": "FlatAdamW (AdamW, one launch)" if args.optimizer == "flat" else "torch.optim.AdamW(fused)",
                                "norm_layer": norm_layer.__name__,
                                "grad_allreduce": None if world == 1 else ("torch DDP buckets" if os.environ.get("CMX_BENCH_TORCH_DDP", "0") == "1"
                                                                           else "two NCCL all-reduces over slices of the flat fp32 gradient buffer, the first overlapped with the backward pass"),
                                "cuda_graph": bool(model.use_cuda_graph)},
                "e2e": {"value": gb * args.steps / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / args.steps,
                        "h2d_bytes_per_step": int(hr.numel() * 4 + hx.numel() * 4 + hg.numel() * 8), "d2h_bytes_per_step": 4},
                "gpu_launches": int(launches_per_step * args.steps) if launches_per_step else 0,
                "gpu_launches_per_step": launches_per_step, "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
                "inference": infer, "ddp_stock": ddp_stock,
                "last_loss": last,
                "top_kernels": [{"kernel": k, "launches": v[1], "ms": round(v[0], 3), "ms_event_pairs": round(v[4], 3)} for k, v in top[:8]]}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _protect_stdout():
    """Libraries (NCCL's version banner, torchrun warnings) print to fd 1; the contract is ONE JSON line on stdout.
    Route fd 1 to stderr for the duration of the run and keep a private handle for the final line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


if __name__ == "__main__":
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="b2_mfnet", choices=sorted(CONFIGS), help="b2_mfnet = the headline configuration "
                    "(BASELINE.json configs[1]/[2]); b4_pst900 = BASELINE.json configs[3]")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel CUDA-event breakdown of one step (csv)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ddp-compare", action="store_true", help="N > 1: skip the extra timing of the stock torch DDP wrapper")
    ap.add_argument("--optimizer", default="flat", choices=["flat", "torch"],
                    help="flat: optim.FlatAdamW (one launch over the flat parameter buffer); torch: torch.optim.AdamW(fused=True)")
    a = ap.parse_args()
    select_config(a.config)
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
