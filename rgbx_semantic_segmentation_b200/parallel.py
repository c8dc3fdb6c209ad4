"""Single-node data parallelism for the CUDA EncoderDecoder (reference: DistributedDataParallel in train.py:145-146,
NCCL process group from engine/engine.py:56).

`torch.nn.parallel.DistributedDataParallel(model)` works unchanged (gradients reach it through autograd), but because
the fused step delivers all 810 gradients at once, DDP's per-parameter bucket copies (~1600 tiny copy kernels) are
fully exposed.  `FlatDataParallel` keeps the same semantics — parameters/buffers broadcast from rank 0 at construction,
gradients averaged over ranks every step — with TWO NCCL all-reduces over contiguous slices of the engine's flat fp32
gradient buffer (266 MB for MiT-B2): the slice holding the decoder and stages 4-3 (91 % of the bytes) is reduced as soon
as the backward pass has finished it, concurrently with the backward of the two high-resolution stages; the rest right
after the last backward kernel (train.py:145-146 relies on DDP's bucketed overlap for the same purpose)."""
import torch
import torch.distributed as dist
import torch.nn as nn


class FlatDataParallel(nn.Module):
    """DistributedDataParallel semantics (train.py:145-146) on the engine's flat gradient buffer: parameters and buffers
    are broadcast from rank 0 at construction, the floating-point buffers (BatchNorm running statistics) again before every
    training forward (DDP's broadcast_buffers=True: with per-rank FFM BatchNorm statistics the ranks would otherwise drift
    apart), and the gradients are averaged over the ranks every step.  An nn.SyncBatchNorm decoder norm (what the
    reference's train.py passes whenever it runs distributed) is synchronised by the engine itself."""

    def __init__(self, module, process_group=None, broadcast_buffers=True):
        super().__init__()
        if not dist.is_initialized():
            raise RuntimeError("FlatDataParallel needs an initialised torch.distributed process group")
        self.module = module
        self.process_group = process_group
        self.world_size = dist.get_world_size(process_group)
        self.broadcast_buffers = broadcast_buffers
        self._src = dist.get_global_rank(process_group, 0) if process_group is not None else 0
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=self._src, group=process_group)
        self._fbufs = [b for b in module.buffers() if b.is_floating_point()]
        module._flat_dp = (process_group, self.world_size)

    def _sync_buffers(self):
        """one coalesced broadcast of the floating-point buffers from rank 0 (3 launches + one small NCCL broadcast)"""
        if not self._fbufs or self.world_size == 1:
            return
        with torch.no_grad():
            flat = torch.cat([b.reshape(-1) for b in self._fbufs])
            dist.broadcast(flat, src=self._src, group=self.process_group)
            torch._foreach_copy_(self._fbufs, [v.view_as(b) for v, b in zip(flat.split([b.numel() for b in self._fbufs]), self._fbufs)])

    def forward(self, *args, **kwargs):
        if self.broadcast_buffers and self.module.training and torch.is_grad_enabled():
            self._sync_buffers()
        return self.module(*args, **kwargs)


class _Pending:
    """asynchronous all-reduces of slices of the flat gradient buffer issued during the step; wait() makes the current
    stream wait for them (no host block) and returns the divisor"""

    def __init__(self, world):
        self.world, self.works = world, []

    def add(self, work):
        self.works.append(work)

    def wait(self):
        for w in self.works:
            w.wait()
        self.works = []
        return self.world


def allreduce_slice_async(model, flat_slice):
    """SUM all-reduce of one contiguous slice of the flat gradient buffer, ordered after the work already enqueued on the
    current stream and running on the process group's own stream, i.e. concurrently with whatever the caller enqueues
    next (the rest of the backward pass).  No-op (returns None) unless the model is wrapped in FlatDataParallel."""
    dp = getattr(model, "_flat_dp", None)
    if dp is None:
        return None
    group, world = dp
    pend = getattr(model, "_flat_pending", None)
    if pend is None:
        pend = model._flat_pending = _Pending(world)
    if flat_slice.numel():
        pend.add(dist.all_reduce(flat_slice, op=dist.ReduceOp.SUM, group=group, async_op=True))
    return pend


def allreduce_flat_grads_(model, flat):
    """finish the gradient reduction: waits for the slices already in flight (issued by the step itself, overlapped with the
    backward pass), or - if none were issued - all-reduces the whole flat buffer now; returns the divisor (world size) or 1"""
    dp = getattr(model, "_flat_dp", None)
    if dp is None:
        return 1
    group, world = dp
    pend = getattr(model, "_flat_pending", None)
    if pend is not None and pend.works:
        return pend.wait()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return world
