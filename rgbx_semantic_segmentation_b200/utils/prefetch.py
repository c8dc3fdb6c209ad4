"""Host->device input staging for the training loop.

The reference moves every minibatch with `.cuda(non_blocking=True)` on the compute stream (train.py:150-152), so the
copy and the step serialise.  `CudaPrefetcher` wraps any iterable of (pinned) host minibatches and uploads batch i+1
on a side stream into the other half of a double buffer while step i runs; `next()` hands out device tensors whose
upload the compute stream has been made to wait for.  Every minibatch is still copied exactly once per step."""
import torch


class CudaPrefetcher:
    def __init__(self, iterable, device=None):
        self.it = iter(iterable)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("CudaPrefetcher needs a CUDA device (cmx_b200 has no CPU path)")
        self.stream = torch.cuda.Stream(self.dev)
        self.bufs = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [None, None]   # event on the compute stream after the step that last read buffer k was enqueued
        self.k = 0
        self.pending = None
        self._upload()

    def _upload(self):
        try:
            host = next(self.it)
        except StopIteration:
            self.pending = None
            return
        k = self.k
        if self.bufs[k] is None:
            self.bufs[k] = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.dev) for t in host)
        with torch.cuda.stream(self.stream):
            if self.consumed[k] is not None:
                self.stream.wait_event(self.consumed[k])   # do not overwrite a buffer a queued step still reads
            for d, h in zip(self.bufs[k], host):
                d.copy_(h, non_blocking=True)
            self.ready[k].record(self.stream)
        self.pending = k
        self.k ^= 1

    def __iter__(self):
        return self

    def __next__(self):
        if self.pending is None:
            raise StopIteration
        k = self.pending
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self.ready[k])
        out = self.bufs[k]
        self._upload()   # start the next upload before the caller launches this step
        return out, k

    def done_with(self, k):
        """call after enqueuing the step that consumes buffer k (marks the point after which it may be overwritten)"""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self.consumed[k] = ev
