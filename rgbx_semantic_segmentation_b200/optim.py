"""FlatAdamW: torch.optim.AdamW (the reference's optimizer, train.py:95-100) as ONE CUDA launch over the engine's flat
parameter storage.

Same constructor, param_groups, state_dict()/load_state_dict() and step()/zero_grad() contract as torch.optim.AdamW,
so the reference's WarmUpPolyLR (which rewrites param_groups[i]['lr'] every iteration, train.py:160-163) and
checkpointing (engine/engine.py:103-111) keep working.  It needs the parameters to live in the EncoderDecoder's flat
buffer (they do after the first forward, or after `model.flatten_parameters()`) and the gradients to be the views
`loss.backward()` hands out; anything else raises - there is no per-tensor fallback on purpose."""
import ctypes

import torch

from . import _lib, ops

_MAXG = 8
_BLOCK = 64


class FlatAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, maximize=False):
        if amsgrad or maximize:
            raise NotImplementedError("FlatAdamW implements plain AdamW (amsgrad=False, maximize=False)")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or not 0.0 <= weight_decay:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) > _MAXG:
            raise ValueError("FlatAdamW supports at most %d parameter groups" % _MAXG)
        b0 = self.param_groups[0]["betas"], self.param_groups[0]["eps"]
        for g in self.param_groups:
            if (g["betas"], g["eps"]) != b0:
                raise ValueError("FlatAdamW needs the same betas/eps in every parameter group")
        self._bound = None
        self._step = 0

    # ---- binding to the flat storage ---------------------------------------------------------------------------
    def _flat_view(self, t):
        st = t.untyped_storage()
        return torch.empty(0, dtype=torch.float32, device=t.device).set_(st, 0, (st.nbytes() // 4,)), st.data_ptr()

    def _bind(self):
        plist = [(gi, p) for gi, g in enumerate(self.param_groups) for p in g["params"]]
        p0 = plist[0][1]
        if not p0.is_cuda:
            raise RuntimeError("FlatAdamW: parameters must be on a CUDA device")
        flat_p, base = self._flat_view(p0)
        n = flat_p.numel()
        if n % _BLOCK or base % 16:
            raise RuntimeError("FlatAdamW: the parameters are not in the engine's flat buffer yet - run one forward "
                               "(or model.flatten_parameters()) before the first optimizer.step()")
        grp = torch.full((n // _BLOCK,), 255, dtype=torch.uint8)
        offs = []
        for gi, p in plist:
            if p.dtype != torch.float32 or p.untyped_storage().data_ptr() != base or not p.is_contiguous():
                raise RuntimeError("FlatAdamW: every parameter must be an fp32 view of the engine's flat buffer "
                                   "(run one forward or model.flatten_parameters() after model.cuda())")
            o = (p.data_ptr() - base) // 4
            if o % _BLOCK:
                raise RuntimeError("FlatAdamW: parameter not aligned to the %d-element block" % _BLOCK)
            grp[o // _BLOCK:(o + p.numel() + _BLOCK - 1) // _BLOCK] = gi
            offs.append(o)
        dev = p0.device
        m = torch.zeros(n, dtype=torch.float32, device=dev)
        v = torch.zeros(n, dtype=torch.float32, device=dev)
        # expose the moments per parameter so that state_dict() has torch.optim.AdamW's layout
        old = {p: self.state.get(p) for _, p in plist}
        for (gi, p), o in zip(plist, offs):
            ea, es = m[o:o + p.numel()].view(p.shape), v[o:o + p.numel()].view(p.shape)
            st = old[p]
            if st:  # restored by load_state_dict() before the first step
                ea.copy_(st["exp_avg"])
                es.copy_(st["exp_avg_sq"])
                self._step = max(self._step, int(st["step"]))
            self.state[p] = {"step": torch.tensor(float(self._step)), "exp_avg": ea, "exp_avg_sq": es}
        self._bound = dict(flat_p=flat_p, base=base, m=m, v=v, grp=grp.to(dev), plist=plist, offs=offs, n=n)

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._bound = None  # re-bind (copies the restored moments into fresh flat buffers) at the next step

    # ---- the step ----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._bound is None or self._bound["plist"][0][1].untyped_storage().data_ptr() != self._bound["base"]:
            self._bind()
        b = self._bound
        plist, offs = b["plist"], b["offs"]
        g0 = plist[0][1].grad
        if g0 is None:
            raise RuntimeError("FlatAdamW.step(): no gradients - call loss.backward() first")
        flat_g, gbase = self._flat_view(g0)
        # the gradients must be the flat views handed out by the fused backward: same relative offsets as the parameters
        g_first = (g0.data_ptr() - gbase) // 4 - offs[0]   # element offset of the flat gradient buffer in its storage
        for k in (0, len(plist) // 2, len(plist) - 1):
            g = plist[k][1].grad
            if g is None or g.dtype != torch.float32 or g.untyped_storage().data_ptr() != gbase or \
                    (g.data_ptr() - gbase) // 4 != g_first + offs[k]:
                raise RuntimeError("FlatAdamW.step(): gradients are not the flat views produced by loss.backward() of the "
                                   "cmx_b200 EncoderDecoder (were .grad tensors replaced?)")
        if g_first < 0 or g_first + b["n"] > flat_g.numel() or (gbase + 4 * g_first) % 16:
            raise RuntimeError("FlatAdamW.step(): unexpected gradient storage layout")
        self._step += 1
        ng = len(self.param_groups)
        lr = (ctypes.c_float * ng)(*[float(g["lr"]) for g in self.param_groups])
        wd = (ctypes.c_float * ng)(*[float(g["weight_decay"]) for g in self.param_groups])
        beta1, beta2 = self.param_groups[0]["betas"]
        ops._call("cmx_adamw_flat", b["flat_p"].data_ptr(), flat_g.data_ptr() + 4 * g_first, b["m"].data_ptr(), b["v"].data_ptr(),
                  None, b["grp"].data_ptr(), b["n"], ctypes.cast(lr, ctypes.c_void_p), ctypes.cast(wd, ctypes.c_void_p), ng,
                  float(beta1), float(beta2), float(self.param_groups[0]["eps"]), 1.0, self._step, ops._stream(),
                  nbytes=28 * b["n"])
        return loss

    def state_dict(self):
        if self._bound is not None:
            for _, p in self._bound["plist"]:
                self.state[p]["step"] = torch.tensor(float(self._step))
        return super().state_dict()
