"""Drop-in replacement for the reference `models/builder.py::EncoderDecoder` (builder.py:14-253) on the
CMX MiT + MLPDecoder hot path.  Same constructor, attributes, state_dict keys and forward contract:

    model = EncoderDecoder(cfg=config, criterion=nn.CrossEntropyLoss(reduction='mean', ignore_index=255),
                           norm_layer=nn.BatchNorm2d)
    loss   = model(rgb, modal_x, label)      # train.py:186
    logits = model(rgb, modal_x)             # engine/evaluator.py:385  -> [B, classes, H, W] fp32

The arithmetic runs on hand-written sm_100a kernels (bf16 operands, fp32 accumulate / residual stream /
statistics / loss) — there is no PyTorch or CPU fallback: CPU inputs raise RuntimeError.
"""
import copy
import os

import torch
import torch.nn as nn

from ..engine import Engine
from .decoders.MLPDecoder import DecoderHead
from .encoders import dual_segformer

_MIT = {"mit_b0": (dual_segformer.mit_b0, [32, 64, 160, 256]), "mit_b1": (dual_segformer.mit_b1, [64, 128, 320, 512]),
        "mit_b2": (dual_segformer.mit_b2, [64, 128, 320, 512]), "mit_b3": (dual_segformer.mit_b3, [64, 128, 320, 512]),
        "mit_b4": (dual_segformer.mit_b4, [64, 128, 320, 512]), "mit_b5": (dual_segformer.mit_b5, [64, 128, 320, 512])}


def _cfg_get(cfg, key, default):
    if cfg is None:
        return default
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def _prep(rgb, modal_x):
    """fp32 NCHW inputs as the reference gets them from its loader - or RAW uint8 images (rgb [B,H,W,3], modal_x [B,H,W] grey or
    [B,H,W,3]): the normalisation / thermal replication / HWC->CHW of dataloader.py:85-112 then happens inside the stage-1
    patch-embed load"""
    if rgb.dtype == torch.uint8:
        return rgb.contiguous(), modal_x.contiguous()
    return rgb.float().contiguous(), modal_x.float().contiguous()


class _CMXStep(torch.autograd.Function):
    """Autograd boundary: forward runs the fused forward+backward step on the engine (optionally as one CUDA
    graph replay); backward only hands the finished gradients (times grad_output) to the parameters, so DDP's
    bucket hooks and torch optimizers see ordinary fp32 .grad tensors."""

    @staticmethod
    def forward(ctx, model, rgb, modal_x, label, *params):
        ctx.model = model
        return model._run_step(rgb, modal_x, label)

    @staticmethod
    def backward(ctx, grad_out):
        eng = ctx.model._engine
        from ..parallel import allreduce_flat_grads_
        world = allreduce_flat_grads_(ctx.model, eng.flat_g)   # no-op unless wrapped in FlatDataParallel
        # one elementwise pass over the flat gradient buffer (fresh tensor each step)
        flat = eng.flat_g * (grad_out / world if world > 1 else grad_out)
        grads = tuple(flat[eng.off[n]:eng.off[n] + eng._numel(n)].view(eng.shape[n]) for n in eng.names)
        return (None, None, None, None) + grads


class EncoderDecoder(nn.Module):
    def __init__(self, cfg=None, criterion=nn.CrossEntropyLoss(reduction='mean', ignore_index=255), norm_layer=nn.BatchNorm2d):
        super().__init__()
        self.norm_layer = norm_layer
        self.cfg = cfg
        backbone = _cfg_get(cfg, "backbone", "mit_b2")
        if backbone not in _MIT:
            # builder.py:146-150 falls back to mit_b2 for unknown names; other families are out of scope here
            if any(backbone.startswith(p) for p in ("swin", "segnext", "resnet")) or backbone.endswith("aspp"):
                raise NotImplementedError("cmx_b200 accelerates the dual MiT (mit_b0..b5) path only; got backbone=%r" % backbone)
            backbone = "mit_b2"
        for key, want in (("feature_rectify_module", "FRM"), ("feature_fusion_module", "FFM")):
            got = _cfg_get(cfg, key, want)
            if got != want:
                raise NotImplementedError("cmx_b200 implements the default %s only (cfg.%s=%r)" % (want, key, got))
        ctor, channels = _MIT[backbone]
        # NOTE: the reference sets channels=[96,192,384,768] for mit_b4/b5 (builder.py:66-75) which crashes in
        # DecoderHead; the channels the backbone really emits are used instead.
        self.channels = list(channels)
        self.backbone = ctor(norm_fuse=norm_layer)
        self.aux_head = None
        decoder = _cfg_get(cfg, "decoder", "MLPDecoder")
        if decoder != "MLPDecoder":
            raise NotImplementedError("cmx_b200 implements cfg.decoder='MLPDecoder' only (got %r)" % decoder)
        self.decode_head = DecoderHead(in_channels=self.channels, num_classes=_cfg_get(cfg, "num_classes", 40),
                                       norm_layer=norm_layer, embed_dim=_cfg_get(cfg, "decoder_embed_dim", 512))
        self.criterion = criterion
        if self.criterion and self.decode_head.num_classes > 64:
            # the fused loss kernels keep one pixel's class scores in registers (two instantiations: <= 16 and <= 64 classes)
            raise NotImplementedError("cmx_b200: the fused loss kernels support up to 64 classes (cfg.num_classes=%d)"
                                      % self.decode_head.num_classes)
        if self.criterion:
            self.init_weights(cfg, pretrained=_cfg_get(cfg, "pretrained_model", None))
        self._engine = None
        self._graphs = {}
        self._flat_dp = None
        # per-channel mean / std of the reference's normalize() for the raw-uint8 input mode (config.norm_mean / norm_std)
        self.norm_mean = list(_cfg_get(cfg, "norm_mean", [0.485, 0.456, 0.406]))
        self.norm_std = list(_cfg_get(cfg, "norm_std", [0.229, 0.224, 0.225]))
        self.use_cuda_graph = os.environ.get("CMX_CUDA_GRAPH", "1") != "0"

    # ---- reference API ---------------------------------------------------------------------------
    def init_weights(self, cfg, pretrained=None):
        """builder.py:199-210 + utils/init_func.py:10-19: kaiming_normal_(fan_in, relu) on decoder convs and
        eps/momentum override on decoder norm layers."""
        if pretrained:
            self.backbone.init_weights(pretrained=pretrained)
        bn_eps, bn_momentum = _cfg_get(cfg, "bn_eps", 1e-3), _cfg_get(cfg, "bn_momentum", 0.1)
        for _, m in self.decode_head.named_modules():
            if isinstance(m, (nn.Conv1d, nn.Conv2d, nn.Conv3d)):
                nn.init.kaiming_normal_(m.weight, mode='fan_in', nonlinearity='relu')
            elif isinstance(m, self.norm_layer):
                m.eps = bn_eps
                m.momentum = bn_momentum
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def encode_decode(self, rgb, modal_x):
        return self._eng().forward_logits(*_prep(rgb, modal_x))

    def forward(self, rgb, modal_x, label=None):
        if label is None:
            with torch.no_grad():
                return self._forward_eval(rgb, modal_x)
        ign, focal = self._criterion_spec()
        if torch.is_grad_enabled():
            params = self._eng_params()
            if any(p.requires_grad for p in params):
                return _CMXStep.apply(self, rgb, modal_x, label, *params)
        with torch.no_grad():
            return self._eng().forward_loss(*_prep(rgb, modal_x), label, ign, with_grad=False, focal=focal)

    def _criterion_spec(self):
        """-> (ignore_index, None | (w_ce, w_focal, gamma, alpha) | ("dice", alpha, smooth)) for the criteria fused into the loss
        kernels (train.py:70-93): DiceCELoss(alpha, ignore_index, 'mean') (utils/loss_opr.py:103-156) ·
        nn.CrossEntropyLoss(mean, ignore_index) · FocalLoss(ignore_label, gamma, alpha, 'mean') (utils/loss_opr.py:157-196, or
        this package's utils.loss_opr.FocalLoss) · the 'CE_Focal' tuple, combined as c0 + 0.2 * c1 (builder.py:246-247)."""
        def is_ce(c):
            return isinstance(c, nn.CrossEntropyLoss) and c.reduction == 'mean' and c.weight is None \
                and getattr(c, "label_smoothing", 0.0) == 0.0

        def is_focal(c):
            return type(c).__name__ == "FocalLoss" and all(hasattr(c, a) for a in ("ignore_label", "gamma", "alpha")) \
                and getattr(c, "reduction", "mean") == 'mean'
        def is_dice_ce(c):
            # reference class: .alpha, .dice (DiceLoss: .smooth, .ignore_index, .reduction), .ce; this package's carrier: flat attributes
            if type(c).__name__ != "DiceCELoss" or not hasattr(c, "alpha"):
                return False
            d = getattr(c, "dice", c)
            return getattr(d, "reduction", "mean") == 'mean' and hasattr(d, "ignore_index")
        crit = self.criterion
        if is_dice_ce(crit):
            d = getattr(crit, "dice", crit)
            return int(d.ignore_index), ("dice", float(crit.alpha), float(getattr(d, "smooth", 1e-6)))
        if is_ce(crit):
            return crit.ignore_index, None
        if is_focal(crit):
            return int(crit.ignore_label), (0.0, 1.0, float(crit.gamma), float(crit.alpha))
        if isinstance(crit, tuple) and len(crit) == 2 and is_ce(crit[0]) and is_focal(crit[1]) \
                and crit[0].ignore_index == int(crit[1].ignore_label):
            return crit[0].ignore_index, (1.0, 0.2, float(crit[1].gamma), float(crit[1].alpha))
        raise NotImplementedError("cmx_b200 fuses nn.CrossEntropyLoss(reduction='mean', ignore_index=k), FocalLoss(mean), "
                                  "DiceCELoss(mean) and the (CrossEntropyLoss, FocalLoss) tuple only (train.py:70-93)")

    # ---- engine plumbing ---------------------------------------------------------------------------
    def _eng(self):
        if self._engine is None:
            self._engine = Engine(self)
            self._engine.set_input_norm(self.__dict__.get("norm_mean", [0.485, 0.456, 0.406]), self.__dict__.get("norm_std", [0.229, 0.224, 0.225]))
        return self._engine

    def flatten_parameters(self):
        """move the parameters into the engine's flat fp32 buffer now (otherwise done by the first forward); call after
        model.cuda() when an optimizer that needs the flat layout (optim.FlatAdamW) is stepped before any forward"""
        dev = next(self.parameters()).device
        self._eng()._ensure_flat(dev)
        return self

    def _eng_params(self):
        """parameters in the engine's flat order; cached (walking named_parameters() costs ~1 ms of host time per call,
        which sits between two steps whenever the caller synchronises on the loss)"""
        eng = self._eng()
        cache = self.__dict__.get("_params_cache")
        if cache is None or cache[0] is not eng.names:
            d = dict(self.named_parameters())
            cache = (eng.names, tuple(d[n] for n in eng.names))
            self.__dict__["_params_cache"] = cache
        return cache[1]

    def _graph_entry(self, key):
        """captured graphs by (mode, shapes, ...) key, least recently used first.  Every entry pins the activation memory of its
        capture, so a caller that keeps feeding new shapes (whole-image evaluation of a dataset with mixed image sizes) must not
        grow the cache without bound: beyond CMX_MAX_GRAPHS (default 24) entries the least recently used one is dropped."""
        g = self._graphs.pop(key, None)
        if g is None:
            cap = max(1, int(os.environ.get("CMX_MAX_GRAPHS", "24")))
            if len(self._graphs) >= cap:
                torch.cuda.synchronize()            # the graph about to be destroyed may still be replaying
                while len(self._graphs) >= cap:
                    self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = {"warm": 1}
            return None
        self._graphs[key] = g                       # most recently used -> last
        return g

    def _forward_eval(self, rgb, modal_x):
        rgb, modal_x = _prep(rgb, modal_x)
        key = ("eval", tuple(rgb.shape), tuple(modal_x.shape), rgb.dtype, self.training, rgb.device.index)
        if not self.use_cuda_graph or self.training:
            return self._eng().forward_logits(rgb, modal_x)
        self._eng()._ensure_flat(rgb.device)  # parameters moved / re-created since the capture -> graphs were dropped
        g = self._graph_entry(key)
        if g is None:
            return self._eng().forward_logits(rgb, modal_x)
        if "graph" not in g:
            g["rgb"], g["x"] = rgb.clone(), modal_x.clone()
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(graph), self._eng().hp_main():
                g["out"] = self._eng().forward_logits(g["rgb"], g["x"])
            g["graph"] = graph
        g["rgb"].copy_(rgb)
        g["x"].copy_(modal_x)
        g["graph"].replay()
        return g["out"].clone()

    def _run_step(self, rgb, modal_x, label):
        """One fused forward+backward step.  The engine generator yields at the points where something outside the kernel
        stream has to happen: the SyncBatchNorm all-reduces of the decoder norm (train.py:64-67 passes nn.SyncBatchNorm in
        every distributed run) and, under FlatDataParallel, the point where the gradients of flat_g[:split_off] are final
        (their all-reduce starts there and overlaps the rest of the backward pass).  With CUDA graphs the step is captured
        as one graph per segment between such points, all sharing one memory pool (NCCL itself is never captured)."""
        from ..parallel import allreduce_slice_async
        rgb, modal_x = _prep(rgb, modal_x)
        ign, focal = self._criterion_spec()
        eng = self._eng()
        eng._ensure_flat(rgb.device)          # parameters moved / re-created since the capture -> graphs were dropped
        eng.split_at_early = self._flat_dp is not None
        key = ("train", tuple(rgb.shape), tuple(modal_x.shape), rgb.dtype, self.training, rgb.device.index, focal, ign, eng.stochastic,
               self._flat_dp is not None)

        def needs_host(ev):
            return ev != "early_gradients_ready" or self._flat_dp is not None

        def on_event(ev):
            if ev == "early_gradients_ready":
                allreduce_slice_async(self, eng.flat_g[:eng.split_off])
            else:
                eng.handle_event(ev)

        def eager(a, b, c):
            gen = eng.forward_loss_steps(a, b, c, ign, with_grad=True, focal=focal)
            try:
                ev = next(gen)
                while True:
                    on_event(ev)
                    ev = next(gen)
            except StopIteration as done:
                loss = done.value
            allreduce_slice_async(self, eng.flat_g[eng.split_off:])
            return loss

        if not self.use_cuda_graph or eng.forced_dp is not None or eng.forced_dropout is not None or eng.sync_hook is not None:
            return eager(rgb, modal_x, label)
        g = self._graph_entry(key)
        if g is None:
            return eager(rgb, modal_x, label)
        if "graphs" not in g:
            g["rgb"], g["x"], g["label"] = rgb.clone(), modal_x.clone(), label.to(torch.int64).clone()
            torch.cuda.synchronize()
            pool = torch.cuda.graph_pool_handle()
            graphs, events = [], []
            gen = eng.forward_loss_steps(g["rgb"], g["x"], g["label"], ign, with_grad=True, focal=focal)
            done = False
            while not done:
                gr = torch.cuda.CUDAGraph()
                ev = None
                with torch.cuda.graph(gr, pool=pool), eng.hp_main():
                    try:
                        ev = next(gen)
                        while not needs_host(ev):
                            ev = next(gen)
                    except StopIteration as fin:
                        g["loss"] = fin.value
                        done = True
                graphs.append(gr)
                if not done:
                    events.append(ev)   # keeps the event's tensor (graph-pool memory) referenced for the replays
            g["graphs"], g["events"] = graphs, events
        g["rgb"].copy_(rgb)
        g["x"].copy_(modal_x)
        g["label"].copy_(label)
        for i, gr in enumerate(g["graphs"]):
            gr.replay()
            if i < len(g["events"]):
                on_event(g["events"][i])
        allreduce_slice_async(self, eng.flat_g[eng.split_off:])
        return g["loss"].clone()

    def __getstate__(self):
        # the evaluator pickles the model into spawned children (engine/evaluator.py:131-137):
        # drop per-process device state (engine buffers, CUDA graphs); it is rebuilt lazily.
        st = self.__dict__.copy()
        st["_engine"] = None
        st["_graphs"] = {}
        st["_flat_dp"] = None
        st.pop("_flat_pending", None)
        st.pop("_params_cache", None)
        eng = self.__dict__.get("_engine")
        if eng is not None and eng.flat_p is not None:
            # after the first step every p.data is a view of the one flat buffer, and plain pickle writes the
            # whole underlying storage once PER TENSOR (837 x 266 MB for MiT-B2): pickle compact clones instead
            memo = {id(p): nn.Parameter(p.detach().clone(), requires_grad=p.requires_grad) for p in self.parameters()}
            st["_modules"] = copy.deepcopy(self._modules, memo)
        return st
