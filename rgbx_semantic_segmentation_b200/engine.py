"""Execution engine of the CMX hot path: explicit forward + hand-derived backward of the dual-branch
MiT encoder, FRM/FFM fusion, MLP decoder and CE loss, expressed as a sequence of sm_100a kernel
launches (ops.py) over token-major bf16 activations with an fp32 residual stream.

Reference semantics restated here (file:line of ynalcakan/RGBX_Semantic_Segmentation):
  encoder  models/encoders/dual_segformer.py:366-442 (stage loop), :116-138 (SR attention),
           :67-74 (Mix-FFN), :176-180 (Block), :217-225 (OverlapPatchEmbed)
  fusion   models/net_utils.py:22-30, 79-83, 147-152 (FRM); :199-214, 273-281, 323-329, 376-384 (FFM)
  decoder  models/decoders/MLPDecoder.py:59-81   loss  models/builder.py:233, 249 (CE, ignore 255)

Design notes
  * parameters are views into ONE flat fp32 buffer (so one cast kernel makes the bf16 operand copies and
    one memset clears all gradients); gradients are accumulated by the wgrad kernels straight into a flat
    fp32 buffer with the same layout.
  * training runs forward AND backward inside `forward_backward` (so the whole step is one stream of
    launches that can be captured in a single CUDA graph); autograd only hands the finished gradients
    to the parameters.
  * linear_fuse (1x1 conv over the 2048-channel concat) is applied per stage at native resolution —
    bilinear interpolation commutes with a per-pixel linear map — so the concat tensor never exists.
"""
import os
import re

import torch

from . import ops
from .ops import ACT_GELU, ACT_NONE, ACT_RELU, ACT_SIGMOID

bf16 = torch.bfloat16
f32 = torch.float32


def conv_out(n, k, s, p):
    return (n + 2 * p - k) // s + 1


class _NS:
    """plain attribute bag for saved activations"""
    pass


class Engine:
    ALIGN = 64  # elements; keeps every parameter 256-byte aligned in the fp32 buffer (128 B in bf16)

    def __init__(self, model):
        self.model = model
        bb = model.backbone
        self.dims, self.heads, self.depths, self.srs = bb.embed_dims, bb.num_heads, bb.depths, bb.sr_ratios
        self.embed = model.decode_head.linear_pred.in_channels
        self.ncls = model.decode_head.num_classes
        self.ncls_ld = (self.ncls + 7) // 8 * 8
        # decoder BatchNorm input: fp32 by default; bf16 (CMX_DECODER_FUSE_BF16=1) saves 4 x 157 MB of traffic per step at
        # batch 8 but measured no step-time difference (A/B on one box: 23.63-23.71 vs 23.63-23.65 ms)
        self.fuse_dtype = bf16 if os.environ.get("CMX_DECODER_FUSE_BF16", "0") == "1" else f32
        # Flat-buffer order = [decoder, stage 4, stage 3 | stage 2, stage 1]: the backward pass finishes the parameter
        # gradients of the first group (91 % of the bytes for MiT-B2) before it starts stage 2, so data-parallel training
        # can all-reduce that contiguous slice while the two high-resolution stages are still running (parallel.py)
        # Within a stage the RGB-branch parameters (patch_embed / block / norm) come first, then the X-branch ones
        # (extra_*) in the SAME order, then the fusion modules: every X-branch tensor lies exactly gs[s] elements behind its
        # RGB twin, in the fp32 parameter, bf16 operand and gradient buffers alike, so that one GROUPED launch (group =
        # modality branch; the kernels take the group stride) serves both branches of a stage.
        names = [n for n, _ in model.named_parameters()]
        idx = {n: i for i, n in enumerate(names)}

        def kind(n):
            if re.match(r"backbone\.extra_(patch_embed|block|norm)\d", n):
                return 1
            return 0 if re.match(r"backbone\.(patch_embed|block|norm)\d", n) else 2
        stage_rank = {4: 0, 3: 1, 2: 2, 1: 3, 0: 4}
        self.names = sorted(names, key=lambda n: (stage_rank[self._stage_of(n)], kind(n), idx[n]))
        self.n_early = sum(1 for n in names if self._stage_of(n) >= 2)
        self._branch_names = [([n for n in self.names if self._stage_of(n) == s and kind(n) == 0],
                               [n for n in self.names if self._stage_of(n) == s and kind(n) == 1]) for s in range(4)]
        self.flat_p = None
        self.forced_dp = None        # test hook: {block prefix: tensor[2,B]}
        self.forced_dropout = None   # test hook: tensor[B, E]
        self.stochastic = True       # DropPath / Dropout2d active in training mode
        self.trace = None            # debug hook: dict filled with fp32 copies of intermediate activations
        self.fused_attention = os.environ.get("CMX_FUSED_ATTENTION", "1") != "0"
        # EXPERIMENTAL, off by default (kernel-level parity is green on B200; model-level tests and the bench A/B under this flag are
        # still to be run): flash-style attention backward - dQ and
        # dK / dV from the two kernels of csrc/attention_dkv.cu that recompute the probabilities from q, k and the forward's
        # lse; the forward then stores no probabilities and the backward has no dS round trip
        self.attn_dkv_recompute = os.environ.get("CMX_ATTN_DKV_RECOMPUTE", "0") == "1"
        # the FFM of stage s feeds only the decoder, the next stage consumes the FRM output: FFM forward / backward run on a
        # side stream concurrently with the block chains of the neighbouring stage (fork / join are captured into the graph)
        self.ffm_stream = os.environ.get("CMX_FFM_STREAM", "1") != "0"
        self._side = None
        self.split_at_early = False   # set by the builder when the step is cut into CUDA-graph segments at the early-gradient event
        # weight-gradient GEMMs never feed the backward chain: they run on a companion stream of whichever stream
        # computes the data gradients and are joined at the end of each module's backward
        self.wgrad_stream = os.environ.get("CMX_WGRAD_STREAM", "1") != "0"
        self.hp_streams = os.environ.get("CMX_HP_STREAMS", "0") == "1"
        self._hp = None
        self._side_forked = False
        self.norm_mean, self.norm_std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]   # config.norm_mean / norm_std defaults
        self._fold_consts = None
        self._wstreams = {}
        self._wkeep = {}
        self._dec_prep = None
        self._bns = [(n, m) for n, m in model.named_modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
        self.keepalive = None        # measurement hook (bench.py): list that keeps every buffer of a step alive for a replay
        self.poison = None           # debug hook: list of (tensor, allocation site) when NaN-poisoning is on
        self.sync_emulate_world = 0  # test hook: treat the decoder SyncBatchNorm as shared by this many ranks ...
        self.sync_hook = None        # ... whose all-reduce is performed by this callable(tensor) (in-process lock-step emulation)

    @staticmethod
    def _stage_of(name):
        """0..3 for backbone parameters of that stage, 4 for the decoder"""
        m = re.match(r"backbone\.(?:extra_)?(?:patch_embed|block|norm)(\d)", name)
        if m:
            return int(m.group(1)) - 1
        m = re.match(r"backbone\.(?:FRMs|FFMs)\.(\d)", name)
        if m:
            return int(m.group(1))
        return 4

    # ------------------------------------------------------------------------------------------
    # flat parameter / gradient storage
    # ------------------------------------------------------------------------------------------
    def _flatten(self, device):
        # captured CUDA graphs hold raw pointers into the previous flat buffers
        self.model.__dict__.get("_graphs", {}).clear()
        self.model.__dict__.pop("_params_cache", None)
        params = dict(self.model.named_parameters())
        off, total = {}, 0
        for n in self.names:
            off[n] = total
            total += (params[n].numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        flat = torch.zeros(total, device=device, dtype=f32)
        for n in self.names:
            p = params[n]
            v = flat[off[n]:off[n] + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
        self.off, self.total = off, total
        self.shape = {n: tuple(params[n].shape) for n in self.names}
        self.gs = []
        for s_, (rgb_n, x_n) in enumerate(self._branch_names):
            assert len(rgb_n) == len(x_n) and len(rgb_n) > 0
            d = off[x_n[0]] - off[rgb_n[0]]
            for a, b in zip(rgb_n, x_n):
                assert b == a.replace("backbone.", "backbone.extra_", 1) and params[a].shape == params[b].shape, (a, b)
                assert off[b] - off[a] == d, "branch twins must lie at a constant distance in the flat buffer"
            self.gs.append(d)
        self.flat_p = flat
        self.flat_g = torch.zeros(total, device=device, dtype=f32)
        self.flat_w = torch.zeros(total, device=device, dtype=bf16)
        self.params = params
        # real (non-depthwise, non-1x1) convolutions - patch embeds and SR convs: (kh,kw,ci)-packed bf16 operands and packed
        # fp32 gradient accumulators live in two flat buffers, described once by a device table, so that packing (start
        # of the step) and gradient unpacking (end of the backward pass) are ONE launch each instead of one per conv
        import struct
        convs, tot = [], 0
        for n in self.names:
            s4 = tuple(params[n].shape)
            if len(s4) == 4 and s4[2] > 1 and s4[1] > 1:
                kpad = (s4[1] * s4[2] * s4[3] + 7) // 8 * 8
                assert s4[1] * s4[2] * s4[3] <= 4608, "conv %s: Ci*kh*kw exceeds the pack kernel's shared-memory row" % n
                convs.append((n, s4, kpad, tot))
                tot += (s4[0] * kpad + 63) // 64 * 64
        self.pk_w = torch.zeros(max(tot, 1), device=device, dtype=bf16)
        self.pk_g = torch.zeros(max(tot, 1), device=device, dtype=f32)
        self.packed, self.packed_g, blob = {}, {}, b""
        for n, s4, kpad, o in convs:
            self.packed[n] = self.pk_w[o:o + s4[0] * kpad].view(s4[0], kpad)
            self.packed_g[n] = self.pk_g[o:o + s4[0] * kpad].view(s4[0], kpad)
            blob += struct.pack("<QQQQiiiiii", flat.data_ptr() + 4 * off[n], self.packed[n].data_ptr(), self.packed_g[n].data_ptr(),
                                self.flat_g.data_ptr() + 4 * off[n], s4[0], s4[1], s4[2], s4[3], kpad, 0)
        pko = {n: o for n, _, _, o in convs}
        self.pk_gs = [0, 0, 0, 0]
        for s_, (rgb_n, x_n) in enumerate(self._branch_names):
            ds = {pko[b] - pko[a] for a, b in zip(rgb_n, x_n) if a in pko}
            assert len(ds) == 1, "packed conv twins must lie at a constant distance"
            self.pk_gs[s_] = ds.pop()
        self.n_convs = len(convs)
        self.n_convs_early = sum(1 for n, _, _, _ in convs if self._stage_of(n) >= 2)
        self.split_off = off[self.names[self.n_early]] if self.n_early < len(self.names) else total   # first late element
        self.conv_table = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device) if convs else None
        self.buffers = dict(self.model.named_buffers())
        # DropPath table: one row per (stage, block index) = the RGB block and its X twin; device tensor built once
        # (graph capture forbids H2D).  Probability 0 (block1.0) just always keeps.
        self._dp_keys, probs = [], []
        for s in range(4):
            for i, blk in enumerate(getattr(self.model.backbone, f"block{s + 1}")):
                xblk = getattr(self.model.backbone, f"extra_block{s + 1}")[i]
                self._dp_keys.append(f"backbone.block{s + 1}.{i}")
                probs.append([getattr(blk.drop_path, "drop_prob", 0.0), getattr(xblk.drop_path, "drop_prob", 0.0)])
        self._dp_pt = torch.tensor(probs, dtype=f32).view(-1, 1, 2, 1).to(device)
        self._dp_any = bool((self._dp_pt > 0).any())

    def _ensure_flat(self, device):
        """(Re)build the flat storage when the parameters moved (model.to()/.cuda() re-creates p.data)."""
        if self.flat_p is not None and self.flat_p.device == device:
            first, last = self.names[0], self.names[-1]
            if (self.params[first].data_ptr() == self.flat_p.data_ptr() + 4 * self.off[first] and
                    self.params[last].data_ptr() == self.flat_p.data_ptr() + 4 * self.off[last]):
                return
        for n, p in self.model.named_parameters():
            if p.device != device:
                raise RuntimeError("cmx_b200: parameter %s is on %s but the inputs are on %s — call model.cuda() first"
                                   % (n, p.device, device))
            break
        self._flatten(device)

    def P(self, name):
        o = self.off[name]
        return self.flat_p[o:o + self._numel(name)].view(self.shape[name])

    def G(self, name):
        o = self.off[name]
        return self.flat_g[o:o + self._numel(name)].view(self.shape[name])

    def W(self, name):
        """bf16 copy of a Linear / 1x1-conv weight as a 2-D [out, in] matrix"""
        o = self.off[name]
        s = self.shape[name]
        return self.flat_w[o:o + self._numel(name)].view(s[0], -1)

    def G2(self, name):
        s = self.shape[name]
        return self.G(name).view(s[0], -1)

    def _numel(self, name):
        n = 1
        for d in self.shape[name]:
            n *= d
        return n

    def refresh_weights(self):
        """fp32 master -> bf16 operand copies (one flat cast) + (kh,kw,ci)-packed conv weights"""
        ops.cast_f32_bf16(self.flat_p, self.flat_w)
        if self.n_convs:
            ops.convw_pack_multi(self.conv_table, self.n_convs)

    # ------------------------------------------------------------------------------------------
    # small helpers
    # ------------------------------------------------------------------------------------------
    def E(self, *shape, dtype=bf16):
        if self.poison is not None:
            # debug aid (tests/tools/gpu_poison.py): NaN-fill every fresh buffer and remember where it was allocated,
            # so that buffers that are read before being (fully) written can be found without compute-sanitizer
            import sys
            if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
                shape = tuple(shape[0])
            t = torch.full(shape, float("nan"), device=self.dev, dtype=dtype) if dtype.is_floating_point \
                else torch.full(shape, -1, device=self.dev, dtype=dtype)
            f = sys._getframe(1)
            self.poison.append((t, "%s:%d" % (f.f_code.co_name, f.f_lineno)))
            return t
        t = torch.empty(*shape, device=self.dev, dtype=dtype)
        if self.keepalive is not None:
            self.keepalive.append(t)
        return t

    def Z(self, *shape, dtype=f32):
        t = torch.zeros(*shape, device=self.dev, dtype=dtype)
        if self.keepalive is not None:
            self.keepalive.append(t)
        return t

    def tr(self, key, t):
        if self.trace is not None:
            self.trace[key] = t.detach().float().clone()

    # ---- companion weight-gradient stream ---------------------------------------------------------
    def _wgrad_ctx(self, *keep):
        """context manager: work issued inside runs on the companion stream of the current stream after everything
        already enqueued on the current stream; `keep` tensors stay referenced until the matching _wgrad_join()."""
        eng = self

        class _W:
            def __enter__(self_):
                self_.ctx = None
                if not eng.wgrad_stream:
                    return
                cur = torch.cuda.current_stream(eng.dev)
                key = cur.cuda_stream
                ws = eng._wstreams.get(key)
                if ws is None:
                    ws = eng._wstreams[key] = torch.cuda.Stream(device=eng.dev)
                eng._wkeep.setdefault(key, []).extend(keep)
                ws.wait_stream(cur)
                self_.ctx = torch.cuda.stream(ws)
                self_.ctx.__enter__()

            def __exit__(self_, *exc):
                if self_.ctx is not None:
                    self_.ctx.__exit__(*exc)
                return False
        return _W()

    def _wgrad_join(self):
        """the current stream waits for its companion weight-gradient stream; operand keep-alives are released"""
        if not self.wgrad_stream:
            return
        cur = torch.cuda.current_stream(self.dev)
        ws = self._wstreams.get(cur.cuda_stream)
        if ws is not None:
            cur.wait_stream(ws)
        self._wkeep.pop(cur.cuda_stream, None)

    def linear_wgrad(self, dy, x, wname, bname=None, gs=None):
        """dW[out,in] += dy[tok,out]^T x[tok,in];  db += colsum(dy)   (on the companion stream).
        gs: dy / x hold the two branches stacked, wname / bname are the RGB-branch parameters (X twin gs elements behind)"""
        g = dict(groups=2, gs_c=gs) if gs is not None else {}
        with self._wgrad_ctx(dy, x):
            ops.mm(dy, x, self.G2(wname), ta=True, tb=True, accumulate=True, **g)
            if bname is not None:
                if gs is not None:
                    ops.colsum(dy, self.G(bname), groups=2, out_gs=gs)
                else:
                    ops.colsum(dy, self.G(bname))

    # ------------------------------------------------------------------------------------------
    # stochastic-depth / dropout multipliers
    # ------------------------------------------------------------------------------------------
    def _make_dp(self, B, training):
        """-> ({RGB block prefix: fp32 [2 (attention | Mix-FFN residual), 2 (RGB | X branch), B] DropPath multipliers}, Dropout2d
        multipliers [B, E] or None)"""
        if not training:
            return {}, None
        if self.forced_dp is not None or self.forced_dropout is not None:
            dp = {}
            for k in self._dp_keys:
                pair = [(self.forced_dp or {}).get(kk) for kk in (k, k.replace("backbone.", "backbone.extra_", 1))]
                if pair[0] is None and pair[1] is None:
                    continue
                pair = [torch.ones(2, B) if t is None else t for t in pair]
                dp[k] = torch.stack([t.to(self.dev, f32) for t in pair], dim=1).contiguous()
            dm = None if self.forced_dropout is None else self.forced_dropout.to(self.dev, f32).contiguous()
            return dp, dm
        if not self.stochastic:
            return {}, None
        dp = {}
        if self._dp_any:
            pt = self._dp_pt
            u = torch.rand(len(self._dp_keys), 2, 2, B, device=self.dev)
            sc = (u >= pt).to(f32) / (1.0 - pt)
            for j, k in enumerate(self._dp_keys):
                dp[k] = sc[j]
        dm = None
        drop = self.model.decode_head.dropout
        if drop is not None and drop.p > 0:
            dm = (torch.rand(B, self.embed, device=self.dev) >= drop.p).to(f32) / (1.0 - drop.p)
        return dp, dm

    # ------------------------------------------------------------------------------------------
    # OverlapPatchEmbed
    # ------------------------------------------------------------------------------------------
    def pe_fwd(self, s, inp, B, H, W, save):
        """both branches in grouped launches.  inp: stage 0 = (rgb, x) NCHW fp32; later stages = the two rectified branch
        tensors stacked [2 * B*H*W, C_prev] bf16.  Returns x0 [2 * M, C] fp32 (RGB rows first)."""
        name = f"backbone.patch_embed{s + 1}"
        C, gs = self.dims[s], self.gs[s]
        wp = self.packed[name + ".proj.weight"]
        if s == 0 and inp[0].dtype == torch.uint8:
            return self.pe1_fwd_u8(inp, B, H, W, save)
        if s == 0:
            k, st, pd = 7, 4, 3
            Ho, Wo = conv_out(H, k, st, pd), conv_out(W, k, st, pd)
            M = B * Ho * Wo
            col = self.E(2 * M, wp.shape[1])
            for g in (0, 1):
                ops.im2col_nchw(inp[g], col[g * M:(g + 1) * M], k, st, pd, Ho, Wo)
        else:
            k, st, pd = 3, 2, 1
            Ho, Wo = conv_out(H, k, st, pd), conv_out(W, k, st, pd)
            M = B * Ho * Wo
            col = self.E(2 * M, wp.shape[1])
            ops.im2col_nhwc(inp, col, 2 * B, H, W, k, st, pd, Ho, Wo)
        y = self.E(2 * M, C, dtype=f32)
        ops.mm(col, wp, y, bias=self.P(name + ".proj.bias"), groups=2, gs_b=self.pk_gs[s], gs_bias=gs)
        x0 = self.E(2 * M, C, dtype=f32)
        mean, rstd = (self.E(2 * M, dtype=f32), self.E(2 * M, dtype=f32)) if save else (None, None)
        ops.layernorm_fwd(y, self.P(name + ".norm.weight"), self.P(name + ".norm.bias"), 1e-5, x0, mean, rstd, groups=2, param_gs=gs)
        c = _NS()
        c.name, c.s, c.col, c.y, c.mean, c.rstd, c.Ho, c.Wo, c.H, c.W, c.wp = name, s, col, y, mean, rstd, Ho, Wo, H, W, wp
        return x0, Ho, Wo, c

    # ---- stage-1 patch embed straight from raw uint8 images: the reference's host input pipeline fused into the load --------
    def set_input_norm(self, mean, std):
        """ImageNet-style per-channel mean / std of the reference's normalize() (config.norm_mean / norm_std)"""
        self.norm_mean, self.norm_std = [float(v) for v in mean], [float(v) for v in std]
        self._fold_consts = None

    def _x_fold(self):
        """grey X: its three replicated channels are x_c = v/255 / std_c - mean_c / std_c (RGBXDataset.py:57-59 + dataloader.py:
        106), so conv(W, x) = conv(Wa, v/255) + conv(Wb, inside-mask) with Wa = sum_c W_c / std_c, Wb = -sum_c W_c mean_c / std_c:
        the 7x7x3 weights folded to the [C, (tap, {value, mask})] operand of cmx_im2col_u8's grey layout (98 of 104 columns)"""
        if self._fold_consts is None or self._fold_consts[0].device != self.dev:
            inv = torch.tensor([1.0 / v for v in self.norm_std], device=self.dev, dtype=f32).view(1, 3, 1, 1)
            nms = torch.tensor([-m / v for m, v in zip(self.norm_mean, self.norm_std)], device=self.dev, dtype=f32).view(1, 3, 1, 1)
            self._fold_consts = (inv, nms)
        inv, nms = self._fold_consts
        Wx = self.P("backbone.extra_patch_embed1.proj.weight")
        C = Wx.shape[0]
        fold = self.Z(C, 104, dtype=f32)
        fold[:, :98].view(C, 49, 2).copy_(torch.stack([(Wx * inv).sum(1), (Wx * nms).sum(1)], dim=-1).view(C, 49, 2))
        return fold.to(bf16)

    def pe1_fwd_u8(self, inp, B, H, W, save):
        """inp = (rgb uint8 [B,H,W,3], x uint8 [B,H,W] grey or [B,H,W,3]) resident on the device"""
        name, xname = "backbone.patch_embed1", "backbone.extra_patch_embed1"
        C, gs = self.dims[0], self.gs[0]
        k, st, pd = 7, 4, 3
        Ho, Wo = conv_out(H, k, st, pd), conv_out(W, k, st, pd)
        M = B * Ho * Wo
        wp = self.packed[name + ".proj.weight"]
        rgb, x = inp
        grey = x.dim() == 3
        y = self.E(2 * M, C, dtype=f32)
        c = _NS()
        if not grey:
            col = self.E(2 * M, wp.shape[1])
            ops.im2col_u8(rgb, col[:M], k, st, pd, Ho, Wo, self.norm_mean, self.norm_std)
            ops.im2col_u8(x, col[M:], k, st, pd, Ho, Wo, self.norm_mean, self.norm_std)
            ops.mm(col, wp, y, bias=self.P(name + ".proj.bias"), groups=2, gs_b=self.pk_gs[0], gs_bias=gs)
            c.col, c.colx = col, None
        else:
            col, colx = self.E(M, wp.shape[1]), self.E(M, 104)
            ops.im2col_u8(rgb, col, k, st, pd, Ho, Wo, self.norm_mean, self.norm_std)
            ops.im2col_u8(x, colx, k, st, pd, Ho, Wo, self.norm_mean, self.norm_std)
            ops.mm(col, wp, y[:M], bias=self.P(name + ".proj.bias"))
            ops.mm(colx, self._x_fold(), y[M:], bias=self.P(xname + ".proj.bias"))
            c.col, c.colx = col, colx
        x0 = self.E(2 * M, C, dtype=f32)
        mean, rstd = (self.E(2 * M, dtype=f32), self.E(2 * M, dtype=f32)) if save else (None, None)
        ops.layernorm_fwd(y, self.P(name + ".norm.weight"), self.P(name + ".norm.bias"), 1e-5, x0, mean, rstd, groups=2, param_gs=gs)
        c.name, c.s, c.y, c.mean, c.rstd, c.Ho, c.Wo, c.H, c.W, c.wp = name, 0, y, mean, rstd, Ho, Wo, H, W, wp
        return x0, Ho, Wo, c

    def pe_bwd(self, c, dx0, B):
        """returns dcol (bf16, both branches stacked) for stages >= 1 (the caller scatters it with col2im), None for stage 0"""
        name, s = c.name, c.s
        gs = self.gs[s]
        M2, C = c.y.shape
        dy = self.E(M2, C)
        ops.layernorm_bwd(dx0, c.y, c.mean, c.rstd, self.P(name + ".norm.weight"), dx=dy,
                          dgamma=self.G(name + ".norm.weight"), dbeta=self.G(name + ".norm.bias"),
                          dbias=self.G(name + ".proj.bias"), groups=2, param_gs=gs)   # conv bias gradient = column sums of dy
        if getattr(c, "colx", None) is not None:
            # grey-X input mode: the two branches have different im2col widths; the folded X weight gradient is chained back to
            # the three channel slices of extra_patch_embed1.proj.weight: dW_c = dWa / std_c - dWb mean_c / std_c
            M = M2 // 2
            with self._wgrad_ctx(dy, c.col, c.colx):
                ops.mm(dy[:M], c.col, self.packed_g[name + ".proj.weight"], ta=True, tb=True, accumulate=True)
                gfold = self.Z(C, 104)
                ops.mm(dy[M:], c.colx, gfold, ta=True, tb=True, accumulate=True)
                inv, nms = self._fold_consts
                g2 = gfold[:, :98].view(C, 49, 2)
                self.G("backbone.extra_patch_embed1.proj.weight").add_(g2[:, :, 0].reshape(C, 1, 7, 7) * inv + g2[:, :, 1].reshape(C, 1, 7, 7) * nms)
            self._wgrad_join()
            return None
        with self._wgrad_ctx(dy, c.col):   # packed gradient; unpacked for all convs at the end of the backward pass
            ops.mm(dy, c.col, self.packed_g[name + ".proj.weight"], ta=True, tb=True, accumulate=True, groups=2, gs_c=self.pk_gs[s])
        dcol = None
        if c.s != 0:
            dcol = self.E(M2, c.wp.shape[1])
            ops.mm(dy, c.wp, dcol, tb=True, groups=2, gs_b=self.pk_gs[s])
        self._wgrad_join()
        return dcol

    # ------------------------------------------------------------------------------------------
    # transformer Block (RGB block and its X twin in grouped launches)
    # ------------------------------------------------------------------------------------------
    def block_fwd(self, p, x, B, H, W, s, dp, save):
        """p: RGB block prefix (the X twin's parameters lie gs elements behind); x: fp32 [2 * B*N, C], RGB rows first;
        dp: None or fp32 [2 (attn | mlp), 2 (branch), B] DropPath multipliers"""
        C, heads, R = self.dims[s], self.heads[s], self.srs[s]
        gs = self.gs[s]
        G2 = dict(groups=2, gs_b=gs, gs_bias=gs)
        B2 = 2 * B
        d = C // heads
        N = H * W
        M = B2 * N
        scale = d ** -0.5
        c = _NS()
        c.p, c.s, c.H, c.W, c.x, c.dp = p, s, H, W, x, dp
        st = (lambda: self.E(M, dtype=f32)) if save else (lambda: None)
        xn1 = self.E(M, C)
        c.m1, c.r1 = st(), st()
        ops.layernorm_fwd(x, self.P(p + ".norm1.weight"), self.P(p + ".norm1.bias"), 1e-6, xn1, c.m1, c.r1, groups=2, param_gs=gs)
        q = self.E(M, C)
        ops.mm(xn1, self.W(p + ".attn.q.weight"), q, bias=self.P(p + ".attn.q.bias"), **G2)
        if R > 1:
            Hk, Wk = conv_out(H, R, R, 0), conv_out(W, R, R, 0)
            Nk = Hk * Wk
            wsr = self.packed[p + ".attn.sr.weight"]
            pat = self.E(B2 * Nk, wsr.shape[1])
            ops.im2col_nhwc(xn1, pat, B2, H, W, R, R, 0, Hk, Wk)
            # few output tiles (B*Nk x C) but K = R*R*C up to 4096: split-K with fp32 atomics (bias added by slice 0)
            sr = self.Z(B2 * Nk, C)
            ops.mm(pat, wsr, sr, bias=self.P(p + ".attn.sr.bias"), accumulate=True, groups=2, gs_b=self.pk_gs[s], gs_bias=gs)
            srn = self.E(B2 * Nk, C)
            c.ms, c.rs = (self.E(B2 * Nk, dtype=f32), self.E(B2 * Nk, dtype=f32)) if save else (None, None)
            ops.layernorm_fwd(sr, self.P(p + ".attn.norm.weight"), self.P(p + ".attn.norm.bias"), 1e-5, srn, c.ms, c.rs,
                              groups=2, param_gs=gs)
            kv_in = srn
            c.pat, c.sr, c.Hk, c.Wk = pat, sr, Hk, Wk
        else:
            Nk = N
            kv_in = xn1
        kv = self.E(B2 * Nk, 2 * C)
        ops.mm(kv_in, self.W(p + ".attn.kv.weight"), kv, bias=self.P(p + ".attn.kv.bias"), **G2)
        Np = (Nk + 7) // 8 * 8   # leading dimension of the score / probability rows (16-byte rows for TMA)
        O = self.E(M, C)
        if d == 64 and Nk <= ops.ATTN_MAX_NK and self.fused_attention:
            # flash-style fused kernel: scores stay in tensor memory; P is only written (by TMA) when backward needs it
            c.lse = self.E(B2 * heads * N, dtype=f32) if save and self.attn_dkv_recompute else None
            # the probabilities are only stored when the (default) backward reads them back
            Pm = self.E(B2 * heads * N, Np)[:, :Nk] if save and c.lse is None else None
            ops.attn_fwd(q, kv, O, B2, N, Nk, heads, scale, p_out=Pm, lse=c.lse)
        elif d == 64 and self.fused_attention:
            # long key axis (Nkv = 880 / 920 at 720x1280): the fused kernel per chunk of <= 320 keys + exact combination; the
            # backward recomputes the probabilities from q, k and the log-sum-exp (no [N, Nkv] tensor in either direction)
            c.lse = self.E(B2 * heads * N, dtype=f32)
            Pm = None
            ops.attn_fwd_chunked(q, kv, O, c.lse, B2, N, Nk, heads, scale)
        else:
            # unfused path (head_dim != 64: mit_b0): S = scale * Q K^T (fp32, transient), P = softmax(S), O = P V
            c.lse = None
            S = self.E(B2 * heads * N, Np, dtype=f32)[:, :Nk]
            ops.gemm_raw(q, kv, S, N, Nk, d, C, 2 * C, Np, batch=(B2, heads), sA=(N * C, d), sB=(Nk * 2 * C, d),
                         sC=(heads * N * Np, N * Np), alpha=scale)
            Pm = self.E(B2 * heads * N, Np)[:, :Nk]
            ops.softmax_rows_fwd(S, Pm)
            del S
            ops.gemm_raw(Pm, kv, O, N, d, Nk, Np, 2 * C, C, b_off=C, trans_b=True, batch=(B2, heads),
                         sA=(heads * N * Np, N * Np), sB=(Nk * 2 * C, d), sC=(N * C, d))
        x1 = self.E(M, C, dtype=f32)
        ops.mm(O, self.W(p + ".attn.proj.weight"), x1, bias=self.P(p + ".attn.proj.bias"), residual=x,
               row_scale=None if dp is None else dp[0], rows_per_sample=N, gs_scale=B, **G2)
        xn2 = self.E(M, C)
        c.m2, c.r2 = st(), st()
        ops.layernorm_fwd(x1, self.P(p + ".norm2.weight"), self.P(p + ".norm2.bias"), 1e-6, xn2, c.m2, c.r2, groups=2, param_gs=gs)
        h = self.E(M, 4 * C)
        ops.mm(xn2, self.W(p + ".mlp.fc1.weight"), h, bias=self.P(p + ".mlp.fc1.bias"), **G2)
        g = self.E(M, 4 * C)
        ops.dwconv3x3_fwd(h, self.P(p + ".mlp.dwconv.dwconv.weight"), self.P(p + ".mlp.dwconv.dwconv.bias"), ACT_GELU, g, B, H, W,
                          groups=2, param_gs=gs)
        x2 = self.E(M, C, dtype=f32)
        ops.mm(g, self.W(p + ".mlp.fc2.weight"), x2, bias=self.P(p + ".mlp.fc2.bias"), residual=x1,
               row_scale=None if dp is None else dp[1], rows_per_sample=N, gs_scale=B, **G2)
        if save:
            c.xn1, c.q, c.kv, c.kv_in, c.Pm, c.O, c.x1, c.xn2, c.h, c.g, c.Nk = xn1, q, kv, kv_in, Pm, O, x1, xn2, h, g, Nk
        return x2, c

    def block_bwd(self, c, dx2, dx2_bf, B, prev_scale, need_bf, prev_fc2_bias=None):
        """dx2: fp32 grad of the block output (both branches stacked); dx2_bf: bf16 copy already multiplied by this block's MLP
        DropPath scale (its column sums = this block's fc2 bias gradient were accumulated by whoever produced it).
        Returns (dx fp32, dx_bf bf16 scaled by `prev_scale` [2, B] or None); `prev_fc2_bias` names the fc2 bias of the
        PREVIOUS block, whose gradient is the column sum of dx_bf and is folded into the norm1 backward kernel."""
        p, s, H, W = c.p, c.s, c.H, c.W
        C, heads, R = self.dims[s], self.heads[s], self.srs[s]
        gs = self.gs[s]
        GD = dict(groups=2, gs_b=gs)           # data-gradient GEMMs: stacked activations x per-branch weight
        GL = dict(groups=2, param_gs=gs)
        B2 = 2 * B
        d = C // heads
        N = H * W
        M = B2 * N
        Nk = c.Nk
        scale = d ** -0.5
        # ---- Mix-FFN
        self.linear_wgrad(dx2_bf, c.g, p + ".mlp.fc2.weight", gs=gs)
        dg = self.E(M, 4 * C)
        ops.mm(dx2_bf, self.W(p + ".mlp.fc2.weight"), dg, tb=True, **GD)
        du = self.E(M, 4 * C)
        ops.dwconv3x3_bwd_pre(c.h, self.P(p + ".mlp.dwconv.dwconv.weight"), self.P(p + ".mlp.dwconv.dwconv.bias"), ACT_GELU,
                              dg, du, self.G(p + ".mlp.dwconv.dwconv.weight").view(4 * C, 9), self.G(p + ".mlp.dwconv.dwconv.bias"),
                              B, H, W, **GL)
        dh = dg
        ops.dwconv3x3_fwd(du, self.P(p + ".mlp.dwconv.dwconv.weight"), None, ACT_NONE, dh, B, H, W, flip=True,
                          ysum=self.G(p + ".mlp.fc1.bias"), **GL)   # fc1 bias gradient = per-channel sums of dh
        del du
        self.linear_wgrad(dh, c.xn2, p + ".mlp.fc1.weight", gs=gs)
        dxn2 = self.E(M, C)
        ops.mm(dh, self.W(p + ".mlp.fc1.weight"), dxn2, tb=True, **GD)
        del dh, dg
        dx1 = self.E(M, C, dtype=f32)
        dx1_bf = self.E(M, C)
        ops.layernorm_bwd(dxn2, c.x1, c.m2, c.r2, self.P(p + ".norm2.weight"), dres=dx2, dx=dx1, dx_bf=dx1_bf,
                          scale=None if c.dp is None else c.dp[0], rows_per_sample=N, scale_gs=B,
                          dgamma=self.G(p + ".norm2.weight"), dbeta=self.G(p + ".norm2.bias"),
                          dbias=self.G(p + ".attn.proj.bias"), **GL)
        # ---- attention
        self.linear_wgrad(dx1_bf, c.O, p + ".attn.proj.weight", gs=gs)
        dO = self.E(M, C)
        ops.mm(dx1_bf, self.W(p + ".attn.proj.weight"), dO, tb=True, **GD)
        # dK / dV contract over all N tokens into a tiny [Nk, 64] tile per (sample, head): split-K with fp32
        # atomics for parallelism, then one small cast to the bf16 GEMM operand
        bs = (B2, heads)
        Np = (Nk + 7) // 8 * 8
        sP = (heads * N * Np, N * Np)
        tiles = B2 * heads * ((Nk + 127) // 128)
        split = max(1, min(N // 512, (2 * ops.NUM_SMS + tiles - 1) // tiles))
        if tiles >= 96:
            split = 1   # the low-resolution stages have enough (sample, head) tiles: bf16 results straight from the
            #             epilogue - no fp32 scratch, memset or cast on the backward chain
        direct = split == 1
        recompute = getattr(c, "lse", None) is not None
        if recompute:
            direct = False
        dkv = self.E(B2 * Nk, 2 * C) if direct else None
        dkv32 = None if direct else self.Z(B2 * Nk, 2 * C)
        dkv_out = dkv if direct else dkv32
        if recompute:
            # EXPERIMENTAL: dK and dV in one key-major kernel, P recomputed from q, k and the forward's lse
            delta = self.E(B2 * heads * N, dtype=f32)
            ops.attn_delta(dO, c.O, delta, B2, N, heads)
            ops.attn_dkv(c.q, dO, c.kv, c.lse, delta, dkv32, B2, N, Nk, heads, scale)
        else:
            # dV = P^T dO
            ops.gemm_raw(c.Pm, dO, dkv_out, Nk, d, N, Np, C, 2 * C, c_off=C, trans_a=True, trans_b=True, batch=bs, sA=sP,
                         sB=(N * C, d), sC=(Nk * 2 * C, d), accumulate=not direct, split_k=split)
        dq = self.E(M, C)
        dS = None if recompute else self.E(B2 * heads * N, Np)[:, :Nk]
        if recompute and Nk > ops.ATTN_MAX_NK:
            ops.attn_dq_chunked(c.q, dO, c.kv, c.lse, delta, dq, B2, N, Nk, heads, scale)
        elif recompute:
            ops.attn_dq(c.q, dO, c.kv, c.lse, delta, dq, B2, N, Nk, heads, scale)
        elif d == 64 and Nk <= ops.ATTN_MAX_NK and self.fused_attention:
            # fused: dP = dO V^T stays in tensor memory, dS in place of the TMA-loaded P tile, dQ = dS K
            ops.attn_bwd(dO, c.kv, c.Pm, dS, dq, B2, N, Nk, heads, scale)
        else:
            # dP = dO V^T
            dP = self.E(B2 * heads * N, Np, dtype=f32)[:, :Nk]
            ops.gemm_raw(dO, c.kv, dP, N, Nk, d, C, 2 * C, Np, b_off=C, batch=bs, sA=(N * C, d), sB=(Nk * 2 * C, d), sC=sP)
            ops.softmax_rows_bwd(c.Pm, dP, scale, dS)
            del dP
            # dQ = dS K
            ops.gemm_raw(dS, c.kv, dq, N, d, Nk, Np, 2 * C, C, trans_b=True, batch=bs, sA=sP, sB=(Nk * 2 * C, d), sC=(N * C, d))
        if not recompute:
            # dK = dS^T Q
            ops.gemm_raw(dS, c.q, dkv_out, Nk, d, N, Np, C, 2 * C, trans_a=True, trans_b=True, batch=bs, sA=sP, sB=(N * C, d),
                         sC=(Nk * 2 * C, d), accumulate=not direct, split_k=split)
        del dS
        if not direct:
            dkv = self.E(B2 * Nk, 2 * C)
            ops.cast_f32_bf16(dkv32, dkv)
            del dkv32
        self.linear_wgrad(dkv, c.kv_in, p + ".attn.kv.weight", p + ".attn.kv.bias", gs=gs)
        dkvin = self.E(B2 * Nk, C)
        ops.mm(dkv, self.W(p + ".attn.kv.weight"), dkvin, tb=True, **GD)
        if R > 1:
            dsr = self.E(B2 * Nk, C)
            ops.layernorm_bwd(dkvin, c.sr, c.ms, c.rs, self.P(p + ".attn.norm.weight"), dx=dsr,
                              dgamma=self.G(p + ".attn.norm.weight"), dbeta=self.G(p + ".attn.norm.bias"),
                              dbias=self.G(p + ".attn.sr.bias"), **GL)
            wsr = self.packed[p + ".attn.sr.weight"]
            with self._wgrad_ctx(dsr, c.pat):
                ops.mm(dsr, c.pat, self.packed_g[p + ".attn.sr.weight"], ta=True, tb=True, accumulate=True, groups=2,
                       gs_c=self.pk_gs[s])
            dpat = self.E(B2 * Nk, wsr.shape[1])
            ops.mm(dsr, wsr, dpat, tb=True, groups=2, gs_b=self.pk_gs[s])
            dxn_b = self.E(M, C)
            ops.col2im_nhwc(dpat, dxn_b, B2, H, W, R, R, 0, c.Hk, c.Wk)
        else:
            dxn_b = dkvin
        self.linear_wgrad(dq, c.xn1, p + ".attn.q.weight", p + ".attn.q.bias", gs=gs)
        dxn_a = self.E(M, C)
        ops.mm(dq, self.W(p + ".attn.q.weight"), dxn_a, tb=True, **GD)
        dx = self.E(M, C, dtype=f32)
        dx_bf = self.E(M, C) if need_bf else None
        ops.layernorm_bwd(dxn_a, c.x, c.m1, c.r1, self.P(p + ".norm1.weight"), dy2=dxn_b, dres=dx1, dx=dx, dx_bf=dx_bf,
                          scale=prev_scale, rows_per_sample=N, scale_gs=B,
                          dgamma=self.G(p + ".norm1.weight"), dbeta=self.G(p + ".norm1.bias"),
                          dbias=None if prev_fc2_bias is None else self.G(prev_fc2_bias), **GL)
        self._wgrad_join()
        return dx, dx_bf

    # ------------------------------------------------------------------------------------------
    # FRM
    # ------------------------------------------------------------------------------------------
    def frm_fwd(self, s, cat12, B, HW, save):
        p = f"backbone.FRMs.{s}"
        C = self.dims[s]
        M = B * HW
        c = _NS()
        y = self.E(B, 4 * C, dtype=f32)
        am = self.E(B, 2 * C, dtype=torch.int32)
        ops.pool_avgmax_fwd(cat12, y, am, B, HW)   # (two-stage reduction; the wrapper allocates the partials workspace)
        hid = self.E(B, 4 * C, dtype=f32)
        ops.smallm_linear_fwd(y, self.P(p + ".channel_weights.mlp.0.weight"), self.P(p + ".channel_weights.mlp.0.bias"), ACT_RELU, hid)
        cw = self.E(B, 2 * C, dtype=f32)
        ops.smallm_linear_fwd(hid, self.P(p + ".channel_weights.mlp.2.weight"), self.P(p + ".channel_weights.mlp.2.bias"), ACT_SIGMOID, cw)
        t = self.E(M, C)
        ops.mm(cat12, self.W(p + ".spatial_weights.mlp.0.weight"), t, bias=self.P(p + ".spatial_weights.mlp.0.bias"), act=ACT_RELU)
        sw = self.E(M, 2, dtype=f32)
        r12 = self.E(2 * M, C)          # the two rectified branches stacked: the next stage's grouped patch embed reads it whole
        r1, r2 = r12[:M], r12[M:]
        ops.frm_rectify_fwd(cat12, t, self.P(p + ".spatial_weights.mlp.2.weight").view(2, C), self.P(p + ".spatial_weights.mlp.2.bias"),
                            cw, sw, r1, r2, B, HW)
        c.p, c.s, c.cat12, c.y, c.am, c.hid, c.cw, c.t, c.sw = p, s, cat12, y, am, hid, cw, t, sw
        return r12, c

    def frm_bwd(self, c, dr1, dr2, B, HW):
        """dr1/dr2 fp32 [M,C] -> dcat fp32 [M,2C] (grad of the two stage-norm outputs)"""
        p, s = c.p, c.s
        C = self.dims[s]
        M = B * HW
        dcat = self.E(M, 2 * C, dtype=f32)
        dt = self.E(M, C)
        dcw = self.Z(B, 2 * C)
        ops.frm_rectify_bwd(dr1, dr2, c.cat12, c.t, self.P(p + ".spatial_weights.mlp.2.weight").view(2, C), c.cw, c.sw, dcat, dt, dcw,
                            self.G(p + ".spatial_weights.mlp.2.weight").view(2, C), self.G(p + ".spatial_weights.mlp.2.bias"), B, HW)
        self.linear_wgrad(dt, c.cat12, p + ".spatial_weights.mlp.0.weight", p + ".spatial_weights.mlp.0.bias")
        ops.mm(dt, self.W(p + ".spatial_weights.mlp.0.weight"), dcat, tb=True, residual=dcat)
        dhid = self.E(B, 4 * C, dtype=f32)
        ws = self.E(B, 4 * C, dtype=f32)
        ops.smallm_linear_bwd(dcw, c.cw, ACT_SIGMOID, c.hid, self.P(p + ".channel_weights.mlp.2.weight"), dhid,
                              self.G(p + ".channel_weights.mlp.2.weight"), self.G(p + ".channel_weights.mlp.2.bias"), ws)
        dy = self.E(B, 4 * C, dtype=f32)
        ops.smallm_linear_bwd(dhid, c.hid, ACT_RELU, c.y, self.P(p + ".channel_weights.mlp.0.weight"), dy,
                              self.G(p + ".channel_weights.mlp.0.weight"), self.G(p + ".channel_weights.mlp.0.bias"), ws)
        ops.pool_avgmax_bwd(dy, c.am, dcat, B, HW)
        self._wgrad_join()
        return dcat

    # ------------------------------------------------------------------------------------------
    # BatchNorm helper (train: batch statistics + running update; eval: running statistics)
    # ------------------------------------------------------------------------------------------
    def bn_stats(self, prefix, module, x, training):
        """training = the BN MODULE's own flag (a sub-module frozen with .eval() keeps its running statistics)"""
        C = x.shape[1]
        mean, invstd = self.E(C, dtype=f32), self.E(C, dtype=f32)
        if training:
            ws = self.Z(2 * C, dtype=torch.float64)
            ops.colstats(x, ws[:C], ws[C:])
            self.bn_finalize(prefix, module, ws, x.shape[0], mean, invstd)
        else:
            ops.bn_eval_stats(self.buffers[prefix + ".running_mean"], self.buffers[prefix + ".running_var"], module.eps, mean, invstd)
        return mean, invstd

    def bn_finalize(self, prefix, module, ws, count, mean, invstd):
        C = mean.numel()
        rm, rv = self.buffers.get(prefix + ".running_mean"), self.buffers.get(prefix + ".running_var")
        nbt = self.buffers.get(prefix + ".num_batches_tracked")
        if module.momentum is None:
            # torch: cumulative moving average, factor 1 / num_batches_tracked (after the increment)
            raise NotImplementedError("cmx_b200: BatchNorm momentum=None (cumulative average) is not supported")
        if not module.track_running_stats:
            rm = rv = nbt = None
        ops.bn_finalize(ws[:C], ws[C:], count, module.eps, module.momentum, rm, rv, nbt, mean, invstd)

    def sync_bn_world(self, bn, training):
        """number of ranks whose batch statistics this BatchNorm shares (0 = plain per-rank BatchNorm).  Same rule as
        torch.nn.SyncBatchNorm.forward: training mode, initialised process group, world size > 1."""
        import torch.nn as nn
        if not training or not isinstance(bn, nn.SyncBatchNorm):
            return 0
        if self.sync_emulate_world:
            return self.sync_emulate_world
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return 0
        w = dist.get_world_size(self.sync_bn_group(bn))
        return w if w > 1 else 0

    @staticmethod
    def sync_bn_group(bn):
        return getattr(bn, "process_group", None)

    def handle_event(self, ev):
        """execute one event yielded by the *_steps generators on the current stream"""
        if isinstance(ev, tuple) and ev[0] == "allreduce_sum":
            if self.sync_hook is not None:
                self.sync_hook(ev[1])
            else:
                import torch.distributed as dist
                dist.all_reduce(ev[1], op=dist.ReduceOp.SUM, group=ev[2])

    def drive(self, gen):
        try:
            ev = next(gen)
            while True:
                self.handle_event(ev)
                ev = next(gen)
        except StopIteration as done:
            return done.value

    # ------------------------------------------------------------------------------------------
    # FFM
    # ------------------------------------------------------------------------------------------
    def ffm_fwd(self, s, r, B, H, W, training, save):
        p = f"backbone.FFMs.{s}"
        C, heads = self.dims[s], self.heads[s]
        d = C // heads
        N = H * W
        M = B * N
        scale = d ** -0.5
        mod = self.model.backbone.FFMs[s]
        c = _NS()
        c.p, c.s, c.H, c.W, c.r = p, s, H, W, r
        c.yv, c.u, c.kv, c.p32, c.p16 = [], [], [], [], []
        for i in (0, 1):
            wcp = self.W(p + f".cross.channel_proj{i + 1}.weight")
            bcp = self.P(p + f".cross.channel_proj{i + 1}.bias")
            yv = self.E(M, 2 * C)
            ops.mm(r[i], wcp[:C], yv[:, :C], bias=bcp[:C], act=ACT_RELU)
            u = self.E(M, C)
            ops.mm(r[i], wcp[C:], u, bias=bcp[C:], act=ACT_RELU)
            kv = self.E(M, 2 * C)
            ops.mm(u, self.W(p + f".cross.cross_attn.kv{i + 1}.weight"), kv)
            ctx = self.Z(B * heads, d, d)
            ops.gemm_raw(kv, kv, ctx, d, d, N, 2 * C, 2 * C, d, b_off=C, trans_a=True, trans_b=True, batch=(B, heads),
                         sA=(N * 2 * C, d), sB=(N * 2 * C, d), sC=(heads * d * d, d * d), accumulate=True,
                         split_k=max(1, min(16, N // 2048)))
            p32, p16 = self.E(B * heads, d, d, dtype=f32), self.E(B * heads, d, d)
            ops.softmax_dim2_fwd(ctx, scale, p32, p16)
            c.yv.append(yv); c.u.append(u); c.kv.append(kv); c.p32.append(p32); c.p16.append(p16)
        merge = self.E(M, 2 * C)
        c.e, c.me, c.re = [], [], []
        for i in (0, 1):
            # v_i = q_i ctx_{other}
            ops.gemm_raw(c.u[i], c.p16[1 - i], c.yv[i], N, d, d, C, d, 2 * C, c_off=C, trans_b=True, batch=(B, heads),
                         sA=(N * C, d), sB=(heads * d * d, d * d), sC=(N * 2 * C, d))
            e = self.E(M, C, dtype=f32)
            ops.mm(c.yv[i], self.W(p + f".cross.end_proj{i + 1}.weight"), e, bias=self.P(p + f".cross.end_proj{i + 1}.bias"),
                   residual=r[i])
            me, re = (self.E(M, dtype=f32), self.E(M, dtype=f32)) if save else (None, None)
            ops.layernorm_fwd(e, self.P(p + f".cross.norm{i + 1}.weight"), self.P(p + f".cross.norm{i + 1}.bias"), 1e-5,
                              merge[:, i * C:(i + 1) * C], me, re)
            c.e.append(e); c.me.append(me); c.re.append(re)
        q = p + ".channel_emb"
        res = self.E(M, C, dtype=f32)
        ops.mm(merge, self.W(q + ".residual.weight"), res)
        c0 = self.E(M, C)
        ops.mm(merge, self.W(q + ".channel_embed.0.weight"), c0, bias=self.P(q + ".channel_embed.0.bias"))
        c1 = self.E(M, C)
        ops.dwconv3x3_fwd(c0, self.P(q + ".channel_embed.1.weight"), self.P(q + ".channel_embed.1.bias"), ACT_RELU, c1, B, H, W)
        c3 = self.E(M, C, dtype=f32)
        ops.mm(c1, self.W(q + ".channel_embed.3.weight"), c3, bias=self.P(q + ".channel_embed.3.bias"))
        bn1, bn2 = mod.channel_emb.channel_embed[4], mod.channel_emb.norm
        c.mean1, c.inv1 = self.bn_stats(q + ".channel_embed.4", bn1, c3, bn1.training)
        z = self.E(M, C, dtype=f32)
        ops.bn_apply(c3, c.mean1, c.inv1, self.P(q + ".channel_embed.4.weight"), self.P(q + ".channel_embed.4.bias"), z, residual=res)
        c.mean2, c.inv2 = self.bn_stats(q + ".norm", bn2, z, bn2.training)
        out = self.E(M, C)
        ops.bn_apply(z, c.mean2, c.inv2, self.P(q + ".norm.weight"), self.P(q + ".norm.bias"), out)
        c.merge, c.c0, c.c1, c.c3, c.z = merge, c0, c1, c3, z
        return out, c

    def ffm_bwd(self, c, dout, B):
        """dout bf16 [M,C] (grad of the fused feature) -> fp32 [2M, C]: grad of the two rectified inputs, stacked"""
        p, s, H, W = c.p, c.s, c.H, c.W
        C, heads = self.dims[s], self.heads[s]
        d = C // heads
        N = H * W
        M = B * N
        scale = d ** -0.5
        q = p + ".channel_emb"
        ws = self.Z(2 * C, dtype=torch.float64)
        dz = self.E(M, C, dtype=f32)
        ops.bn_bwd(dout, c.z, c.mean2, c.inv2, self.P(q + ".norm.weight"), self.P(q + ".norm.bias"), dz,
                   self.G(q + ".norm.weight"), self.G(q + ".norm.bias"), ws)
        ws2 = self.Z(2 * C, dtype=torch.float64)
        dc3, dz_bf = self.E(M, C), self.E(M, C)
        ops.bn_bwd(dz, c.c3, c.mean1, c.inv1, self.P(q + ".channel_embed.4.weight"), self.P(q + ".channel_embed.4.bias"), dc3,
                   self.G(q + ".channel_embed.4.weight"), self.G(q + ".channel_embed.4.bias"), ws2, dres=dz_bf)
        del dz
        self.linear_wgrad(dc3, c.c1, q + ".channel_embed.3.weight", q + ".channel_embed.3.bias")
        dc1 = self.E(M, C)
        ops.mm(dc3, self.W(q + ".channel_embed.3.weight"), dc1, tb=True)
        du = self.E(M, C)
        ops.dwconv3x3_bwd_pre(c.c0, self.P(q + ".channel_embed.1.weight"), self.P(q + ".channel_embed.1.bias"), ACT_RELU, dc1, du,
                              self.G(q + ".channel_embed.1.weight").view(C, 9), self.G(q + ".channel_embed.1.bias"), B, H, W)
        dc0 = dc1
        ops.dwconv3x3_fwd(du, self.P(q + ".channel_embed.1.weight"), None, ACT_NONE, dc0, B, H, W, flip=True,
                          ysum=self.G(q + ".channel_embed.0.bias"))
        self.linear_wgrad(dc0, c.merge, q + ".channel_embed.0.weight")
        self.linear_wgrad(dz_bf, c.merge, q + ".residual.weight")
        dmerge = self.E(M, 2 * C)
        ops.mm(dc0, self.W(q + ".channel_embed.0.weight"), dmerge, tb=True)
        ops.mm(dz_bf, self.W(q + ".residual.weight"), dmerge, tb=True, residual=dmerge)
        de, dyv = [], []
        for i in (0, 1):
            de_i, de_bf = self.E(M, C, dtype=f32), self.E(M, C)
            ops.layernorm_bwd(dmerge[:, i * C:(i + 1) * C], c.e[i], c.me[i], c.re[i], self.P(p + f".cross.norm{i + 1}.weight"),
                              dx=de_i, dx_bf=de_bf, dgamma=self.G(p + f".cross.norm{i + 1}.weight"),
                              dbeta=self.G(p + f".cross.norm{i + 1}.bias"), dbias=self.G(p + f".cross.end_proj{i + 1}.bias"))
            self.linear_wgrad(de_bf, c.yv[i], p + f".cross.end_proj{i + 1}.weight")
            dyv_i = self.E(M, 2 * C)
            ops.mm(de_bf, self.W(p + f".cross.end_proj{i + 1}.weight"), dyv_i, tb=True)
            de.append(de_i); dyv.append(dyv_i)
        bs = (B, heads)
        sctx = (heads * d * d, d * d)
        split = max(1, min(16, N // 2048))
        dPc = [None, None]
        du = [None, None]
        for i in (0, 1):
            o = 1 - i
            # v_i = q_i ctx_o  =>  dctx_o = q_i^T dv_i ;  du_i = dv_i ctx_o^T
            dp_ = self.Z(B * heads, d, d)
            ops.gemm_raw(c.u[i], dyv[i], dp_, d, d, N, C, 2 * C, d, b_off=C, trans_a=True, trans_b=True, batch=bs,
                         sA=(N * C, d), sB=(N * 2 * C, d), sC=sctx, accumulate=True, split_k=split)
            dPc[o] = dp_
            du_i = self.E(M, C)
            ops.gemm_raw(dyv[i], c.p16[o], du_i, N, d, d, 2 * C, d, C, a_off=C, batch=bs, sA=(N * 2 * C, d), sB=sctx, sC=(N * C, d))
            du[i] = du_i
        drs = self.E(2 * M, C, dtype=f32)   # stacked: the next stage's patch-embed gradient is scattered into it in one launch
        dr = []
        for i in (0, 1):
            dC = self.E(B * heads, d, d)
            ops.softmax_dim2_bwd(c.p32[i], dPc[i], scale, dC)
            dkv = self.E(M, 2 * C)
            # ctx = K^T V : dK = V dC^T ; dV = K dC
            ops.gemm_raw(c.kv[i], dC, dkv, N, d, d, 2 * C, d, 2 * C, a_off=C, batch=bs, sA=(N * 2 * C, d), sB=sctx, sC=(N * 2 * C, d))
            ops.gemm_raw(c.kv[i], dC, dkv, N, d, d, 2 * C, d, 2 * C, c_off=C, trans_b=True, batch=bs, sA=(N * 2 * C, d), sB=sctx,
                         sC=(N * 2 * C, d))
            self.linear_wgrad(dkv, c.u[i], p + f".cross.cross_attn.kv{i + 1}.weight")
            ops.mm(dkv, self.W(p + f".cross.cross_attn.kv{i + 1}.weight"), du[i], tb=True, residual=du[i])
            ops.relu_bwd_(du[i], c.u[i])
            dy_i = dyv[i][:, :C]
            ops.relu_bwd_(dy_i, c.yv[i][:, :C])
            wname, bname = p + f".cross.channel_proj{i + 1}.weight", p + f".cross.channel_proj{i + 1}.bias"
            gw, gb = self.G2(wname), self.G(bname)
            with self._wgrad_ctx(dyv[i], du[i], c.r[i]):
                ops.mm(dy_i, c.r[i], gw[:C], ta=True, tb=True, accumulate=True)
                ops.mm(du[i], c.r[i], gw[C:], ta=True, tb=True, accumulate=True)
                ops.colsum(dy_i, gb[:C])
                ops.colsum(du[i], gb[C:])
            wcp = self.W(wname)
            dr_i = drs[i * M:(i + 1) * M]
            ops.mm(dy_i, wcp[:C], dr_i, tb=True, residual=de[i])
            ops.mm(du[i], wcp[C:], dr_i, tb=True, residual=dr_i)
            dr.append(dr_i)
        self._wgrad_join()
        return drs

    # ------------------------------------------------------------------------------------------
    # decoder + loss
    # ------------------------------------------------------------------------------------------
    def decoder_fwd(self, feats, sizes, B, training, dropmask, save):
        """eager driver of decoder_fwd_steps (a SyncBatchNorm all-reduce, if any, is executed in place)"""
        return self.drive(self.decoder_fwd_steps(feats, sizes, B, training, dropmask, save))

    def decoder_fwd_steps(self, feats, sizes, B, training, dropmask, save):
        """generator (yields only at a SyncBatchNorm all-reduce).  MLPDecoder.py:59-81.  linear_c{s} (a Linear) and its 512-column slice of the 1x1 linear_fuse conv are two
        linear maps with nothing in between (reshape, bilinear upsample and concat are linear and commute), so they
        are folded per step into one [E, C_s] matrix  Wcomb_s = Wf[:, slice_s] @ Wc_s  and one bias
        btot = bf + Wf @ concat(bc4, bc3, bc2, bc1):  z_s = x_s Wcomb_s^T replaces the [M_s, E] x [E, E] GEMM per
        stage (40.3 of the decoder's 42.9 GFLOP/img) and the e_s intermediates; the parameter gradients are recovered
        from dWcomb_s by two tiny GEMMs in decoder_bwd.  The 2048-channel concat is never materialised."""
        E_ = self.embed
        hd = self.model.decode_head
        p = "decode_head"
        c = _NS()
        c.wcomb, c.bcat, btot = self._dec_prep
        self._dec_prep = None
        self._wgrad_join()      # the folded weights were computed on the companion stream while the encoder ran
        c.zs = []
        for s in range(4):
            z = self.E(feats[s].shape[0], E_)
            ops.mm(feats[s], c.wcomb[s], z)
            c.zs.append(z)
        M0 = feats[0].shape[0]
        fuse = self.E(M0, E_, dtype=self.fuse_dtype)
        ops.upsample_sum_fwd(c.zs, sizes, btot.view(E_), fuse, B, E_)
        bn = hd.linear_fuse[1]
        # decoder norm = nn.SyncBatchNorm in every distributed run of the reference (train.py:64-67): the per-channel sums
        # (and therefore mean / biased variance / the unbiased running variance over the GLOBAL batch) are all-reduced
        c.sync = self.sync_bn_world(bn, bn.training)
        if c.sync:
            ws = self.Z(2 * E_, dtype=torch.float64)
            ops.colstats(fuse, ws[:E_], ws[E_:])
            yield ("allreduce_sum", ws, self.sync_bn_group(bn))
            c.mean, c.inv = self.E(E_, dtype=f32), self.E(E_, dtype=f32)
            self.bn_finalize(p + ".linear_fuse.1", bn, ws, fuse.shape[0] * c.sync, c.mean, c.inv)
        else:
            c.mean, c.inv = self.bn_stats(p + ".linear_fuse.1", bn, fuse, bn.training)
        yb = self.E(M0, E_)
        N0 = sizes[0][0] * sizes[0][1]
        ops.bn_apply(fuse, c.mean, c.inv, self.P(p + ".linear_fuse.1.weight"), self.P(p + ".linear_fuse.1.bias"), yb, relu=True,
                     mask=dropmask, rows_per_sample=N0)
        # class axis padded to a multiple of 8 (row stride only: the pad columns are never read) so that the prediction
        # layer and its dgrad/wgrad satisfy the 16-byte row-stride rule of the tcgen05 path
        logits = self.E(M0, self.ncls_ld, dtype=f32)[:, :self.ncls]
        ops.mm(yb, self.W(p + ".linear_pred.weight"), logits, bias=self.P(p + ".linear_pred.bias"))
        c.feats, c.sizes, c.fuse, c.yb, c.dropmask, c.N0 = feats, sizes, fuse, yb, dropmask, N0
        if not save:
            c.zs = None
        self.tr("decode_head.logits", logits)
        return logits, c

    def decoder_prep(self):
        """weights-only part of decoder_fwd (the folded matrices and bias): issued at the start of the step on the
        companion stream, so it overlaps the encoder instead of sitting on the decoder's critical path"""
        E_ = self.embed
        p = "decode_head"
        wf = self.W(p + ".linear_fuse.0.weight")  # [E, 4E]; concat order c4, c3, c2, c1 (MLPDecoder.py:77)
        wcombs = [self.E(E_, self.W(p + f".linear_c{s + 1}.proj.weight").shape[1]) for s in range(4)]
        btot = self.E(1, E_, dtype=f32)
        bcat = self.E(1, 4 * E_, dtype=f32)
        with self._wgrad_ctx(wcombs, btot, bcat):
            for s in range(4):
                ops.mm(wf[:, (3 - s) * E_:(4 - s) * E_], self.W(p + f".linear_c{s + 1}.proj.weight"), wcombs[s], tb=True)
            torch.cat([self.P(p + f".linear_c{s + 1}.proj.bias") for s in (3, 2, 1, 0)], out=bcat.view(4 * E_))
            ops.smallm_linear_fwd(bcat, self.P(p + ".linear_fuse.0.weight").view(E_, 4 * E_), self.P(p + ".linear_fuse.0.bias"), 0, btot)
        self._dec_prep = (wcombs, bcat, btot)

    def decoder_bwd(self, c, dlog, B):
        return self.drive(self.decoder_bwd_steps(c, dlog, B))

    def decoder_bwd_steps(self, c, dlog, B):
        """generator (yields only at the SyncBatchNorm all-reduce).  dlog bf16 [M0, ncls] -> list of df_s bf16 [M_s, C_s]"""
        E_ = self.embed
        p = "decode_head"
        M0 = c.fuse.shape[0]
        self.linear_wgrad(dlog, c.yb, p + ".linear_pred.weight", p + ".linear_pred.bias")
        dyb = self.E(M0, E_)
        ops.mm(dlog, self.W(p + ".linear_pred.weight"), dyb, tb=True)
        ws = self.Z(2 * E_, dtype=torch.float64)
        dfuse = self.E(M0, E_)
        bnargs = (dyb, c.fuse, c.mean, c.inv, self.P(p + ".linear_fuse.1.weight"), self.P(p + ".linear_fuse.1.bias"))
        bnkw = dict(relu=True, mask=c.dropmask, rows_per_sample=c.N0)
        ops.bn_bwd_reduce(*bnargs, ws, **bnkw)
        if c.sync:
            # SyncBatchNorm backward: the two per-channel sums are averaged over the ranks, so that the mean terms of dx are
            # taken over the global batch; the parameter gradients become the rank average of the local sums, which is what
            # DDP's gradient averaging makes of torch's local sums (identical on every rank, so averaging them again is a no-op)
            self._wgrad_join()   # a yield may end a CUDA-graph segment: no side stream may hold unjoined work
            yield ("allreduce_sum", ws, self.sync_bn_group(self.model.decode_head.linear_fuse[1]))
            ws.mul_(1.0 / c.sync)
        ops.bn_bwd_apply(*bnargs, dfuse, self.G(p + ".linear_fuse.1.weight"), self.G(p + ".linear_fuse.1.bias"), ws, **bnkw)
        self.tr("grad.decode_head.logits", dlog)
        self.tr("grad.decode_head.post_bn", dyb)
        self.tr("grad.decode_head.fuse", dfuse)
        del dyb
        c.zs = None
        # bias path (parameter gradients only -> companion stream): d btot = colsum(dfuse);  d bf = d btot;
        # d Wf += d btot (x) bcat;  d bcat = Wf^T d btot
        with self._wgrad_ctx(dfuse, c.bcat):
            dbt = self.Z(1, E_)
            ops.colsum(dfuse, dbt.view(E_))
            dbcat = self.E(1, 4 * E_, dtype=f32)
            ops.smallm_linear_bwd(dbt, dbt, 0, c.bcat, self.P(p + ".linear_fuse.0.weight").view(E_, 4 * E_), dbcat,
                                  self.G(p + ".linear_fuse.0.weight").view(E_, 4 * E_), self.G(p + ".linear_fuse.0.bias"),
                                  self.E(1, E_, dtype=f32))
            for j, s in enumerate((3, 2, 1, 0)):
                self.G(p + f".linear_c{s + 1}.proj.bias").add_(dbcat[0, j * E_:(j + 1) * E_])
        wf = self.W(p + ".linear_fuse.0.weight")
        gf = self.G2(p + ".linear_fuse.0.weight")
        H0, W0 = c.sizes[0]
        dfs = []
        # adjoint of the three upsampled branches in one pass over dfuse (coarsest destination first = band index)
        dzs = [dfuse] + [self.E(c.feats[s].shape[0], E_) for s in (1, 2, 3)]
        if E_ % 64 == 0:
            ops.upsample_bwd_multi(dfuse, H0, W0, dzs[:0:-1], [c.sizes[s] for s in (3, 2, 1)], B, E_)
        else:
            for s in (1, 2, 3):
                ops.upsample_bwd(dfuse, H0, W0, dzs[s], c.sizes[s][0], c.sizes[s][1], B, E_)
        for s in range(4):
            Ms, Cs = c.feats[s].shape
            dz = dzs[s]
            sl = slice((3 - s) * E_, (4 - s) * E_)
            wc = self.W(p + f".linear_c{s + 1}.proj.weight")
            with self._wgrad_ctx(dz, c.feats[s]):
                gcomb = self.Z(E_, Cs)
                ops.mm(dz, c.feats[s], gcomb, ta=True, tb=True, accumulate=True)      # dWcomb_s = dz^T x_s
                gcomb_bf = self.E(E_, Cs)
                ops.cast_f32_bf16(gcomb, gcomb_bf)
                ops.mm(gcomb_bf, wc, gf[:, sl], accumulate=True)                        # dWf[:, slice] += dWcomb Wc^T
                ops.mm(wf[:, sl], gcomb_bf, self.G2(p + f".linear_c{s + 1}.proj.weight"), ta=True, tb=True,
                       accumulate=True)                                                # dWc += Wf[:, slice]^T dWcomb
            df = self.E(Ms, Cs)
            ops.mm(dz, c.wcomb[s], df, tb=True)
            dfs.append(df)
        self._wgrad_join()
        return dfs

    # ------------------------------------------------------------------------------------------
    # whole network
    # ------------------------------------------------------------------------------------------
    def _encode(self, rgb, x, training, save, dp):
        """dual_segformer.py:366-442.  The RGB and X branch of a stage run as ONE chain of grouped launches over the stacked
        tensors [2, B*N, C] (per-branch weights selected by the group stride), not as two chains of half-sized launches."""
        B, H, W = (rgb.shape[0], rgb.shape[1], rgb.shape[2]) if rgb.dtype == torch.uint8 else (rgb.shape[0], rgb.shape[2], rgb.shape[3])
        inp = (rgb, x)
        ctx = _NS()
        ctx.stages = []
        feats, sizes = [], []
        Hc, Wc = H, W
        for s in range(4):
            C = self.dims[s]
            st = _NS()
            st.blocks = []
            x0, Ho, Wo, st.pe = self.pe_fwd(s, inp, B, Hc, Wc, save)
            N = Ho * Wo
            M = B * N
            if self.trace is not None:
                self.tr(f"backbone.patch_embed{s + 1}", x0[:M])
                self.tr(f"backbone.extra_patch_embed{s + 1}", x0[M:])
            xcur = x0
            for i in range(self.depths[s]):
                bp = f"backbone.block{s + 1}.{i}"
                xcur, cb = self.block_fwd(bp, xcur, B, Ho, Wo, s, dp.get(bp), save)
                if self.trace is not None:
                    self.tr(bp, xcur[:M])
                    self.tr(f"backbone.extra_block{s + 1}.{i}", xcur[M:])
                if save:
                    st.blocks.append(cb)
            cat12 = self.E(M, 2 * C)
            st.mn, st.rn = [], []
            for br, nname in enumerate(("norm", "extra_norm")):   # the stage norms write the two column halves of one tensor
                mn, rn = (self.E(M, dtype=f32), self.E(M, dtype=f32)) if save else (None, None)
                ops.layernorm_fwd(xcur[br * M:(br + 1) * M], self.P(f"backbone.{nname}{s + 1}.weight"),
                                  self.P(f"backbone.{nname}{s + 1}.bias"), 1e-6, cat12[:, br * C:(br + 1) * C], mn, rn)
                st.mn.append(mn); st.rn.append(rn)
            st.xs = xcur
            self.tr(f"backbone.norm{s + 1}", cat12[:, :C])
            self.tr(f"backbone.extra_norm{s + 1}", cat12[:, C:])
            r12, st.frm = self.frm_fwd(s, cat12, B, N, save)
            self.tr(f"backbone.FRMs.{s}.out1", r12[:M])
            self.tr(f"backbone.FRMs.{s}.out2", r12[M:])
            (fused, st.ffm), _ = self._on_side(self.ffm_fwd, s, [r12[:M], r12[M:]], B, Ho, Wo, training, save)
            self.tr(f"backbone.FFMs.{s}", fused)
            feats.append(fused)
            sizes.append((Ho, Wo))
            st.H, st.W = Ho, Wo
            ctx.stages.append(st)
            inp = r12
            Hc, Wc = Ho, Wo
        self._side_join()
        return feats, sizes, ctx

    def _side_stream(self):
        if self._side is None or self._side.device != self.dev:
            self._side = torch.cuda.Stream(device=self.dev)
        return self._side

    def _on_side(self, fn, *a, **k):
        """run fn on the side stream, ordered after everything already enqueued on the current stream; returns
        (result, event recorded on the side stream after fn).  Works eagerly and under CUDA-graph capture."""
        if not self.ffm_stream or self.trace is not None:
            return fn(*a, **k), None
        main, side = torch.cuda.current_stream(self.dev), self._side_stream()
        side.wait_stream(main)
        self._side_forked = True
        with torch.cuda.stream(side):
            out = fn(*a, **k)
            ev = torch.cuda.Event()
            ev.record(side)
        return out, ev

    def _side_join(self):
        """the current stream waits for all side-stream work (required before a yield that may end a graph segment)"""
        if self.ffm_stream and self._side_forked:
            torch.cuda.current_stream(self.dev).wait_stream(self._side_stream())
            self._side_forked = False

    def hp_main(self):
        """context manager (CMX_HP_STREAMS=1): run the enclosed step on an engine-owned HIGH-priority stream (forked from and
        joined back into the current stream; the X-branch side stream is high priority too) so that the kernels of the
        dependent forward/backward chains are scheduled ahead of the weight-gradient kernels of the companion streams,
        which keep the default (lowest) priority.  Kernel-node priorities survive CUDA-graph capture.
        Measured on one B200 (scripts/gpu_runs/gpu_run63.sh): 22.75 ms/step against 21.77 with equal priorities - delaying the
        weight-gradient GEMMs pushes them into the tail of each stage where nothing is left to overlap them with - so it
        stays off by default."""
        eng = self

        class _HP:
            def __enter__(self_):
                self_.ctx = None
                if not eng.hp_streams:
                    return
                if eng._hp is None or eng._hp.device != eng.dev:
                    eng._hp = torch.cuda.Stream(device=eng.dev, priority=-1)
                self_.cur = torch.cuda.current_stream(eng.dev)
                eng._hp.wait_stream(self_.cur)
                self_.ctx = torch.cuda.stream(eng._hp)
                self_.ctx.__enter__()

            def __exit__(self_, *exc):
                if self_.ctx is not None:
                    self_.ctx.__exit__(*exc)
                    self_.cur.wait_stream(eng._hp)
                return False
        return _HP()

    def forward_logits(self, rgb, x):
        """eval / inference path: full-resolution NCHW fp32 logits (builder.py:212-238)"""
        self._begin(rgb, x)
        self.decoder_prep()
        training = self.model.training
        B, H, W = (rgb.shape[0], rgb.shape[1], rgb.shape[2]) if rgb.dtype == torch.uint8 else (rgb.shape[0], rgb.shape[2], rgb.shape[3])
        dp, dm = self._make_dp(B, training)
        feats, sizes, _ = self._encode(rgb, x, training, False, dp)
        logits, _ = self.decoder_fwd(feats, sizes, B, training, dm, False)
        out = self.E(B, self.ncls, H, W, dtype=f32)
        ops.logits_upsample_nchw(logits, out, B, sizes[0][0], sizes[0][1], H, W, self.ncls)
        return out

    def forward_loss(self, rgb, x, label, ignore_index, with_grad, focal=None):
        """loss (0-d fp32).  with_grad: also runs the complete backward pass, leaving d loss / d theta in flat_g.
        focal = (w_ce, w_focal, gamma, alpha) selects w_ce * CE + w_focal * FocalLoss instead of plain cross entropy."""
        return self.drive(self.forward_loss_steps(rgb, x, label, ignore_index, with_grad, focal))

    def forward_loss_steps(self, rgb, x, label, ignore_index, with_grad, focal=None):
        """generator form of forward_loss; returns the loss.  Yields (the caller executes the event, e.g. through
        handle_event, and may switch CUDA graphs there: every side stream is joined at a yield):
          ("allreduce_sum", tensor, group)  SyncBatchNorm statistics of the decoder norm (forward: once, backward: once) -
                                            only when the decoder norm layer is an nn.SyncBatchNorm shared by > 1 ranks;
          "early_gradients_ready"           (with_grad only) every gradient of flat_g[:split_off] (decoder, stages 4 and 3)
                                            is final: a data-parallel caller starts the all-reduce of that slice here."""
        self._begin(rgb, x)
        self.decoder_prep()
        training = self.model.training
        B, H, W = (rgb.shape[0], rgb.shape[1], rgb.shape[2]) if rgb.dtype == torch.uint8 else (rgb.shape[0], rgb.shape[2], rgb.shape[3])
        assert B <= 16, "per-GPU batch > 16 is not supported by the FRM small-M kernels"
        label = label.to(torch.int64).contiguous()
        if with_grad:
            for n_, m_ in self._bns:
                if not m_.training:
                    # the backward kernels implement the batch-statistics form; with frozen statistics the mean terms vanish
                    raise NotImplementedError("cmx_b200: gradients through a BatchNorm in eval mode (%s) are not implemented; "
                                              "call model.train() or wrap the call in torch.no_grad()" % n_)
        dp, dm = self._make_dp(B, training)
        feats, sizes, ctx = self._encode(rgb, x, training, with_grad, dp)
        logits, cdec = yield from self.decoder_fwd_steps(feats, sizes, B, training, dm, with_grad)
        h0, w0 = sizes[0]
        acc = self.Z(2, dtype=torch.float64)
        loss = self.E((), dtype=f32)
        dice = focal is not None and focal[0] == "dice"     # ("dice", alpha, smooth): DiceCELoss, two passes over the pixels
        if dice:
            dstats = self.Z(B * 3 * self.ncls, dtype=torch.float64)
            coef = self.E(B * 2 * self.ncls + 1, dtype=f32) if with_grad else None
            ops.dice_ce_stats(logits, label, ignore_index, acc, dstats, B, h0, w0, H, W, self.ncls)
            ops.dice_ce_finalize(acc, dstats, B, self.ncls, focal[1], focal[2], loss, coef)
            if not with_grad:
                return loss
        elif not with_grad:
            if focal is None:
                ops.ce_upsampled(logits, label, ignore_index, acc, None, B, h0, w0, H, W, self.ncls)
            else:
                ops.ce_focal_upsampled(logits, label, ignore_index, acc, None, B, h0, w0, H, W, self.ncls, *focal)
            ops.ce_finalize(acc, loss)
            return loss
        self.flat_g.zero_()
        self.pk_g.zero_()
        dl = self.Z(B * h0 * w0, self.ncls_ld)
        dlog = self.E(B * h0 * w0, self.ncls_ld)
        if dice:
            ops.dice_ce_grad(logits, label, ignore_index, coef, dl[:, :self.ncls], B, h0, w0, H, W, self.ncls)
            ops.cast_f32_bf16(dl, dlog)
        else:
            if focal is None:
                ops.ce_upsampled(logits, label, ignore_index, acc, dl[:, :self.ncls], B, h0, w0, H, W, self.ncls)
            else:
                ops.ce_focal_upsampled(logits, label, ignore_index, acc, dl[:, :self.ncls], B, h0, w0, H, W, self.ncls, *focal)
            ops.ce_finalize(acc, loss, dl, None, dlog)   # element-wise over the padded buffer (pad columns stay 0)
        dfs = yield from self.decoder_bwd_steps(cdec, dlog[:, :self.ncls], B)
        del cdec
        pending = None  # dcol (both branches stacked) of the next stage's patch embeds, to be scattered into this stage's dr
        # the FFM backward of a stage depends on the decoder only: all of them are issued on the side stream now and run
        # concurrently with the block chains of the stages above (when the step is cut at the early-gradient event, the two
        # high-resolution ones are issued after the cut - a graph segment must not end with unjoined side-stream work)
        ffm_out = {}
        for s in ((3, 2) if self.split_at_early else (3, 2, 1, 0)):
            ffm_out[s] = self._on_side(self.ffm_bwd, ctx.stages[s].ffm, dfs[s], B)
        for s in (3, 2, 1, 0):
            st = ctx.stages[s]
            C = self.dims[s]
            N = st.H * st.W
            M = B * N
            drs, ev = ffm_out.pop(s)
            if ev is not None:
                torch.cuda.current_stream(self.dev).wait_event(ev)
            if pending is not None:
                nH, nW = sizes[s + 1]
                ops.col2im_nhwc(pending, drs, 2 * B, st.H, st.W, 3, 2, 1, nH, nW, add=drs)
            dcat = self.frm_bwd(st.frm, drs[:M], drs[M:], B, N)
            blocks = st.blocks
            last = blocks[-1]
            dx = self.E(2 * M, C, dtype=f32)
            dx_bf = self.E(2 * M, C)
            for br, nname in enumerate(("norm", "extra_norm")):
                rows = slice(br * M, (br + 1) * M)
                ops.layernorm_bwd(dcat[:, br * C:(br + 1) * C], st.xs[rows], st.mn[br], st.rn[br],
                                  self.P(f"backbone.{nname}{s + 1}.weight"), dx=dx[rows], dx_bf=dx_bf[rows],
                                  scale=None if last.dp is None else last.dp[1][br], rows_per_sample=N,
                                  dgamma=self.G(f"backbone.{nname}{s + 1}.weight"), dbeta=self.G(f"backbone.{nname}{s + 1}.bias"),
                                  dbias=self.G(last.p.replace("backbone.", "backbone.extra_", 1) + ".mlp.fc2.bias" if br else last.p + ".mlp.fc2.bias"))
            for i in range(len(blocks) - 1, -1, -1):
                prev = blocks[i - 1] if i > 0 else None
                prev_scale = None if (prev is None or prev.dp is None) else prev.dp[1]
                dx, dx_bf = self.block_bwd(blocks[i], dx, dx_bf, B, prev_scale, need_bf=(i > 0),
                                           prev_fc2_bias=None if prev is None else prev.p + ".mlp.fc2.bias")
                blocks[i] = None
            pending = self.pe_bwd(st.pe, dx, B)
            if s == 0:
                pending = None
            ctx.stages[s] = None
            if s == 2:
                # every weight-gradient stream has been joined: add the packed conv gradients of stages 3-4 into flat_g
                if self.n_convs_early:
                    ops.convw_unpack_grad_multi(self.conv_table, self.n_convs_early)
                if self.split_at_early:
                    self._side_join()
                yield "early_gradients_ready"
                if self.split_at_early:
                    for s2 in (1, 0):
                        ffm_out[s2] = self._on_side(self.ffm_bwd, ctx.stages[s2].ffm, dfs[s2], B)
        self._side_join()
        if self.n_convs > self.n_convs_early:
            ops.convw_unpack_grad_multi(self.conv_table[56 * self.n_convs_early:], self.n_convs - self.n_convs_early)
        return loss

    def _begin(self, rgb, x):
        if not rgb.is_cuda:
            raise RuntimeError("cmx_b200: inputs must be CUDA tensors — the hot path has no CPU fallback")
        if rgb.dtype == torch.uint8:   # raw images: [B,H,W,3] + [B,H,W,3] or grey [B,H,W] (input pipeline fused into patch embed 1)
            assert x.dtype == torch.uint8 and rgb.dim() == 4 and rgb.shape[3] == 3 and tuple(x.shape[:3]) == tuple(rgb.shape[:3]) and \
                (x.dim() == 3 or tuple(x.shape) == tuple(rgb.shape)), "expected uint8 [B,H,W,3] and [B,H,W] / [B,H,W,3] inputs"
            assert rgb.is_contiguous() and x.is_contiguous()
        else:
            assert rgb.shape == x.shape and rgb.dim() == 4 and rgb.shape[1] == 3, "expected two [B,3,H,W] inputs"
        if rgb.device.type == "cuda" and torch.cuda.current_device() != rgb.device.index:
            # the C launchers enqueue on the current device's stream and never call cudaSetDevice
            raise RuntimeError("cmx_b200: inputs are on %s but the current CUDA device is %d - call torch.cuda.set_device(%d) "
                               "(one process per GPU, as train.py does)" % (rgb.device, torch.cuda.current_device(), rgb.device.index))
        self.dev = rgb.device
        self._ensure_flat(rgb.device)
        self.refresh_weights()
