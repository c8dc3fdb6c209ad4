"""Host-orchestration dry run of the whole training / inference step on CPU (no kernel runs, nothing is computed).

The engine (rgbx_semantic_segmentation_b200/engine.py) is ~1700 C-ABI calls per step that are otherwise only exercised on a
GPU box.  Here the ctypes library is replaced by a mock built from include/cmx_b200.h that checks every call against the
header prototype (arity, pointer / integer / float convertibility) and returns 0, and the inputs are a Tensor subclass
that reports `is_cuda` so the product's loud "no CPU fallback" check lets the dry run through.  This is test
infrastructure, not a CPU path: outputs are uninitialised memory and are never looked at.  It catches Python-level
mistakes (wrong argument lists, undefined names, shape bookkeeping) in GPU-only code before a GPU call is spent on them,
for the default path and for the experimental CMX_ATTN_DKV_RECOMPUTE path."""
import collections

import pytest
import torch
import torch.nn as nn

from rgbx_semantic_segmentation_b200 import _lib, ops
from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder


class _Cfg:
    backbone = "mit_b1"          # dims [64, 128, 320, 512], heads [1, 2, 5, 8] => head_dim 64: the fused attention path
    decoder = "MLPDecoder"
    decoder_embed_dim = 64
    num_classes = 9
    pretrained_model = None
    bn_eps = 1e-3
    bn_momentum = 0.1


class _ReportsCuda(torch.Tensor):
    is_cuda = property(lambda self: True)


@pytest.fixture
def mock_lib(monkeypatch):
    protos = _lib.parse_header()
    calls = collections.Counter()

    class Mock:
        def __getattr__(self, name):
            if name not in protos:
                raise AttributeError("%s is not declared in include/cmx_b200.h" % name)
            restype, argtypes = protos[name]

            def fn(*args):
                assert len(args) == len(argtypes), "%s: called with %d arguments, the header declares %d" % (name, len(args), len(argtypes))
                for i, (a, t) in enumerate(zip(args, argtypes)):
                    try:
                        t.from_param(a)
                    except Exception as e:  # noqa: BLE001 - report which argument
                        raise AssertionError("%s: argument %d (%r) does not convert to %s (%s)" % (name, i, a, t.__name__, e))
                calls[name] += 1
                return 0
            return fn
    mock = Mock()
    monkeypatch.setattr(_lib, "load", lambda: mock)
    monkeypatch.setattr(ops, "_cuda", lambda *ts: None)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    # CUDA streams do not exist on CPU: single-stream orchestration (the stream fork/join helpers are no-ops when disabled)
    monkeypatch.setenv("CMX_WGRAD_STREAM", "0")
    monkeypatch.setenv("CMX_FFM_STREAM", "0")
    return calls


def _model_and_inputs():
    torch.manual_seed(0)
    m = EncoderDecoder(cfg=_Cfg, criterion=nn.CrossEntropyLoss(reduction='mean', ignore_index=255), norm_layer=nn.BatchNorm2d)
    B, H, W = 2, 64, 96
    rgb = torch.randn(B, 3, H, W).as_subclass(_ReportsCuda)
    x = torch.randn(B, 3, H, W).as_subclass(_ReportsCuda)
    lab = torch.randint(0, 9, (B, H, W))
    return m, rgb, x, lab


@pytest.mark.parametrize("recompute", ["0", "1"])
def test_training_step_orchestration(mock_lib, monkeypatch, recompute):
    monkeypatch.setenv("CMX_ATTN_DKV_RECOMPUTE", recompute)
    m, rgb, x, lab = _model_and_inputs()
    m.train()
    eng = m._eng()
    loss = eng.forward_loss(rgb, x, lab, 255, with_grad=True, focal=None)
    assert loss.numel() == 1 and loss.dtype == torch.float32
    assert eng.flat_g.shape == eng.flat_p.shape and 0 < eng.split_off < eng.flat_g.numel()
    c = mock_lib
    n_attn = sum(m.backbone.depths)                          # one grouped launch per (RGB block, X twin) pair
    assert c["cmx_attn_fwd"] == n_attn
    assert c["cmx_layernorm_fwd"] == c["cmx_layernorm_bwd"] and c["cmx_dwconv3x3_fwd"] == 2 * c["cmx_dwconv3x3_bwd_pre"]
    assert c["cmx_ce_upsampled_fwd_bwd"] == 1 and c["cmx_upsample_sum_fwd"] == 1 and c["cmx_upsample_bwd_multi"] == 1
    if recompute == "1":
        assert c["cmx_attn_delta"] == c["cmx_attn_dkv"] == c["cmx_attn_dq"] == n_attn and c["cmx_attn_bwd"] == 0
    else:
        assert c["cmx_attn_bwd"] == n_attn and c["cmx_attn_dkv"] == 0 and c["cmx_attn_dq"] == 0 and c["cmx_attn_delta"] == 0


def test_recompute_mode_drops_the_probability_tensors_and_two_gemms_per_block(mock_lib, monkeypatch):
    counts = {}
    for flag in ("0", "1"):
        mock_lib.clear()
        monkeypatch.setenv("CMX_ATTN_DKV_RECOMPUTE", flag)
        m, rgb, x, lab = _model_and_inputs()
        m.train()
        m._eng().forward_loss(rgb, x, lab, 255, with_grad=True, focal=None)
        counts[flag] = dict(mock_lib)
    n_attn = sum(m.backbone.depths)
    assert counts["0"]["cmx_gemm"] - counts["1"]["cmx_gemm"] == 2 * n_attn          # dV = P^T dO and dK = dS^T Q


def test_inference_and_focal_orchestration(mock_lib):
    m, rgb, x, lab = _model_and_inputs()
    m.eval()
    eng = m._eng()
    out = eng.forward_logits(rgb, x)
    assert tuple(out.shape) == (2, 9, 64, 96) and out.dtype == torch.float32
    assert mock_lib["cmx_attn_fwd"] == sum(m.backbone.depths) and mock_lib["cmx_attn_bwd"] == 0 and mock_lib["cmx_layernorm_bwd"] == 0
    mock_lib.clear()
    m.train()
    loss = eng.forward_loss(rgb, x, lab, 255, with_grad=True, focal=(1.0, 0.2, 2.0, 0.25))   # the CE_Focal tuple (builder.py:246-247)
    assert loss.numel() == 1 and mock_lib["cmx_ce_focal_upsampled_fwd_bwd"] == 1 and mock_lib["cmx_ce_upsampled_fwd_bwd"] == 0


def test_autograd_handoff_through_the_module_call(mock_lib, monkeypatch):
    """loss = model(rgb, x, label); loss.backward(): the fused step hands every parameter an fp32 gradient that is a view of
    the engine's flat gradient buffer (what DDP's bucket hooks and FlatAdamW rely on; train.py:186-201)."""
    monkeypatch.setenv("CMX_CUDA_GRAPH", "0")
    m, rgb, x, lab = _model_and_inputs()
    m.train()
    loss = m(rgb, x, lab)
    assert loss.requires_grad and loss.numel() == 1
    loss.backward()
    n = 0
    for name, p in m.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape and p.grad.dtype == torch.float32, name
        n += 1
    assert n == len(m._eng().names)
    # parameters now live in the flat buffer; a second step reuses it (no re-flattening)
    base = m._eng().flat_p.data_ptr()
    m(rgb, x, lab).backward()
    assert m._eng().flat_p.data_ptr() == base
    # no_grad / eval call: logits at input resolution, no backward kernels
    mock_lib.clear()
    m.eval()
    with torch.no_grad():
        out = m(rgb, x)
    assert tuple(out.shape) == (2, 9, 64, 96) and mock_lib["cmx_layernorm_bwd"] == 0


@pytest.mark.parametrize("recompute", ["0", "1"])
def test_head_dim_32_takes_the_unfused_attention_path(mock_lib, monkeypatch, recompute):
    """mit_b0 (head_dim 32, odd sizes): S-materialising batched GEMM + softmax path, whatever the experimental flag says"""
    monkeypatch.setenv("CMX_ATTN_DKV_RECOMPUTE", recompute)

    class Cfg(_Cfg):
        pass
    Cfg.backbone = "mit_b0"
    torch.manual_seed(0)
    m = EncoderDecoder(cfg=Cfg, criterion=nn.CrossEntropyLoss(reduction='mean', ignore_index=255), norm_layer=nn.BatchNorm2d)
    m.train()
    rgb = torch.randn(1, 3, 72, 104).as_subclass(_ReportsCuda)
    x = torch.randn(1, 3, 72, 104).as_subclass(_ReportsCuda)
    lab = torch.randint(0, 9, (1, 72, 104))
    loss = m._eng().forward_loss(rgb, x, lab, 255, with_grad=True, focal=None)
    assert loss.numel() == 1
    n_attn = sum(m.backbone.depths)
    assert mock_lib["cmx_attn_fwd"] == 0 and mock_lib["cmx_attn_dkv"] == 0 and mock_lib["cmx_attn_dq"] == 0
    assert mock_lib["cmx_softmax_rows_fwd"] == mock_lib["cmx_softmax_rows_bwd"] == n_attn


def test_long_key_axis_takes_the_key_chunked_fused_path(mock_lib):
    """mit_b1 at 608x608: Nkv = 361 > 320 in every stage -> fused kernel per key chunk + combine in forward, recompute kernels
    (dkv once, dq per chunk + sum) in backward; no score / probability tensor, no softmax kernels"""
    class Cfg(_Cfg):
        pass
    Cfg.backbone = "mit_b1"
    torch.manual_seed(0)
    m = EncoderDecoder(cfg=Cfg, criterion=nn.CrossEntropyLoss(reduction='mean', ignore_index=255), norm_layer=nn.BatchNorm2d)
    m.train()
    rgb = torch.randn(1, 3, 608, 608).as_subclass(_ReportsCuda)
    x = torch.randn(1, 3, 608, 608).as_subclass(_ReportsCuda)
    lab = torch.randint(0, 9, (1, 608, 608))
    loss = m._eng().forward_loss(rgb, x, lab, 255, with_grad=True, focal=None)
    assert loss.numel() == 1
    n_attn = sum(m.backbone.depths)
    assert mock_lib["cmx_attn_fwd"] == 2 * n_attn and mock_lib["cmx_attn_combine"] == n_attn
    assert mock_lib["cmx_attn_dkv"] == n_attn and mock_lib["cmx_attn_dq"] == 2 * n_attn and mock_lib["cmx_sum_parts_bf16"] == n_attn
    assert mock_lib["cmx_softmax_rows_fwd"] == 0 and mock_lib["cmx_softmax_rows_bwd"] == 0 and mock_lib["cmx_attn_bwd"] == 0


def test_sync_batchnorm_orchestration_two_allreduce_events(mock_lib):
    """norm_layer=nn.SyncBatchNorm (train.py:64-67): the step generator yields exactly one all-reduce of the decoder norm's
    (sum, sumsq) in forward and one of (sum dy, sum dy*xhat) in backward - both double[2*E] - before the early-gradient
    event; with a plain BatchNorm2d (or a single rank) it yields neither."""
    for norm, emulate, want in ((nn.SyncBatchNorm, 2, 2), (nn.SyncBatchNorm, 0, 0), (nn.BatchNorm2d, 2, 0)):
        torch.manual_seed(0)
        m = EncoderDecoder(cfg=_Cfg, criterion=nn.CrossEntropyLoss(reduction='mean', ignore_index=255), norm_layer=norm)
        _, rgb, x, lab = _model_and_inputs()
        m.train()
        eng = m._eng()
        eng.sync_emulate_world = emulate
        events = list(_events(eng.forward_loss_steps(rgb, x, lab, 255, with_grad=True)))
        sync = [e for e in events if isinstance(e, tuple) and e[0] == "allreduce_sum"]
        assert len(sync) == want, (norm.__name__, emulate, events)
        assert events[-1] == "early_gradients_ready"
        for e in sync:
            assert e[1].dtype == torch.float64 and e[1].numel() == 2 * _Cfg.decoder_embed_dim


def _events(gen):
    try:
        while True:
            yield next(gen)
    except StopIteration:
        return
