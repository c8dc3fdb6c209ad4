"""GPU parity of the tcgen05/TMA GEMM (impl=2 forces the tensor-memory path) against a PyTorch fp32
reference of the same bf16 operands: forward (K-major x K-major), dgrad (B MN-major) and wgrad
(A and B MN-major, split-K fp32 accumulate), with every epilogue feature."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from rgbx_semantic_segmentation_b200 import ops

DEV = "cuda"
bf = torch.bfloat16


@pytest.fixture(autouse=True, params=["auto", "pad"])
def tile_policy(request, monkeypatch):
    """every case runs under both tile-width policies: "auto" (the two-CTA-per-SM BN = 64 / 128 shapes the engine uses)
    and "pad" (least padding: also exercises the whole-SM BN = 160 / 256 instantiations)"""
    monkeypatch.setenv("CMX_GEMM_TILE_POLICY", request.param)
    yield request.param


def _ref(a, b, ta, tb):
    A = a.float().t() if ta else a.float()
    Bm = b.float() if tb else b.float().t()
    return A @ Bm


def _check(out, ref, K, what):
    err = (out.float() - ref).abs().max().item()
    tol = 2e-2 * max(1.0, ref.abs().max().item()) if out.dtype == bf else 1e-3 * K ** 0.5
    assert err <= tol, "%s: max err %.4g > %.4g" % (what, err, tol)


FWD = [(128, 64, 64), (300, 64, 64), (1000, 128, 128), (520, 320, 320), (256, 512, 2048), (384, 1280, 320),
       (4800, 256, 64), (130, 2048, 512), (200, 160, 160), (640, 640, 320), (19200, 64, 256), (77, 1024, 512), (333, 72, 152)]


@pytest.mark.parametrize("M,N,K", FWD)
def test_tc_forward_plain(M, N, K):
    torch.manual_seed(0)
    a = torch.randn(M, K, device=DEV).to(bf)
    b = torch.randn(N, K, device=DEV).to(bf)
    out = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.mm(a, b, out, impl=2)
    _check(out, _ref(a, b, False, False), K, "tc fwd %dx%dx%d" % (M, N, K))


@pytest.mark.parametrize("M,N,K", [(200, 300, 64), (130, 64, 300), (257, 300, 300)])
def test_tc_forward_ragged(M, N, K):
    """N and K that are not multiples of 8 (Nkv = 300): padded leading dimensions, TMA zero fill, scalar tail stores"""
    torch.manual_seed(6)
    Kp, Np = (K + 7) // 8 * 8, (N + 7) // 8 * 8
    a = torch.randn(M, Kp, device=DEV).to(bf)[:, :K]
    b = torch.randn(N, Kp, device=DEV).to(bf)[:, :K]
    out = torch.full((M, Np), 3.0, device=DEV)[:, :N]
    ops.mm(a, b, out, impl=2)
    _check(out, a.float() @ b.float().t(), K, "ragged")
    if Np > N:
        assert float((out.as_strided((M, Np - N), (Np, 1), N) - 3.0).abs().max()) == 0.0   # pad columns untouched


@pytest.mark.parametrize("M,N,K", [(300, 64, 64), (520, 320, 320), (1000, 256, 128)])
def test_tc_forward_epilogue(M, N, K):
    torch.manual_seed(1)
    a = torch.randn(M, K, device=DEV).to(bf)
    b = torch.randn(N, K, device=DEV).to(bf)
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    rps = 100
    scale = torch.rand((M + rps - 1) // rps, device=DEV) * 2
    sc = scale.repeat_interleave(rps)[:M, None]
    base = _ref(a, b, False, False)
    out = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.mm(a, b, out, bias=bias, residual=res, row_scale=scale, rows_per_sample=rps, alpha=0.5, impl=2)
    _check(out, res + sc * (0.5 * base + bias), K, "bias+res+scale")
    outb = torch.empty(M, N, device=DEV, dtype=bf)
    ops.mm(a, b, outb, bias=bias, act=ops.ACT_RELU, impl=2)
    _check(outb, torch.relu(base + bias), K, "relu bf16 out")
    ops.mm(a, b, outb, bias=bias, act=ops.ACT_RELU, residual=res.to(bf), impl=2)
    _check(outb, torch.relu(base + bias) + res.to(bf).float(), K, "relu + bf16 residual")
    # strided output / strided A (views into wider buffers)
    wide = torch.zeros(M, 2 * N, device=DEV, dtype=bf)
    ops.mm(a, b, wide[:, N:], impl=2)
    _check(wide[:, N:], base, K, "strided C")
    assert float(wide[:, :N].abs().max()) == 0.0
    awide = torch.randn(M, 2 * K, device=DEV).to(bf)
    ops.mm(awide[:, K:], b, out, impl=2)
    _check(out, awide[:, K:].float() @ b.float().t(), K, "strided A")


@pytest.mark.parametrize("M,N,K", [(300, 64, 64), (1000, 64, 256), (520, 320, 1280), (256, 2048, 512), (4800, 128, 256), (200, 576, 64)])
def test_tc_dgrad_b_mn_major(M, N, K):
    """dX[M,N] = dY[M,K] @ W[K,N]  with W stored [K,N] (= nn.Linear weight [out=K, in=N])"""
    torch.manual_seed(2)
    a = torch.randn(M, K, device=DEV).to(bf)
    w = torch.randn(K, N, device=DEV).to(bf)
    out = torch.empty(M, N, device=DEV, dtype=bf)
    ops.mm(a, w, out, tb=True, impl=2)
    _check(out, _ref(a, w, False, True), K, "tc dgrad")


@pytest.mark.parametrize("M,N,K", [(64, 64, 5000), (256, 64, 19200), (320, 1280, 2400), (512, 2048, 600), (128, 152, 3000), (2048, 512, 777)])
def test_tc_wgrad_mn_major_splitk(M, N, K):
    """dW[M,N] += dY[K,M]^T @ X[K,N]  (both operands token-major as stored), fp32 atomic split-K"""
    torch.manual_seed(3)
    a = torch.randn(K, M, device=DEV).to(bf)
    b = torch.randn(K, N, device=DEV).to(bf)
    out = torch.ones(M, N, device=DEV, dtype=torch.float32)
    ops.mm(a, b, out, ta=True, tb=True, accumulate=True, impl=2)
    _check(out, 1 + _ref(a, b, True, True), K, "tc wgrad")


@pytest.mark.parametrize("N,Nk,heads", [(1200, 300, 5), (400, 300, 2), (333, 77, 1)])
def test_tc_batched_attention_shapes(N, Nk, heads):
    """the six batched GEMMs of (spatial-reduction) self-attention forward + backward on strided head views,
    Nkv = 300 (ragged: not a multiple of 8 -> padded leading dimension, TMA zero fill)"""
    torch.manual_seed(4)
    B, d = 2, 64
    C = heads * d
    Np = (Nk + 7) // 8 * 8
    q = torch.randn(B * N, C, device=DEV).to(bf)
    kv = torch.randn(B * Nk, 2 * C, device=DEV).to(bf)
    dO = torch.randn(B * N, C, device=DEV).to(bf)
    qf = q.float().view(B, N, heads, d).permute(0, 2, 1, 3)
    kf = kv.float().view(B, Nk, 2, heads, d)[:, :, 0].permute(0, 2, 1, 3)
    vf = kv.float().view(B, Nk, 2, heads, d)[:, :, 1].permute(0, 2, 1, 3)
    dOf = dO.float().view(B, N, heads, d).permute(0, 2, 1, 3)
    bs = (B, heads)
    sP = (heads * N * Np, N * Np)
    S = torch.full((B * heads * N, Np), 7.0, device=DEV)[:, :Nk]
    ops.gemm_raw(q, kv, S, N, Nk, d, C, 2 * C, Np, batch=bs, sA=(N * C, d), sB=(Nk * 2 * C, d), sC=sP, alpha=0.125, impl=2)
    _check(S.reshape(B, heads, N, Nk), 0.125 * qf @ kf.transpose(-1, -2), d, "S = QK^T")
    P = torch.softmax(S.float(), -1)
    Pb = torch.zeros(B * heads * N, Np, device=DEV, dtype=bf)[:, :Nk]
    Pb.copy_(P)
    Pf = Pb.float().reshape(B, heads, N, Nk)
    O = torch.empty(B * N, C, device=DEV, dtype=bf)
    ops.gemm_raw(Pb, kv, O, N, d, Nk, Np, 2 * C, C, b_off=C, trans_b=True, batch=bs, sA=sP, sB=(Nk * 2 * C, d), sC=(N * C, d), impl=2)
    _check(O.view(B, N, heads, d).permute(0, 2, 1, 3), Pf @ vf, Nk, "O = PV")
    dkv = torch.zeros(B * Nk, 2 * C, device=DEV)
    ops.gemm_raw(Pb, dO, dkv, Nk, d, N, Np, C, 2 * C, c_off=C, trans_a=True, trans_b=True, batch=bs, sA=sP, sB=(N * C, d),
                 sC=(Nk * 2 * C, d), accumulate=True, split_k=3, impl=2)
    ops.gemm_raw(Pb, q, dkv, Nk, d, N, Np, C, 2 * C, trans_a=True, trans_b=True, batch=bs, sA=sP, sB=(N * C, d),
                 sC=(Nk * 2 * C, d), accumulate=True, split_k=2, impl=2)
    got = dkv.view(B, Nk, 2, heads, d)
    _check(got[:, :, 1].permute(0, 2, 1, 3), Pf.transpose(-1, -2) @ dOf, N, "dV = P^T dO")
    _check(got[:, :, 0].permute(0, 2, 1, 3), Pf.transpose(-1, -2) @ qf, N, "dK-like = P^T Q")
    # the low-resolution stages skip split-K: bf16 results straight from the epilogue (TMA store clipped at Nk rows)
    dkvb = torch.full((B * Nk, 2 * C), 3.0, device=DEV, dtype=bf)
    ops.gemm_raw(Pb, dO, dkvb, Nk, d, N, Np, C, 2 * C, c_off=C, trans_a=True, trans_b=True, batch=bs, sA=sP, sB=(N * C, d),
                 sC=(Nk * 2 * C, d), impl=2)
    ops.gemm_raw(Pb, q, dkvb, Nk, d, N, Np, C, 2 * C, trans_a=True, trans_b=True, batch=bs, sA=sP, sB=(N * C, d),
                 sC=(Nk * 2 * C, d), impl=2)
    gb = dkvb.view(B, Nk, 2, heads, d)
    _check(gb[:, :, 1].permute(0, 2, 1, 3), Pf.transpose(-1, -2) @ dOf, N, "dV = P^T dO (bf16, no split)")
    _check(gb[:, :, 0].permute(0, 2, 1, 3), Pf.transpose(-1, -2) @ qf, N, "dK-like = P^T Q (bf16, no split)")
    dP = torch.empty(B * heads * N, Np, device=DEV)[:, :Nk]
    ops.gemm_raw(dO, kv, dP, N, Nk, d, C, 2 * C, Np, b_off=C, batch=bs, sA=(N * C, d), sB=(Nk * 2 * C, d), sC=sP, impl=2)
    _check(dP.reshape(B, heads, N, Nk), dOf @ vf.transpose(-1, -2), d, "dP = dO V^T")
    dq = torch.empty(B * N, C, device=DEV, dtype=bf)
    ops.gemm_raw(Pb, kv, dq, N, d, Nk, Np, 2 * C, C, trans_b=True, batch=bs, sA=sP, sB=(Nk * 2 * C, d), sC=(N * C, d), impl=2)
    _check(dq.view(B, N, heads, d).permute(0, 2, 1, 3), Pf @ kf, Nk, "dQ-like = P K")


def test_tc_batched_ffm_context():
    """FFM cross-attention context: ctx[b,h] = K^T V over all tokens (64x64 output, split-K), then q @ ctx"""
    torch.manual_seed(5)
    B, N, heads, d = 2, 2500, 2, 64
    C = heads * d
    kv = torch.randn(B * N, 2 * C, device=DEV).to(bf)
    ctx = torch.zeros(B * heads, d, d, device=DEV)
    ops.gemm_raw(kv, kv, ctx, d, d, N, 2 * C, 2 * C, d, b_off=C, trans_a=True, trans_b=True, batch=(B, heads),
                 sA=(N * 2 * C, d), sB=(N * 2 * C, d), sC=(heads * d * d, d * d), accumulate=True, split_k=4, impl=2)
    k = kv.float().view(B, N, 2, heads, d)[:, :, 0].permute(0, 2, 1, 3)
    v = kv.float().view(B, N, 2, heads, d)[:, :, 1].permute(0, 2, 1, 3)
    _check(ctx.view(B, heads, d, d), k.transpose(-1, -2) @ v, N, "ctx = K^T V")
    p16 = torch.softmax(ctx * 0.125, dim=-2).to(bf)
    u = torch.randn(B * N, C, device=DEV).to(bf)
    yv = torch.zeros(B * N, 2 * C, device=DEV, dtype=bf)
    ops.gemm_raw(u, p16, yv, N, d, d, C, d, 2 * C, c_off=C, trans_b=True, batch=(B, heads), sA=(N * C, d),
                 sB=(heads * d * d, d * d), sC=(N * 2 * C, d), impl=2)
    ref = u.float().view(B, N, heads, d).permute(0, 2, 1, 3) @ p16.float().view(B, heads, d, d)
    _check(yv[:, C:].reshape(B, N, heads, d).permute(0, 2, 1, 3), ref, d, "v = q ctx")
    assert float(yv[:, :C].abs().max()) == 0.0


def test_auto_dispatch_prefers_tc():
    assert ops.gemm_which(1000, 64, 64, 64, 64, 64) == 2
    assert ops.gemm_which(1000, 9, 512, 512, 512, 9) == 1
