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
    ops.mm(a, b, outb, bias=bias, act=ops.ACT_GELU, residual=res.to(bf), impl=2)
    _check(outb, torch.nn.functional.gelu(base + bias) + res.to(bf).float(), K, "gelu + bf16 residual")
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


def test_auto_dispatch_prefers_tc():
    assert ops.gemm_which(1000, 64, 64, 64, 64, 64) == 2
    assert ops.gemm_which(1000, 9, 512, 512, 512, 9) == 1
