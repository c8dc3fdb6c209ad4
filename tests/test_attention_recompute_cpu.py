"""CPU restatement of the arithmetic of the EXPERIMENTAL flash-style attention backward (csrc/attention_dkv.cu), with the
kernels' rounding points (bf16 q / kv / dO / O operands, fp32 scores, P^T and dS^T rounded to bf16 before the gradient
MMAs, fp32 accumulation, +inf normaliser / zero probability for padded queries and keys), checked against fp32 autograd of
softmax(scale q k^T) v (dual_segformer.py:127-134).  It pins the tolerance the gated GPU parity test uses
(tests/test_ops_gpu.py::test_attention_dkv_recompute: 2 % of the gradient's max-abs) - no CUDA code runs here."""
import pytest
import torch

bf = torch.bfloat16
LOG2E = 1.4426950408889634


def _emulate(q, kv, d_o, o, lse, B, N, Nk, heads, scale, BK=128, BQ=128):
    d = 64
    C = heads * d
    qf = q.float().view(B, N, heads, d).permute(0, 2, 1, 3)
    dof = d_o.float().view(B, N, heads, d).permute(0, 2, 1, 3)
    kf = kv.float().view(B, Nk, 2, heads, d)[:, :, 0].permute(0, 2, 1, 3)
    vf = kv.float().view(B, Nk, 2, heads, d)[:, :, 1].permute(0, 2, 1, 3)
    delta = (dof * o.float().view(B, N, heads, d).permute(0, 2, 1, 3)).sum(-1)            # cmx_attn_delta
    nkb, ntq = (Nk + BK - 1) // BK, (N + BQ - 1) // BQ
    pad = lambda t, n: torch.cat([t, t.new_zeros(*t.shape[:2], n - t.shape[2], d)], 2)      # TMA zero fill
    qp, dop, kp, vp = pad(qf, ntq * BQ), pad(dof, ntq * BQ), pad(kf, nkb * BK), pad(vf, nkb * BK)
    l2 = torch.full((B, heads, ntq * BQ), float("inf"))
    l2[:, :, :N] = lse * LOG2E
    dl = torch.zeros(B, heads, ntq * BQ)
    dl[:, :, :N] = delta
    key_ok = (torch.arange(nkb * BK) < Nk).float()
    dk, dv, dq = torch.zeros_like(kp), torch.zeros_like(vp), torch.zeros_like(qp)
    for j in range(nkb):
        ks = slice(j * BK, (j + 1) * BK)
        for i in range(ntq):
            qs = slice(i * BQ, (i + 1) * BQ)
            st = kp[:, :, ks] @ qp[:, :, qs].transpose(-1, -2)                               # S^T  [keys, queries]
            dpt = vp[:, :, ks] @ dop[:, :, qs].transpose(-1, -2)                             # dP^T
            pt = torch.exp2(st * (scale * LOG2E) - l2[:, :, None, qs]) * key_ok[ks, None]
            dst = scale * pt * (dpt - dl[:, :, None, qs])
            pt_b, dst_b = pt.to(bf).float(), dst.to(bf).float()
            dv[:, :, ks] += pt_b @ dop[:, :, qs]
            dk[:, :, ks] += dst_b @ qp[:, :, qs]
            dq[:, :, qs] += dst_b.transpose(-1, -2) @ kp[:, :, ks]                           # the query-major kernel's product
    return dq[:, :, :N].to(bf).float(), dk[:, :, :Nk], dv[:, :, :Nk]


@pytest.mark.parametrize("B,N,Nk,heads", [(1, 130, 4, 2), (1, 333, 77, 1), (2, 600, 300, 2)])
def test_recompute_backward_matches_autograd_within_the_gpu_test_tolerance(B, N, Nk, heads):
    torch.manual_seed(12)
    d, C, scale = 64, heads * 64, 64 ** -0.5
    q, kv, d_o = (torch.randn(B * N, C).to(bf), torch.randn(B * Nk, 2 * C).to(bf), torch.randn(B * N, C).to(bf))
    qf = q.float().view(B, N, heads, d).permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    kf = kv.float().view(B, Nk, 2, heads, d)[:, :, 0].permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    vf = kv.float().view(B, Nk, 2, heads, d)[:, :, 1].permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    s = (qf @ kf.transpose(-1, -2)) * scale
    of = torch.softmax(s, -1) @ vf
    of.backward(d_o.float().view(B, N, heads, d).permute(0, 2, 1, 3))
    o = of.detach().permute(0, 2, 1, 3).reshape(B * N, C).to(bf)
    lse = torch.logsumexp(s.detach(), -1)
    dq, dk, dv = _emulate(q, kv, d_o, o, lse, B, N, Nk, heads, scale)
    for got, ref, what in ((dq, qf.grad, "dQ"), (dk, kf.grad, "dK"), (dv, vf.grad, "dV")):
        err = (got - ref).abs()
        tol = 2e-2 * ref.abs() + 2e-2 * float(ref.abs().max())
        assert bool((err <= tol).all()), "%s: max err %.4g vs max |ref| %.4g" % (what, float(err.max()), float(ref.abs().max()))
        # and it is not a vacuous bound: the emulation is much closer than the tolerance
        assert float(err.max()) < 0.5 * 2e-2 * float(ref.abs().max()), what
