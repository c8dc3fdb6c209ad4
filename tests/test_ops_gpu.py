"""GPU parity tests, kernel by kernel, through the C ABI (ctypes) against plain PyTorch fp32 references
of the same op on the same seeded inputs.  Tolerances: bf16 storage => 2^-8 relative on outputs of
magnitude ~1; fp32 paths 1e-4/1e-5; integer paths bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from rgbx_semantic_segmentation_b200 import ops

DEV = "cuda"
bf = torch.bfloat16


def rnd(*shape, scale=1.0, dtype=torch.float32, seed=None):
    if seed is not None:
        torch.manual_seed(seed)
    return (torch.randn(*shape, device=DEV) * scale).to(dtype)


def close(a, b, rtol, atol, what=""):
    a, b = a.float(), b.float()
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bad.any(), "%s: %d/%d mismatches, max err %.4g (ref max %.4g)" % (
        what, int(bad.sum()), bad.numel(), float(err.max()), float(b.abs().max()))


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("impl", [1])
@pytest.mark.parametrize("M,N,K,ta,tb", [
    (300, 64, 64, False, False), (130, 9, 512, False, False), (257, 72, 100, False, False),
    (192, 96, 333, True, True), (100, 64, 40, False, True), (64, 2, 64, False, False)])
def test_gemm_fallback(M, N, K, ta, tb, impl):
    torch.manual_seed(0)
    a = rnd(K, M, dtype=bf) if ta else rnd(M, K, dtype=bf)
    b = rnd(K, N, dtype=bf) if tb else rnd(N, K, dtype=bf)
    bias = rnd(N)
    res = rnd(M, N)
    out = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.mm(a, b, out, ta=ta, tb=tb, bias=bias, residual=res, act=ops.ACT_RELU, alpha=0.5, impl=impl)
    A = a.float().t() if ta else a.float()
    Bm = b.float() if tb else b.float().t()
    ref = res + torch.relu(0.5 * (A @ Bm) + bias)
    close(out, ref, 1e-4, 1e-3 * K ** 0.5 / 8, "gemm fallback")


def test_gemm_fallback_batched_splitk():
    torch.manual_seed(1)
    B1, B2, M, N, K = 2, 3, 70, 40, 900
    a = rnd(B1, B2, K, M, dtype=bf)   # stored [K, M]  (trans_a)
    b = rnd(B1, B2, K, N, dtype=bf)   # stored [K, N]  (trans_b)
    out = torch.zeros(B1, B2, M, N, device=DEV)
    ops.gemm_raw(a, b, out, M, N, K, M, N, N, trans_a=True, trans_b=True, batch=(B1, B2), sA=(B2 * K * M, K * M),
                 sB=(B2 * K * N, K * N), sC=(B2 * M * N, M * N), split_k=4, accumulate=True, alpha=0.125)
    ref = 0.125 * torch.matmul(a.float().transpose(-1, -2), b.float())
    close(out, ref, 1e-4, 2e-3, "batched split-K")


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C,xdt,ydt", [(64, torch.float32, bf), (320, torch.float32, torch.float32), (512, bf, bf), (160, bf, torch.float32)])
def test_layernorm_fwd_bwd(C, xdt, ydt):
    torch.manual_seed(2)
    M = 777
    x = rnd(M, C, scale=2.0).add_(0.5).to(xdt)
    g = (1 + 0.1 * rnd(C))
    b = 0.1 * rnd(C)
    y = torch.empty(M, C, device=DEV, dtype=ydt)
    mean = torch.empty(M, device=DEV)
    rstd = torch.empty(M, device=DEV)
    ops.layernorm_fwd(x, g, b, 1e-6, y, mean, rstd)
    xr = x.float().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (C,), gr, br, 1e-6)
    close(y, ref, 1e-2 if ydt == bf else 1e-5, 1e-2 if ydt == bf else 1e-5, "ln fwd")
    # backward with two upstream grads, residual grad, per-sample scale on the bf16 copy
    dy = rnd(M, C, dtype=bf)
    dy2 = rnd(M, C, dtype=bf)
    dres = rnd(M, C)
    rows_per_sample = 111
    scale = torch.tensor([1.0, 0.0, 1.5, 1.0, 2.0, 1.0, 0.5], device=DEV)
    dx = torch.empty(M, C, device=DEV)
    dxbf = torch.empty(M, C, device=DEV, dtype=bf)
    dg = torch.zeros(C, device=DEV)
    db = torch.zeros(C, device=DEV)
    dbias = torch.zeros(C, device=DEV)
    ops.layernorm_bwd(dy, x, mean, rstd, g, dy2=dy2, dres=dres, dx=dx, dx_bf=dxbf, scale=scale, rows_per_sample=rows_per_sample,
                      dgamma=dg, dbeta=db, dbias=dbias)
    close(dbias, dxbf.float().sum(0), 1e-2, 5e-2 * M ** 0.5, "ln fused bias-gradient column sum")
    up = dy.float() + dy2.float()
    ref.backward(up)
    close(dx, xr.grad + dres, 1e-3, 2e-3, "ln dx")
    sc = scale.repeat_interleave(rows_per_sample)[:M, None]
    close(dxbf, (xr.grad + dres) * sc, 1e-2, 2e-2, "ln dx bf16")
    close(dg, gr.grad, 1e-3, 1e-2, "ln dgamma")
    close(db, br.grad, 1e-3, 1e-2, "ln dbeta")


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,HW,C,xdt,use_res", [(3, 50, 64, torch.float32, True), (2, 1201, 320, bf, False),
                                                 (8, 300, 512, torch.float32, True), (3, 77, 24, torch.float32, True),
                                                 (1, 5, 128, bf, True)])
def test_batchnorm_train_fwd_bwd(B, HW, C, xdt, use_res):
    torch.manual_seed(3)
    M = B * HW
    x = (rnd(M, C, scale=1.5) + 0.3).to(xdt)
    res = rnd(M, C, dtype=bf) if use_res else None
    g = 1 + 0.1 * rnd(C)
    b = 0.1 * rnd(C)
    mask = ((torch.rand(B, C, device=DEV) > 0.2).float() / 0.8)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    ws = torch.zeros(2 * C, dtype=torch.float64, device=DEV)
    mean, invstd = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    ops.colstats(x, ws[:C], ws[C:])
    ops.bn_finalize(ws[:C], ws[C:], M, 1e-3, 0.1, rm, rv, nbt, mean, invstd)
    y = torch.empty(M, C, device=DEV, dtype=bf)
    ops.bn_apply(x, mean, invstd, g, b, y, residual=res, relu=True, mask=mask, rows_per_sample=HW)
    xr = x.float().clone().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    xb = F.batch_norm(xr.t().reshape(1, C, M), rm2, rv2, gr, br, True, 0.1, 1e-3).reshape(C, M).t()
    pre = xb + (res.float() if use_res else 0.)
    ref = torch.relu(pre) * mask.repeat_interleave(HW, 0)
    close(y, ref, 1e-2, 1e-2, "bn apply")
    close(rm, rm2, 1e-5, 1e-6, "running mean")
    close(rv, rv2, 1e-5, 1e-6, "running var")
    assert int(nbt) == 1
    dy = rnd(M, C, dtype=bf)
    ref.backward(dy.float())
    dx = torch.empty(M, C, device=DEV)
    dres = torch.empty(M, C, device=DEV)
    dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    ws.zero_()
    ops.bn_bwd(dy, x, mean, invstd, g, b, dx, dg, db, ws, residual=res, relu=True, mask=mask, rows_per_sample=HW, dres=dres)
    close(dx, xr.grad, 1e-3, 1e-3, "bn dx")
    close(dres, dy.float() * mask.repeat_interleave(HW, 0) * (pre > 0), 1e-6, 1e-6, "bn effective upstream gradient")
    close(dg, gr.grad, 1e-3, 1e-2, "bn dgamma")
    close(db, br.grad, 1e-3, 1e-2, "bn dbeta")


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("act", [2, 1, 0])
@pytest.mark.parametrize("B,H,W,C", [(2, 15, 20, 256), (1, 7, 5, 64), (2, 30, 41, 1280)])
def test_dwconv(act, B, H, W, C):
    torch.manual_seed(4)
    x = rnd(B * H * W, C, dtype=bf)
    w = rnd(C, 1, 3, 3, scale=0.3)
    bias = rnd(C, scale=0.1)
    y = torch.empty_like(x)
    ops.dwconv3x3_fwd(x, w, bias, act, y, B, H, W)
    xr = x.float().reshape(B, H, W, C).permute(0, 3, 1, 2).requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    u = F.conv2d(xr, wr, br, padding=1, groups=C)
    ref = F.gelu(u) if act == 2 else (F.relu(u) if act == 1 else u)
    close(y, ref.permute(0, 2, 3, 1).reshape(-1, C), 1e-2, 1e-2, "dwconv fwd")
    dy = rnd(B * H * W, C, dtype=bf)
    ref.backward(dy.float().reshape(B, H, W, C).permute(0, 3, 1, 2))
    du = torch.empty_like(x)
    dw, db = torch.zeros(C, 9, device=DEV), torch.zeros(C, device=DEV)
    ops.dwconv3x3_bwd_pre(x, w, bias, act, dy, du, dw, db, B, H, W)
    dx = torch.empty_like(x)
    ysum = torch.zeros(C, device=DEV)
    ops.dwconv3x3_fwd(du, w, None, 0, dx, B, H, W, flip=True, ysum=ysum)
    close(ysum, dx.float().sum(0), 1e-3, 1e-2, "dwconv dgrad fused column sum")
    close(dx, xr.grad.permute(0, 2, 3, 1).reshape(-1, C), 2e-2, 2e-2, "dwconv dx")
    n = (B * H * W) ** 0.5
    close(dw.reshape(C, 1, 3, 3), wr.grad, 1e-2, 2e-2 * n, "dwconv dw")
    close(db, br.grad, 1e-2, 2e-2 * n, "dwconv db")


# ---------------------------------------------------------------------------------------------
def test_im2col_paths():
    torch.manual_seed(5)
    B, H, W = 2, 37, 50
    img = rnd(B, 3, H, W)
    Ho, Wo = (H + 6 - 7) // 4 + 1, (W + 6 - 7) // 4 + 1
    col = torch.empty(B * Ho * Wo, 160, device=DEV, dtype=bf)
    ops.im2col_nchw(img, col, 7, 4, 3, Ho, Wo)
    w = rnd(64, 3, 7, 7, scale=0.1)
    wp = torch.empty(64, 160, device=DEV, dtype=bf)
    ops.convw_pack(w, wp)
    got = col.float() @ wp.float().t()
    ref = F.conv2d(img.to(bf).float(), w.to(bf).float(), stride=4, padding=3).permute(0, 2, 3, 1).reshape(-1, 64)
    close(got, ref, 1e-3, 1e-3, "im2col_nchw + pack == conv")
    # NHWC 3x3 s2 p1 and its adjoint
    C = 64
    x = rnd(B * H * W, C, dtype=bf)
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    col = torch.empty(B * Ho * Wo, 9 * C, device=DEV, dtype=bf)
    ops.im2col_nhwc(x, col, B, H, W, 3, 2, 1, Ho, Wo)
    xr = x.float().reshape(B, H, W, C).permute(0, 3, 1, 2).requires_grad_(True)
    unf = F.unfold(xr, 3, padding=1, stride=2)  # [B, C*9, L] with (c, kh, kw) ordering
    ref = unf.reshape(B, C, 9, Ho * Wo).permute(0, 3, 2, 1).reshape(B * Ho * Wo, 9 * C)
    assert torch.equal(col.float(), ref.detach())
    dcol = rnd(B * Ho * Wo, 9 * C, dtype=bf)
    ref.backward(dcol.float())
    add = rnd(B * H * W, C)
    dx = torch.empty(B * H * W, C, device=DEV)
    ops.col2im_nhwc(dcol, dx, B, H, W, 3, 2, 1, Ho, Wo, add=add)
    close(dx, xr.grad.permute(0, 2, 3, 1).reshape(-1, C) + add, 1e-5, 1e-5, "col2im")
    # SR patchify k=s=4, H not divisible
    Hk, Wk = (H - 4) // 4 + 1, (W - 4) // 4 + 1
    col = torch.empty(B * Hk * Wk, 16 * C, device=DEV, dtype=bf)
    ops.im2col_nhwc(x, col, B, H, W, 4, 4, 0, Hk, Wk)
    unf = F.unfold(x.float().reshape(B, H, W, C).permute(0, 3, 1, 2), 4, stride=4)
    ref = unf.reshape(B, C, 16, Hk * Wk).permute(0, 3, 2, 1).reshape(B * Hk * Wk, 16 * C)
    assert torch.equal(col.float(), ref)
    # conv weight grad unpack
    gp = rnd(64, 160)
    gw = torch.ones(64, 3, 7, 7, device=DEV)
    ops.convw_unpack_grad(gp, gw)
    ref = 1 + gp[:, :147].reshape(64, 7, 7, 3).permute(0, 3, 1, 2)
    close(gw, ref, 0, 1e-6, "unpack grad")
    # the one-launch variants over a device descriptor table (what the engine uses) == the per-conv calls
    import struct
    shapes = [(64, 3, 7, 7), (128, 64, 3, 3), (64, 64, 8, 8), (40, 24, 2, 2)]
    ws = [rnd(*sh, scale=0.1) for sh in shapes]
    kp = [(sh[1] * sh[2] * sh[3] + 7) // 8 * 8 for sh in shapes]
    wps = [torch.full((sh[0], k), 7.0, device=DEV, dtype=bf) for sh, k in zip(shapes, kp)]
    gps = [rnd(sh[0], k) for sh, k in zip(shapes, kp)]
    gws = [rnd(*sh) for sh in shapes]
    gws0 = [g.clone() for g in gws]
    blob = b"".join(struct.pack("<QQQQiiiiii", w.data_ptr(), wp.data_ptr(), gp.data_ptr(), gw.data_ptr(), *sh, k, 0)
                    for w, wp, gp, gw, sh, k in zip(ws, wps, gps, gws, shapes, kp))
    table = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(DEV)
    ops.convw_pack_multi(table, len(shapes))
    ops.convw_unpack_grad_multi(table, len(shapes))
    for w, wp, gp, gw, gw0, sh, k in zip(ws, wps, gps, gws, gws0, shapes, kp):
        one = torch.empty_like(wp)
        ops.convw_pack(w, one)
        assert torch.equal(wp, one)
        ops.convw_unpack_grad(gp, gw0)
        assert torch.equal(gw, gw0)


def test_casts_colsum_relu():
    torch.manual_seed(6)
    x = rnd(1003)
    pad = torch.empty(1008, device=DEV, dtype=bf)
    ops.cast_f32_bf16(x, pad[:1003])
    assert torch.equal(pad[:1003], x.to(bf))
    m = rnd(999, 72, dtype=bf)
    out = torch.ones(72, device=DEV)
    ops.colsum(m, out)
    close(out, 1 + m.float().sum(0), 1e-4, 1e-3, "colsum")
    y = rnd(40, 64, dtype=bf)
    dy = rnd(40, 64, dtype=bf)
    ref = dy.float() * (y.float() > 0)
    ops.relu_bwd_(dy, y)
    assert torch.equal(dy.float(), ref)


def test_softmax_kernels():
    torch.manual_seed(7)
    rows, n = 1000, 300
    s = rnd(rows, n, scale=3.0)
    p = torch.empty(rows, n, device=DEV, dtype=bf)
    ops.softmax_rows_fwd(s, p)
    ref = torch.softmax(s, -1)
    close(p, ref, 1e-2, 1e-4, "softmax rows")
    dp = rnd(rows, n)
    ds = torch.empty(rows, n, device=DEV, dtype=bf)
    ops.softmax_rows_bwd(p, dp, 0.125, ds)
    pf = p.float()
    refd = 0.125 * pf * (dp - (pf * dp).sum(-1, keepdim=True))
    close(ds, refd, 1e-2, 1e-4, "softmax rows bwd")
    nb, d = 6, 64
    c = rnd(nb, d, d, scale=20.0)
    p32 = torch.empty(nb, d, d, device=DEV)
    p16 = torch.empty(nb, d, d, device=DEV, dtype=bf)
    ops.softmax_dim2_fwd(c, 0.125, p32, p16)
    ref = torch.softmax(c * 0.125, dim=-2)
    close(p32, ref, 1e-4, 1e-6, "softmax dim2")
    close(p16, ref, 1e-2, 1e-5, "softmax dim2 bf16")
    dpp = rnd(nb, d, d)
    dc = torch.empty(nb, d, d, device=DEV, dtype=bf)
    ops.softmax_dim2_bwd(p32, dpp, 0.125, dc)
    refd = 0.125 * p32 * (dpp - (p32 * dpp).sum(-2, keepdim=True))
    close(dc, refd, 1e-2, 1e-5, "softmax dim2 bwd")


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,HW,C", [(3, 77, 64), (2, 301, 32), (2, 1200, 320), (8, 300, 512), (1, 53, 48)])
def test_frm_kernels(B, HW, C):
    torch.manual_seed(8)
    M = B * HW
    a = rnd(M, 2 * C, dtype=bf)
    y = torch.empty(B, 4 * C, device=DEV)
    am = torch.empty(B, 2 * C, device=DEV, dtype=torch.int32)
    ops.pool_avgmax_fwd(a, y, am, B, HW)
    af = a.float().reshape(B, HW, 2 * C)
    close(y[:, :2 * C], af.mean(1), 1e-5, 1e-5, "avg pool")
    assert torch.equal(y[:, 2 * C:], af.amax(1))
    assert torch.equal(am.long(), af.argmax(1)) or torch.equal(torch.gather(af, 1, am.long()[:, None])[:, 0], af.amax(1))
    # small-M MLP
    w0, b0 = rnd(4 * C, 4 * C, scale=0.1), rnd(4 * C, scale=0.1)
    hid = torch.empty(B, 4 * C, device=DEV)
    ops.smallm_linear_fwd(y, w0, b0, 1, hid)
    yr = y.clone().requires_grad_(True)
    w0r, b0r = w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    refh = torch.relu(F.linear(yr, w0r, b0r))
    close(hid, refh, 1e-4, 1e-4, "smallm fwd")
    dh = rnd(B, 4 * C)
    refh.backward(dh)
    dx = torch.empty(B, 4 * C, device=DEV)
    dw, db = torch.zeros_like(w0), torch.zeros_like(b0)
    ws = torch.empty(B, 4 * C, device=DEV)
    ops.smallm_linear_bwd(dh, hid, 1, y, w0, dx, dw, db, ws)
    close(dx, yr.grad, 1e-4, 1e-4, "smallm dx")
    close(dw, w0r.grad, 1e-4, 1e-4, "smallm dw")
    close(db, b0r.grad, 1e-4, 1e-4, "smallm db")
    # pool backward
    dyp = rnd(B, 4 * C)
    dxa = torch.zeros(M, 2 * C, device=DEV)
    ops.pool_avgmax_bwd(dyp, am, dxa, B, HW)
    ar = a.float().reshape(B, HW, 2 * C).requires_grad_(True)
    # the reference pools with nn.AdaptiveMaxPool2d (net_utils.py:15): gradient goes to the FIRST maximum
    mx = F.adaptive_max_pool2d(ar.permute(0, 2, 1).reshape(B, 2 * C, HW, 1), 1).reshape(B, 2 * C)
    (torch.cat([ar.mean(1), mx], 1) * dyp).sum().backward()
    close(dxa, ar.grad.reshape(M, 2 * C), 1e-5, 1e-6, "pool bwd")
    # rectify
    t = torch.relu(rnd(M, C)).to(bf)
    w2, b2 = rnd(2, C, scale=0.2), rnd(2, scale=0.1)
    cw = torch.sigmoid(rnd(B, 2 * C))
    sw = torch.empty(M, 2, device=DEV)
    r1 = torch.empty(M, C, device=DEV, dtype=bf)
    r2 = torch.empty(M, C, device=DEV, dtype=bf)
    ops.frm_rectify_fwd(a, t, w2, b2, cw, sw, r1, r2, B, HW)
    a1 = a[:, :C].float().clone().requires_grad_(True)
    a2 = a[:, C:].float().clone().requires_grad_(True)
    tr = t.float().clone().requires_grad_(True)
    w2r, b2r, cwr = w2.clone().requires_grad_(True), b2.clone().requires_grad_(True), cw.clone().requires_grad_(True)
    swr = torch.sigmoid(tr @ w2r.t() + b2r)
    cwm = cwr.repeat_interleave(HW, 0)
    o1 = a1 + 0.5 * cwm[:, C:] * a2 + 0.5 * swr[:, 1:2] * a2
    o2 = a2 + 0.5 * cwm[:, :C] * a1 + 0.5 * swr[:, 0:1] * a1
    close(sw, swr, 1e-4, 1e-5, "spatial weights")
    close(r1, o1, 1e-2, 1e-2, "rectify out1")
    close(r2, o2, 1e-2, 1e-2, "rectify out2")
    d1, d2 = rnd(M, C), rnd(M, C)
    (o1 * d1 + o2 * d2).sum().backward()
    da = torch.empty(M, 2 * C, device=DEV)
    dt = torch.empty(M, C, device=DEV, dtype=bf)
    dcw, dw2, db2 = torch.zeros(B, 2 * C, device=DEV), torch.zeros(2, C, device=DEV), torch.zeros(2, device=DEV)
    ops.frm_rectify_bwd(d1, d2, a, t, w2, cw, sw, da, dt, dcw, dw2, db2, B, HW)
    close(da[:, :C], a1.grad, 1e-4, 1e-4, "rectify da1")
    close(da[:, C:], a2.grad, 1e-4, 1e-4, "rectify da2")
    close(dt, tr.grad * (t.float() > 0), 1e-2, 1e-3, "rectify dt")
    close(dcw, cwr.grad, 1e-3, 1e-3, "rectify dcw")
    close(dw2, w2r.grad, 1e-3, 1e-3, "rectify dw2")
    close(db2, b2r.grad, 1e-3, 1e-3, "rectify db2")


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sizes", [[(24, 32), (12, 16), (6, 8), (3, 4)], [(45, 80), (23, 40), (12, 20), (6, 10)]])
def test_upsample_sum_and_adjoint(sizes):
    torch.manual_seed(9)
    B, C = 2, 64
    zs = [rnd(B * h * w, C, dtype=bf) for h, w in sizes]
    bias = rnd(C)
    H0, W0 = sizes[0]
    out = torch.empty(B * H0 * W0, C, device=DEV)
    ops.upsample_sum_fwd(zs, sizes, bias, out, B, C)
    zr = [z.float().reshape(B, h, w, C).permute(0, 3, 1, 2).requires_grad_(True) for z, (h, w) in zip(zs, sizes)]
    ref = zr[0] + bias[None, :, None, None]
    for z in zr[1:]:
        ref = ref + F.interpolate(z, size=(H0, W0), mode="bilinear", align_corners=False)
    close(out, ref.permute(0, 2, 3, 1).reshape(-1, C), 1e-5, 1e-5, "upsample sum")
    out_bf = torch.empty(B * H0 * W0, C, device=DEV, dtype=bf)     # the engine's bf16 BatchNorm-input variant
    ops.upsample_sum_fwd(zs, sizes, bias, out_bf, B, C)
    assert torch.equal(out_bf, out.to(bf)), "bf16 output must be the rounded fp32 result"
    dout = rnd(B * H0 * W0, C, dtype=bf)
    ref.backward(dout.float().reshape(B, H0, W0, C).permute(0, 3, 1, 2))
    for i in (1, 2, 3):
        h, w = sizes[i]
        dz = torch.empty(B * h * w, C, device=DEV, dtype=bf)
        ops.upsample_bwd(dout, H0, W0, dz, h, w, B, C)
        close(dz, zr[i].grad.permute(0, 2, 3, 1).reshape(-1, C), 1e-2, 2e-2, "upsample adjoint %d" % i)
    # the three adjoints in one pass over dout (any order / subset of destinations)
    for order in ((3, 2, 1), (1, 2, 3), (2,), (3, 1)):
        dzs = [torch.full((B * sizes[i][0] * sizes[i][1], C), float("nan"), device=DEV, dtype=bf) for i in order]
        ops.upsample_bwd_multi(dout, H0, W0, dzs, [sizes[i] for i in order], B, C)
        for dz, i in zip(dzs, order):
            close(dz, zr[i].grad.permute(0, 2, 3, 1).reshape(-1, C), 1e-2, 2e-2, "fused upsample adjoint %d" % i)


@pytest.mark.parametrize("h,w,H,W,ncls,ld", [(12, 16, 48, 64, 9, 16), (23, 40, 90, 160, 5, 8), (5, 7, 19, 27, 9, 9)])
def test_ce_upsampled(h, w, H, W, ncls, ld):
    """ld > ncls: the engine's padded class axis (pad columns hold garbage and must be neither read nor written)"""
    torch.manual_seed(10)
    B = 2
    base = torch.full((B * h * w, ld), float("nan"), device=DEV)
    logits = base[:, :ncls]
    logits.copy_(rnd(B * h * w, ncls, scale=2.0))
    label = torch.randint(0, ncls, (B, H, W), device=DEV)
    label[torch.rand(B, H, W, device=DEV) < 0.1] = 255
    label[:, :3] = 255
    acc = torch.zeros(2, dtype=torch.float64, device=DEV)
    dl_full = torch.zeros(B * h * w, ld, device=DEV)
    dl = dl_full[:, :ncls]
    ops.ce_upsampled(logits, label, 255, acc, dl, B, h, w, H, W, ncls)
    assert float(dl_full[:, ncls:].abs().sum()) == 0.0
    loss = torch.empty((), device=DEV)
    gs = torch.tensor(2.0, device=DEV)
    dout = torch.empty_like(dl_full)
    ops.ce_finalize(acc, loss, dl_full, gs, dout)
    lr = logits.reshape(B, h, w, ncls).permute(0, 3, 1, 2).clone().requires_grad_(True)
    up = F.interpolate(lr, size=(H, W), mode="bilinear", align_corners=False)
    ref = F.cross_entropy(up, label, ignore_index=255)
    (2.0 * ref).backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1, abs(float(ref)))
    assert float(acc[1]) == float((label != 255).sum())
    close(dout[:, :ncls], lr.grad.permute(0, 2, 3, 1).reshape(-1, ncls), 1e-3, 1e-6, "ce dlogits")
    full = torch.empty(B, ncls, H, W, device=DEV)
    ops.logits_upsample_nchw(logits, full, B, h, w, H, W, ncls)
    close(full, up, 1e-5, 1e-5, "logits upsample")


@pytest.mark.parametrize("name,mode", [("g2", "focal"), ("g4", "focal"), ("g1", "focal"), ("g2", "cefocal"), ("g4", "cefocal")])
def test_focal_loss_kernel_vs_reference_golden(golden_dir, name, mode):
    """fused FocalLoss / CE_Focal (train.py:70-93) against values produced by the reference's FocalLoss class; identity
    'upsample' (H = h) so the golden logits are the kernel's logits"""
    import os
    z = np.load(os.path.join(golden_dir, "focal.npz"))
    ncls, gamma, alpha, _ = z[name + "_meta"]
    ncls = int(ncls)
    lg = torch.from_numpy(z[name + "_logits"]).to(DEV)                     # [B, C, h, w]
    B, _, h, w = lg.shape
    ld = (ncls + 7) // 8 * 8
    base = torch.full((B * h * w, ld), float("nan"), device=DEV)
    logits = base[:, :ncls]
    logits.copy_(lg.permute(0, 2, 3, 1).reshape(-1, ncls))
    label = torch.from_numpy(z[name + "_target"]).to(DEV)
    acc = torch.zeros(2, dtype=torch.float64, device=DEV)
    dl_full = torch.zeros(B * h * w, ld, device=DEV)
    w_ce, w_f = (0.0, 1.0) if mode == "focal" else (1.0, 0.2)
    ops.ce_focal_upsampled(logits, label, 255, acc, dl_full[:, :ncls], B, h, w, h, w, ncls, w_ce, w_f, float(gamma), float(alpha))
    loss = torch.empty((), device=DEV)
    dout = torch.empty_like(dl_full)
    ops.ce_finalize(acc, loss, dl_full, None, dout)
    ref_loss = float(z[name + "_" + mode])
    assert abs(float(loss) - ref_loss) < 2e-5 * max(1.0, abs(ref_loss)), (float(loss), ref_loss)
    ref_grad = torch.from_numpy(z[name + "_" + mode + "_grad"]).to(DEV).permute(0, 2, 3, 1).reshape(-1, ncls)
    close(dout[:, :ncls], ref_grad, 2e-3, 1e-7, "focal dlogits")


@pytest.mark.parametrize("name", ["d5", "d9", "d40"])
def test_dice_ce_loss_kernels_vs_reference_golden(golden_dir, name):
    """fused DiceCELoss (train.py:79-80; utils/loss_opr.py:103-156) against values produced by the reference class; identity
    'upsample' (H = h) so the golden logits are the kernel's logits; d40 = the 40-class (NYUDv2 default) register-array variant"""
    import os
    z = np.load(os.path.join(golden_dir, "dice.npz"))
    ncls, alpha, _ = z[name + "_meta"]
    ncls = int(ncls)
    lg = torch.from_numpy(z[name + "_logits"]).to(DEV)
    B, _, h, w = lg.shape
    ld = (ncls + 7) // 8 * 8
    base = torch.full((B * h * w, ld), float("nan"), device=DEV)
    logits = base[:, :ncls]
    logits.copy_(lg.permute(0, 2, 3, 1).reshape(-1, ncls))
    label = torch.from_numpy(z[name + "_target"]).to(DEV)
    acc = torch.zeros(2, dtype=torch.float64, device=DEV)
    dstats = torch.zeros(B * 3 * ncls, dtype=torch.float64, device=DEV)
    coef = torch.empty(B * 2 * ncls + 1, device=DEV)
    loss = torch.empty((), device=DEV)
    ops.dice_ce_stats(logits, label, 255, acc, dstats, B, h, w, h, w, ncls)
    ops.dice_ce_finalize(acc, dstats, B, ncls, float(alpha), 1e-6, loss, coef)
    dl = torch.zeros(B * h * w, ld, device=DEV)
    ops.dice_ce_grad(logits, label, 255, coef, dl[:, :ncls], B, h, w, h, w, ncls)
    ref_loss = float(z[name + "_loss"])
    assert abs(float(loss) - ref_loss) < 2e-5 * max(1.0, abs(ref_loss)), (float(loss), ref_loss)
    ref_grad = torch.from_numpy(z[name + "_grad"]).to(DEV).permute(0, 2, 3, 1).reshape(-1, ncls)
    close(dl[:, :ncls], ref_grad, 2e-3, 1e-7, "dice-ce dlogits")


def test_ce_kernel_40_classes_upsampled():
    """num_classes = 40 is the reference's default config (NYUDv2, config.py): the > 16-class instantiation of the fused
    bilinear x4 + CE kernel against F.interpolate + F.cross_entropy"""
    B, h, w, H, W, ncls = 2, 12, 16, 48, 64, 40
    g = torch.Generator().manual_seed(5)
    lr = (2.0 * torch.randn(B, ncls, h, w, generator=g)).to(DEV).requires_grad_(True)
    label = torch.randint(0, ncls, (B, H, W), generator=g)
    label[torch.rand(B, H, W, generator=g) < 0.1] = 255
    label = label.to(DEV)
    logits = lr.detach().permute(0, 2, 3, 1).reshape(-1, ncls).contiguous()
    acc = torch.zeros(2, dtype=torch.float64, device=DEV)
    dl = torch.zeros_like(logits)
    ops.ce_upsampled(logits, label, 255, acc, dl, B, h, w, H, W, ncls)
    loss = torch.empty((), device=DEV)
    dout = torch.empty_like(dl)
    ops.ce_finalize(acc, loss, dl, None, dout)
    up = torch.nn.functional.interpolate(lr, size=(H, W), mode="bilinear", align_corners=False)
    ref = torch.nn.functional.cross_entropy(up, label, ignore_index=255)
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1, abs(float(ref)))
    close(dout, lr.grad.permute(0, 2, 3, 1).reshape(-1, ncls), 1e-3, 1e-6, "ce40 dlogits")


def test_confusion_bit_exact():
    from oracle import metric_ref
    rng = np.random.default_rng(0)
    for n_cl, shape in [(9, (480, 640)), (5, (37, 53)), (40, (64, 64)), (9, (1, 1))]:
        pred = rng.integers(0, n_cl, shape).astype(np.int64)
        gt = rng.integers(0, n_cl, shape).astype(np.uint8)
        gt[rng.random(shape) < 0.1] = 255
        h_ref, lab, cor = metric_ref.hist_info(n_cl, pred, gt)
        hist = torch.zeros(n_cl, n_cl, dtype=torch.int64, device=DEV)
        stats = torch.zeros(2, dtype=torch.int64, device=DEV)
        ops.confusion(torch.from_numpy(pred).to(DEV), torch.from_numpy(gt).to(DEV), n_cl, hist, stats)
        assert np.array_equal(hist.cpu().numpy(), h_ref)
        assert stats.tolist() == [lab, cor]
        scores = torch.randn(n_cl, *shape, device=DEV)
        hist.zero_(); stats.zero_()
        pred8 = torch.empty(shape, dtype=torch.uint8, device=DEV)
        ops.argmax_confusion(scores, torch.from_numpy(gt).to(DEV), n_cl, hist, stats, pred8)
        p_ref = scores.cpu().numpy().argmax(0)
        h_ref, lab, cor = metric_ref.hist_info(n_cl, p_ref, gt)
        assert np.array_equal(pred8.cpu().numpy(), p_ref)
        assert np.array_equal(hist.cpu().numpy(), h_ref) and stats.tolist() == [lab, cor]


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N,Nk,heads", [(2, 1200, 300, 5), (1, 333, 77, 1), (2, 4800, 300, 2), (1, 19200, 300, 1), (3, 300, 300, 8),
                                          (1, 130, 4, 2)])
def test_fused_attention_forward(B, N, Nk, heads):
    torch.manual_seed(11)
    d = 64
    C = heads * d
    q = rnd(B * N, C, dtype=bf)
    kv = rnd(B * Nk, 2 * C, dtype=bf)
    o = torch.full((B * N, C), 7.0, device=DEV, dtype=bf)
    Np = (Nk + 7) // 8 * 8
    pbuf = torch.full((B * heads * N, Np), 3.0, device=DEV, dtype=bf)
    lse = torch.empty(B * heads * N, device=DEV)
    scale = d ** -0.5
    ops.attn_fwd(q, kv, o, B, N, Nk, heads, scale, p_out=pbuf[:, :Nk], lse=lse)
    qf = q.float().view(B, N, heads, d).permute(0, 2, 1, 3)
    kf = kv.float().view(B, Nk, 2, heads, d)[:, :, 0].permute(0, 2, 1, 3)
    vf = kv.float().view(B, Nk, 2, heads, d)[:, :, 1].permute(0, 2, 1, 3)
    s = (qf @ kf.transpose(-1, -2)) * scale
    pr = torch.softmax(s, -1)
    ref = (pr @ vf).permute(0, 2, 1, 3).reshape(B * N, C)
    close(pbuf[:, :Nk].reshape(B, heads, N, Nk), pr, 2e-2, 2e-3, "attention probabilities")
    if Np > Nk:   # pad columns are either untouched or zero (TMA bulk stores clip the inner dimension at 16-byte granularity)
        pad = pbuf[:, Nk:].float()
        assert bool(((pad == 3.0) | (pad == 0.0)).all())
    close(o, ref, 2e-2, 2e-2, "attention output")
    close(lse.view(B, heads, N), torch.logsumexp(s, -1), 1e-3, 1e-3, "lse")
    # inference mode (no P, no LSE): the probabilities stay unnormalised in shared memory and the 64 output columns are scaled
    # by 1 / rowsum instead - the same result up to the bf16 rounding of P
    o2 = torch.empty_like(o)
    ops.attn_fwd(q, kv, o2, B, N, Nk, heads, scale)
    close(o2, ref, 2e-2, 2e-2, "attention output (no stored P)")
    close(o2, o, 2e-2, 4e-3, "attention output: stored-P vs scaled-O mode")
    # fused backward core: dS and dQ from (dO, K, V, P)
    dO = rnd(B * N, C, dtype=bf)
    dsb = torch.full((B * heads * N, Np), 5.0, device=DEV, dtype=bf)
    dq = torch.full((B * N, C), 9.0, device=DEV, dtype=bf)
    ops.attn_bwd(dO, kv, pbuf[:, :Nk], dsb[:, :Nk], dq, B, N, Nk, heads, scale)
    Pf = pbuf[:, :Nk].float().reshape(B, heads, N, Nk)
    dOf = dO.float().view(B, N, heads, d).permute(0, 2, 1, 3)
    dP = dOf @ vf.transpose(-1, -2)
    dS = scale * Pf * (dP - (Pf * dP).sum(-1, keepdim=True))
    close(dsb[:, :Nk].reshape(B, heads, N, Nk), dS, 2e-2, 2e-3 * float(dS.abs().max()) + 1e-6, "dS")
    dq_ref = (dsb[:, :Nk].float().reshape(B, heads, N, Nk) @ kf).permute(0, 2, 1, 3).reshape(B * N, C)
    close(dq, dq_ref, 2e-2, 2e-2 * float(dq_ref.abs().max()) + 1e-6, "dQ")


# ---------------------------------------------------------------------------------------------
# Flash-style attention backward with recomputed probabilities (csrc/attention_dkv.cu).  Kernel-level parity is green on B200
# (profiles/r1_exp_attention_bwd_parity.log); the ENGINE still uses it only under CMX_ATTN_DKV_RECOMPUTE=1 because the model-level
# tests and the bench A/B under that flag have not been run yet.
@pytest.mark.parametrize("B,N,Nk,heads", [(1, 128, 128, 1), (1, 130, 4, 2), (1, 333, 77, 1), (2, 1200, 300, 5), (3, 300, 300, 8),
                                          (2, 4800, 300, 2), (1, 19200, 300, 1), (1, 920, 920, 8),
                                          (5, 640, 300, 8)])   # last: 200 tiles on 148 CTAs => (sample, head) changes inside a CTA
def test_attention_dkv_recompute(B, N, Nk, heads):
    """key-major dK / dV and query-major dQ with recomputed probabilities (cmx_attn_delta + cmx_attn_dkv + cmx_attn_dq) against fp32 autograd of
    softmax(scale q k^T) v on the same bf16 inputs.  Tolerance: 2 % of the gradient's max-abs (bf16 P / dS operands)."""
    torch.manual_seed(12)
    d = 64
    C = heads * d
    scale = d ** -0.5
    q = rnd(B * N, C, dtype=bf)
    kv = rnd(B * Nk, 2 * C, dtype=bf)
    dO = rnd(B * N, C, dtype=bf)
    qf = q.float().view(B, N, heads, d).permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    kf = kv.float().view(B, Nk, 2, heads, d)[:, :, 0].permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    vf = kv.float().view(B, Nk, 2, heads, d)[:, :, 1].permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    s = (qf @ kf.transpose(-1, -2)) * scale
    of = torch.softmax(s, -1) @ vf
    dOf = dO.float().view(B, N, heads, d).permute(0, 2, 1, 3)
    of.backward(dOf)
    o = of.detach().permute(0, 2, 1, 3).reshape(B * N, C).to(bf)
    lse = torch.logsumexp(s.detach(), -1).reshape(-1).contiguous()
    delta = torch.full((B * heads * N,), 7.0, device=DEV)
    ops.attn_delta(dO, o, delta, B, N, heads)
    close(delta.view(B, heads, N), (dOf * o.float().view(B, N, heads, d).permute(0, 2, 1, 3)).sum(-1), 1e-4, 1e-4, "delta")
    dkv32 = torch.zeros(B * Nk, 2 * C, device=DEV)
    ops.attn_dkv(q, dO, kv, lse, delta, dkv32, B, N, Nk, heads, scale)
    got = dkv32.view(B, Nk, 2, heads, d)
    dk_ref, dv_ref = kf.grad.permute(0, 2, 1, 3), vf.grad.permute(0, 2, 1, 3)
    close(got[:, :, 1], dv_ref, 2e-2, 2e-2 * float(dv_ref.abs().max()), "dV")
    close(got[:, :, 0], dk_ref, 2e-2, 2e-2 * float(dk_ref.abs().max()), "dK")
    if Nk <= 384:   # the query-major companion keeps K / V of one (sample, head) in shared memory
        dq = torch.full((B * N, C), 9.0, device=DEV, dtype=bf)
        ops.attn_dq(q, dO, kv, lse, delta, dq, B, N, Nk, heads, scale)
        dq_ref = qf.grad.permute(0, 2, 1, 3).reshape(B * N, C)
        close(dq, dq_ref, 2e-2, 2e-2 * float(dq_ref.abs().max()), "dQ")


@pytest.mark.parametrize("B,N,Nk,heads", [(2, 700, 880, 2), (1, 920, 920, 8), (1, 3600, 880, 5), (2, 300, 321, 1)])
def test_attention_key_chunked_fwd_bwd(B, N, Nk, heads):
    """Nkv above the fused kernels' resident-K/V limit (BASELINE configs[3]: 880 / 920 keys at 720x1280): fused forward per chunk
    of <= 320 keys + exact log-sum-exp combination, dQ as the sum of per-chunk shares - against fp32 softmax attention and its
    autograd on the same bf16 inputs.  No [N, Nkv] tensor is allocated by the product path."""
    torch.manual_seed(13)
    d = 64
    C = heads * d
    scale = d ** -0.5
    q = rnd(B * N, C, dtype=bf)
    kv = rnd(B * Nk, 2 * C, dtype=bf)
    dO = rnd(B * N, C, dtype=bf)
    qf = q.float().view(B, N, heads, d).permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    kf = kv.float().view(B, Nk, 2, heads, d)[:, :, 0].permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    vf = kv.float().view(B, Nk, 2, heads, d)[:, :, 1].permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    s = (qf @ kf.transpose(-1, -2)) * scale
    of = torch.softmax(s, -1) @ vf
    dOf = dO.float().view(B, N, heads, d).permute(0, 2, 1, 3)
    of.backward(dOf)
    o = torch.full((B * N, C), 5.0, device=DEV, dtype=bf)
    lse = torch.full((B * heads * N,), 5.0, device=DEV)
    ops.attn_fwd_chunked(q, kv, o, lse, B, N, Nk, heads, scale)
    o_ref = of.detach().permute(0, 2, 1, 3).reshape(B * N, C)
    close(o, o_ref, 2e-2, 2e-2 * float(o_ref.abs().max()), "chunked attention O")
    close(lse, torch.logsumexp(s.detach(), -1).reshape(-1), 2e-3, 2e-3, "chunked attention lse")
    delta = torch.empty(B * heads * N, device=DEV)
    ops.attn_delta(dO, o, delta, B, N, heads)
    dq = torch.full((B * N, C), 9.0, device=DEV, dtype=bf)
    ops.attn_dq_chunked(q, dO, kv, lse, delta, dq, B, N, Nk, heads, scale)
    dq_ref = qf.grad.permute(0, 2, 1, 3).reshape(B * N, C)
    close(dq, dq_ref, 2e-2, 2e-2 * float(dq_ref.abs().max()), "chunked dQ")
    dkv32 = torch.zeros(B * Nk, 2 * C, device=DEV)
    ops.attn_dkv(q, dO, kv, lse, delta, dkv32, B, N, Nk, heads, scale)
    got = dkv32.view(B, Nk, 2, heads, d)
    dk_ref, dv_ref = kf.grad.permute(0, 2, 1, 3), vf.grad.permute(0, 2, 1, 3)
    close(got[:, :, 1], dv_ref, 2e-2, 2e-2 * float(dv_ref.abs().max()), "dV")
    close(got[:, :, 0], dk_ref, 2e-2, 2e-2 * float(dk_ref.abs().max()), "dK")


# ---------------------------------------------------------------------------------------------
# grouped launches (group = modality branch): one launch over two stacked problems with per-group parameters lying a
# fixed number of elements apart in one flat buffer must equal two separate launches
# ---------------------------------------------------------------------------------------------
def _flat_pair(shape, gs, seed, scale=1.0):
    """a flat fp32 buffer holding two parameter tensors of `shape` gs elements apart -> (flat, view0, view1)"""
    n = int(np.prod(shape))
    assert gs >= n
    torch.manual_seed(seed)
    flat = torch.randn(gs + n, device=DEV) * scale
    return flat, flat[:n].view(shape), flat[gs:gs + n].view(shape)


@pytest.mark.parametrize("M,N,K", [(300, 64, 64), (1200, 320, 320), (257, 128, 512)])
def test_grouped_gemm_fwd_dgrad_wgrad(M, N, K):
    gs = 64 * ((N * K + 1000) // 64)
    wf, w0, w1 = _flat_pair((N, K), gs, 0, K ** -0.5)
    wb = wf.to(bf)
    w0b, w1b = wb[:N * K].view(N, K), wb[gs:gs + N * K].view(N, K)
    bfl, b0, b1 = _flat_pair((N,), gs, 1)
    x = rnd(2 * M, K, dtype=bf, seed=2)
    res = rnd(2 * M, N, seed=3)
    rs = torch.rand(2, 4, device=DEV) + 0.5       # 4 "samples" per group
    rps = (M + 3) // 4
    out = torch.empty(2 * M, N, device=DEV)
    ops.mm(x, w0b, out, bias=b0, residual=res, row_scale=rs, rows_per_sample=rps, groups=2, gs_b=gs, gs_bias=gs, gs_scale=4)
    for g, (w, b) in enumerate(((w0b, b0), (w1b, b1))):
        ref = torch.empty(M, N, device=DEV)
        ops.mm(x[g * M:(g + 1) * M], w, ref, bias=b, residual=res[g * M:(g + 1) * M], row_scale=rs[g], rows_per_sample=rps)
        assert torch.equal(out[g * M:(g + 1) * M], ref), "grouped forward differs from the single launch (group %d)" % g
        close(ref, res[g * M:(g + 1) * M] + (x[g * M:(g + 1) * M].float() @ w.float().t() + b) *
              rs[g].repeat_interleave(rps)[:M, None], 2e-2, 2e-2, "grouped fwd vs torch")
    # dgrad: dx[2M, K] = dy[2M, N] @ W_g[N, K]
    dy = rnd(2 * M, N, dtype=bf, seed=4)
    dx = torch.empty(2 * M, K, device=DEV, dtype=bf)
    ops.mm(dy, w0b, dx, tb=True, groups=2, gs_b=gs)
    for g, w in enumerate((w0b, w1b)):
        ref = torch.empty(M, K, device=DEV, dtype=bf)
        ops.mm(dy[g * M:(g + 1) * M], w, ref, tb=True)
        assert torch.equal(dx[g * M:(g + 1) * M], ref), "grouped dgrad (group %d)" % g
    # wgrad: dW_g[N, K] += dy_g^T x_g into a flat gradient buffer
    gfl = torch.zeros_like(wf)
    ops.mm(dy, x, gfl[:N * K].view(N, K), ta=True, tb=True, accumulate=True, groups=2, gs_c=gs)
    for g in (0, 1):
        ref = dy[g * M:(g + 1) * M].float().t() @ x[g * M:(g + 1) * M].float()
        got = gfl[g * gs:g * gs + N * K].view(N, K)
        close(got, ref, 1e-3, 1e-3 * M ** 0.5, "grouped wgrad (group %d)" % g)
    assert float(gfl[N * K:gs].abs().max()) == 0.0, "grouped wgrad wrote outside its two slots"


@pytest.mark.parametrize("M,C", [(300, 64), (75, 320), (40, 512), (33, 160)])
def test_grouped_layernorm_fwd_bwd(M, C):
    gs = 64 * ((C + 200) // 64)
    gf, g0, g1 = _flat_pair((C,), gs, 0)
    bfl, b0, b1 = _flat_pair((C,), gs, 1)
    x = rnd(2 * M, C, seed=2)
    y = torch.empty(2 * M, C, device=DEV, dtype=bf)
    mean, rstd = torch.empty(2 * M, device=DEV), torch.empty(2 * M, device=DEV)
    ops.layernorm_fwd(x, g0, b0, 1e-6, y, mean, rstd, groups=2, param_gs=gs)
    for g, (ga, be) in enumerate(((g0, b0), (g1, b1))):
        r = slice(g * M, (g + 1) * M)
        yr = torch.empty(M, C, device=DEV, dtype=bf)
        m2, r2 = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
        ops.layernorm_fwd(x[r], ga, be, 1e-6, yr, m2, r2)
        assert torch.equal(y[r], yr) and torch.equal(mean[r], m2) and torch.equal(rstd[r], r2), "grouped LN fwd (group %d)" % g
    dy = rnd(2 * M, C, dtype=bf, seed=3)
    dres = rnd(2 * M, C, seed=4)
    sc = torch.rand(2, 3, device=DEV) + 0.5
    rps = (M + 2) // 3
    dx, dxb = torch.empty(2 * M, C, device=DEV), torch.empty(2 * M, C, device=DEV, dtype=bf)
    dg, db, dbi = torch.zeros_like(gf), torch.zeros_like(gf), torch.zeros_like(gf)
    ops.layernorm_bwd(dy, x, mean, rstd, g0, dres=dres, dx=dx, dx_bf=dxb, scale=sc, rows_per_sample=rps,
                      dgamma=dg[:C], dbeta=db[:C], dbias=dbi[:C], groups=2, param_gs=gs, scale_gs=3)
    for g, ga in enumerate((g0, g1)):
        r = slice(g * M, (g + 1) * M)
        dx2, dxb2 = torch.empty(M, C, device=DEV), torch.empty(M, C, device=DEV, dtype=bf)
        dg2, db2, dbi2 = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        ops.layernorm_bwd(dy[r], x[r], mean[r], rstd[r], ga, dres=dres[r], dx=dx2, dx_bf=dxb2, scale=sc[g], rows_per_sample=rps,
                          dgamma=dg2, dbeta=db2, dbias=dbi2)
        assert torch.equal(dx[r], dx2) and torch.equal(dxb[r], dxb2), "grouped LN bwd dx (group %d)" % g
        for a, b_, what in ((dg, dg2, "dgamma"), (db, db2, "dbeta"), (dbi, dbi2, "dbias")):
            close(a[g * gs:g * gs + C], b_, 1e-4, 1e-4 * M ** 0.5, "grouped LN bwd %s (group %d)" % (what, g))


@pytest.mark.parametrize("B,H,W,C", [(2, 15, 20, 64), (1, 9, 13, 128)])
def test_grouped_dwconv_and_colsum(B, H, W, C):
    gs = 64 * ((9 * C + 300) // 64)
    wf, w0, w1 = _flat_pair((C, 9), gs, 0, 0.3)
    bfl, b0, b1 = _flat_pair((C,), gs, 1)
    M = B * H * W
    x = rnd(2 * M, C, dtype=bf, seed=2)
    y = torch.empty(2 * M, C, device=DEV, dtype=bf)
    ops.dwconv3x3_fwd(x, w0, b0, ops.ACT_GELU, y, B, H, W, groups=2, param_gs=gs)
    dy = rnd(2 * M, C, dtype=bf, seed=3)
    du = torch.empty(2 * M, C, device=DEV, dtype=bf)
    dwf, dbf = torch.zeros_like(wf), torch.zeros_like(bfl)
    ops.dwconv3x3_bwd_pre(x, w0, b0, ops.ACT_GELU, dy, du, dwf[:9 * C].view(C, 9), dbf[:C], B, H, W, groups=2, param_gs=gs)
    dxg = torch.empty(2 * M, C, device=DEV, dtype=bf)
    ysum = torch.zeros_like(bfl)
    ops.dwconv3x3_fwd(du, w0, None, ops.ACT_NONE, dxg, B, H, W, flip=True, ysum=ysum[:C], groups=2, param_gs=gs)
    cs = torch.zeros_like(bfl)
    ops.colsum(dy, cs[:C], groups=2, out_gs=gs)
    for g, (w, b) in enumerate(((w0, b0), (w1, b1))):
        r = slice(g * M, (g + 1) * M)
        yr = torch.empty(M, C, device=DEV, dtype=bf)
        ops.dwconv3x3_fwd(x[r], w, b, ops.ACT_GELU, yr, B, H, W)
        assert torch.equal(y[r], yr), "grouped dwconv fwd (group %d)" % g
        dur = torch.empty(M, C, device=DEV, dtype=bf)
        dwr, dbr = torch.zeros(C, 9, device=DEV), torch.zeros(C, device=DEV)
        ops.dwconv3x3_bwd_pre(x[r], w, b, ops.ACT_GELU, dy[r], dur, dwr, dbr, B, H, W)
        assert torch.equal(du[r], dur), "grouped dwconv bwd_pre du (group %d)" % g
        close(dwf[g * gs:g * gs + 9 * C].view(C, 9), dwr, 1e-4, 1e-3, "grouped dwconv dW (group %d)" % g)
        close(dbf[g * gs:g * gs + C], dbr, 1e-4, 1e-3, "grouped dwconv db (group %d)" % g)
        dxr = torch.empty(M, C, device=DEV, dtype=bf)
        ysr = torch.zeros(C, device=DEV)
        ops.dwconv3x3_fwd(dur, w, None, ops.ACT_NONE, dxr, B, H, W, flip=True, ysum=ysr)
        assert torch.equal(dxg[r], dxr), "grouped dwconv dgrad (group %d)" % g
        close(ysum[g * gs:g * gs + C], ysr, 1e-4, 1e-3, "grouped dwconv ysum (group %d)" % g)
        close(cs[g * gs:g * gs + C], dy[r].float().sum(0), 1e-4, 1e-3, "grouped colsum (group %d)" % g)
