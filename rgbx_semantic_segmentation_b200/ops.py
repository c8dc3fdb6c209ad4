"""Thin Python wrappers over the C ABI (include/cmx_b200.h).  torch is used only for device memory
and the current stream; every function below enqueues hand-written sm_100a kernels and raises
RuntimeError on a non-zero return code.  There is no CPU path."""
import ctypes

import torch

from . import _lib
from ._lib import CmxGemm

BF16, F32 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU, ACT_SIGMOID = 0, 1, 2, 3
NUM_SMS = 148


def _dt(t):
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise TypeError("cmx_b200: unsupported dtype %s" % t.dtype)


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ld(t):
    assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1), "need a row-major 2-D view"
    return t.stride(0)


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cmx_b200: tensors must live on a CUDA device (no CPU fallback exists)")


def launch_count():
    return int(_lib.load().cmx_launch_count())


# Optional per-launch profiling (bench.py / debugging): when PROFILE is a list every C-ABI call is bracketed
# by CUDA events on the launching stream and (name, start, stop, flops, bytes) is appended.
PROFILE = None
# Optional call recording (bench.py): when RECORD is a list every C-ABI call is appended as (tag, function, argument tuple) so
# that the launches of one kernel class can be replayed back to back between ONE pair of events (the caller keeps every
# buffer alive through Engine.keepalive)
RECORD = None
PROFILE_SHAPES = bool(int(__import__('os').environ.get('CMX_PROFILE_SHAPES', '0')))


def _call(name, *args, tag=None, flops=0, nbytes=0):
    fn = getattr(_lib.load(), name)
    if RECORD is not None:
        RECORD.append((tag or name, fn, args))
    if PROFILE is None:
        rc = fn(*args)
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        PROFILE.append((tag or name, e0, e1, flops, nbytes))
    if rc != 0:
        raise RuntimeError("cmx_b200.%s failed (rc=%d): %s" % (name, rc, _lib.last_error()))


def _tg(name, *dims):
    """kernel-class tag of the per-launch profile; CMX_PROFILE_SHAPES=1 appends the problem size"""
    return name + ("_" + "x".join(str(d) for d in dims) if PROFILE_SHAPES else "")


def _nb(*ts):
    return sum(t.numel() * t.element_size() for t in ts if t is not None)


# ------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------
def _auto_split(M, N, K, want):
    if not want:
        return 1
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    s = (2 * NUM_SMS) // max(tiles, 1)
    return max(1, min(s, K // 512 if K >= 1024 else 1))


def gemm_raw(A, B, C, M, N, K, lda, ldb, ldc, *, a_off=0, b_off=0, c_off=0, trans_a=False, trans_b=False, bias=None,
             residual=None, ldr=0, r_off=0, row_scale=None, rows_per_sample=0, act=ACT_NONE, alpha=1.0, split_k=1,
             accumulate=False, batch=(1, 1), sA=(0, 0), sB=(0, 0), sC=(0, 0), impl=0, s_bias=0, s_res=0, s_scale=0):
    """A, B, C (and residual) are base tensors; *_off are element offsets into them.  s_bias / s_res / s_scale: element
    strides of bias / residual / row_scale per batch[0] index (grouped launches: one batch index per modality branch)."""
    _cuda(A, B, C)
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    g = CmxGemm()
    g.A = A.data_ptr() + 2 * a_off
    g.B = B.data_ptr() + 2 * b_off
    g.C = C.data_ptr() + C.element_size() * c_off
    g.bias = _p(bias)
    g.residual = None if residual is None else residual.data_ptr() + residual.element_size() * r_off
    g.row_scale = _p(row_scale)
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldc, g.ldr = lda, ldb, ldc, ldr
    g.trans_a, g.trans_b = int(trans_a), int(trans_b)
    g.batch1, g.batch2 = batch
    g.sA1, g.sA2 = sA
    g.sB1, g.sB2 = sB
    g.sC1, g.sC2 = sC
    g.c_dtype = _dt(C)
    g.r_dtype = _dt(residual) if residual is not None else F32
    g.act = act
    g.alpha = alpha
    g.accumulate = int(accumulate)
    g.split_k = split_k
    g.rows_per_sample = rows_per_sample
    g.impl = impl
    g.sBias1, g.sR1, g.sS1 = s_bias, s_res, s_scale
    nb = batch[0] * batch[1]
    which = "tc" if int(_lib.load().cmx_gemm_which(ctypes.byref(g))) == 2 else ("fb_batched" if nb > 1 else "fb")
    kind = "wgrad" if trans_a else ("dgrad" if trans_b else "fwd")   # by operand layout (scripts/ncu_launch_summary.py agrees)
    cbytes = C.element_size() * (2 if (accumulate or split_k > 1) else 1)
    tag = "gemm_%s_%s" % (which, kind)
    if PROFILE_SHAPES:
        tag += "_%dx%dx%d" % (M, N, K) + ("_b%d" % nb if nb > 1 else "") + ("_s%d" % split_k if split_k > 1 else "")
    _call("cmx_gemm", ctypes.byref(g), _stream(), tag=tag, flops=2 * M * N * K * nb,
          nbytes=nb * (2 * (M * K + N * K) + cbytes * M * N + (residual.element_size() * M * N if residual is not None else 0)))
    return C


def mm(a, b, out, *, ta=False, tb=False, bias=None, residual=None, row_scale=None, rows_per_sample=0, act=ACT_NONE,
       alpha=1.0, accumulate=False, split_k=None, impl=0, groups=1, gs_a=None, gs_b=None, gs_c=None, gs_bias=0, gs_scale=0):
    """out[M,N] = residual + row_scale * act(alpha * op(a) @ op(b) + bias) on 2-D row-major views.
    ta=False: a is [M,K]; ta=True: a is stored [K,M].  tb=False: b is [N,K] (nn.Linear weight layout);
    tb=True: b is stored [K,N].  accumulate=True adds into fp32 `out` (split-K chosen automatically).
    groups > 1: ONE launch over `groups` independent problems of the same shape (the RGB and X branches of a stage).  An
    operand whose group stride gs_* (elements) is given is the group-0 view of a per-branch parameter (the same tensor of the
    other branch lies gs_* elements further in the flat buffers); an operand without one is STACKED along its dim 0
    ([groups * rows, cols]; residual like out).  bias / row_scale: group-0 views with strides gs_bias / gs_scale."""
    def per_group(t, gs):
        if groups == 1:
            return t.shape, 0
        if gs is not None:
            return t.shape, gs
        assert t.shape[0] % groups == 0, (t.shape, groups)
        return (t.shape[0] // groups, t.shape[1]), (t.shape[0] // groups) * _ld(t)
    (sa0, sa1), sA1 = per_group(a, gs_a)
    (sb0, sb1), sB1 = per_group(b, gs_b)
    (sc0, sc1), sC1 = per_group(out, gs_c)
    K, M = (sa0, sa1) if ta else (sa1, sa0)
    Kb, N = (sb0, sb1) if tb else (sb1, sb0)
    assert K == Kb, (a.shape, b.shape, ta, tb, groups)
    assert (sc0, sc1) == (M, N), (out.shape, M, N, groups)
    if split_k is None:
        split_k = _auto_split(M * groups, N, K, accumulate)
    s_res = 0
    if residual is not None and groups > 1:
        assert residual.shape == out.shape and gs_c is None
        s_res = M * _ld(residual)
    return gemm_raw(a, b, out, M, N, K, _ld(a), _ld(b), _ld(out), trans_a=ta, trans_b=tb, bias=bias, residual=residual,
                    ldr=_ld(residual) if residual is not None else 0, row_scale=row_scale, rows_per_sample=rows_per_sample,
                    act=act, alpha=alpha, split_k=split_k, accumulate=accumulate, impl=impl, batch=(groups, 1),
                    sA=(sA1, 0), sB=(sB1, 0), sC=(sC1, 0), s_bias=gs_bias, s_res=s_res, s_scale=gs_scale)


def gemm_which(M, N, K, lda, ldb, ldc, trans_a=False, trans_b=False, c_dtype=BF16, ptr=256):
    g = CmxGemm()
    g.A = g.B = g.C = ptr
    g.M, g.N, g.K, g.lda, g.ldb, g.ldc = M, N, K, lda, ldb, ldc
    g.trans_a, g.trans_b, g.batch1, g.batch2, g.c_dtype = int(trans_a), int(trans_b), 1, 1, c_dtype
    return int(_lib.load().cmx_gemm_which(ctypes.byref(g)))


# ------------------------------------------------------------------------------------------------
# normalisation
# ------------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, eps, y, mean=None, rstd=None, groups=1, param_gs=0):
    """groups > 1: x / y / mean / rstd are `groups` stacked blocks of rows; gamma / beta = group-0 views, param_gs apart"""
    M, C = x.shape
    assert M % groups == 0
    M //= groups
    _cuda(x, y)
    _call("cmx_layernorm_fwd", x.data_ptr(), _dt(x), _ld(x), gamma.data_ptr(), beta.data_ptr(), eps,
          y.data_ptr(), _dt(y), _ld(y), _p(mean), _p(rstd), M, C, groups, param_gs, _stream(),
          tag=_tg("cmx_layernorm_fwd", M * groups, C), nbytes=_nb(x, y))
    return y


def layernorm_bwd(dy, x, mean, rstd, gamma, *, dy2=None, dres=None, dx=None, dx_bf=None, scale=None, rows_per_sample=0,
                  dgamma=None, dbeta=None, dbias=None, groups=1, param_gs=0, scale_gs=0):
    M, C = x.shape
    assert M % groups == 0
    M //= groups
    _call("cmx_layernorm_bwd",
        dy.data_ptr(), _dt(dy), _ld(dy), _p(dy2), _ld(dy2) if dy2 is not None else 0, x.data_ptr(), _dt(x), _ld(x),
        mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), _p(dres), _ld(dres) if dres is not None else 0,
        _p(dx), _dt(dx) if dx is not None else F32, _ld(dx) if dx is not None else 0,
        _p(dx_bf), _ld(dx_bf) if dx_bf is not None else 0, _p(scale), rows_per_sample,
        _p(dgamma), _p(dbeta), _p(dbias), M, C, groups, param_gs, scale_gs, _stream(),
        tag=_tg("cmx_layernorm_bwd", M * groups, C), nbytes=_nb(dy, dy2, x, dres, dx, dx_bf))


def colstats(x, sum_, sumsq):
    M, C = x.shape
    _call("cmx_colstats", x.data_ptr(), _dt(x), _ld(x), sum_.data_ptr(), sumsq.data_ptr(), M, C, _stream(), nbytes=_nb(x))


def bn_finalize(sum_, sumsq, count, eps, momentum, running_mean, running_var, nbt, mean, invstd):
    C = mean.numel()
    _call("cmx_bn_finalize", sum_.data_ptr(), sumsq.data_ptr(), count, eps, momentum, _p(running_mean),
                                           _p(running_var), _p(nbt), mean.data_ptr(), invstd.data_ptr(), C, _stream())


def bn_eval_stats(running_mean, running_var, eps, mean, invstd):
    _call("cmx_bn_eval_stats", running_mean.data_ptr(), running_var.data_ptr(), eps, mean.data_ptr(),
                                             invstd.data_ptr(), mean.numel(), _stream())


def bn_apply(x, mean, invstd, gamma, beta, y, *, residual=None, relu=False, mask=None, rows_per_sample=0):
    M, C = x.shape
    _call("cmx_bn_apply", x.data_ptr(), _dt(x), _ld(x), mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(),
                                        beta.data_ptr(), _p(residual), _dt(residual) if residual is not None else F32,
                                        _ld(residual) if residual is not None else 0, int(relu), _p(mask), rows_per_sample,
                                        y.data_ptr(), _dt(y), _ld(y), M, C, _stream(), nbytes=_nb(x, residual, y))
    return y


def _bn_bwd_common(dy, x, mean, invstd, gamma, beta, residual, relu, mask, rows_per_sample):
    return (dy.data_ptr(), _dt(dy), _ld(dy), x.data_ptr(), _dt(x), _ld(x), mean.data_ptr(), invstd.data_ptr(),
            gamma.data_ptr(), beta.data_ptr(), _p(residual), _dt(residual) if residual is not None else F32,
            _ld(residual) if residual is not None else 0, int(relu), _p(mask), rows_per_sample)


def bn_bwd_reduce(dy, x, mean, invstd, gamma, beta, ws, *, residual=None, relu=False, mask=None, rows_per_sample=0):
    """pass 1: ws (zeroed double[2*C]) += [sum dy_eff, sum dy_eff * xhat] over the rows"""
    M, C = x.shape
    common = _bn_bwd_common(dy, x, mean, invstd, gamma, beta, residual, relu, mask, rows_per_sample)
    _call("cmx_bn_bwd_reduce", *common, ws[:C].data_ptr(), ws[C:].data_ptr(), M, C, _stream(), nbytes=_nb(dy, x))


def bn_bwd_apply(dy, x, mean, invstd, gamma, beta, dx, dgamma, dbeta, ws, *, residual=None, relu=False, mask=None,
                 rows_per_sample=0, dres=None):
    """pass 2: dx (and dres = effective upstream grad; same dtype) from the sums in ws; dgamma / dbeta += the sums"""
    M, C = x.shape
    common = _bn_bwd_common(dy, x, mean, invstd, gamma, beta, residual, relu, mask, rows_per_sample)
    _call("cmx_bn_bwd_apply", *common, ws[:C].data_ptr(), ws[C:].data_ptr(), dx.data_ptr(), _dt(dx), _ld(dx), _p(dres),
          _dt(dres) if dres is not None else _dt(dx), _ld(dres) if dres is not None else 0,
          _p(dgamma), _p(dbeta), M, C, _stream(), nbytes=_nb(dy, x, dx, dres))


def bn_bwd(dy, x, mean, invstd, gamma, beta, dx, dgamma, dbeta, ws, *, residual=None, relu=False, mask=None,
           rows_per_sample=0, dres=None):
    """ws: zeroed double[2*C] workspace.  dx (and dres = effective upstream grad) share a dtype."""
    kw = dict(residual=residual, relu=relu, mask=mask, rows_per_sample=rows_per_sample)
    bn_bwd_reduce(dy, x, mean, invstd, gamma, beta, ws, **kw)
    bn_bwd_apply(dy, x, mean, invstd, gamma, beta, dx, dgamma, dbeta, ws, dres=dres, **kw)


# ------------------------------------------------------------------------------------------------
# depthwise conv
# ------------------------------------------------------------------------------------------------
def dwconv3x3_fwd(x, w, bias, act, y, B, H, W, flip=False, ysum=None, groups=1, param_gs=0):
    """groups > 1: x / y are `groups` stacked blocks of B samples; w / bias / ysum = group-0 views, param_gs elements apart"""
    C = x.shape[1]
    _call("cmx_dwconv3x3_fwd", x.data_ptr(), _ld(x), w.data_ptr(), _p(bias), act, int(flip), y.data_ptr(), _ld(y), _p(ysum),
          B, H, W, C, groups, param_gs, _stream(),
          tag=_tg("cmx_dwconv3x3_%s" % ("dgrad" if flip else "fwd"), groups * B * H * W, C), nbytes=_nb(x, y))
    return y


def dwconv3x3_bwd_pre(x, w, bias, act, dy, du, dw, db, B, H, W, groups=1, param_gs=0):
    C = x.shape[1]
    _call("cmx_dwconv3x3_bwd_pre", x.data_ptr(), _ld(x), w.data_ptr(), _p(bias), act, dy.data_ptr(), _ld(dy),
          du.data_ptr(), _ld(du), dw.data_ptr(), _p(db), B, H, W, C, groups, param_gs, _stream(),
          tag=_tg("cmx_dwconv3x3_bwd_pre", groups * B * H * W, C), nbytes=_nb(x, dy, du))


# ------------------------------------------------------------------------------------------------
# movers
# ------------------------------------------------------------------------------------------------
def im2col_nchw(x, col, k, s, p, Ho, Wo):
    B, Cin, H, W = x.shape
    assert x.is_contiguous() and x.dtype == torch.float32
    _call("cmx_im2col_nchw", x.data_ptr(), col.data_ptr(), B, Cin, H, W, k, s, p, Ho, Wo, col.shape[1], _stream(), nbytes=_nb(x, col))
    return col


def im2col_u8(x, col, k, s, p, Ho, Wo, mean, std):
    """raw uint8 image [B,H,W,3] or grey [B,H,W] on the device -> bf16 im2col rows of the normalised image (float64 normalisation
    in the gather; grey: one column per tap)"""
    assert x.dtype == torch.uint8 and x.is_contiguous()
    B, H, W = x.shape[:3]
    ch = 1 if x.dim() == 3 else x.shape[3]
    _call("cmx_im2col_u8", x.data_ptr(), col.data_ptr(), B, ch, H, W, k, s, p, Ho, Wo, col.shape[1], mean[0], mean[1], mean[2],
          std[0], std[1], std[2], _stream(), nbytes=_nb(x, col))
    return col


def im2col_nhwc(x, col, B, H, W, k, s, p, Ho, Wo):
    C = x.shape[1]
    _call("cmx_im2col_nhwc", x.data_ptr(), _ld(x), col.data_ptr(), B, H, W, C, k, s, p, Ho, Wo, _stream(), nbytes=_nb(x, col))
    return col


def col2im_nhwc(dcol, dx, B, H, W, k, s, p, Ho, Wo, add=None):
    C = dx.shape[1]
    _call("cmx_col2im_nhwc", dcol.data_ptr(), _p(add), _dt(add) if add is not None else F32,
                                           _ld(add) if add is not None else 0, dx.data_ptr(), _dt(dx), _ld(dx), B, H, W, C, k, s, p,
                                           Ho, Wo, _stream(), nbytes=_nb(dcol, add, dx))
    return dx


def convw_pack(w, wp):
    Co, Ci, kh, kw = w.shape
    _call("cmx_convw_pack", w.data_ptr(), wp.data_ptr(), Co, Ci, kh, kw, wp.shape[1], _stream())
    return wp


def convw_pack_multi(table, n):
    """table: device uint8 tensor holding n CmxConvDesc records (include/cmx_b200.h)"""
    _call("cmx_convw_pack_multi", table.data_ptr(), n, _stream())


def convw_unpack_grad_multi(table, n):
    _call("cmx_convw_unpack_grad_multi", table.data_ptr(), n, _stream())


def convw_unpack_grad(gp, gw):
    Co, Ci, kh, kw = gw.shape
    _call("cmx_convw_unpack_grad", gp.data_ptr(), gw.data_ptr(), Co, Ci, kh, kw, gp.shape[1], _stream())


def cast_f32_bf16(x, y):
    _call("cmx_cast_f32_bf16", x.data_ptr(), y.data_ptr(), x.numel(), _stream(), nbytes=_nb(x, y))
    return y


def cast_bf16_f32(x, y):
    _call("cmx_cast_bf16_f32", x.data_ptr(), y.data_ptr(), x.numel(), _stream())
    return y


def colsum(x, out, groups=1, out_gs=0):
    """out[n] += sum_m x[m, n]; groups > 1: x = `groups` stacked row blocks, out of group g lies g * out_gs elements further"""
    M, N = x.shape
    assert M % groups == 0
    M //= groups
    _call("cmx_colsum", x.data_ptr(), _dt(x), _ld(x), out.data_ptr(), M, N, groups, out_gs, _stream(),
          tag=_tg("cmx_colsum", M * groups, N), nbytes=_nb(x))


def relu_bwd_(dy, y):
    M, N = dy.shape
    _call("cmx_relu_bwd", dy.data_ptr(), _ld(dy), y.data_ptr(), _ld(y), M, N, _stream(), nbytes=2 * _nb(dy) + _nb(y))
    return dy


def axpby(a, x, b, y, out):
    _call("cmx_axpby_f32", a, x.data_ptr(), b, _p(y), out.data_ptr(), x.numel(), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# fused attention forward
# ------------------------------------------------------------------------------------------------
ATTN_MAX_NK = 320


def attn_fwd(q, kv, o, B, N, Nk, heads, scale, p_out=None, lse=None, kv_rows=0, kv_row0=0):
    """q [B*N, C], kv [B*Nk, 2C], o [B*N, C] (bf16, head_dim 64); p_out: bf16 view [B*heads*N, Nk] with padded ld.
    kv_rows / kv_row0: this call attends to the Nk keys starting at key row kv_row0 of every sample, whose kv holds kv_rows key
    rows (key-chunked attention, see attn_fwd_chunked)"""
    _cuda(q, kv, o)
    _call("cmx_attn_fwd", q.data_ptr(), _ld(q), kv.data_ptr() + 2 * kv_row0 * _ld(kv), _ld(kv), o.data_ptr(), _ld(o), _p(p_out),
          _ld(p_out) if p_out is not None else 0, _p(lse), B, N, Nk, heads, scale, kv_rows, _stream(), tag=_tg("cmx_attn_fwd", B, N, Nk, heads),
          flops=4 * B * heads * N * Nk * 64, nbytes=_nb(q, o) + 2 * B * Nk * kv.shape[1] + (B * heads * N * Nk * 2 if p_out is not None else 0))
    return o


ATTN_CHUNK = 320


def attn_fwd_chunked(q, kv, o, lse, B, N, Nk, heads, scale):
    """fused attention over a key axis longer than ATTN_MAX_NK (e.g. Nkv = 880 / 920 at 720x1280): the fused kernel per chunk of
    <= 320 keys (normalised chunk output + chunk log-sum-exp), then the exact combination over the chunks - no [N, Nkv]
    score / probability tensor exists.  o (bf16 [B*N, C]) and lse (fp32 [B*heads*N]) are written."""
    nc = (Nk + ATTN_CHUNK - 1) // ATTN_CHUNK
    M, C = o.shape
    o_parts = torch.empty(nc, M, C, device=o.device, dtype=o.dtype)
    lse_parts = torch.empty(nc, B * heads * N, device=o.device, dtype=torch.float32)
    for c in range(nc):
        k0 = c * ATTN_CHUNK
        attn_fwd(q, kv, o_parts[c], B, N, min(ATTN_CHUNK, Nk - k0), heads, scale, lse=lse_parts[c], kv_rows=Nk, kv_row0=k0)
    _call("cmx_attn_combine", o_parts.data_ptr(), M * C, C, lse_parts.data_ptr(), B * heads * N, nc, o.data_ptr(), _ld(o), lse.data_ptr(),
          B, N, heads, _stream(), nbytes=_nb(o_parts, o))
    return o


def attn_bwd(d_o, kv, p, ds, dq, B, N, Nk, heads, scale):
    """d_o [B*N, C], kv [B*Nk, 2C], p / ds: bf16 views [B*heads*N, Nk] (padded ld), dq [B*N, C]"""
    _cuda(d_o, kv, p, ds, dq)
    _call("cmx_attn_bwd", d_o.data_ptr(), _ld(d_o), kv.data_ptr(), _ld(kv), p.data_ptr(), _ld(p), ds.data_ptr(), _ld(ds),
          dq.data_ptr(), _ld(dq), B, N, Nk, heads, scale, _stream(), tag=_tg("cmx_attn_bwd", B, N, Nk, heads),
          flops=4 * B * heads * N * Nk * 64, nbytes=_nb(d_o, kv, dq) + 2 * (B * heads * N * Nk * 2))
    return dq


def attn_delta(d_o, o, delta, B, N, heads):
    """(engine: CMX_ATTN_DKV_RECOMPUTE=1 only) delta [B*heads*N] fp32 = rowsum over each head's 64 channels of dO .* O (d_o, o: bf16 [B*N, C])"""
    _cuda(d_o, o, delta)
    _call("cmx_attn_delta", d_o.data_ptr(), _ld(d_o), o.data_ptr(), _ld(o), delta.data_ptr(), B, N, heads, _stream(),
          nbytes=_nb(d_o, o, delta))
    return delta


def attn_dkv(q, d_o, kv, lse, delta, dkv32, B, N, Nk, heads, scale):
    """(engine: CMX_ATTN_DKV_RECOMPUTE=1 only) key-major dK / dV with recomputed probabilities, added into the zero-initialised fp32 dkv32 [B*Nk, 2C]
    (q, d_o: bf16 [B*N, C]; kv: bf16 [B*Nk, 2C]; lse, delta: fp32 [B*heads*N])"""
    _cuda(q, d_o, kv, lse, delta, dkv32)
    if dkv32.dtype != torch.float32 or lse.dtype != torch.float32 or delta.dtype != torch.float32:
        raise TypeError("attn_dkv: lse, delta and dkv32 must be fp32")
    nkb = (Nk + 127) // 128
    _call("cmx_attn_dkv", q.data_ptr(), _ld(q), d_o.data_ptr(), _ld(d_o), kv.data_ptr(), _ld(kv), lse.data_ptr(), delta.data_ptr(),
          dkv32.data_ptr(), _ld(dkv32), B, N, Nk, heads, scale, _stream(),
          flops=8 * B * heads * N * nkb * 128 * 64, nbytes=nkb * _nb(q, d_o) + _nb(kv, dkv32, lse, delta))
    return dkv32


def attn_dq(q, d_o, kv, lse, delta, dq, B, N, Nk, heads, scale, kv_rows=0, kv_row0=0):
    """query-major dQ (bf16 [B*N, C]) with recomputed probabilities (Nk <= 384); kv_rows / kv_row0 as in attn_fwd: with the
    log-sum-exp and delta of the FULL key axis, a call over one key chunk gives that chunk's additive share of dQ"""
    _cuda(q, d_o, kv, lse, delta, dq)
    if lse.dtype != torch.float32 or delta.dtype != torch.float32:
        raise TypeError("attn_dq: lse and delta must be fp32")
    nkb = (Nk + 127) // 128
    _call("cmx_attn_dq", q.data_ptr(), _ld(q), d_o.data_ptr(), _ld(d_o), kv.data_ptr() + 2 * kv_row0 * _ld(kv), _ld(kv), lse.data_ptr(),
          delta.data_ptr(), dq.data_ptr(), _ld(dq), B, N, Nk, heads, scale, kv_rows, _stream(),
          flops=6 * B * heads * N * nkb * 128 * 64, nbytes=_nb(q, d_o, dq, lse, delta) + 2 * B * Nk * kv.shape[1])
    return dq


def attn_dq_chunked(q, d_o, kv, lse, delta, dq, B, N, Nk, heads, scale):
    """dQ over a key axis longer than the dq kernel's 384 keys: one call per chunk (full-axis lse / delta), partials summed"""
    nc = (Nk + ATTN_CHUNK - 1) // ATTN_CHUNK
    M, C = dq.shape
    parts = torch.empty(nc, M, C, device=dq.device, dtype=dq.dtype)
    for c in range(nc):
        k0 = c * ATTN_CHUNK
        attn_dq(q, d_o, kv, lse, delta, parts[c], B, N, min(ATTN_CHUNK, Nk - k0), heads, scale, kv_rows=Nk, kv_row0=k0)
    _call("cmx_sum_parts_bf16", parts.data_ptr(), M * C, nc, dq.data_ptr(), M * C, _stream(), nbytes=_nb(parts, dq))
    return dq


# ------------------------------------------------------------------------------------------------
# softmax
# ------------------------------------------------------------------------------------------------
def softmax_rows_fwd(s, p):
    rows, n = s.shape
    _call("cmx_softmax_rows_fwd", s.data_ptr(), _ld(s), p.data_ptr(), _ld(p), rows, n, _stream(), nbytes=_nb(s, p))
    return p


def softmax_rows_bwd(p, dp, scale, ds):
    rows, n = p.shape
    _call("cmx_softmax_rows_bwd", p.data_ptr(), _ld(p), dp.data_ptr(), _ld(dp), scale, ds.data_ptr(), _ld(ds), rows, n,
                                                _stream(), nbytes=_nb(p, dp, ds))
    return ds


def softmax_dim2_fwd(c, scale, p32, p16):
    nb, d, _ = c.shape
    _call("cmx_softmax_dim2_fwd", c.data_ptr(), scale, p32.data_ptr(), p16.data_ptr(), nb, d, _stream())


def softmax_dim2_bwd(p32, dp, scale, dc16):
    nb, d, _ = p32.shape
    _call("cmx_softmax_dim2_bwd", p32.data_ptr(), dp.data_ptr(), scale, dc16.data_ptr(), nb, d, _stream())


# ------------------------------------------------------------------------------------------------
# FRM
# ------------------------------------------------------------------------------------------------
def pool_avgmax_fwd(x, y, argmax, B, HW, ws=None):
    C2 = x.shape[1]
    if ws is None:
        ws = torch.empty(int(_lib.load().cmx_pool_avgmax_ws_bytes(B, C2)), dtype=torch.uint8, device=x.device)
    _call("cmx_pool_avgmax_fwd", x.data_ptr(), _ld(x), y.data_ptr(), argmax.data_ptr(), ws.data_ptr(), B, HW, C2, _stream(),
          nbytes=_nb(x))


def pool_avgmax_bwd(dy, argmax, dx, B, HW):
    C2 = dx.shape[1]
    _call("cmx_pool_avgmax_bwd", dy.data_ptr(), argmax.data_ptr(), dx.data_ptr(), _ld(dx), B, HW, C2, _stream(), nbytes=2 * _nb(dx))


def smallm_linear_fwd(x, w, b, act, y):
    Mb, K = x.shape
    N = w.shape[0]
    _call("cmx_smallm_linear_fwd", x.data_ptr(), w.data_ptr(), _p(b), act, y.data_ptr(), Mb, N, K, _stream())
    return y


def smallm_linear_bwd(dy, y, act, x, w, dx, dw, db, ws):
    Mb, K = x.shape
    N = w.shape[0]
    _call("cmx_smallm_linear_bwd", dy.data_ptr(), y.data_ptr(), act, x.data_ptr(), w.data_ptr(), _p(dx), _p(dw), _p(db),
                                                 ws.data_ptr(), Mb, N, K, _stream())


def frm_rectify_fwd(a, t, w2, b2, cw, sw, r1, r2, B, HW):
    C = t.shape[1]
    _call("cmx_frm_rectify_fwd", a.data_ptr(), _ld(a), t.data_ptr(), _ld(t), w2.data_ptr(), b2.data_ptr(), cw.data_ptr(),
                                               sw.data_ptr(), r1.data_ptr(), _ld(r1), r2.data_ptr(), _ld(r2), B, HW, C, _stream(), nbytes=_nb(a, t, r1, r2))


def frm_rectify_bwd(dr1, dr2, a, t, w2, cw, sw, da, dt, dcw, dw2, db2, B, HW):
    C = t.shape[1]
    _call("cmx_frm_rectify_bwd", dr1.data_ptr(), _ld(dr1), dr2.data_ptr(), _ld(dr2), a.data_ptr(), _ld(a), t.data_ptr(),
                                               _ld(t), w2.data_ptr(), cw.data_ptr(), sw.data_ptr(), da.data_ptr(), _ld(da),
                                               dt.data_ptr(), _ld(dt), dcw.data_ptr(), dw2.data_ptr(), db2.data_ptr(), B, HW, C,
                                               _stream(), nbytes=_nb(dr1, dr2, a, t, da, dt))


# ------------------------------------------------------------------------------------------------
# decoder / loss / metric
# ------------------------------------------------------------------------------------------------
def upsample_sum_fwd(zs, sizes, bias, out, B, C):
    """zs: [z0, z1, z2, z3] bf16 (z0 at output resolution; later entries may be None); sizes: [(H,W)]*4"""
    z = list(zs) + [None] * (4 - len(zs))
    sz = list(sizes) + [(1, 1)] * (4 - len(sizes))
    _call("cmx_upsample_sum_fwd", _p(z[0]), _p(z[1]), _p(z[2]), _p(z[3]), sz[0][0], sz[0][1], sz[1][0], sz[1][1],
                                                sz[2][0], sz[2][1], sz[3][0], sz[3][1], _p(bias), out.data_ptr(), _dt(out), B, C, _stream(), nbytes=_nb(*zs) + _nb(out))
    return out


def upsample_bwd(dout, Ho, Wo, dz, Hi, Wi, B, C):
    _call("cmx_upsample_bwd", dout.data_ptr(), Ho, Wo, dz.data_ptr(), Hi, Wi, B, C, _stream(), nbytes=_nb(dout, dz))
    return dz


def upsample_bwd_multi(dout, Ho, Wo, dzs, sizes, B, C):
    """adjoint of 1..3 bilinear sources in one pass over dout: dzs[i] bf16 [B*h_i*w_i, C], sizes[i] = (h_i, w_i)"""
    assert 1 <= len(dzs) <= 3 and len(dzs) == len(sizes)
    args = []
    for i in range(3):
        if i < len(dzs):
            args += [dzs[i].data_ptr(), sizes[i][0], sizes[i][1]]
        else:
            args += [None, 0, 0]
    _call("cmx_upsample_bwd_multi", dout.data_ptr(), Ho, Wo, *args, B, C, _stream(), nbytes=_nb(dout, *dzs))
    return dzs


def ce_upsampled(logits, label, ignore_index, acc, dlogits, B, h, w, H, W, ncls):
    """logits / dlogits: [B*h*w, ncls] row-major views (row stride >= ncls, identical for both)"""
    assert label.dtype == torch.int64 and label.is_contiguous()
    assert logits.dtype == torch.float32 and tuple(logits.shape) == (B * h * w, ncls)
    if dlogits is not None:
        assert dlogits.dtype == torch.float32 and dlogits.shape == logits.shape and _ld(dlogits) == _ld(logits)
    _call("cmx_ce_upsampled_fwd_bwd", logits.data_ptr(), _ld(logits), label.data_ptr(), ignore_index, acc.data_ptr(), _p(dlogits),
                                                    B, h, w, H, W, ncls, _stream())


def ce_focal_upsampled(logits, label, ignore_index, acc, dlogits, B, h, w, H, W, ncls, w_ce, w_focal, gamma, alpha):
    """w_ce * CE + w_focal * FocalLoss(gamma, alpha) (utils/loss_opr.py:157-196) in the same single pass as ce_upsampled"""
    assert label.dtype == torch.int64 and label.is_contiguous()
    assert logits.dtype == torch.float32 and tuple(logits.shape) == (B * h * w, ncls)
    if dlogits is not None:
        assert dlogits.dtype == torch.float32 and dlogits.shape == logits.shape and _ld(dlogits) == _ld(logits)
    _call("cmx_ce_focal_upsampled_fwd_bwd", logits.data_ptr(), _ld(logits), label.data_ptr(), ignore_index, acc.data_ptr(), _p(dlogits),
          B, h, w, H, W, ncls, float(w_ce), float(w_focal), float(gamma), float(alpha), _stream())


def dice_ce_stats(logits, label, ignore_index, acc, dstats, B, h, w, H, W, ncls):
    """DiceCELoss pass 0 (utils/loss_opr.py:103-156): acc (double[2]) += (CE sum, valid count); dstats (double[B, 3, ncls]) +=
    per (sample, class) (sum p, sum p * onehot, sum onehot) over the valid pixels of the upsampled logits"""
    assert label.dtype == torch.int64 and label.is_contiguous()
    assert logits.dtype == torch.float32 and tuple(logits.shape) == (B * h * w, ncls)
    assert dstats.dtype == torch.float64 and dstats.numel() == B * 3 * ncls
    _call("cmx_dice_ce_stats", logits.data_ptr(), _ld(logits), label.data_ptr(), ignore_index, acc.data_ptr(), dstats.data_ptr(),
          B, h, w, H, W, ncls, _stream())


def dice_ce_finalize(acc, dstats, B, ncls, alpha, smooth, loss=None, coef=None):
    """loss = alpha * (1 - mean dice) + (1 - alpha) * CE; coef (float[B*2*ncls + 1]) = gradient coefficients for pass 1"""
    assert coef is None or (coef.dtype == torch.float32 and coef.numel() == B * 2 * ncls + 1)
    _call("cmx_dice_ce_finalize", acc.data_ptr(), dstats.data_ptr(), B, ncls, float(alpha), float(smooth), _p(loss), _p(coef), _stream())


def dice_ce_grad(logits, label, ignore_index, coef, dlogits, B, h, w, H, W, ncls):
    """DiceCELoss pass 1: dlogits (fp32, zeroed, same row stride as logits) += d loss / d logits (final scale)"""
    assert dlogits.dtype == torch.float32 and dlogits.shape == logits.shape and _ld(dlogits) == _ld(logits)
    _call("cmx_dice_ce_grad", logits.data_ptr(), _ld(logits), label.data_ptr(), ignore_index, coef.data_ptr(), dlogits.data_ptr(),
          B, h, w, H, W, ncls, _stream())


def ce_finalize(acc, loss, dlogits=None, gscale=None, out=None):
    n = dlogits.numel() if dlogits is not None else 0
    _call("cmx_ce_finalize", acc.data_ptr(), _p(loss), _p(dlogits), _p(gscale), _p(out),
                                           _dt(out) if out is not None else F32, n, _stream())


def logits_upsample_nchw(logits, out, B, h, w, H, W, ncls):
    assert logits.dtype == torch.float32 and tuple(logits.shape) == (B * h * w, ncls)
    _call("cmx_logits_upsample_nchw", logits.data_ptr(), _ld(logits), out.data_ptr(), B, h, w, H, W, ncls, _stream())
    return out


_INT_TAG = {torch.uint8: 0, torch.int32: 1, torch.int64: 2}


def confusion(pred, gt, n_cl, hist, stats):
    _cuda(pred, gt, hist, stats)
    assert pred.is_contiguous() and gt.is_contiguous() and pred.numel() == gt.numel()
    _call("cmx_confusion", pred.data_ptr(), _INT_TAG[pred.dtype], gt.data_ptr(), _INT_TAG[gt.dtype], pred.numel(), n_cl,
                                         hist.data_ptr(), stats.data_ptr(), _stream())


def argmax_confusion(scores, gt, n_cl, hist, stats, pred_out=None):
    """scores: [n_cl, H, W] fp32 or fp64 contiguous (one image); gt may be None (argmax only)."""
    _cuda(scores)
    assert scores.is_contiguous() and scores.dtype in (torch.float32, torch.float64) and scores.shape[0] == n_cl
    npix = scores[0].numel()
    _call("cmx_argmax_confusion", scores.data_ptr(), int(scores.dtype == torch.float64), _p(gt),
          _INT_TAG[gt.dtype] if gt is not None else 0, npix, n_cl, _p(pred_out), _p(hist), _p(stats), _stream())


def eval_pack_crop(img_u8, pad, win, out_off, mean, std, flip, out):
    """one normalised network crop out [ch, crop_h, crop_w] fp32 from a device uint8 image [rows, cols(, 3)] (include/cmx_b200.h:
    cmx_eval_pack_crop).  pad = (pad_top, pad_left) of the zero-padded raw canvas, win = (s_y, s_x, win_h, win_w) in canvas
    coordinates, out_off = (top, left) zero margin inside the crop, mean / std: 3 python floats (float64 arithmetic)"""
    _cuda(img_u8, out)
    assert img_u8.dtype == torch.uint8 and img_u8.is_contiguous() and out.dtype == torch.float32 and out.is_contiguous()
    rows, cols = img_u8.shape[:2]
    ch = 1 if img_u8.dim() == 2 else img_u8.shape[2]
    assert out.shape[0] == ch
    _call("cmx_eval_pack_crop", img_u8.data_ptr(), rows, cols, ch, pad[0], pad[1], win[0], win[1], win[2], win[3], out_off[0], out_off[1],
          mean[0], mean[1], mean[2], std[0], std[1], std[2], int(flip), out.data_ptr(), out.shape[1], out.shape[2], _stream(),
          nbytes=_nb(out))
    return out


def eval_accumulate_scale(logits, logits_flip, tiles, margin, rows, cols, processed):
    """processed[ncls, H, W] (fp64) += bilinear resize of one scale's score map (exp of the tile logits summed on the canvas,
    margins sliced off): logits [n, ncls, ch, cw] fp32, tiles: int32 device table [n_tiles, 8] of CmxEvalTile records"""
    _cuda(logits, tiles, processed)
    assert logits.dtype == torch.float32 and logits.is_contiguous() and processed.dtype == torch.float64 and processed.is_contiguous()
    assert tiles.dtype == torch.int32 and tiles.dim() == 2 and tiles.shape[1] == 8 and tiles.is_contiguous()
    assert logits_flip is None or (logits_flip.shape == logits.shape and logits_flip.is_contiguous())
    n, ncls, ch, cw = logits.shape
    _call("cmx_eval_accumulate_scale", logits.data_ptr(), _p(logits_flip), ch, cw, ncls, tiles.data_ptr(), tiles.shape[0], margin[0], margin[1],
          rows, cols, processed.data_ptr(), processed.shape[1], processed.shape[2], _stream(), nbytes=_nb(processed) * 2)
    return processed
