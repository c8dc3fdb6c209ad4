"""Batched sliding-window / multi-scale inference: the caller of the hot path in evaluation
(reference: engine/evaluator.py:306-395, `Evaluator.sliding_eval_rgbX` -> `scale_process_rgbX` ->
`val_func_process_rgbX`).

The reference runs one batch-1 forward per crop (6 per image for scales [0.75, 1, 1.25] on 480x640 with a 480x640 crop:
H2D, forward, exp, slice, accumulate - each on its own).  All crops of all scales have the crop shape, so here they are
collected first and pushed through the model as ONE batch (chunks of `max_batch`), flip-TTA included; exp, margin
slicing and the per-scale accumulation stay on the device in the reference's order and dtype.  Everything else is
kept as the reference does it - `cv2.resize` of the inputs per scale and of each scale's score map back to the original
size (host), float64 accumulation over scales, `argmax(2)` - including the tile indexing of evaluator.py:347-352, which
derives x offsets from the HEIGHT stride/crop: predictions are identical to the reference's whenever
the model is batch-invariant.

Use:  `pred = sliding_eval_rgbX_batched(evaluator, img, modal_x, config.eval_crop_size, config.eval_stride_rate, device)`
in place of `evaluator.sliding_eval_rgbX(...)`; `evaluator` only needs `.multi_scales`, `.class_num`, `.is_flip`,
`.val_func` and the reference's own `.process_image_rgbX` (normalisation + padding, unchanged)."""
import numpy as np
import torch

try:
    import cv2
except ImportError:  # pragma: no cover - cv2 is part of the reference's environment
    cv2 = None


def _to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def _pad_margin(shape_hw, crop):
    """utils/transforms.py:61-75 margins (top, bottom, left, right) that pad an image up to `crop`"""
    ph = max(crop[0] - shape_hw[0], 0)
    pw = max(crop[1] - shape_hw[1], 0)
    return ph // 2, ph // 2 + ph % 2, pw // 2, pw // 2 + pw % 2


def _pad(img, m):
    return cv2.copyMakeBorder(img, int(m[0]), int(m[1]), int(m[2]), int(m[3]), cv2.BORDER_CONSTANT, value=0)


class SlidingEvalContext:
    """The slice of the reference `Evaluator` the driver needs, for callers that do not construct the reference class
    (bench.py, tests): `.class_num`, `.multi_scales`, `.is_flip`, `.val_func` and `process_image_rgbX`, which restates
    engine/evaluator.py:397-430 + utils/transforms.py:61-75,182-187 (float64 normalisation, grey X normalised with
    mean 0 / std 1, zero padding to the crop split as evenly as possible, HWC -> CHW)."""

    def __init__(self, val_func, class_num, multi_scales=(1.0,), is_flip=False, norm_mean=(0.485, 0.456, 0.406),
                 norm_std=(0.229, 0.224, 0.225)):
        self.val_func, self.class_num, self.multi_scales, self.is_flip = val_func, class_num, list(multi_scales), is_flip
        self.norm_mean, self.norm_std = np.asarray(norm_mean, np.float64), np.asarray(norm_std, np.float64)

    @staticmethod
    def _norm(img, mean, std):
        return (img.astype(np.float64) / 255.0 - mean) / std

    def process_image_rgbX(self, img, modal_x, crop_size):
        p = self._norm(img, self.norm_mean, self.norm_std)
        q = self._norm(modal_x, 0, 1) if modal_x.ndim == 2 else self._norm(modal_x, self.norm_mean, self.norm_std)
        m = np.array(_pad_margin(p.shape[:2], crop_size), np.uint32)
        p, q = _pad(p, m), _pad(q, m)
        return p.transpose(2, 0, 1), (q[np.newaxis, ...] if modal_x.ndim == 2 else q.transpose(2, 0, 1)), m


def sliding_eval_rgbX_batched(evaluator, img, modal_x, crop_size, stride_rate, device=None, max_batch=8):
    if cv2 is None:
        raise RuntimeError("sliding_eval_rgbX_batched needs OpenCV (cv2), like the reference evaluator")
    crop = _to_2tuple(crop_size)
    dev = torch.device("cuda" if device is None else device) if not isinstance(device, torch.device) else device
    ori_rows, ori_cols = img.shape[:2]
    ncls = evaluator.class_num
    crops_a, crops_b, plan = [], [], []   # plan: per scale -> dict(whole | tiles)
    for s in evaluator.multi_scales:
        img_s = cv2.resize(img, None, fx=s, fy=s, interpolation=cv2.INTER_LINEAR)
        interp = cv2.INTER_NEAREST if modal_x.ndim == 2 else cv2.INTER_LINEAR
        mx_s = cv2.resize(modal_x, None, fx=s, fy=s, interpolation=interp)
        rows, cols = img_s.shape[:2]
        if cols <= crop[1] or rows <= crop[0]:
            a, b, margin = evaluator.process_image_rgbX(img_s, mx_s, crop)
            plan.append(dict(whole=len(crops_a), margin=[int(v) for v in margin]))
            crops_a.append(a); crops_b.append(b)
            continue
        stride = (int(np.ceil(crop[0] * stride_rate)), int(np.ceil(crop[1] * stride_rate)))
        margin = _pad_margin((rows, cols), crop)
        img_pad, mx_pad = _pad(img_s, margin), _pad(mx_s, margin)
        pad_rows, pad_cols = img_pad.shape[:2]
        r_grid = int(np.ceil((pad_rows - crop[0]) / stride[0])) + 1
        c_grid = int(np.ceil((pad_cols - crop[1]) / stride[1])) + 1
        tiles = []
        for gy in range(r_grid):
            for gx in range(c_grid):
                # evaluator.py:347-352 verbatim (x uses stride[0]/crop[0], y uses stride[1]/crop[1])
                s_x = gx * stride[0]
                s_y = gy * stride[1]
                e_x = min(s_x + crop[0], pad_cols)
                e_y = min(s_y + crop[1], pad_rows)
                s_x = e_x - crop[0]
                s_y = e_y - crop[1]
                a, b, tm = evaluator.process_image_rgbX(img_pad[s_y:e_y, s_x:e_x, :],
                                                        mx_pad[s_y:e_y, s_x:e_x] if mx_pad.ndim == 2 else mx_pad[s_y:e_y, s_x:e_x, :],
                                                        crop)
                tiles.append((len(crops_a), s_y, e_y, s_x, e_x, [int(v) for v in tm]))
                crops_a.append(a); crops_b.append(b)
        plan.append(dict(tiles=tiles, margin=[int(v) for v in margin], pad=(pad_rows, pad_cols)))
    shapes = {c.shape for c in crops_a}
    if len(shapes) != 1:
        # the reference's x/y swap can produce crops of another shape for non-square crops on some image sizes; the
        # reference then calls the model per crop with whatever shape results - do the same, grouped by shape
        groups = {}
        for i, c in enumerate(crops_a):
            groups.setdefault(c.shape, []).append(i)
    else:
        groups = {next(iter(shapes)): list(range(len(crops_a)))}
    scores = [None] * len(crops_a)
    net = evaluator.val_func
    net.eval()
    with torch.no_grad():
        for idxs in groups.values():
            for k in range(0, len(idxs), max_batch):
                sel = idxs[k:k + max_batch]
                a = torch.from_numpy(np.ascontiguousarray(np.stack([crops_a[i] for i in sel]), dtype=np.float32)).to(dev, non_blocking=True)
                b = torch.from_numpy(np.ascontiguousarray(np.stack([crops_b[i] for i in sel]), dtype=np.float32)).to(dev, non_blocking=True)
                sc = net(a, b)
                if evaluator.is_flip:
                    sc = sc + net(a.flip(-1), b.flip(-1)).flip(-1)
                sc = torch.exp(sc)
                for j, i in enumerate(sel):
                    scores[i] = sc[j]
    processed = np.zeros((ori_rows, ori_cols, ncls))
    for pl in plan:
        m = pl["margin"]
        if "whole" in pl:
            score = scores[pl["whole"]]
        else:
            score = torch.zeros(ncls, pl["pad"][0], pl["pad"][1], device=dev)
            for i, s_y, e_y, s_x, e_x, tm in pl["tiles"]:
                t = scores[i]
                t = t[:, tm[0]:(t.shape[1] - tm[1]), tm[2]:(t.shape[2] - tm[3])]
                score[:, s_y:e_y, s_x:e_x] += t
        score = score[:, m[0]:(score.shape[1] - m[1]), m[2]:(score.shape[2] - m[3])]
        out = cv2.resize(score.permute(1, 2, 0).cpu().numpy(), (ori_cols, ori_rows), interpolation=cv2.INTER_LINEAR)
        processed += out
    return processed.argmax(2)


import weakref  # noqa: E402

_UPLOADERS = weakref.WeakKeyDictionary()


class _PinnedUploader:
    """host -> device copies of the per-image uint8 arrays through a small ring of PINNED staging buffers: a copy from pageable
    memory makes the host wait until the stream has drained (the driver stages it chunk by chunk in stream order), which would
    serialise the host work of image i+1 behind the kernels of image i; from pinned memory `non_blocking=True` really is
    asynchronous.  A slot is reused only after the copies issued from it (RING calls ago) have completed (CUDA event)."""
    RING = 3

    def __init__(self):
        self.bufs, self.events, self.call = {}, [None] * self.RING, 0

    def begin(self):
        self.slot = self.call % self.RING
        self.call += 1
        if self.events[self.slot] is not None:
            self.events[self.slot].synchronize()
        self.n = 0

    def put(self, arr, dev):
        arr = np.ascontiguousarray(arr)
        key = (self.slot, self.n, arr.shape, arr.dtype.str)
        self.n += 1
        pin = self.bufs.get(key)
        if pin is None:
            pin = self.bufs[key] = torch.empty(arr.shape, dtype=torch.from_numpy(arr[:0]).dtype).pin_memory()
        pin.numpy()[...] = arr
        return pin.to(dev, non_blocking=True)

    def end(self):
        ev = self.events[self.slot]
        if ev is None:
            ev = self.events[self.slot] = torch.cuda.Event()
        ev.record()


def sliding_eval_rgbX_gpu(evaluator, img, modal_x, crop_size, stride_rate, device=None, max_batch=8, return_device=False, gt=None,
                          accum=None):
    """Device-resident variant of `sliding_eval_rgbX_batched` on the library's own kernels: only the uint8
    `cv2.resize` of the two inputs per scale stays on the host (its fixed-point arithmetic defines the reference's inputs).
    Per scale the resized uint8 images are uploaded once; `cmx_eval_pack_crop` builds every network crop (float64
    normalisation - bit-identical to utils/transforms.py:182-187 -, zero padding of the raw canvas and of the normalised crop,
    tile cut with the reference's index arithmetic, flip-TTA mirror); the model runs once over all crops;
    `cmx_eval_accumulate_scale` does exp, the tile sum, the margin slice, the bilinear resize to the original size and the
    fp64 sum over scales in one pass per scale; `cmx_argmax_confusion` takes the argmax (and, when `gt` is given, the
    confusion matrix of utils/metric.py:8-15) on the device.
    Returns the [H, W] int64 prediction map (host ndarray, or a uint8 device tensor with return_device=True); with `gt`
    (uint8 / int64 ndarray or device tensor) returns (pred, hist[n, n] int64 ndarray, labeled, correct).
    Dataset streaming: with `gt` and `accum=(hist, stats)` - caller-owned device tensors int64 [n, n] and [2] - the confusion
    matrix and the (labeled, correct) counts are ADDED into them on the device and nothing is read back (the uint8 device
    prediction is returned): no host synchronisation per image, so the host work of the next image (cv2 resizes, uploads,
    launches) overlaps the kernels of this one; read `hist` / `stats` once after the last image (engine/evaluator.py:117-137 only
    needs the per-dataset sums).
    Difference to the reference: the score-map resize evaluates cv2.INTER_LINEAR's sampling positions and weights in fp32 on
    the device instead of inside cv2 (rounding differs by ~1 ulp), so the prediction can differ only where two class scores
    tie to ~1e-6 relative - the tie rule the tests state."""
    from .. import ops
    if cv2 is None:
        raise RuntimeError("sliding_eval_rgbX_gpu needs OpenCV (cv2) for the uint8 input resize, like the reference evaluator")
    crop = _to_2tuple(crop_size)
    dev = torch.device("cuda" if device is None else device) if not isinstance(device, torch.device) else device
    ori_rows, ori_cols = img.shape[:2]
    ncls = evaluator.class_num
    grey = modal_x.ndim == 2
    mean = [float(v) for v in np.asarray(evaluator.norm_mean, np.float64)]
    std = [float(v) for v in np.asarray(evaluator.norm_std, np.float64)]
    xmean, xstd = ([0.0] * 3, [1.0] * 3) if grey else (mean, std)
    up = _UPLOADERS.get(evaluator)   # per evaluator, outside its __dict__ (the reference pickles the evaluator into workers)
    if up is None:
        up = _UPLOADERS[evaluator] = _PinnedUploader()
    up.begin()
    # ---- plan (host integers only): per scale the uploaded uint8 pair, its crops and its tile table
    scales, n_crops = [], 0
    for s in evaluator.multi_scales:
        img_s = cv2.resize(img, None, fx=s, fy=s, interpolation=cv2.INTER_LINEAR)
        mx_s = cv2.resize(modal_x, None, fx=s, fy=s, interpolation=cv2.INTER_NEAREST if grey else cv2.INTER_LINEAR)
        rows, cols = img_s.shape[:2]
        sc = dict(a=up.put(img_s, dev), b=up.put(mx_s, dev), rows=rows, cols=cols, crops=[], tiles=[])
        margin = _pad_margin((rows, cols), crop)
        sc["margin"] = margin
        if cols <= crop[1] or rows <= crop[0]:
            # process_image_rgbX on the whole image: normalise, THEN zero-pad (evaluator.py:409-415)
            if rows > crop[0] or cols > crop[1]:
                raise NotImplementedError("sliding_eval_rgbX_gpu: image side larger than the crop on one axis only; use sliding_eval_rgbX_batched")
            sc["crops"].append(dict(pad=(0, 0), win=(0, 0, rows, cols), out=(margin[0], margin[2])))
            sc["tiles"].append((n_crops, 0, 0, crop[0], crop[1], 0, 0))
            sc["canvas_margin"] = (margin[0], margin[2])
            n_crops += 1
        else:
            # scale_process_rgbX: the RAW image is zero-padded (black), tiles are cut, then normalised (evaluator.py:337-360)
            stride = (int(np.ceil(crop[0] * stride_rate)), int(np.ceil(crop[1] * stride_rate)))
            pad_rows, pad_cols = rows + margin[0] + margin[1], cols + margin[2] + margin[3]
            r_grid = int(np.ceil((pad_rows - crop[0]) / stride[0])) + 1
            c_grid = int(np.ceil((pad_cols - crop[1]) / stride[1])) + 1
            for gy in range(r_grid):
                for gx in range(c_grid):
                    s_x = gx * stride[0]           # evaluator.py:347-352 verbatim (x/y swap included)
                    s_y = gy * stride[1]
                    e_x = min(s_x + crop[0], pad_cols)
                    e_y = min(s_y + crop[1], pad_rows)
                    s_x = e_x - crop[0]
                    s_y = e_y - crop[1]
                    # the reference slices numpy arrays with these numbers: negative starts wrap, stops clamp
                    ys, ye, _ = slice(s_y, e_y).indices(pad_rows)
                    xs, xe, _ = slice(s_x, e_x).indices(pad_cols)
                    wh, ww = max(ye - ys, 0), max(xe - xs, 0)
                    if wh > crop[0] or ww > crop[1] or wh == 0 or ww == 0:
                        raise NotImplementedError("sliding_eval_rgbX_gpu: the reference's tile window (%d x %d) does not fit the crop "
                                                  "%s for this image size; use sliding_eval_rgbX_batched" % (wh, ww, crop))
                    tm = _pad_margin((wh, ww), crop)
                    sc["crops"].append(dict(pad=(margin[0], margin[2]), win=(ys, xs, wh, ww), out=(tm[0], tm[2])))
                    sc["tiles"].append((n_crops, ys, xs, ye, xe, tm[0], tm[2]))
                    n_crops += 1
            sc["canvas_margin"] = (margin[0], margin[2])
        scales.append(sc)
    # ---- network inputs: every crop written straight into its slot of the batch tensors
    xch = 1 if grey else 3
    A = torch.empty(n_crops, 3, crop[0], crop[1], device=dev)
    Bx = torch.empty(n_crops, xch, crop[0], crop[1], device=dev)
    flips = (0, 1) if evaluator.is_flip else (0,)
    Af = torch.empty_like(A) if evaluator.is_flip else None
    Bf = torch.empty_like(Bx) if evaluator.is_flip else None
    i = 0
    for sc in scales:
        for c in sc["crops"]:
            for fl in flips:
                ops.eval_pack_crop(sc["a"], c["pad"], c["win"], c["out"], mean, std, fl, (Af if fl else A)[i])
                ops.eval_pack_crop(sc["b"], c["pad"], c["win"], c["out"], xmean, xstd, fl, (Bf if fl else Bx)[i])
            i += 1
    net = evaluator.val_func
    net.eval()
    with torch.no_grad():
        def run(a, b):
            if n_crops <= max_batch:
                return net(a, b)
            return torch.cat([net(a[k:k + max_batch], b[k:k + max_batch]) for k in range(0, n_crops, max_batch)])
        logits = run(A, Bx).contiguous()
        logits_f = run(Af, Bf).contiguous() if evaluator.is_flip else None
    processed = torch.zeros(ncls, ori_rows, ori_cols, dtype=torch.float64, device=dev)
    for sc in scales:
        table = up.put(np.asarray([list(t) + [0] for t in sc["tiles"]], dtype=np.int32), dev)
        ops.eval_accumulate_scale(logits, logits_f, table, sc["canvas_margin"], sc["rows"], sc["cols"], processed)
    pred = torch.empty(ori_rows, ori_cols, dtype=torch.uint8, device=dev)
    gt_dev = None
    if gt is not None:
        gt_dev = up.put(gt, dev) if isinstance(gt, np.ndarray) else gt.to(dev, non_blocking=True)
    up.end()
    if gt is None:
        ops.argmax_confusion(processed, None, ncls, None, None, pred_out=pred)
        return pred if return_device else pred.cpu().numpy().astype(np.int64)
    if accum is not None:
        hist, stats = accum
        assert hist.is_cuda and hist.dtype == torch.int64 and tuple(hist.shape) == (ncls, ncls) and stats.dtype == torch.int64
        ops.argmax_confusion(processed, gt_dev, ncls, hist, stats, pred_out=pred)
        return pred
    hist = torch.zeros(ncls, ncls, dtype=torch.int64, device=dev)
    stats = torch.zeros(2, dtype=torch.int64, device=dev)
    ops.argmax_confusion(processed, gt_dev, ncls, hist, stats, pred_out=pred)
    st = stats.cpu()
    return (pred if return_device else pred.cpu().numpy().astype(np.int64)), hist.cpu().numpy(), int(st[0]), int(st[1])
