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


def _normalize_dev(u8, mean, std):
    """utils/transforms.py:182-187 on the device: float64 (x/255 - mean)/std, then the float32 cast the reference applies in
    val_func_process_rgbX (evaluator.py:375) - IEEE double arithmetic, so bit-identical to the numpy result."""
    x = u8.to(torch.float64) / 255.0
    return ((x - mean) / std).to(torch.float32)


def sliding_eval_rgbX_gpu(evaluator, img, modal_x, crop_size, stride_rate, device=None, max_batch=8, return_device=False):
    """Device-resident variant of `sliding_eval_rgbX_batched`: only the uint8 `cv2.resize` of the two
    inputs per scale stays on the host (its fixed-point arithmetic defines the reference's inputs).  The resized uint8
    images are uploaded once per scale; normalisation (float64, bit-identical), padding, tiling, flip-TTA, exp, tile
    accumulation, the bilinear resize of each scale's score map back to the original size, the float64 sum over scales and
    the argmax all run on the device; one [H, W] int64 map comes back (or stays on the device with return_device=True,
    ready for `utils.metric` / cmx_argmax_confusion).
    Difference to the reference: the score-map resize is `F.interpolate(bilinear, align_corners=False)` instead of
    `cv2.resize(INTER_LINEAR)` on float32 - the same sampling positions and weights, results equal to ~5e-7 - so the
    prediction can differ only where two class scores tie to that precision."""
    import torch.nn.functional as F
    if cv2 is None:
        raise RuntimeError("sliding_eval_rgbX_gpu needs OpenCV (cv2) for the uint8 input resize, like the reference evaluator")
    crop = _to_2tuple(crop_size)
    dev = torch.device("cuda" if device is None else device) if not isinstance(device, torch.device) else device
    ori_rows, ori_cols = img.shape[:2]
    ncls = evaluator.class_num
    mean = torch.as_tensor(np.asarray(evaluator.norm_mean, np.float64), device=dev)
    std = torch.as_tensor(np.asarray(evaluator.norm_std, np.float64), device=dev)
    grey = modal_x.ndim == 2

    def norm_pair(a_u8, b_u8):       # HWC (or HW) uint8 device tensors -> CHW float32
        a = _normalize_dev(a_u8, mean, std).permute(2, 0, 1)
        b = _normalize_dev(b_u8, 0.0, 1.0)[None] if grey else _normalize_dev(b_u8, mean, std).permute(2, 0, 1)
        return a, b

    def pad_to(t, m):                # CHW, zero padding (top, bottom, left, right)
        return F.pad(t, (m[2], m[3], m[0], m[1])) if any(m) else t

    crops_a, crops_b, plan = [], [], []
    for s in evaluator.multi_scales:
        img_s = cv2.resize(img, None, fx=s, fy=s, interpolation=cv2.INTER_LINEAR)
        mx_s = cv2.resize(modal_x, None, fx=s, fy=s, interpolation=cv2.INTER_NEAREST if grey else cv2.INTER_LINEAR)
        rows, cols = img_s.shape[:2]
        a_u8 = torch.from_numpy(np.ascontiguousarray(img_s)).to(dev, non_blocking=True)
        b_u8 = torch.from_numpy(np.ascontiguousarray(mx_s)).to(dev, non_blocking=True)
        margin = _pad_margin((rows, cols), crop)
        if cols <= crop[1] or rows <= crop[0]:
            # process_image_rgbX: normalise, THEN zero-pad (evaluator.py:409-415)
            a, b = norm_pair(a_u8, b_u8)
            plan.append(dict(whole=len(crops_a), margin=margin))
            crops_a.append(pad_to(a, margin)); crops_b.append(pad_to(b, margin))
            continue
        # scale_process_rgbX: the RAW image is zero-padded (black), tiles are cut, then normalised (evaluator.py:337-360)
        stride = (int(np.ceil(crop[0] * stride_rate)), int(np.ceil(crop[1] * stride_rate)))
        a_pad = F.pad(a_u8, (0, 0, margin[2], margin[3], margin[0], margin[1]))
        b_pad = F.pad(b_u8, ((margin[2], margin[3], margin[0], margin[1]) if grey else (0, 0, margin[2], margin[3], margin[0], margin[1])))
        a_n, b_n = norm_pair(a_pad, b_pad)
        pad_rows, pad_cols = a_pad.shape[:2]
        r_grid = int(np.ceil((pad_rows - crop[0]) / stride[0])) + 1
        c_grid = int(np.ceil((pad_cols - crop[1]) / stride[1])) + 1
        tiles = []
        for gy in range(r_grid):
            for gx in range(c_grid):
                s_x = gx * stride[0]           # evaluator.py:347-352 verbatim (x/y swap included)
                s_y = gy * stride[1]
                e_x = min(s_x + crop[0], pad_cols)
                e_y = min(s_y + crop[1], pad_rows)
                s_x = e_x - crop[0]
                s_y = e_y - crop[1]
                ta, tb = a_n[:, s_y:e_y, s_x:e_x], b_n[:, s_y:e_y, s_x:e_x]
                tm = _pad_margin(ta.shape[1:], crop)
                tiles.append((len(crops_a), s_y, e_y, s_x, e_x, tm))
                crops_a.append(pad_to(ta, tm)); crops_b.append(pad_to(tb, tm))
        plan.append(dict(tiles=tiles, margin=margin, pad=(pad_rows, pad_cols)))
    groups = {}
    for i, c in enumerate(crops_a):
        groups.setdefault(tuple(c.shape), []).append(i)
    scores = [None] * len(crops_a)
    net = evaluator.val_func
    net.eval()
    with torch.no_grad():
        for idxs in groups.values():
            for k in range(0, len(idxs), max_batch):
                sel = idxs[k:k + max_batch]
                a = torch.stack([crops_a[i] for i in sel]).contiguous()
                b = torch.stack([crops_b[i] for i in sel]).contiguous()
                sc = net(a, b)
                if evaluator.is_flip:
                    sc = sc + net(a.flip(-1), b.flip(-1)).flip(-1)
                sc = torch.exp(sc)
                for j, i in enumerate(sel):
                    scores[i] = sc[j]
        processed = torch.zeros(ncls, ori_rows, ori_cols, dtype=torch.float64, device=dev)
        for pl in plan:
            m = pl["margin"]
            if "whole" in pl:
                score = scores[pl["whole"]]
            else:
                score = torch.zeros(ncls, pl["pad"][0], pl["pad"][1], device=dev)
                for i, s_y, e_y, s_x, e_x, tm in pl["tiles"]:
                    t = scores[i]
                    score[:, s_y:e_y, s_x:e_x] += t[:, tm[0]:(t.shape[1] - tm[1]), tm[2]:(t.shape[2] - tm[3])]
            score = score[:, m[0]:(score.shape[1] - m[1]), m[2]:(score.shape[2] - m[3])]
            if tuple(score.shape[1:]) != (ori_rows, ori_cols):
                score = F.interpolate(score[None].float(), size=(ori_rows, ori_cols), mode="bilinear", align_corners=False)[0]
            processed += score.double()
        pred = processed.argmax(0)
    return pred if return_device else pred.cpu().numpy()
