"""Drop-in for the reference utils/metric.py:8-30.  `hist_info` keeps the numpy-in / numpy-out contract of
eval.py:35 but counts on the GPU (single-pass shared-memory histogram kernel, int64, bit-exact); it also accepts
CUDA tensors directly (no host round trip).  `compute_score` stays on the host in float64 like the reference."""
import numpy as np
import torch

from .. import ops

np.seterr(divide='ignore', invalid='ignore')


def _to_dev(a, device):
    if isinstance(a, torch.Tensor):
        t = a
    else:
        a = np.ascontiguousarray(a)
        if a.dtype not in (np.uint8, np.int32, np.int64):
            a = a.astype(np.int64)
        t = torch.from_numpy(a)
    if t.dtype not in (torch.uint8, torch.int32, torch.int64):
        t = t.to(torch.int64)
    return t.to(device, non_blocking=True).contiguous()


def hist_info_device(n_cl, pred, gt, device=None):
    """-> (hist int64 [n_cl,n_cl] CUDA tensor, stats int64[2] = (labeled, correct)) without synchronising."""
    if device is None:
        device = pred.device if isinstance(pred, torch.Tensor) and pred.is_cuda else torch.device("cuda", torch.cuda.current_device())
    assert tuple(pred.shape) == tuple(gt.shape)
    p, g = _to_dev(pred, device), _to_dev(gt, device)
    hist = torch.zeros(n_cl, n_cl, dtype=torch.int64, device=device)
    stats = torch.zeros(2, dtype=torch.int64, device=device)
    ops.confusion(p, g, n_cl, hist, stats)
    return hist, stats


def hist_info(n_cl, pred, gt):
    assert (pred.shape == gt.shape)
    if not torch.cuda.is_available():
        raise RuntimeError("cmx_b200.utils.metric.hist_info needs a CUDA device (no CPU fallback)")
    hist, stats = hist_info_device(n_cl, pred, gt)
    s = stats.cpu()
    return hist.cpu().numpy(), int(s[0]), int(s[1])


def compute_score(hist, correct, labeled):
    if isinstance(hist, torch.Tensor):
        hist = hist.cpu().numpy()
    iou = np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist))
    mean_IoU = np.nanmean(iou)
    mean_IoU_no_back = np.nanmean(iou[1:])
    freq = hist.sum(1) / hist.sum()
    freq_IoU = (iou[freq > 0] * freq[freq > 0]).sum()
    classAcc = np.diag(hist) / hist.sum(axis=1)
    mean_pixel_acc = np.nanmean(classAcc)
    pixel_acc = correct / labeled
    return iou, mean_IoU, mean_IoU_no_back, freq_IoU, mean_pixel_acc, pixel_acc
