"""Host-side mirror of the reference criteria that `EncoderDecoder` fuses into its loss kernel besides
`nn.CrossEntropyLoss` (reference: utils/loss_opr.py:103-196, selected in train.py:70-93).

`FocalLoss` here is only a parameter carrier with the reference's constructor (`ignore_label, gamma, alpha, reduction`):
pass it (or the reference's own class - it is recognised by name and attributes) as `criterion=`, or as the second
element of the 'CE_Focal' tuple `(nn.CrossEntropyLoss(...), FocalLoss(...))` that builder.py:246-247 combines as
`c0 + 0.2 * c1`.  Calling it on tensors evaluates the same formula in PyTorch (any device) - used by the tests."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class FocalLoss(nn.Module):
    def __init__(self, ignore_label, gamma=2.0, alpha=0.25, reduction='mean'):
        super().__init__()
        self.ignore_label, self.gamma, self.alpha, self.reduction = ignore_label, gamma, alpha, reduction

    def forward(self, pred, target):
        """all-classes focal loss: for every class k, pt = p_k if k is the label else 1 - p_k,
        -a_k (1 - pt)^gamma log(pt + 1e-8) with a_k = alpha / (1 - alpha); summed over classes, averaged over valid pixels"""
        b, c = pred.shape[:2]
        p = torch.softmax(pred.reshape(b, c, -1), dim=1)
        t = target.reshape(b, -1)
        valid = (t != self.ignore_label).to(p.dtype)
        hot = torch.zeros_like(p).scatter_(1, t.clamp(0, c - 1).unsqueeze(1), 1.0)
        pt = hot * p + (1 - hot) * (1 - p)
        a = hot * self.alpha + (1 - hot) * (1 - self.alpha)
        loss = -a * (1 - pt) ** self.gamma * torch.log(pt + 1e-8) * valid.unsqueeze(1)
        if self.reduction == 'mean':
            return loss.sum() / (valid.sum() + 1e-8)
        if self.reduction == 'sum':
            return loss.sum()
        return loss.sum(1)


class DiceCELoss(nn.Module):
    """Parameter carrier with the reference's constructor (utils/loss_opr.py:146-156: `alpha, ignore_index, reduction`);
    `EncoderDecoder` recognises it (or the reference's own class) and runs the fused two-pass kernels.  Calling it on tensors
    evaluates the reference formula in PyTorch (tests): alpha * DiceLoss + (1 - alpha) * CrossEntropy, Dice per (sample, class)
    over the valid pixels with smooth = 1e-6, labels clamped to the class range, mean over (sample, class)."""

    def __init__(self, alpha=0.5, ignore_index=255, reduction='mean', smooth=1e-6):
        super().__init__()
        self.alpha, self.ignore_index, self.reduction, self.smooth = alpha, ignore_index, reduction, smooth

    def forward(self, preds, targets):
        c = preds.shape[1]
        valid = (targets != self.ignore_index).to(preds.dtype).unsqueeze(1)
        p = torch.softmax(preds, dim=1) * valid
        hot = F.one_hot(targets.clamp(0, c - 1), c).permute(0, 3, 1, 2).to(preds.dtype) * valid
        inter = (p * hot).sum(dim=(2, 3))
        union = p.sum(dim=(2, 3)) + hot.sum(dim=(2, 3))
        dice = (2.0 * inter + self.smooth) / (union + self.smooth)
        ce = F.cross_entropy(preds, targets, ignore_index=self.ignore_index, reduction='mean')
        return self.alpha * (1 - dice.mean()) + (1 - self.alpha) * ce
