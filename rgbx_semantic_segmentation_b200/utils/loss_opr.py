"""Host-side mirror of the reference criteria that `EncoderDecoder` fuses into its loss kernel besides
`nn.CrossEntropyLoss` (reference: utils/loss_opr.py:157-196, selected in train.py:70-93).

`FocalLoss` here is only a parameter carrier with the reference's constructor (`ignore_label, gamma, alpha, reduction`):
pass it (or the reference's own class - it is recognised by name and attributes) as `criterion=`, or as the second
element of the 'CE_Focal' tuple `(nn.CrossEntropyLoss(...), FocalLoss(...))` that builder.py:246-247 combines as
`c0 + 0.2 * c1`.  Calling it on tensors evaluates the same formula in PyTorch (any device) - used by the tests."""
import torch
import torch.nn as nn


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
