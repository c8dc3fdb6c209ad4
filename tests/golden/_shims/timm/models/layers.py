"""Minimal stand-in for timm.models.layers: DropPath, to_2tuple, trunc_normal_ (the only symbols
the reference imports: models/encoders/dual_segformer.py:6, models/net_utils.py:5, engine/evaluator.py:6)."""
import collections.abc
import torch
import torch.nn as nn

trunc_normal_ = nn.init.trunc_normal_


def to_2tuple(x):
    if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
        return tuple(x)
    return (x, x)


class DropPath(nn.Module):
    """Per-sample stochastic depth.  ``forced`` (class attribute) lets the golden generator inject
    the per-sample multipliers instead of sampling them: dict id(module) -> Tensor[B] queue."""
    forced = None

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if DropPath.forced is not None and id(self) in DropPath.forced:
            s = DropPath.forced[id(self)].pop(0)
            return x * s.view(-1, *([1] * (x.dim() - 1)))
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep
