"""MaskCriterion drop-in (utils.py:6-26) on the sm_100a cross-entropy kernel."""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from .lib import dense


class _MeanCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits2d, target1d):
        R, V = logits2d.shape
        loss = torch.empty((), device=logits2d.device)
        ops.ce_f32(logits2d, R, V, target1d, 0, dense(1), loss)
        ctx.save_for_backward(logits2d, target1d)
        return loss

    @staticmethod
    def backward(ctx, g):
        logits2d, target1d = ctx.saved_tensors
        R, V = logits2d.shape
        dl = torch.empty_like(logits2d)
        scratch = torch.empty((), device=logits2d.device)
        ops.ce_f32(logits2d, R, V, target1d, 0, dense(1), scratch, dlogits=dl, gscale=g.contiguous().to(torch.float32))
        return dl, None


class MaskCriterion(nn.Module):
    """calculate the CrossEntropyLoss "in mask=1 area" -- exactly as the reference does it, i.e. the mean CE
    over ALL B*(L-1) positions: nn.CrossEntropyLoss() already reduced to a scalar, so the mask cancels in
    sum(loss*mask)/sum(mask) (utils.py:19-26).  The division is kept so that an all-zero mask still gives NaN."""

    def forward(self, logits, target, mask):
        item_sum = logits.shape[0] * logits.shape[1]
        target, mask = target[:, 1:], mask[:, 1:]
        loss = _MeanCEFn.apply(logits.contiguous().view(item_sum, -1), target.contiguous().view(-1).to(torch.int64))
        mask_loss = loss * mask.contiguous().view(-1)
        return torch.sum(mask_loss) / torch.sum(mask)
