"""MaskCriterion drop-in (utils.py:6-26) on the sm_100a cross-entropy kernel."""
from __future__ import annotations

import weakref

import torch
from torch import nn

from . import ops
from .lib import dense, rowmap


class _MeanCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits2d, target1d):
        R, V = logits2d.shape
        loss = torch.empty((), device=logits2d.device)
        # logits produced by this package's tensor-core forward(mode='train'): single-pass CE that keeps each row's log-sum-exp, and a
        # backward that writes dL/dlogits once, as time-major bf16, straight into the producer (model._TrainLogitsBf16Fn)
        from . import model as M
        from . import engine_bf16 as EB
        prod = M.lookup_logits_producer(logits2d) if ctx.needs_input_grad[0] else None
        ctx.prod = None
        if prod is not None:
            lse = torch.empty(R, dtype=torch.float32, device=logits2d.device)
            EB.ce_bf16(logits2d, R, V, target1d, 0, dense(1), loss=loss, row_lse=lse)
            ctx.prod, ctx.lse = weakref.ref(prod), lse
        else:
            ops.ce_f32(logits2d, R, V, target1d, 0, dense(1), loss)
        ctx.save_for_backward(logits2d, target1d)
        return loss

    @staticmethod
    def backward(ctx, g):
        logits2d, target1d = ctx.saved_tensors
        R, V = logits2d.shape
        prod = ctx.prod() if ctx.prod is not None else None
        if prod is not None:
            from . import engine_bf16 as EB
            B, Lm1 = prod.targets.shape
            dl = EB.dlogits_buffer(R, V, logits2d.device)
            ldv = dl.stride(0)
            # logits row r = b*(L-1) + t  ->  gradient row t*B + b
            EB.ce_bf16(logits2d, R, V, target1d, 0, dense(1), dlogits=dl, gscale=g.contiguous().to(torch.float32), row_lse=ctx.lse,
                       have_lse=True, omap=rowmap(Lm1, ldv, B * ldv))
            prod.dl_bf16 = dl if prod.dl_bf16 is None else prod.dl_bf16 + dl
            return torch.zeros((), device=logits2d.device).expand(R, V), None      # zero-stride placeholder (see the producer's backward)
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
