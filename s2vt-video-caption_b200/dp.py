"""Data-parallel training of the S2VT hot path: one process per GPU, identical replicas, the batch sharded
across ranks, and ONE exchange step -- a bucketed all-reduce (average) of the flat gradient buffer over NCCL
(NVLink 5 / NVSwitch), issued bucket by bucket on a side stream as soon as backward has produced each group.

The reference has no distributed code; because its loss is a mean over B*(L-1) positions and
every rank sees the same B, averaging rank gradients equals the single-process gradient of the concatenated batch.

Bucket order = order in which backward finishes them: out_linear -> word_rnn -> embedding -> vid_rnn -> feat_linear.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

BUCKETS: Tuple[Tuple[str, Tuple[str, ...]], ...] = (
    ("out_linear", ("out_linear.weight", "out_linear.bias")),
    ("word_rnn", ("word_rnn.weight_ih_l0", "word_rnn.weight_hh_l0", "word_rnn.bias_ih_l0", "word_rnn.bias_hh_l0")),
    ("embedding", ("embedding.weight",)),
    ("vid_rnn", ("vid_rnn.weight_ih_l0", "vid_rnn.weight_hh_l0", "vid_rnn.bias_ih_l0", "vid_rnn.bias_hh_l0")),
    ("feat_linear", ("feat_linear.weight", "feat_linear.bias")),
)


def bucket_ranges(names: Sequence[str], offsets: Sequence[int], sizes: Sequence[int]) -> Dict[str, Tuple[int, int]]:
    """Contiguous [begin, end) element range of each bucket inside the flat gradient buffer.  Raises if a bucket's
    tensors are not adjacent (they are, for the reference's registration order)."""
    pos = {n: (o, o + s) for n, o, s in zip(names, offsets, sizes)}
    out = {}
    for bname, members in BUCKETS:
        spans = sorted(pos[m] for m in members)
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            if b0 - a1 > 7:                      # alignment padding only
                raise ValueError("bucket %s is not contiguous in the flat buffer" % bname)
        out[bname] = (spans[0][0], spans[-1][1])
    return out


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard of n_items for `rank` (sizes differ by at most one; 1970 videos / 8 -> 247,247,246,...)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradAllReducer:
    """Averages a flat gradient buffer across ranks, one bucket at a time.  Device-agnostic (gloo on CPU in tests,
    NCCL on the GPUs): `ready(bucket)` may be called as backward progresses; `finish()` joins everything."""

    def __init__(self, flat_grad: torch.Tensor, ranges: Dict[str, Tuple[int, int]], group=None, overlap: bool = True):
        self.flat, self.ranges, self.group = flat_grad, ranges, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cuda = flat_grad.is_cuda
        self.overlap = overlap and self.cuda and self.world > 1
        self.comm_stream = torch.cuda.Stream(device=flat_grad.device) if self.overlap else None
        self.pending: List = []
        self.done: set = set()
        self.bytes_reduced = 0

    def ready(self, bucket: str) -> None:
        if self.world == 1 or bucket in self.done:
            self.done.add(bucket)
            return
        a, b = self.ranges[bucket]
        view = self.flat[a:b]
        self.done.add(bucket)
        self.bytes_reduced += view.numel() * view.element_size()
        if self.overlap:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group)      # NCCL averages in the collective
        else:
            if self.cuda:
                dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group)
            else:                                # gloo (CPU tests) has no AVG
                work = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                self.pending.append((work, view))

    def finish(self) -> None:
        for bname, _ in BUCKETS:
            if bname not in self.done:
                self.ready(bname)
        if self.overlap:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.comm_stream)
        for work, view in self.pending:
            work.wait()
            view.mul_(1.0 / self.world)
        self.pending.clear()
        self.done.clear()


class DataParallelTrainer:
    """model + FusedAdam + gradient all-reduce: `step(feats, targets)` is the reference's train-loop body
    (train.py:116-127) on this rank's shard."""

    def __init__(self, model, optimizer, group=None, overlap: bool = True):
        self.model, self.opt = model, optimizer
        f = optimizer._ensure_flat()
        names = [n for n, _ in model.named_parameters()]
        sizes = [p.numel() for p in f["params"]]
        self.ranges = bucket_ranges(names, f["offsets"], sizes)
        self.reducer = GradAllReducer(f["g"], self.ranges, group=group, overlap=overlap)
        # Adam per bucket, right behind that bucket's all-reduce: needs gradients that backward writes in place (CUDA path)
        self.early_adam = f["g"].is_cuda
        optimizer.attach(model, on_bucket_ready=self._bucket_ready)

    def _bucket_ready(self, bucket: str) -> None:
        """Called by backward once every kernel producing `bucket`'s gradients has been enqueued (and nothing later in the step
        reads that bucket's weights): all-reduce it, then update it, both beside the rest of backward."""
        self.reducer.ready(bucket)
        if not self.early_adam:
            return
        a, b = self.ranges[bucket]
        a, b = a // 8 * 8, (b + 7) // 8 * 8
        if self.reducer.overlap:
            with torch.cuda.stream(self.reducer.comm_stream):        # stream order: after this bucket's all-reduce
                self.opt.step_range(a, b)
        else:
            self.opt.step_range(a, b)

    def step(self, feats, targets, mask=None):
        self.opt.zero_grad(set_to_none=True)
        loss = self.model.forward_loss(feats, targets, mask)
        self.opt.begin_step()
        loss.backward()
        self.reducer.finish()
        self.opt.finish_step()
        return loss
