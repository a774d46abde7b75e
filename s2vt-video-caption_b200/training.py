"""Training-loop policy around the step (train.py:95-168, utils.py:29-80): ReduceLROnPlateau, the validation pass, EarlyStopping,
whole-module checkpoints -- host-side only; every step runs on the sm_100a kernels through DataParallelTrainer / the drop-in modules.

    store = DeviceFeatureStore(captions, feats_dir);  valid = DeviceFeatureStore(captions, feats_dir, mode='valid')
    model = S2VT(len(store.word2ix), 4096, 80, dim_hid=512, dim_embed=512).cuda()
    fit(model, store, valid, epochs=300, batch_size=16, save_path='checkpoint/')
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from .criterion import MaskCriterion
from .dp import DataParallelTrainer
from .optim import FusedAdam


class EarlyStopping:
    """utils.py:29-80 (itself from Bjarten/early-stopping-pytorch): stop when the validation loss has not improved by `delta` for
    `patience` epochs; every improvement saves the WHOLE module with torch.save (train.py loads it back with torch.load)."""

    def __init__(self, patience=7, verbose=False, delta=0, path="checkpoint.pt", trace_func=print):
        self.patience, self.verbose, self.delta, self.path, self.trace_func = patience, verbose, delta, path, trace_func
        self.counter = 0
        self.best_score = None
        self.early_stop = False
        self.val_loss_min = np.inf                      # the reference's np.Inf was removed in numpy 2

    def __call__(self, val_loss, model):
        score = -val_loss
        if self.best_score is None:
            self.best_score = score
            self.save_checkpoint(val_loss, model)
        elif score < self.best_score + self.delta:
            self.counter += 1
            self.trace_func(f"EarlyStopping counter: {self.counter} out of {self.patience}")
            if self.counter >= self.patience:
                self.early_stop = True
        else:
            self.best_score = score
            self.save_checkpoint(val_loss, model)
            self.counter = 0

    def save_checkpoint(self, val_loss, model):
        if self.verbose:
            self.trace_func(f"Validation loss decreased ({self.val_loss_min:.6f} --> {val_loss:.6f}).  Saving model ...")
        torch.save(model, self.path)
        self.val_loss_min = val_loss


@torch.no_grad()
def validate(model, store, batch_size: int) -> float:
    """train.py:138-149: mean over batches of criterion(model(feats, targets[:, :-1], 'train'), targets, masks) under no_grad."""
    crit = MaskCriterion()
    total, n = 0.0, 0
    model.eval()
    for feats, targets, _, masks in store.batches(batch_size, shuffle=False):
        if hasattr(model, "forward_loss"):
            loss = model.forward_loss(feats, targets, masks)
        else:
            loss = crit(model(feats, targets=targets[:, :-1], mode="train"), targets, masks)
        total += float(loss.item())
        n += 1
    if next(model.parameters()).is_cuda:
        from .lib import load, S2VTLibraryError
        code = load().s2vt_device_error_flag(None)
        if code != 0:
            load().s2vt_device_error_clear()
            raise S2VTLibraryError("a tensor-core kernel reported a timed-out barrier wait (code %d) during validation" % code)
    return total / max(1, n)


def _mean_over_ranks(x: float, device, group) -> float:
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item()) / dist.get_world_size(group)


def fit(model, train_store, valid_store, epochs: int = 300, batch_size: int = 16, lr: float = 1e-4, lr_patience: int = 20,
        early_stopping_patience: int = 30, save_freq: int = 100, save_path: Optional[str] = None, tag: str = "",
        log: Optional[Callable[[Dict], None]] = None, group=None, seed: int = 0) -> List[Dict]:
    """The reference's train() (train.py:56-168) with its Opt() defaults: Adam(lr) -> ReduceLROnPlateau(patience) on the validation
    loss -> EarlyStopping(patience) -> whole-module checkpoints every `save_freq` epochs and at the end.  Returns the per-epoch
    history (the scalars the reference writes to TensorBoard: train_loss, valid_loss, lr).

    Under torch.distributed (one process per GPU) every rank draws the SAME epoch permutation from `seed` and takes every world-th
    batch of it (the tail that does not fill a round is dropped, so all ranks issue the same collectives); the validation loss is
    averaged over ranks before the scheduler and the stopper see it, so every rank takes the same decisions; only rank 0 writes
    checkpoints."""
    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    dev = next(model.parameters()).device
    opt = FusedAdam(model.parameters(), lr=lr)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, patience=lr_patience)        # train.py:95-97 (verbose= was removed from torch)
    stopper = None
    if save_path is not None:
        if rank == 0:
            os.makedirs(save_path, exist_ok=True)
        stopper = EarlyStopping(patience=early_stopping_patience, verbose=False,
                                path=os.path.join(save_path, tag + "stop.pth") if rank == 0 else os.devnull)
    fused = hasattr(model, "forward_loss")
    trainer = DataParallelTrainer(model, opt, group=group) if fused else None
    crit = MaskCriterion()
    history = []
    for epoch in range(epochs):
        model.train()
        running, n = 0.0, 0
        gen = torch.Generator().manual_seed(seed + epoch) if distributed else None
        n_batches = (len(train_store) + batch_size - 1) // batch_size
        usable = n_batches - n_batches % world
        for bi, (feats, targets, _, masks) in enumerate(train_store.batches(batch_size, shuffle=True, generator=gen, into=trainer,
                                                                             only=(rank, world, usable) if distributed else None)):
            if fused:
                loss = trainer.step(feats, targets, masks)
            else:                                                                        # e.g. Att_Baseline: module + criterion
                opt.zero_grad(set_to_none=True)
                loss = crit(model(feats, targets=targets[:, :-1], mode="train"), targets, masks)
                loss.backward()
                opt.step()
            running += float(loss.item())                                               # train.py:127 reads the loss every step
            n += 1
        if trainer is not None and dev.type == "cuda":
            trainer.check_device_errors()               # raises if a tensor-core kernel gave up on a barrier during this epoch
        rec = {"epoch": epoch, "train_loss": _mean_over_ranks(running / max(1, n), dev, group),
               "valid_loss": _mean_over_ranks(validate(model, valid_store, batch_size), dev, group), "lr": opt.param_groups[0]["lr"]}
        history.append(rec)
        if log is not None:
            log(rec)
        sched.step(rec["valid_loss"])
        if stopper is not None:
            stopper(rec["valid_loss"], model)
            if stopper.early_stop:
                break
            if epoch % save_freq == 0 and rank == 0:
                torch.save(model, os.path.join(save_path, tag + str(epoch) + ".pth"))
    if save_path is not None and rank == 0:
        torch.save(model, os.path.join(save_path, tag + "final.pth"))
    return history
