"""Fused Adam over a flat parameter buffer (train.py:89-93,125: Adam(lr=1e-4), torch defaults)."""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional

import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (no weight decay, no amsgrad) in ONE kernel launch per step.

    On first use the parameters are re-homed into one contiguous fp32 buffer (param.data become views of it, so
    state_dict / named_parameters are unchanged) with a matching flat gradient buffer.  `attach(model)` lets the
    model's backward write gradients straight into that buffer (p.grad become views of it): no gather copy, and
    the data-parallel all-reduce works on contiguous bucket slices."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8):
        params = list(params)
        if params and isinstance(params[0], dict) and len(params) > 1:
            raise ValueError("FusedAdam keeps one flat buffer with one set of hyper-parameters: pass a single parameter group "
                             "(the reference uses Adam(model.parameters(), lr), train.py:89)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._flat = None

    # m, v and the step count live in the flat buffers, not in self.state: carry them through (load_)state_dict explicitly
    def state_dict(self):
        sd = super().state_dict()
        f = self._flat
        if f is not None:
            sd["fused"] = {"step": int(f["step"]), "m": f["m"].clone(), "v": f["v"].clone()}
        return sd

    def load_state_dict(self, state_dict):
        fused = state_dict.get("fused")
        super().load_state_dict({k: v for k, v in state_dict.items() if k != "fused"})
        if fused is not None:
            f = self._ensure_flat()
            if fused["m"].numel() != f["m"].numel():
                raise ValueError("FusedAdam.load_state_dict: flat state has %d elements, this optimizer %d" % (fused["m"].numel(), f["m"].numel()))
            f["m"].copy_(fused["m"]); f["v"].copy_(fused["v"]); f["step"] = int(fused["step"])
            if "step_dev" in f:
                f["step_dev"].fill_(f["step"])

    def _ensure_flat(self):
        if self._flat is not None:
            base = self._flat["p"]
            if all(p.data.data_ptr() == base.data_ptr() + o * 4 for p, o in zip(self._flat["params"], self._flat["offsets"])):
                return self._flat
        ps: List[torch.nn.Parameter] = [p for g in self.param_groups for p in g["params"]]
        dev = ps[0].device
        offsets, n = [], 0
        for p in ps:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FusedAdam needs float32 parameters on one device")
            offsets.append(n)
            n += (p.numel() + 7) // 8 * 8          # keep every tensor 16-byte aligned, in fp32 and in the bf16 shadow
        flat = torch.zeros(n, device=dev)
        old = self._flat
        for p, o in zip(ps, offsets):
            v = flat[o:o + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
        g = torch.zeros(n, device=dev)
        m = torch.zeros(n, device=dev)
        v2 = torch.zeros(n, device=dev)
        step = 0
        if old is not None and old["p"].numel() == n:
            m.copy_(old["m"]); v2.copy_(old["v"]); step = old["step"]
        # bf16 shadow of the weights, refreshed by the Adam kernel itself (the tensor-core path reads weights as bf16)
        shadow = torch.zeros(n, dtype=torch.bfloat16, device=dev)
        self._flat = dict(params=ps, offsets=offsets, p=flat, g=g, m=m, v=v2, step=step, n=n, shadow=shadow, shadow_state=None)
        if dev.type == "cuda":
            # the step counter, lr and this step's bias-corrected scalars also live on the device (ops.adam_prepare): nothing that
            # changes from step to step is a kernel argument, so a captured CUDA graph of the train step can be replayed
            self._flat.update(step_dev=torch.tensor([step], dtype=torch.int32, device=dev), lr_host=None,
                              lr_dev=torch.zeros(1, device=dev), hyper_dev=torch.zeros(2, device=dev))
        return self._flat

    def flat_grad_views(self):
        f = self._ensure_flat()
        return [f["g"][o:o + p.numel()].view(p.shape) for p, o in zip(f["params"], f["offsets"])]

    def attach(self, model, on_bucket_ready: Optional[Callable[[str], None]] = None) -> None:
        """Let `model`'s backward write its gradients directly into this optimizer's flat buffer."""
        views = self.flat_grad_views()
        names = [n for n, _ in model.named_parameters()]
        params = [p for _, p in model.named_parameters()]
        f = self._flat
        if len(params) != len(f["params"]) or any(a is not b for a, b in zip(params, f["params"])):
            raise ValueError("attach(): the optimizer must have been built from model.parameters()")
        model._grad_views = dict(zip(names, views))
        model._on_bucket_ready = on_bucket_ready
        model._adam_shadow = self            # engine_bf16.ShadowCache asks shadow_views() for weights the last step() already cast
        f["names"] = names

    def shadow_views(self, named_params):
        """{name: bf16 view} of the weights as written by the last Adam kernel, or None when they are stale (no step yet, another
        optimizer stepped since, or a parameter was modified in place / re-homed since)."""
        f = self._flat
        if f is None or f["shadow_state"] is None:
            return None
        epoch, versions = f["shadow_state"]
        params = [p for _, p in named_params]
        if epoch != ops.WEIGHT_EPOCH or len(params) != len(f["params"]) or any(a is not b for a, b in zip(params, f["params"])):
            return None
        base = f["p"].data_ptr()
        if versions != tuple(p._version for p in params) or any(p.data_ptr() != base + o * 4 for p, o in zip(params, f["offsets"])):
            return None
        views = {n: f["shadow"][o:o + p.numel()].view(p.shape) for (n, p), o in zip(named_params, f["offsets"])}
        if f.get("derived_ok"):
            views.update(f["derived"])      # W_hh^T (bf16) and b_ih + b_hh of both LSTMs, refreshed right behind their buckets' updates
        return views

    # ---- bucket-wise stepping: Adam on a slice of the flat buffer as soon as that slice's gradient is final, so that most of
    # the (HBM-bound) update runs beside the BPTT sweeps instead of after them.  begin_step() / step_range()* / finish_step().
    def sync_lr(self) -> None:
        """Copies the group's learning rate to the device when it changed (a scheduler wrote it); stream-ordered, outside any graph."""
        f = self._ensure_flat()
        lr = float(self.param_groups[0]["lr"])
        if "lr_dev" in f and f["lr_host"] != lr:
            f["lr_dev"].fill_(lr)
            f["lr_host"] = lr

    @torch.no_grad()
    def gather_grads(self) -> int:
        """Copies every parameter's .grad that is not already a view of the flat gradient buffer into it (zero for a parameter
        without a gradient: its update is then pure momentum decay -- zero for a parameter that never had one).  Returns how many
        tensors had to be copied; 0 on the normal path where backward wrote in place."""
        f = self._ensure_flat()
        moved = 0
        for p, o in zip(f["params"], f["offsets"]):
            gv = f["g"][o:o + p.numel()].view(p.shape)
            if p.grad is None:
                gv.zero_()
                moved += 1
            elif p.grad.data_ptr() != gv.data_ptr():
                gv.copy_(p.grad)
                moved += 1
        return moved

    @torch.no_grad()
    def begin_step(self) -> None:
        f = self._ensure_flat()
        f["step"] += 1
        f["stepped"] = []
        f["active"] = True                      # bucket callbacks arriving outside begin_step()..finish_step() are ignored (dp.py)
        if "step_dev" in f:
            if not torch.cuda.is_current_stream_capturing():
                self.sync_lr()
            group = self.param_groups[0]
            ops.adam_prepare(f["step_dev"], f["lr_dev"], float(group["betas"][0]), float(group["betas"][1]), f["hyper_dev"])

    @torch.no_grad()
    def refresh_derived(self, bucket: str) -> None:
        """Right behind the Adam update of an LSTM's bucket (same stream): what the tensor-core path derives from its weights -- the
        transposed bf16 W_hh for the BPTT kernel and b_ih + b_hh -- is rebuilt into persistent buffers, in the tail of THIS step
        instead of at the head of the next one (where it sits on the serial chain).  In place is safe for the reason it is for the
        bf16 shadows: the bucket is released only after the last reader of its weights."""
        f = self._flat
        if f is None or "step_dev" not in f or "names" not in f or bucket not in ("vid_rnn", "word_rnn"):
            return
        from . import lib as L
        P = dict(zip(f["names"], f["params"]))
        w = P[bucket + ".weight_hh_l0"]
        d = f.setdefault("derived", {})
        kt, kb = bucket + ".weight_hh_l0.T", ("b1" if bucket == "vid_rnn" else "b2")
        if kt not in d:
            d[kt] = torch.empty(w.shape[1], w.shape[0], dtype=torch.bfloat16, device=w.device)
            d[kb] = torch.empty_like(P[bucket + ".bias_ih_l0"])
        L.check(L.load().s2vt_cast_bf16(L.stream_ptr(w.device), L.ptr(w), None, L.ptr(d[kt]), w.shape[0], w.shape[1]), "s2vt_cast_bf16")
        ops.add_f32(P[bucket + ".bias_ih_l0"], P[bucket + ".bias_hh_l0"], d[kb])
        f.setdefault("derived_done", set()).add(bucket)

    def note_replayed_step(self) -> None:
        """Host-side bookkeeping for one replay of a captured train step (the device did begin_step .. finish_step itself)."""
        f = self._flat
        f["step"] += 1
        ops.WEIGHT_EPOCH += 1
        f["shadow_state"] = (ops.WEIGHT_EPOCH, tuple(p._version for p in f["params"]))

    @torch.no_grad()
    def step_range(self, a: int, b: int, grad_scale: float = 1.0) -> None:
        """Adam on flat elements [a, b) (a multiple of 8) with this step's bias correction; launches on the current stream."""
        f = self._flat
        group = self.param_groups[0]
        if "step_dev" in f:
            ops.adam_f32_dev(f["p"][a:b], f["g"][a:b], f["m"][a:b], f["v"][a:b], float(group["betas"][0]), float(group["betas"][1]),
                             float(group["eps"]), f["hyper_dev"], grad_scale=grad_scale, bf16_copy=f["shadow"][a:b])
        else:
            ops.adam_f32(f["p"][a:b], f["g"][a:b], f["m"][a:b], f["v"][a:b], float(group["lr"]), float(group["betas"][0]),
                         float(group["betas"][1]), float(group["eps"]), f["step"], grad_scale=grad_scale, bf16_copy=f["shadow"][a:b])
        f["stepped"].append((a, b))

    @torch.no_grad()
    def finish_step(self, grad_scale: float = 1.0) -> None:
        """Steps whatever begin_step()/step_range() has not covered yet and marks the bf16 shadows current."""
        f = self._flat
        pos = 0
        for a, b in sorted(f["stepped"]) + [(f["n"], f["n"])]:
            if a > pos:
                self.step_range(pos, a, grad_scale)
            pos = max(pos, b)
        f["stepped"] = []
        f["active"] = False
        f["shadow_state"] = (ops.WEIGHT_EPOCH, tuple(p._version for p in f["params"]))
        f["derived_ok"] = f.pop("derived_done", set()) == {"vid_rnn", "word_rnn"}     # (a replayed graph repeats what its capture did)

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        f = self._ensure_flat()
        if f["g"].is_cuda and torch.cuda.is_current_stream_capturing():
            # inside a CUDA-graph capture (dp.GraphedLoopBody, or a user's own torch.cuda.graph around the loop body): nothing that
            # changes from step to step may be a kernel argument, so take the route whose step count / lr / bias corrections live on
            # the device (the one DataParallelTrainer uses)
            self.gather_grads()
            self.begin_step()
            self.step_range(0, f["n"], grad_scale)
            self.finish_step(grad_scale)
            f["derived_ok"] = False
            return loss
        for p, o in zip(f["params"], f["offsets"]):
            gv = f["g"][o:o + p.numel()].view(p.shape)
            if p.grad is None:
                gv.zero_()
            elif p.grad.data_ptr() != gv.data_ptr():
                gv.copy_(p.grad)
        group = self.param_groups[0]
        f["step"] += 1
        if "step_dev" in f:
            f["step_dev"].fill_(f["step"])          # (the bucket-wise path counts on the device; keep both in step)
        ops.adam_f32(f["p"], f["g"], f["m"], f["v"], float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]),
                     float(group["eps"]), f["step"], grad_scale=grad_scale, bf16_copy=f["shadow"])
        f["shadow_state"] = (ops.WEIGHT_EPOCH, tuple(p._version for p in f["params"]))
        f["derived_ok"] = False
        return loss
