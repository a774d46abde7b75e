"""Data-parallel training of the S2VT hot path: one process per GPU, identical replicas, the batch sharded
across ranks, and ONE exchange step -- a bucketed all-reduce (average) of the flat gradient buffer over NCCL
(NVLink 5 / NVSwitch), issued bucket by bucket on a side stream as soon as backward has produced each group.

The reference has no distributed code; because its loss is a mean over B*(L-1) positions and
every rank sees the same B, averaging rank gradients equals the single-process gradient of the concatenated batch.

Buckets, in the order the exact path finishes them: out_linear (weight) -> word_rnn -> embedding (+ out_linear.bias) -> vid_rnn ->
feat_linear; the tensor-core path releases the embedding bucket before word_rnn.  Every rank runs the same code, hence the same order.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

BUCKETS: Tuple[Tuple[str, Tuple[str, ...]], ...] = (
    # out_linear's bias (a column sum over the 131 MB of dlogits) travels with the embedding table it is adjacent to in the flat
    # buffer: the first all-reduce, 26.6 MB, then starts right behind the weight-gradient product instead of one more pass later
    ("out_linear", ("out_linear.weight",)),
    ("word_rnn", ("word_rnn.weight_ih_l0", "word_rnn.weight_hh_l0", "word_rnn.bias_ih_l0", "word_rnn.bias_hh_l0")),
    ("embedding", ("out_linear.bias", "embedding.weight")),
    ("vid_rnn", ("vid_rnn.weight_ih_l0", "vid_rnn.weight_hh_l0", "vid_rnn.bias_ih_l0", "vid_rnn.bias_hh_l0")),
    ("feat_linear", ("feat_linear.weight", "feat_linear.bias")),
)


def bucket_ranges(names: Sequence[str], offsets: Sequence[int], sizes: Sequence[int], buckets=None) -> Dict[str, Tuple[int, int]]:
    """Contiguous [begin, end) element range of each bucket inside the flat gradient buffer.  Raises if a bucket's
    tensors are not adjacent (they are, for the reference's registration order).  `buckets`: ((name, members), ...), default the
    S2VT module's BUCKETS; a model may bring its own as `DP_BUCKETS` (Att_Baseline)."""
    pos = {n: (o, o + s) for n, o, s in zip(names, offsets, sizes)}
    out = {}
    for bname, members in (buckets if buckets is not None else BUCKETS):
        spans = sorted(pos[m] for m in members)
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            if b0 - a1 > 7:                      # alignment padding only
                raise ValueError("bucket %s is not contiguous in the flat buffer" % bname)
        out[bname] = (spans[0][0], spans[-1][1])
    return out


def bind_to_local_numa(device_index: int) -> Optional[int]:
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off (sysfs: the PCI device's numa_node and that node's cpulist).
    Pinned host buffers allocated afterwards are first-touched on that node, so a rank's H2D copies do not cross the socket
    interconnect -- with 8 ranks each feeding 84 MB per step that link, not PCIe, is what saturates.  Returns the node, or None when
    the topology is not exposed (single node, container without sysfs, ...); never raises."""
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard of n_items for `rank` (sizes differ by at most one; 1970 videos / 8 -> 247,247,246,...)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradAllReducer:
    """Averages a flat gradient buffer across ranks, one bucket at a time.  Device-agnostic (gloo on CPU in tests,
    NCCL on the GPUs): `ready(bucket)` may be called as backward progresses; `finish()` joins everything."""

    def __init__(self, flat_grad: torch.Tensor, ranges: Dict[str, Tuple[int, int]], group=None, overlap: bool = True,
                 transport_dtype: Optional[torch.dtype] = None):
        """transport_dtype=torch.bfloat16: a bucket crosses NVLink as bf16 (cast, all-reduce, cast back: half the bytes on the wire for
        two extra passes over the bucket in HBM); the ranks still end with bit-identical gradients.  Default: fp32 as stored."""
        self.flat, self.ranges, self.group = flat_grad, ranges, group
        self.transport_dtype = transport_dtype if flat_grad.is_cuda else None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cuda = flat_grad.is_cuda
        self.overlap = overlap and self.cuda and self.world > 1
        self.comm_stream = torch.cuda.Stream(device=flat_grad.device) if self.overlap else None
        # Communication lanes: each lane = its own NCCL communicator + stream, and the buckets alternate between the lanes in the
        # order backward releases them.  On one communicator the five all-reduces form a serial chain of latency-bound collectives
        # (8-27 MB each at 150-210 GB/s, far below NVLink 5) that ends 0.2 ms after the step would; two lanes overlap neighbouring
        # all-reduces (measured at 2 GPUs: 1.555 -> 1.498 ms per step; bf16 transport, by contrast, made it slower: 1.587).
        # S2VT_COMM_GROUPS sets the number of lanes (default 3: 8 GPUs 1.710 -> 1.580 ms with two lanes, 1.562 with three; 1 = the single chain).
        self.lanes = [(self.comm_stream, group)]
        n_lanes = max(1, int(os.environ.get("S2VT_COMM_GROUPS", "3")))
        if self.overlap and group is None:
            for _ in range(n_lanes - 1):
                self.lanes.append((torch.cuda.Stream(device=flat_grad.device), dist.new_group(backend="nccl")))
        self.n_released = 0                      # buckets alternate between the lanes in the order backward releases them
        self.last_stream = self.comm_stream
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
            cstream, cgroup = self.lanes[self.n_released % len(self.lanes)]
            self.n_released += 1
            self.last_stream = cstream
            with torch.cuda.stream(cstream):
                cstream.wait_event(ev)
                from . import ops
                if ops.MARKS is not None:                                          # tools/timeline_step.py
                    ops._mark("B all_reduce[%s %.1f MB]" % (bucket, view.numel() * 4 / 1e6))
                if self.transport_dtype is not None:
                    wire = view.to(self.transport_dtype)
                    dist.all_reduce(wire, op=dist.ReduceOp.AVG, group=cgroup)
                    view.copy_(wire)
                else:
                    dist.all_reduce(view, op=dist.ReduceOp.AVG, group=cgroup)      # NCCL averages in the collective
                if ops.MARKS is not None:
                    ops._mark("E all_reduce[%s %.1f MB]" % (bucket, view.numel() * 4 / 1e6))
        else:
            if self.cuda:
                dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group)
            else:                                # gloo (CPU tests) has no AVG
                work = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                self.pending.append((work, view))

    def finish(self) -> None:
        for bname in self.ranges:
            if bname not in self.done:
                self.ready(bname)
        if self.overlap:
            for cstream, _ in self.lanes:
                torch.cuda.current_stream(self.flat.device).wait_stream(cstream)
        for work, view in self.pending:
            work.wait()
            view.mul_(1.0 / self.world)
        self.pending.clear()
        self.done.clear()
        self.n_released = 0


class DataParallelTrainer:
    """model + FusedAdam + gradient all-reduce: `step(feats, targets)` is the reference's train-loop body
    (train.py:116-127) on this rank's shard."""

    def __init__(self, model, optimizer, group=None, overlap: bool = True, cuda_graph: Optional[bool] = None,
                 grad_transport_dtype: Optional[torch.dtype] = None):
        self.model, self.opt = model, optimizer
        if grad_transport_dtype is None and os.environ.get("S2VT_GRAD_BF16") == "1":
            grad_transport_dtype = torch.bfloat16
        f = optimizer._ensure_flat()
        names = [n for n, _ in model.named_parameters()]
        sizes = [p.numel() for p in f["params"]]
        self.ranges = bucket_ranges(names, f["offsets"], sizes, getattr(model, "DP_BUCKETS", None))
        self.reducer = GradAllReducer(f["g"], self.ranges, group=group, overlap=overlap, transport_dtype=grad_transport_dtype)
        # S2VT_EARLY_OUT_WGRAD=1: backward produces out_linear's weight gradient (the first, biggest bucket) ahead of the serial chain so
        # that its all-reduce starts ~0.3 ms earlier.  Measured at 2 GPUs: 1.602 ms per step against 1.561 without -- the product costs
        # the chain more than the earlier all-reduce saves, and its NCCL CTAs then compete with the sweeps' clusters -- so it is off.
        model._dp_world = self.reducer.world if os.environ.get("S2VT_EARLY_OUT_WGRAD") == "1" else 1
        # Adam per bucket, right behind that bucket's all-reduce: needs gradients that backward writes in place (CUDA path)
        self.early_adam = f["g"].is_cuda
        optimizer.attach(model, on_bucket_ready=self._bucket_ready)
        # The step is ~70 kernel launches on 6 streams; enqueueing them from Python takes longer than the GPU needs to run them.
        # After two eager steps (kernel code loaded, caches warm) the whole step -- forward, backward, all-reduce, Adam -- is captured
        # into a CUDA graph per distinct (feats, targets) buffer pair and replayed; nothing step-dependent is baked into it
        # (Adam's step count / lr live on the device).  S2VT_CUDA_GRAPH=0 or cuda_graph=False keeps every step eager.
        if cuda_graph is None:
            cuda_graph = os.environ.get("S2VT_CUDA_GRAPH", "1") != "0"
        from . import ops
        self.use_graph = bool(cuda_graph) and f["g"].is_cuda and not ops.serialising_profiler_attached()
        self._graphs: Dict[tuple, tuple] = {}
        self._eager_steps = 0
        self._pool = None
        self.max_graphs = 8
        self._adam_stream = None
        self._execs: List[int] = []
        # A captured step is tied to the addresses of its inputs.  Batches from a loader live in fresh tensors every step, so the
        # trainer owns one static (feats, targets) pair per input shape: step() copies the batch in (or the loader gathers straight
        # into input_buffers()) and replays that pair's graph.  Buffers passed to register_inputs() are replayed in place.
        self._static: Dict[tuple, Tuple[torch.Tensor, torch.Tensor]] = {}
        self._registered: set = set()
        self.replays = 0
        self.steps_since_check = 0

    def __del__(self):
        try:
            from .lib import load
            for e in getattr(self, "_execs", []):
                load().s2vt_graph_exec_destroy(e)
        except Exception:                        # interpreter shutdown
            pass

    def _bucket_ready(self, bucket: str) -> None:
        """Called by backward once every kernel producing `bucket`'s gradients has been enqueued (and nothing later in the step
        reads that bucket's weights): all-reduce it, then update it, both beside the rest of backward."""
        if not (self.opt._flat or {}).get("active"):
            return               # a plain loss.backward() outside trainer.step(): FusedAdam.step() gathers and steps everything itself
        if bucket.endswith(":reduce"):                   # gradients final, weights still being read: start the all-reduce, update later
            self.reducer.ready(bucket[:-7])
            return
        self.reducer.ready(bucket)
        if not self.early_adam:
            return
        a, b = self.ranges[bucket]
        a, b = a // 8 * 8, (b + 7) // 8 * 8
        if self.reducer.overlap:
            # behind this bucket's all-reduce, but on a stream of its own: the next bucket's all-reduce does not wait for this update
            ev = torch.cuda.Event()
            ev.record(self.reducer.last_stream)          # the stream this bucket's all-reduce went to
            ev_here = torch.cuda.Event()                 # (the all-reduce may have been started earlier than this call: ':reduce')
            ev_here.record(torch.cuda.current_stream(self.reducer.flat.device))
            if self._adam_stream is None:
                self._adam_stream = torch.cuda.Stream(device=self.reducer.flat.device)
            with torch.cuda.stream(self._adam_stream):
                self._adam_stream.wait_event(ev)
                self._adam_stream.wait_event(ev_here)
                self.opt.step_range(a, b)
                if self._needs_derived():
                    self.opt.refresh_derived(bucket)
        else:
            self.opt.step_range(a, b)
            if self._needs_derived():
                self.opt.refresh_derived(bucket)

    def _needs_derived(self) -> bool:
        """Does the model's engine read the optimizer-maintained bf16 shadows / derived weight forms (engine_bf16)?  Engines that cast
        their own private copies at the head of every step (engine_step, Att_Baseline) do not, and their captured steps need no
        'shadows are current' precondition."""
        fn = getattr(self.model, "_needs_derived_shadows", None)
        return True if fn is None else bool(fn())

    # ---- input buffers
    def input_buffers(self, feats_shape, targets_shape, dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
        """The trainer-owned (feats, targets) pair for this input shape; a loader that fills these and passes them to step() pays no
        copy (data.DeviceFeatureStore.batches(into=...))."""
        key = (tuple(feats_shape), dtype, tuple(targets_shape))
        ent = self._static.get(key)
        if ent is None:
            dev = self.reducer.flat.device
            ent = (torch.empty(feats_shape, dtype=dtype, device=dev), torch.empty(targets_shape, dtype=torch.int64, device=dev))
            self._static[key] = ent
            self._registered.add((ent[0].data_ptr(), ent[1].data_ptr()))
        return ent

    def register_inputs(self, feats: torch.Tensor, targets: torch.Tensor) -> None:
        """Declare a caller-owned, long-lived (feats, targets) pair (e.g. double buffers behind an H2D copy stream): steps on it are
        captured and replayed in place, without the copy into the trainer's own buffers.  At most `max_graphs` pairs."""
        self._registered.add((feats.data_ptr(), targets.data_ptr()))

    def release_graphs(self) -> None:
        """Drop every captured step (they hold NCCL kernels when world > 1): call before dist.destroy_process_group()."""
        if self._graphs:
            torch.cuda.synchronize()
        self._graphs.clear()
        self._registered.clear()
        self._static.clear()
        self._pool = None
        from .lib import load
        for e in self._execs:
            load().s2vt_graph_exec_destroy(e)
        self._execs = []

    def check_device_errors(self) -> None:
        """Synchronises and raises if a tensor-core kernel flagged a lost barrier arrival (sm100_ptx.cuh: a wait that gives up after
        2 s sets a sticky flag and the kernel carries on with garbage).  fit() / validate() call this once per epoch."""
        from .lib import load, S2VTLibraryError
        self.steps_since_check = 0
        code = load().s2vt_device_error_flag(None)
        if code != 0:
            load().s2vt_device_error_clear()
            raise S2VTLibraryError("a tensor-core kernel reported a timed-out barrier wait (code %d): the results of the steps since the "
                                   "last check are not trustworthy" % code)

    def _step_eager(self, feats, targets, mask=None):
        self.opt.zero_grad(set_to_none=True)
        loss = self.model.forward_loss(feats, targets, mask)
        self.opt.begin_step()
        loss.backward()
        # backward normally writes each gradient straight into the optimizer's flat buffer (and fires the bucket callbacks); when it
        # could not (a frozen parameter, a device mismatch) autograd produced ordinary .grad tensors: bring them into the flat buffer
        # before it is reduced and stepped, instead of training on whatever the buffer held
        self.opt.gather_grads()
        self.reducer.finish()
        if self._adam_stream is not None:
            torch.cuda.current_stream(self.reducer.flat.device).wait_stream(self._adam_stream)
        self.opt.finish_step()
        # (detached: a caller that keeps the loss must not keep this step's autograd graph -- and with it the parameters' AccumulateGrad
        # nodes, which are bound to the stream they were created on -- alive into a later step's graph capture)
        return loss.detach()

    def step(self, feats, targets, mask=None):
        from . import ops
        if not (self.use_graph and feats.is_cuda and self.model._use_bf16()) or ops._PROFILE is not None:
            self._eager_steps += 1
            return self._step_eager(feats, targets, mask)
        # A captured step reads the optimizer's bf16 weight shadows (and the derived W_hh^T / bias sums) as the previous step left them.
        # If anything else touched the weights since (load_state_dict, an in-place edit, another optimizer), run this step eagerly: it
        # re-derives them from the fp32 masters and leaves everything current again.
        f = self.opt._flat
        if self._needs_derived() and (f.get("shadow_state") != (ops.WEIGHT_EPOCH, tuple(p._version for p in f["params"])) or
                                      not f.get("derived_ok")):
            self._eager_steps += 1
            return self._step_eager(feats, targets, mask)
        if self._eager_steps < 2 or feats.requires_grad:         # (feats.grad belongs to the caller's tensor: no static copy of it)
            self._eager_steps += 1
            return self._step_eager(feats, targets, mask)
        if (feats.data_ptr(), targets.data_ptr()) not in self._registered or \
                (len(self._graphs) >= self.max_graphs and
                 (feats.data_ptr(), targets.data_ptr(), tuple(feats.shape), tuple(targets.shape), False, feats.dtype) not in self._graphs):
            sf, st_ = self.input_buffers(feats.shape, targets.shape, feats.dtype)
            sf.copy_(feats, non_blocking=True)
            st_.copy_(targets, non_blocking=True)
            feats, targets = sf, st_
        key = (feats.data_ptr(), targets.data_ptr(), tuple(feats.shape), tuple(targets.shape), False, feats.dtype)
        ent = self._graphs.get(key)
        if ent is None:
            ent = self._capture(key, feats, targets, mask)
        graph, loss, n_launch, _keep, exec_ = ent
        self.opt.sync_lr()
        from .lib import load, check, stream_ptr
        if exec_ is not None:
            check(load().s2vt_graph_launch(exec_, stream_ptr(feats.device)), "s2vt_graph_launch")
        else:
            graph.replay()
        self.opt.note_replayed_step()
        load().s2vt_add_launch_count(n_launch)
        self.replays += 1
        return loss

    def _capture(self, key, feats, targets, mask):
        from .lib import launch_count
        f = self.opt._flat
        host_step, epoch = f["step"], None
        self.opt.sync_lr()
        torch.cuda.synchronize(feats.device)
        # S2VT_GRAPH_NODE_PRIORITY=1: instantiate the executable graph here with cudaGraphInstantiateFlagUseNodePriority (torch passes no
        # such flag, so every node runs at the launch stream's priority).  Measured (tools/probe_priority.py): CTAs are dispatched in
        # launch order with or without node priorities, so the default is torch's own replay; the wave front does not rely on priorities.
        own = os.environ.get("S2VT_GRAPH_NODE_PRIORITY", "0") == "1"
        graph = torch.cuda.CUDAGraph(keep_graph=True) if own else torch.cuda.CUDAGraph()
        n0 = launch_count()
        with torch.cuda.graph(graph, pool=self._pool):
            loss = self._step_eager(feats, targets, mask)
        n_launch = launch_count() - n0
        if self._pool is None:
            self._pool = graph.pool()
        f["step"] = host_step                    # capturing enqueued nothing: the step happens at replay
        exec_ = None
        if own:
            import ctypes
            from .lib import load, check
            out = ctypes.c_void_p()
            check(load().s2vt_graph_instantiate(graph.raw_cuda_graph(), 1, ctypes.byref(out)), "s2vt_graph_instantiate")
            exec_ = out.value
            self._execs.append(exec_)
        ent = (graph, loss, n_launch, (feats, targets), exec_)
        self._graphs[key] = ent
        return ent


class GraphedLoopBody:
    """The reference's UNCHANGED train-loop body (train.py:116-127) captured once into a CUDA graph and replayed:

        def body(feats, targets, masks):
            optimizer.zero_grad()
            loss = criterion(model(feats, targets[:, :-1], mode='train'), targets, masks)
            loss.backward()
            optimizer.step()
            return loss
        step = GraphedLoopBody(body, (feats, targets, masks), optimizer=optimizer)
        loss = step(feats, targets, masks)          # copies the batch into the captured buffers, replays

    What an eager step loses on this path is not host time but launch order and gaps on the device (1.73 ms against 1.34 ms for the
    same kernels); a replayed graph gets them back without touching the body.  Requirements: fixed input shapes,
    a FusedAdam optimizer (its step() keeps the step count / lr on the device while capturing) and no host reads (.item()) inside the
    body.  The returned loss is the captured output tensor (overwritten by the next call)."""

    def __init__(self, body, example_inputs, optimizer=None, warmup: int = 3):
        self.body, self.opt = body, optimizer
        self.static = [x.clone() if torch.is_tensor(x) else x for x in example_inputs]
        dev = next(x.device for x in self.static if torch.is_tensor(x))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                        # (warm-up off the capture stream, as torch's CUDA-graph notes ask)
            for _ in range(max(2, warmup)):
                out = body(*self.static)
                del out                                      # no autograd graph of a warm-up step may outlive it (AccumulateGrad streams)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        host_step = optimizer._flat["step"] if optimizer is not None and getattr(optimizer, "_flat", None) else None
        if optimizer is not None and hasattr(optimizer, "sync_lr"):
            optimizer.sync_lr()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = body(*self.static)
        if host_step is not None:
            optimizer._flat["step"] = host_step              # capturing enqueued nothing: the step happens at replay
        self.out = self.out.detach() if torch.is_tensor(self.out) else self.out
        self.replays = 0

    def __call__(self, *inputs):
        for s_, x in zip(self.static, inputs):
            if torch.is_tensor(s_) and x.data_ptr() != s_.data_ptr():
                s_.copy_(x, non_blocking=True)
        if self.opt is not None and hasattr(self.opt, "sync_lr"):
            self.opt.sync_lr()
        self.graph.replay()
        if self.opt is not None and hasattr(self.opt, "note_replayed_step"):
            self.opt.note_replayed_step()
        self.replays += 1
        return self.out
