"""Training step of the reference (C:780-786 / U:342-348) as a reusable, sync-free callable.

    optimizer.zero_grad(); out = model(x); loss = BCELoss(out, y); loss.backward(); optimizer.step()

Precision policy (the reference is fp32-only): ``precision="bf16"`` keeps fp32 master
weights and runs the forward under ``torch.autocast(bfloat16)`` — GEMMs, convolutions and the csb200
attention / SimAM kernels see bf16 activations, LayerNorm / softmax statistics stay fp32 — while the
sigmoid and the BCE loss are evaluated in fp32 outside the autocast region (CUDA autocast refuses
BCELoss on probabilities).  No ``.item()`` is called here: the reference's five host syncs per step
(C:797-806) are the caller's choice, not the step's.
"""
import contextlib
import os
from typing import Tuple

import torch
import torch.nn.functional as F

from . import functional as csbF


def synthetic_batch(batch: int, size: int, device, seed: int = 0, first_index: int = 0,
                    pin: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Images ~ U[0,1) (B,3,S,S) and binary masks (B,1,S,S), as BASELINE.md §4 specifies.

    Sample i of the GLOBAL batch is generated from (seed, first_index + i) alone, so an n-GPU run and
    a 1-GPU run see the same global batch however it is sharded.
    """
    imgs = torch.empty((batch, 3, size, size), dtype=torch.float32, pin_memory=pin)
    masks = torch.empty((batch, 1, size, size), dtype=torch.float32, pin_memory=pin)
    for i in range(batch):
        g = torch.Generator().manual_seed(seed * 1_000_003 + first_index + i)
        imgs[i] = torch.rand((3, size, size), generator=g)
        masks[i] = (torch.rand((1, size, size), generator=g) > 0.5).float()
    if torch.device(device).type == "cpu":
        return imgs, masks
    return imgs.to(device, non_blocking=pin), masks.to(device, non_blocking=pin)


def bce_from_logits_as_probabilities(probs: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """nn.BCELoss() on sigmoid outputs, fp32 (C:936, C:782)."""
    return F.binary_cross_entropy(probs.float(), target.float())


class InferStep:
    """``model(images)`` in eval mode under ``torch.no_grad()`` (the validation pass C:795-806, the high-resolution
    inference of BASELINE config 5), optionally replayed from a CUDA graph.

    A 1024^2 forward is ~1 000 kernel launches that take 6.6 ms on a B200 and 13.7 ms to ISSUE from Python, so
    eager inference is host-bound; ``cuda_graph=True`` captures the forward for the first input shape it sees (one
    graph per shape; the input is copied into a static buffer, the result is a view of the graph's output buffer
    that the next call overwrites — clone it to keep it).  Attention / path dropout are identities in eval mode.
    """

    def __init__(self, model: torch.nn.Module, precision: str = "bf16", cuda_graph: bool = True):
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.model, self.precision, self.cuda_graph = model, precision, cuda_graph
        self._graphs = {}  # (shape, dtype) -> (graph, static input, static output)

    def _forward(self, images):
        with torch.no_grad():
            if self.precision == "bf16":
                with torch.autocast(device_type=images.device.type, dtype=torch.bfloat16):
                    return self.model(images)
            return self.model(images)

    def __call__(self, images: torch.Tensor) -> torch.Tensor:
        if self.model.training:
            raise RuntimeError("InferStep needs model.eval() (BatchNorm statistics, dropout)")
        if not (self.cuda_graph and images.is_cuda):
            return self._forward(images)
        key = (tuple(images.shape), images.dtype)
        entry = self._graphs.get(key)
        if entry is None:
            static_in = images.clone()
            side = torch.cuda.Stream(device=images.device)
            side.wait_stream(torch.cuda.current_stream(images.device))
            with torch.cuda.stream(side):  # lazy initialisation (cuDNN plans, tensor maps, bf16 shadows) outside the capture
                for _ in range(2):
                    self._forward(static_in)
            torch.cuda.current_stream(images.device).wait_stream(side)
            torch.cuda.synchronize(images.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self._forward(static_in)
            entry = self._graphs[key] = (graph, static_in, static_out)
        graph, static_in, static_out = entry
        static_in.copy_(images, non_blocking=True)
        graph.replay()
        return static_out


class TrainStep:
    """One optimisation step; returns the loss as a device tensor (no host sync).

    ``cuda_graph=True`` captures zero_grad -> forward -> loss -> backward -> [all-reduce] -> optimizer
    into ONE CUDA graph on the first call and replays it afterwards: the step is
    ~5000 kernel launches, which the Python host cannot issue as fast as a B200 executes them.  Requires
    static shapes, a capturable optimizer (``AdamW(..., capturable=True)``) and no host syncs inside the
    model (true for the models of this package).  The lazy-initialisation warm-up runs that capture
    needs are undone (parameters, buffers and optimizer state are restored), so the first replay is
    step 1 of training exactly as in eager mode.
    """

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, precision: str = "bf16",
                 reducer=None, cuda_graph: bool = False, capture_collectives: bool = True):
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.model, self.optimizer, self.precision, self.reducer = model, optimizer, precision, reducer
        self.cuda_graph = cuda_graph
        self.capture_collectives = capture_collectives  # data parallel: NCCL all-reduces inside the captured backward
        self.defer_sums = os.environ.get("CSB200_DEFER_SUMS", "1") != "0"  # see _backward
        self._zero_arena_numel = 0
        self._reduce_in_graph = False
        self._graph = self._graph_opt = None
        self._shadow = None  # (fp32 masters, bf16 shadows): see functional.shadow_params

    def _optimizer_step(self):
        """optimizer.step() + ONE multi-tensor refresh of the bf16 shadows the Linear / conv layers read
        (instead of one cast kernel per layer per step)."""
        self.optimizer.step()
        if self.precision == "bf16":
            if self._shadow is None:
                params = [p for g in self.optimizer.param_groups for p in g["params"]]
                self._shadow = csbF.shadow_params(params)
                if getattr(self.optimizer, "writes_shadows", False):
                    self.optimizer.attach_shadows(*self._shadow)  # csb200_adam_step refreshes them from now on
            elif not getattr(self.optimizer, "writes_shadows", False):
                csbF.refresh_shadows(*self._shadow)
            elif not torch.cuda.is_current_stream_capturing():
                # the kernel wrote the shadows; re-stamp them in case something bumped a parameter's version
                # counter since (load_state_dict), else every layer would cast per call from then on
                csbF.restamp_shadows(*self._shadow)

    def _autocast(self, device_type: str):
        if self.precision == "bf16":
            return torch.autocast(device_type=device_type, dtype=torch.bfloat16)
        return contextlib.nullcontext()

    def _backward(self, loss: torch.Tensor) -> None:
        """loss.backward() with the ~100 "sum the per-CTA partials" launches of the pass (LayerNorm parameter
        gradients, bias column sums) batched into one at its end (``functional.deferred_sums``).  The conditions
        hold here: every ``.grad`` was just set to None, the gradients are first read after this returns (the
        reducer's hooks only launch collectives mid-backward in overlap mode, which therefore keeps the immediate
        sums), and the whole step runs on the current stream."""
        overlapped = self.reducer is not None and self.reducer.world > 1 and self.reducer.overlap
        if loss.is_cuda and self.defer_sums and not overlapped:
            # the weight-gradient kernels accumulate into zeroed outputs: one arena, zeroed once, sized by what
            # the previous pass asked for (the first pass zeroes per call)
            with csbF.deferred_sums(loss.device, zero_arena_numel=self._zero_arena_numel) as block:
                loss.backward()
            self._zero_arena_numel = block.arena_demand
        else:
            loss.backward()

    def forward_loss(self, images: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        with self._autocast(images.device.type):
            probs = self.model(images)
        return bce_from_logits_as_probabilities(probs, masks)

    def __call__(self, images: torch.Tensor = None, masks: torch.Tensor = None) -> torch.Tensor:
        """One step on (images, masks); with no arguments, on the batch staged by ``prefetch``."""
        slot = None
        if images is None:
            images, masks, slot = self._take_staged()
        loss = self._graph_step(images, masks, slot) if self.cuda_graph else self._eager_step(images, masks)
        if slot is not None and not self.cuda_graph:
            self._mark_consumed(slot)  # eager: the staged tensors are read until the end of the step
        return loss

    # ---- input pipeline: host -> device copy of the NEXT batch under the current step ------------
    def prefetch(self, images: torch.Tensor, masks: torch.Tensor) -> None:
        """Start the host->device copy of the next batch (pinned host tensors) on a side stream; the
        following ``step()`` call (no arguments) consumes it.  The copy engine works while the SMs run
        the current step, which is how a training loop with a data loader behaves."""
        dev = next(self.model.parameters()).device
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staged = None
            self._stage_bufs = [None, None]
            self._stage_idx = 0
            self._consumed = [None, None]
        i = self._stage_idx = self._stage_idx ^ 1
        cs = self._copy_stream
        # two staging pairs alternate; overwriting pair i only has to wait for the step that last READ it
        # (not for the step running now, which reads the other pair): that is the whole overlap
        if self._consumed[i] is not None:
            cs.wait_event(self._consumed[i])
        with torch.cuda.stream(cs):
            if self._stage_bufs[i] is None:
                self._stage_bufs[i] = (torch.empty(images.shape, dtype=images.dtype, device=dev),
                                       torch.empty(masks.shape, dtype=masks.dtype, device=dev))
            x, y = self._stage_bufs[i]
            x.copy_(images, non_blocking=True)
            y.copy_(masks, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        self._staged = (x, y, ev, i)

    def _take_staged(self):
        if getattr(self, "_staged", None) is None:
            raise RuntimeError("step() without arguments needs a batch staged by prefetch()")
        x, y, ev, slot = self._staged
        self._staged = None
        torch.cuda.current_stream(x.device).wait_event(ev)
        return x, y, slot

    def _mark_consumed(self, slot):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._consumed[slot] = ev

    # ---- CUDA-graph path -----------------------------------------------------------------------
    def _graph_step(self, images, masks, slot=None):
        if self._graph is None:
            self._capture(images, masks)
        self._x.copy_(images, non_blocking=True)
        self._y.copy_(masks, non_blocking=True)
        if slot is not None:
            self._mark_consumed(slot)  # the graph reads its own static copies from here on
        self._graph.replay()
        if self._graph_opt is not None:
            if self.reducer is not None and not self._reduce_in_graph:
                self.reducer.finish_step()  # fallback: gradients averaged eagerly between the two graphs
            if hasattr(self.optimizer, "sync_hyperparameters"):
                self.optimizer.sync_hyperparameters()  # a scheduler may have changed lr since the capture
            self._graph_opt.replay()
        return self._loss

    def _capture(self, images, masks):
        dev = next(self.model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("cuda_graph=True needs the model on a CUDA device")
        self._x, self._y = images.to(dev, copy=True), masks.to(dev, copy=True)
        images = self._x
        saved_model = {k: v.clone() for k, v in self.model.state_dict().items()}
        saved_opt = self._snapshot_optimizer()  # restored state / eager steps before the capture survive it
        side = torch.cuda.Stream(device=images.device)
        side.wait_stream(torch.cuda.current_stream(images.device))
        with torch.cuda.stream(side):
            for _ in range(3):  # lazy init: cuBLAS/cuDNN handles and autotune, optimizer state
                self._eager_step(self._x, self._y)
        torch.cuda.current_stream(images.device).wait_stream(side)
        # undo the warm-up: weights / buffers and the optimizer state (moments, step counters) back to what
        # they were before it; state the warm-up created is zeroed (== "not yet stepped")
        with torch.no_grad():
            self.model.load_state_dict(saved_model)
            self._restore_optimizer(saved_opt)
            if self._shadow is not None:
                csbF.refresh_shadows(*self._shadow)
        self._graph = torch.cuda.CUDAGraph()
        self._graph_opt = None
        own_tables = hasattr(self.optimizer, "prepare")  # csb200 FusedAdamW: pointer tables built on the host
        if self.reducer is None and not own_tables:
            self.optimizer.zero_grad(set_to_none=True)
            with torch.cuda.graph(self._graph):
                self._loss = self._eager_step(self._x, self._y)
            return
        if self.reducer is None:
            # graph 1 = forward + backward (its gradient buffers keep their addresses over replays),
            # host: pointer tables for exactly those buffers, graph 2 = the one-launch optimizer step
            self.optimizer.zero_grad(set_to_none=True)
            with torch.cuda.graph(self._graph):
                loss = self.forward_loss(self._x, self._y)
                self._backward(loss)
                self._loss = loss.detach()
            self.optimizer.prepare(freeze=True)
            torch.cuda.synchronize(dev)
            self._graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_opt, pool=self._graph.pool()):
                self._optimizer_step()
            return
        # Data parallel: graph 1 = forward + backward + the bucket all-reduces, CAPTURED (NCCL's stream is forked
        # from / joined to the capture by events), so nothing is issued from the host between the graphs; graph 2 =
        # optimizer step.  With ``reducer.overlap`` the all-reduce of a bucket is recorded as soon as backward has
        # produced its last gradient (it then runs under the rest of backward), otherwise after backward.  If the
        # process group cannot be captured (gloo, an old NCCL) the all-reduces stay eager between the two graphs.
        self._reduce_in_graph = self.reducer.world > 1 and self.reducer.capturable and self.capture_collectives
        try:
            self._capture_dp_graph(in_graph=self._reduce_in_graph)
        except RuntimeError as e:
            if not self._reduce_in_graph:
                raise
            import warnings
            warnings.warn("the NCCL all-reduce could not be captured into the backward graph "
                          f"({str(e).splitlines()[0][:200]}); it is issued eagerly between the two graphs instead")
            torch.cuda.synchronize(dev)
            self._reduce_in_graph = False
            self._graph = torch.cuda.CUDAGraph()
            self._capture_dp_graph(in_graph=False)
        if own_tables:
            self.optimizer.prepare(freeze=True)
            torch.cuda.synchronize(dev)
        self._graph_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph_opt, pool=self._graph.pool()):
            self._optimizer_step()

    def _capture_dp_graph(self, in_graph: bool):
        if not in_graph:
            self.reducer.overlap = False  # the hooks must not launch eager collectives inside a capture
        with torch.cuda.graph(self._graph):
            self.reducer.begin_step()
            loss = self.forward_loss(self._x, self._y)
            self._backward(loss)
            if in_graph:
                self.reducer.finish_step()  # launches what the hooks have not, joins NCCL's stream to the capture
            else:
                self.reducer.pack()  # gradients -> flat buckets, inside the graph
            self._loss = loss.detach()

    def _snapshot_optimizer(self):
        return {p: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                for p, st in self.optimizer.state.items()}

    def _restore_optimizer(self, snap):
        """In place (a later capture records these very tensors): copy the snapshot back, zero what is new."""
        done = set()  # FusedAdamW shares one device step counter between all parameters
        for p, st in self.optimizer.state.items():
            old = snap.get(p, {})
            for k, v in st.items():
                if not torch.is_tensor(v) or id(v) in done:
                    continue
                done.add(id(v))
                if torch.is_tensor(old.get(k)):
                    v.copy_(old[k])
                else:
                    v.zero_()

    def refresh_shadows(self):
        """Re-synchronise the bf16 parameter shadows with the fp32 masters (and re-stamp them as current).
        Call after writing parameters behind the optimizer's back — ``model.load_state_dict``, EMA swaps,
        ``p.data.copy_`` — while a bf16 TrainStep is live; parameters must not be modified through ``.data``
        without it (a stale shadow would be read by the next forward)."""
        if self._shadow is not None:
            csbF.refresh_shadows(*self._shadow)

    def _eager_step(self, images: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        if self.reducer is not None:
            self.reducer.begin_step()  # == zero_grad(set_to_none=True)
        else:
            self.optimizer.zero_grad(set_to_none=True)
        loss = self.forward_loss(images, masks)
        self._backward(loss)
        if self.reducer is not None:
            self.reducer.finish_step()  # wait for the bucketed all-reduces
        self._optimizer_step()
        return loss.detach()
