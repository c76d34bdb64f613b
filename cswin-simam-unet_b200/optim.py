"""``optimizer.step()`` of the reference train loops (C:786 with AdamW C:937-941; U:348 with Adam U:486-490)
as ONE kernel launch over every parameter tensor (csb200_adam_step), which also writes the bf16 shadows
the Linear / conv layers read under autocast.  Drop-in for ``torch.optim.AdamW`` / ``torch.optim.Adam``
(amsgrad=False, maximize=False): same constructor arguments, ``param_groups``, per-parameter
``state[p] = {"step", "exp_avg", "exp_avg_sq"}`` (checkpoints interchange with torch's).

Hyper-parameters and the step count live in device memory: a captured CUDA graph follows a learning-rate
schedule (the reference drives ReduceLROnPlateau, C:943-949) after ``sync_hyperparameters()``.
"""
import ctypes
from typing import Optional

import torch

from . import capi


class _AdamTensor(ctypes.Structure):  # == csb200_adam_tensor
    _fields_ = [("param", ctypes.c_void_p), ("grad", ctypes.c_void_p), ("exp_avg", ctypes.c_void_p),
                ("exp_avg_sq", ctypes.c_void_p), ("shadow", ctypes.c_void_p), ("numel", ctypes.c_int64),
                ("group", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class FusedAdamW(torch.optim.Optimizer):
    """AdamW (``decoupled=True``, default) or Adam with L2 weight decay (``decoupled=False``)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, decoupled=True):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("FusedAdamW: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay,
                                      decoupled=bool(decoupled)))
        self._steps = None          # device scalar: number of steps taken
        self._hyper = None          # device [groups][8]
        self._hyper_host = None     # what the device copy holds
        self._table_key = None
        self._keep = None           # host staging tensors (pinned) + device tables of the current key
        self._shadows = {}          # id(param) -> bf16 shadow written by the kernel
        self._frozen = {}           # pointer-set key -> tables owned by a captured CUDA graph

    # ---- checkpoints --------------------------------------------------------------------------------
    def load_state_dict(self, state_dict):
        """torch's layout (one ``step`` per parameter): the kernel reads ONE device counter, restored here from
        the loaded steps (they are equal for every parameter that has taken part in every step)."""
        super().load_state_dict(state_dict)
        for g in self.param_groups:  # a torch.optim.Adam(W) checkpoint: no "decoupled" key
            if "decoupled" not in g:
                g["decoupled"] = bool(g.get("decoupled_weight_decay", self.defaults["decoupled"]))
        steps = [st["step"] for st in self.state.values() if "step" in st]
        if steps:
            first = steps[0]
            dev = next(st["exp_avg"].device for st in self.state.values() if "exp_avg" in st)
            value = float(first.item() if torch.is_tensor(first) else first)
            if any(float(t.item() if torch.is_tensor(t) else t) != value for t in steps):
                raise RuntimeError("FusedAdamW: parameters with different step counts are not supported")
            self._steps = torch.full((), value, dtype=torch.float32, device=dev)
            for st in self.state.values():
                if "step" in st:
                    st["step"] = self._steps
            if self._hyper is None:
                self._hyper_host = self._hyper_rows()
                self._hyper = torch.tensor(self._hyper_host, dtype=torch.float32, device=dev)
        self._table_key = None
        if self._frozen:
            raise RuntimeError("FusedAdamW: load_state_dict after a step has been captured in a CUDA graph "
                               "(the graph holds the old moment buffers); load before the first captured step")

    # ---- bf16 shadows (functional.shadow_params) ------------------------------------------------
    def attach_shadows(self, masters, shadows):
        """From now on the kernel writes ``shadow <- bf16(master)`` for these pairs after every update."""
        new = {id(p): s for p, s in zip(masters, shadows)}
        if {k: v.data_ptr() for k, v in new.items()} != {k: v.data_ptr() for k, v in self._shadows.items()}:
            if self._frozen:
                raise RuntimeError("FusedAdamW: the shadows cannot change after a step has been captured")
            self._shadows = new
            self._table_key = None

    writes_shadows = True

    # ---- device-resident hyper-parameters -------------------------------------------------------
    def _hyper_rows(self):
        return [[float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                 float(g["weight_decay"]), 1.0 if g["decoupled"] else 0.0, 0.0, 0.0] for g in self.param_groups]

    def sync_hyperparameters(self):
        """Copy lr / betas / eps / weight_decay of ``param_groups`` to the device if they changed (call
        before replaying a captured step when a scheduler may have touched them)."""
        rows = self._hyper_rows()
        if self._hyper is not None and rows != self._hyper_host:
            self._hyper.copy_(torch.tensor(rows, dtype=torch.float32), non_blocking=False)
            self._hyper_host = rows

    def _init_state(self, dev):
        if self._steps is None:
            self._steps = torch.zeros((), dtype=torch.float32, device=dev)
            self._hyper_host = self._hyper_rows()
            self._hyper = torch.tensor(self._hyper_host, dtype=torch.float32, device=dev)
        for g in self.param_groups:
            for p in g["params"]:
                st = self.state[p]
                if "exp_avg" not in st:
                    st["step"] = self._steps  # shared device scalar (torch keeps one per parameter)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)

    def _tables(self, items, dev, freeze=False):
        """(device table of tensors, device table of chunks, chunk count) for this set of pointers."""
        key = tuple((p.data_ptr(), p.grad.data_ptr(), gi) for p, gi in items)
        if key in self._frozen:
            return self._frozen[key]
        if key == self._table_key and not freeze:
            return self._keep["dev_t"], self._keep["dev_c"], self._keep["n_chunks"]
        chunk = int(capi.lib().csb200_adam_chunk_elems())
        arr = (_AdamTensor * len(items))()
        numels = []
        for i, (p, gi) in enumerate(items):
            st = self.state[p]
            sh = self._shadows.get(id(p))
            arr[i] = _AdamTensor(p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(),
                                 st["exp_avg_sq"].data_ptr(), 0 if sh is None else sh.data_ptr(), p.numel(), gi, 0)
            numels.append(p.numel())
        nbytes = ctypes.sizeof(arr)
        k = self._keep
        if k is None or k["numels"] != numels or k["dev_t"].device != dev:
            chunks = [v for i, n in enumerate(numels) for c in range(-(-n // chunk)) for v in (i, c)]
            dev_c = torch.tensor(chunks, dtype=torch.int32).to(dev)
            k = self._keep = {"numels": numels, "dev_c": dev_c, "n_chunks": len(chunks) // 2, "flip": 0,
                              "uploaded": [None, None],
                              "host": [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)],
                              "dev_t": torch.empty(nbytes, dtype=torch.uint8, device=dev)}
        k["flip"] ^= 1  # two pinned staging buffers: the previous upload may still be queued on the stream
        host = k["host"][k["flip"]]
        if k["uploaded"][k["flip"]] is not None:
            k["uploaded"][k["flip"]].synchronize()  # the upload that last read THIS buffer (two rebuilds ago)
        ctypes.memmove(host.data_ptr(), ctypes.addressof(arr), nbytes)
        if freeze:
            # a table a captured CUDA graph will read on every replay: its own device copy, never rewritten
            # (an eager step() of the same optimizer with other gradient tensors builds another table)
            dev_t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            dev_t.copy_(host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            k["uploaded"][k["flip"]] = ev
            self._frozen[key] = (dev_t, k["dev_c"], k["n_chunks"])
            return self._frozen[key]
        k["dev_t"].copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        k["uploaded"][k["flip"]] = ev
        self._table_key = key
        return k["dev_t"], k["dev_c"], k["n_chunks"]

    def _collect(self):
        items, dev = [], None
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise RuntimeError("FusedAdamW needs fp32 CUDA parameters and gradients "
                                       "(there is no CPU or mixed-dtype fallback)")
                # the update is elementwise in STORAGE order: any dense layout works (contiguous, or
                # channels-last conv weights) as long as parameter, gradient and moments share it
                dense = p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last))
                same = all(a == b for a, b, n in zip(p.stride(), p.grad.stride(), p.shape) if n > 1)  # size-1 dims: any stride
                if not dense or not same:
                    raise RuntimeError("FusedAdamW needs dense parameters whose gradients have the same strides "
                                       f"(parameter {tuple(p.shape)}: strides {p.stride()} vs {p.grad.stride()})")
                dev = dev or p.device
                items.append((p, gi))
        return items, dev

    @torch.no_grad()
    def prepare(self, freeze=False):
        """Allocate the moments and upload the pointer tables for the CURRENT ``.grad`` tensors.  ``step()``
        does this itself; call it with ``freeze=True`` before capturing ``step()`` into a CUDA graph (host-side
        table building must not happen inside the capture): the captured step then finds everything in
        place as long as the gradients keep their addresses — true for the gradients of a captured backward."""
        items, dev = self._collect()
        if not items:
            return None
        self._init_state(dev)
        if not torch.cuda.is_current_stream_capturing():
            self.sync_hyperparameters()
        return self._tables(items, dev, freeze)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        tables = self.prepare()
        if tables is None:
            return loss
        dev_t, dev_c, n_chunks = tables
        self._steps.add_(1.0)
        with torch.cuda.device(dev_t.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev_t.device).cuda_stream)
            capi.check(capi.lib().csb200_adam_step(ctypes.c_void_p(dev_t.data_ptr()), ctypes.c_void_p(dev_c.data_ptr()),
                                                   n_chunks, ctypes.c_void_p(self._hyper.data_ptr()),
                                                   ctypes.c_void_p(self._steps.data_ptr()), stream), "csb200_adam_step")
        return loss


def fused_adam(params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0) -> FusedAdamW:
    """torch.optim.Adam semantics (L2 weight decay added to the gradient), U:486-490."""
    return FusedAdamW(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=False)
