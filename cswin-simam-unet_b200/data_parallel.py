"""Batch-sharded data parallelism: one process per GPU, gradients-only all-reduce.

The reference has no distributed code at all.  Every op of both models is per-sample independent
(BatchNorm in the UNet uses per-rank batch statistics), so the only exchange per step
is the gradient average: 23.6 M values for the CSWin-UNet.  Gradients are packed into a few flat
buckets (one multi-tensor copy each); a bucket's all-reduce is launched asynchronously as soon as autograd
has produced its last gradient, so communication overlaps the rest of backward on NCCL's own stream;
NVSwitch makes the cost latency- rather than link-bound, hence few large buckets.
"""
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def shard_of_global_batch(global_batch: int, rank: int, world_size: int) -> range:
    """Contiguous shard [lo, hi) of the global batch owned by `rank`."""
    if global_batch % world_size != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world_size}")
    per = global_batch // world_size
    return range(rank * per, (rank + 1) * per)


class _Bucket:
    def __init__(self, params: List[torch.nn.Parameter], wire_dtype: Optional[torch.dtype] = None):
        self.params = params
        p0 = params[0]
        self.flat = torch.zeros(sum(p.numel() for p in params), dtype=p0.dtype, device=p0.device)
        # the buffer the collective runs on: the bucket itself, or a narrower copy of it (reduce_dtype)
        self.wire = None if wire_dtype in (None, p0.dtype) else torch.zeros_like(self.flat, dtype=wire_dtype)
        self.views, self.wire_views, off = [], [], 0
        for p in params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            if self.wire is not None:
                self.wire_views.append(self.wire[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.pending = len(params)
        self.work = None
        self.packed = False

    def bind(self):
        for p, v in zip(self.params, self.views):
            p.grad = v  # from here on the gradient IS a bucket view


class GradientAllReducer:
    """Averages gradients over the process group, bucket by bucket, overlapped with backward.

    Usage per step:  ``begin_step()`` -> forward / ``loss.backward()`` -> ``finish_step()`` ->
    ``optimizer.step()``; afterwards every ``p.grad`` is a view into a flat bucket.  With world_size == 1
    it only manages the flat gradient buffers.

    Gradients are PACKED after autograd has produced them (one multi-tensor copy per bucket) rather than
    accumulated into pre-bound views: with ``.grad`` unset autograd hands over its own buffer, whereas a
    pre-bound view costs one ``grad += new`` kernel per parameter (463 per step for the CSWin-UNet) plus a
    memset of the buckets.

    ``reduce_dtype=torch.bfloat16`` halves the bytes on the wire: the pack copy converts the gradients into a
    bf16 image of the bucket, the all-reduce averages that, and one more copy widens the result back into the
    fp32 bucket the optimizer reads.  The average is then rounded to bf16 (2^-9 relative);
    the default (None) reduces in the parameters' own type and keeps the 1e-5 gradient parity of SURVEY.md §8(e).
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20,
                 process_group: Optional[dist.ProcessGroup] = None, overlap: bool = True,
                 reduce_dtype: Optional[torch.dtype] = None):
        self.group = process_group
        self.reduce_dtype = reduce_dtype
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.overlap = overlap
        wire = reduce_dtype if self.world > 1 else None  # a single process has nothing to put on a wire
        params = [p for p in params if p.requires_grad]
        # backward produces gradients roughly in reverse registration order (decoder first)
        self.buckets: List[_Bucket] = []
        cur, cur_bytes, cur_key = [], 0, None
        for p in reversed(params):
            key = (p.dtype, p.device)
            if cur and (key != cur_key or cur_bytes + p.numel() * p.element_size() > bucket_bytes):
                self.buckets.append(_Bucket(cur, wire))
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_key = key
            cur_bytes += p.numel() * p.element_size()
        if cur:
            self.buckets.append(_Bucket(cur, wire))
        for b in self.buckets:
            b.bind()
        self._handles = []
        if self.world > 1:
            for b in self.buckets:
                for p in b.params:
                    self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(b)))
        # the backend may lack a native average (gloo): sum, then scale once in finish_step
        nccl = self.world > 1 and dist.get_backend(process_group) == "nccl"
        self._avg = dist.ReduceOp.AVG if nccl else None
        self.capturable = nccl  # NCCL collectives can be recorded into a CUDA graph; gloo's cannot

    def _launch(self, b: _Bucket):
        op = self._avg if self._avg is not None else dist.ReduceOp.SUM
        b.work = dist.all_reduce(b.flat if b.wire is None else b.wire, op=op, group=self.group, async_op=True)

    def _make_hook(self, b: _Bucket):
        def hook(_param):
            b.pending -= 1
            if b.pending == 0 and self.overlap:
                self._pack(b)
                self._launch(b)
        return hook

    def begin_step(self):
        """Unset every gradient (== ``zero_grad(set_to_none=True)``): autograd then hands over its own buffers."""
        for b in self.buckets:
            b.pending = len(b.params)
            b.work = None
            b.packed = False
            for p in b.params:
                p.grad = None

    @staticmethod
    def _pack(b: _Bucket):
        srcs, dsts = [], []
        for p, v in zip(b.params, b.views if b.wire is None else b.wire_views):
            if p.grad is None:
                v.zero_()  # took no part in this backward
            elif p.grad.data_ptr() != v.data_ptr():  # (a wire view never aliases a gradient: always copied)
                srcs.append(p.grad)
                dsts.append(v)
        if srcs:
            torch._foreach_copy_(dsts, srcs)
        b.bind()
        b.packed = True

    def pack(self):
        """Copy the gradients autograd produced into the flat buckets (stream-ordered, capturable in a CUDA
        graph) and re-point every ``p.grad`` at its bucket view.  ``finish_step()`` calls it."""
        with torch.no_grad():
            for b in self.buckets:
                if not b.packed:
                    self._pack(b)

    def finish_step(self):
        self.pack()
        if self.world == 1:
            return
        for b in self.buckets:
            if b.work is None:  # overlap off, or a parameter took no part in this backward
                self._launch(b)
        for b in self.buckets:
            b.work.wait()
            if b.wire is not None:
                b.flat.copy_(b.wire)  # widen the averaged bf16 image back into the bucket the optimizer reads
            if self._avg is None:
                b.flat.div_(self.world)
            # a captured train step replays backward WITHOUT calling begin_step() again: the next
            # finish_step() must launch a fresh all-reduce, not wait on this finished one
            b.work = None

    def gradient_bytes(self) -> int:
        """Bytes one rank hands to the collective per step."""
        return sum((b.flat if b.wire is None else b.wire).numel() * (b.flat if b.wire is None else b.wire).element_size()
                   for b in self.buckets)

    def close(self):
        for h in self._handles:
            h.remove()
        self._handles = []
