"""torch.autograd bindings of the csb200 kernels.

Tensors stay ``torch.Tensor``; device pointers come from ``data_ptr()`` and kernels are enqueued on
torch's current stream.  Nothing here computes on the host and no csb200 kernel has a
CPU or ATen stand-in.  What deliberately stays on library code is said where it happens: convolutions
(cuDNN, ``conv2d``), GEMMs the tcgen05 token-path kernels do not tile (cuBLAS, ``linear``), and the column
sums / bias adds of widths that are not a multiple of the 16-byte vector (``_bias_grad``, ``add_row_bias``).
"""
import ctypes
import os
from typing import Optional, Sequence, Tuple

import torch

from . import capi

_vp = ctypes.c_void_p


class KernelTimer:
    """Optional CUDA-event stopwatch around every csb200 call, with the call's ALGORITHMIC work
    (bytes that must cross HBM once, FLOPs of the contraction) — bench.py's roofline source.
    Events are recorded on the stream the kernels are launched on."""

    def __init__(self):
        self.spans = []  # (family, shape tag, start, end, bytes, flops)

    def summary(self, by_shape: bool = False):
        """Totals per kernel family, or per (family, shape) when ``by_shape`` — the latter is one
        physical launch configuration, which is what a roofline fraction should be quoted for."""
        torch.cuda.synchronize()
        out = {}
        for fam, tag, a, b, nbytes, flops in self.spans:
            key = f"{fam}|{tag}" if by_shape and tag else fam
            r = out.setdefault(key, {"calls": 0, "ms": 0.0, "bytes": 0, "flops": 0})
            r["calls"] += 1
            r["ms"] += a.elapsed_time(b)
            r["bytes"] += nbytes
            r["flops"] += flops
        return out


_timer: Optional[KernelTimer] = None


def set_kernel_timer(t: Optional[KernelTimer]):
    global _timer
    _timer = t


class _span:
    __slots__ = ("fam", "nbytes", "flops", "start", "tag")

    def __init__(self, fam, nbytes, flops=0, tag=""):
        self.fam, self.nbytes, self.flops, self.tag = fam, nbytes, flops, tag

    def __enter__(self):
        if _timer is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record()

    def __exit__(self, *exc):
        if _timer is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            _timer.spans.append((self.fam, self.tag, self.start, end, self.nbytes, self.flops))
        return False


def _ptr(t: Optional[torch.Tensor], elem_offset: int = 0):
    if t is None:
        return None
    return _vp(t.data_ptr() + elem_offset * t.element_size())


# ------------------------------------------------------------------------------------------------
# SimAM
# ------------------------------------------------------------------------------------------------
# ---- deferred final sums (include/csb200.h "Deferred final sums", csrc/sum_rows.cu) ---------------------------
# The LayerNorm parameter gradients, the column-sum bias gradients and the fc1 bias gradient of the fused Mlp all
# end with a 4-5 us "sum the per-CTA partial rows" launch: 106 of them on the critical path of one 512^2 backward
# pass, for vectors nobody reads before the optimizer step.  Inside ``with deferred_sums(device):`` those call
# sites stop after their main kernel, record the sum, and ONE launch per 120 records performs them all when the
# block exits — the gradient tensors autograd hands to the parameters are filled at that point.
# Only valid when (TrainStep guarantees all three): every ``.grad`` is None when backward starts (a second
# contribution would be ADDED to a vector that is not filled yet), nothing reads the gradients before the block
# exits (no post-accumulate hooks launching all-reduces), and backward runs on the current stream.
_deferred: Optional[list] = None  # off: None; on: tensors that must stay allocated until the flush
_deferred_outs: list = []         # the sum vectors among them (checked at the flush: were they adopted?)


_zero_arena = None  # [buffer or None, offset, demand]: see deferred_sums(zero_arena_numel=...)


class deferred_sums:
    """``zero_arena_numel``: also serve the accumulate-into outputs of the pass (csb200_linear_wgrad's grad_w /
    grad_b, two memset nodes in front of each of the 35 calls of a 512^2 backward) from ONE fp32 buffer zeroed
    once on entry; ``arena_demand`` afterwards = the elements the pass asked for (size the next arena with it)."""

    def __init__(self, device, zero_arena_numel: int = 0):
        self.device = torch.device(device)
        self.zero_arena_numel = int(zero_arena_numel)
        self.arena_demand = 0

    def __enter__(self):
        global _deferred
        if _deferred is not None:
            raise RuntimeError("deferred_sums blocks do not nest")
        capi.lib().csb200_sum_rows_discard()
        _deferred = []
        del _deferred_outs[:]
        global _zero_arena
        buf = None
        if self.zero_arena_numel > 0:
            with torch.cuda.device(self.device):
                buf = torch.zeros(self.zero_arena_numel, dtype=torch.float32, device=self.device)
        _zero_arena = [buf, 0, 0]
        return self

    def __exit__(self, etype, evalue, tb):
        global _deferred
        global _zero_arena
        self.arena_demand, _zero_arena = _zero_arena[2], None
        keep, _deferred = _deferred, None
        outs = list(_deferred_outs)
        del _deferred_outs[:]
        lib = capi.lib()
        if etype is not None:
            lib.csb200_sum_rows_discard()
            return False
        # Every recorded vector went to autograd as a view; if that view is gone, AccumulateGrad did not adopt it
        # but cloned it (before it was filled) or the gradient was dropped: refuse loudly rather than train on
        # garbage.  (storage use count: this list's tensor + the handle below = 2, an adopted view makes 3)
        for out in outs:
            if torch._C._storage_Use_Count(out.untyped_storage()._cdata) <= 2:
                lib.csb200_sum_rows_discard()
                raise RuntimeError("deferred_sums: a deferred gradient vector was not adopted by autograd (gradient "
                                   "accumulation into an existing .grad, or a copy): use the immediate sums")
        with torch.cuda.device(self.device):
            capi.check(lib.csb200_sum_rows_flush(_vp(torch.cuda.current_stream(self.device).cuda_stream)),
                       "csb200_sum_rows_flush")
        keep.clear()  # freed in stream order: after the flush kernel
        return False


def _zeroed(numel: int, device) -> Optional[torch.Tensor]:
    """``numel`` zero-filled fp32 elements out of the arena of the enclosing deferred_sums block (256-byte
    aligned), or None (no block, no arena yet, or the arena is exhausted: the caller zeroes its own buffer)."""
    if _zero_arena is None:
        return None
    need = (numel + 63) // 64 * 64
    buf, off, _ = _zero_arena
    _zero_arena[2] += need
    if buf is None or buf.device != torch.device(device) or off + need > buf.numel():
        return None
    _zero_arena[1] = off + need
    return buf[off:off + numel]


def _defer_sum(partials: int, partial_rows: int, cols: int, out: torch.Tensor, *keep: torch.Tensor) -> torch.Tensor:
    """Record out[c] = sum of the partial rows; returns a VIEW of ``out`` for the caller to hand to autograd.
    ``out`` itself stays referenced here until the flush (its storage must not be recycled if autograd drops the
    gradient), and AccumulateGrad only adopts a gradient tensor nobody else references — a tensor that is also
    held here would be CLONED at accumulation time, before the flush has filled it; a view is adopted as it is."""
    capi.check(capi.lib().csb200_sum_rows_deferred(partials, partial_rows, cols, cols, _ptr(out)),
               "csb200_sum_rows_deferred")
    _deferred.extend(keep)
    _deferred.append(out)
    _deferred_outs.append(out)
    return out.view(-1)


def _layernorm_bwd_call(lib, s, gy, gres, w, stats, gx, want_rb: bool, rows: int, C: int, defer_ok: bool = True):
    """LayerNorm backward (plain when gres is None and not want_rb); returns (grad_gamma, grad_beta,
    grad_res_bias or None) — immediate, or views of one vector filled at the deferred flush."""
    nws = lib.csb200_layernorm_bwd_workspace_bytes(rows, C)
    wsp = torch.empty(nws, dtype=torch.uint8, device=s.device)
    st = _vp(capi.stream_of(s))
    xc, gc = capi.dtype_code(s), capi.dtype_code(gy)
    if _deferred is not None and defer_ok and rows > 0:
        K = 3 if want_rb else 2
        out = torch.empty(K * C, dtype=torch.float32, device=s.device)
        pp, pr = ctypes.c_void_p(), ctypes.c_int32()
        capi.check(lib.csb200_layernorm_bwd_partials(_ptr(s), _ptr(gy), _ptr(gres), _ptr(w), _ptr(stats), _ptr(gx),
                                                     int(want_rb), _ptr(wsp), nws, rows, C, xc, gc, ctypes.byref(pp),
                                                     ctypes.byref(pr), st), "csb200_layernorm_bwd_partials")
        out = _defer_sum(pp.value, pr.value, K * C, out, wsp)
        return out[:C], out[C:2 * C], (out[2 * C:] if want_rb else None)
    gw, gb = torch.empty_like(w), torch.empty_like(w)
    if want_rb:
        grb = torch.empty_like(w)
        capi.check(lib.csb200_add_layernorm_bwd_rb(_ptr(s), _ptr(gy), _ptr(gres), _ptr(w), _ptr(stats), _ptr(gx),
                                                   _ptr(gw), _ptr(gb), _ptr(grb), _ptr(wsp), nws, rows, C, xc, gc, st),
                   "csb200_add_layernorm_bwd_rb")
        return gw, gb, grb
    if gres is None:
        capi.check(lib.csb200_layernorm_bwd(_ptr(s), _ptr(gy), _ptr(w), _ptr(stats), _ptr(gx), _ptr(gw), _ptr(gb),
                                            _ptr(wsp), nws, rows, C, xc, gc, st), "csb200_layernorm_bwd")
    else:
        capi.check(lib.csb200_add_layernorm_bwd(_ptr(s), _ptr(gy), _ptr(gres), _ptr(w), _ptr(stats), _ptr(gx),
                                                _ptr(gw), _ptr(gb), _ptr(wsp), nws, rows, C, xc, gc, st),
                   "csb200_add_layernorm_bwd")
    return gw, gb, None


def _simam_dims(x: torch.Tensor, layout: str) -> Tuple[int, int, int, int]:
    if layout == "NCHW":
        if x.dim() != 4:
            raise ValueError("SimAM NCHW expects (B, C, H, W)")
        B, C, H, W = x.shape
        return B, C, H * W, capi.NCHW
    if layout == "NLC":
        if x.dim() != 3:
            raise ValueError("SimAM NLC expects (B, L, C)")
        B, L, C = x.shape
        return B, C, L, capi.NLC
    raise ValueError(f"unknown SimAM layout {layout!r}")


def _simam_plan(x: torch.Tensor, layout: str):
    """Pick the physical kernel layout.  A channels-last (B, C, H, W) tensor IS a (B, H*W, C) token
    matrix in memory, so it takes the NLC kernel in place instead of a transposing copy."""
    if layout == "NCHW" and x.dim() == 4 and not x.is_contiguous() \
            and x.is_contiguous(memory_format=torch.channels_last):
        B, C, H, W = x.shape
        return x, (B, C, H * W, capi.NLC), torch.channels_last
    x = x.contiguous()
    return x, _simam_dims(x, layout), torch.contiguous_format


_simam_ws = {}  # device index -> zero-initialised workspace of the grid-resident kernels (csb200_simam_*_ws)
# The grid-resident SimAM kernels (csrc/simam_grid.cuh) are correct but, as measured on B200, slower than the
# cluster kernels on every config-3 shape: off unless CSB200_SIMAM_GRID=1.
SIMAM_GRID_KERNELS = os.environ.get("CSB200_SIMAM_GRID", "0") == "1"


def _simam_workspace(x: torch.Tensor, B, C, S, lay):
    """(pointer, bytes) of this device's SimAM workspace, grown on demand; (None, 0) when the grid-resident
    kernels do not apply to the shape.  The library keeps it clean between calls (include/csb200.h), so it is
    zeroed only when (re)allocated; calls on one device are expected to be stream-ordered."""
    if not SIMAM_GRID_KERNELS:
        return None, 0
    need = capi.lib().csb200_simam_workspace_bytes(B, C, S, lay, capi.dtype_code(x))
    if need == 0:
        return None, 0
    idx = x.device.index if x.device.index is not None else torch.cuda.current_device()
    ws = _simam_ws.get(idx)
    if ws is None or ws.numel() < need:
        ws = _simam_ws[idx] = torch.zeros(need, dtype=torch.uint8, device=x.device)
    return ws.data_ptr(), ws.numel()


class _SimAMFn(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, e_lambda, layout):
        capi.require_cuda(x)
        x, (B, C, S, lay), fmt = _simam_plan(x, layout)
        y = torch.empty_like(x, memory_format=fmt)
        stats = torch.empty((B * C, 2), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device), _span("simam_fwd", 2 * x.numel() * x.element_size(), 0, f"B{B}xC{C}xS{S}"):
            ws, ws_bytes = _simam_workspace(x, B, C, S, lay)
            capi.check(capi.lib().csb200_simam_fwd_ws(_ptr(x), _ptr(y), _ptr(stats), B, C, S, lay,
                                                      capi.dtype_code(x), float(e_lambda), ws, ws_bytes,
                                                      _vp(capi.stream_of(x))),
                       "csb200_simam_fwd_ws")
        ctx.save_for_backward(x, stats)
        ctx.cfg = (B, C, S, lay, float(e_lambda), fmt)
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gy):
        x, stats = ctx.saved_tensors
        B, C, S, lay, e_lambda, fmt = ctx.cfg
        gy = gy.contiguous(memory_format=fmt)
        if gy.dtype != x.dtype:
            gy = gy.to(x.dtype)
        gx = torch.empty_like(x, memory_format=fmt)
        with torch.cuda.device(x.device), _span("simam_bwd", 3 * x.numel() * x.element_size(), 0, f"B{B}xC{C}xS{S}"):
            ws, ws_bytes = _simam_workspace(x, B, C, S, lay)
            capi.check(capi.lib().csb200_simam_bwd_ws(_ptr(x), _ptr(gy), _ptr(stats), _ptr(gx), B, C, S, lay,
                                                      capi.dtype_code(x), e_lambda, ws, ws_bytes,
                                                      _vp(capi.stream_of(x))),
                       "csb200_simam_bwd_ws")
        return gx, None, None


def simam(x: torch.Tensor, e_lambda: float = 1e-4, layout: str = "NCHW") -> torch.Tensor:
    """Fused SimAM: ``x * sigmoid((x-mean)^2 / (4 (var_n + e_lambda)) + 0.5)`` per (image, channel).

    layout "NCHW": x (B, C, H, W); "NLC": x (B, L, C) tokens.  float32 or bfloat16, CUDA only.
    """
    return _SimAMFn.apply(x, e_lambda, layout)


# ------------------------------------------------------------------------------------------------
# LayerNorm (token-major, last dimension)
# ------------------------------------------------------------------------------------------------
class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, weight, bias, eps, out_dtype):
        capi.require_cuda(x)
        x = x.contiguous()
        C = x.shape[-1]
        rows = x.numel() // C
        w, b = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
        stats = torch.empty((rows, 2), dtype=torch.float32, device=x.device)
        nbytes = x.numel() * (x.element_size() + y.element_size())
        with torch.cuda.device(x.device), _span("layernorm_fwd", nbytes, 0, f"{rows}x{C}"):
            capi.check(capi.lib().csb200_layernorm_fwd(_ptr(x), _ptr(w), _ptr(b), _ptr(y), _ptr(stats), rows, C,
                                                       capi.dtype_code(x), capi.dtype_code(y), float(eps),
                                                       _vp(capi.stream_of(x))), "csb200_layernorm_fwd")
        ctx.save_for_backward(x, w, stats)
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gy):
        x, w, stats = ctx.saved_tensors
        C = x.shape[-1]
        rows = x.numel() // C
        gy = gy.contiguous()
        gx = torch.empty_like(x)
        nbytes = x.numel() * (2 * x.element_size() + gy.element_size())
        with torch.cuda.device(x.device), _span("layernorm_bwd", nbytes, 0, f"{rows}x{C}"):
            gw, gb, _ = _layernorm_bwd_call(capi.lib(), x, gy, None, w, stats, gx, False, rows, C,
                                            defer_ok=ctx.needs_input_grad[1] and ctx.needs_input_grad[2])
        return gx, gw, gb, None, None


class _AddLayerNormFn(torch.autograd.Function):
    """(s, y) = (x + r, LayerNorm(x + r)) in one pass; backward folds the gradient arriving on s into
    the LayerNorm backward pass (csb200_add_layernorm_fwd / _bwd).

    ``r_bias``: when r is the output of a Linear whose bias gradient was deferred (``linear(...,
    defer_bias_grad=True)``), its bias parameter.  The forward value is not touched (the GEMM already
    added the bias); backward returns its gradient — the column sums of grad_r, accumulated by the same
    LayerNorm backward pass — so no column-sum pass over that tensor runs."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, r, weight, bias, eps, out_dtype, r_bias=None):
        capi.require_cuda(x, r)
        x, r = x.contiguous(), r.contiguous()
        C = x.shape[-1]
        rows = x.numel() // C
        w, b = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        s = torch.empty_like(x)
        y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
        stats = torch.empty((rows, 2), dtype=torch.float32, device=x.device)
        nbytes = x.numel() * (3 * x.element_size() + y.element_size())
        with torch.cuda.device(x.device), _span("layernorm_fwd", nbytes, 0, f"{rows}x{C}"):
            capi.check(capi.lib().csb200_add_layernorm_fwd(
                _ptr(x), _ptr(r), _ptr(s), _ptr(w), _ptr(b), _ptr(y), _ptr(stats), rows, C, capi.dtype_code(x),
                capi.dtype_code(y), float(eps), _vp(capi.stream_of(x))), "csb200_add_layernorm_fwd")
        ctx.save_for_backward(s, w, stats)
        ctx.rb_dtype = None if r_bias is None else r_bias.dtype
        return s, y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gs, gy):
        s, w, stats = ctx.saved_tensors
        C = s.shape[-1]
        rows = s.numel() // C
        want_rb = ctx.rb_dtype is not None and ctx.needs_input_grad[6]
        if gy is None:  # the normalised branch was unused: only the residual stream carries gradient
            grb = None
            if want_rb and gs is not None:
                grb = _bias_grad(gs.reshape(rows, C).contiguous(), C, ctx.rb_dtype)
            return gs, gs, None, None, None, None, grb
        gy = gy.contiguous()
        gres = None if gs is None else gs.to(s.dtype).contiguous()
        gx = torch.empty_like(s)
        nbytes = s.numel() * ((2 + (gres is not None)) * s.element_size() + gy.element_size())
        with torch.cuda.device(s.device), _span("layernorm_bwd", nbytes, 0, f"{rows}x{C}"):
            # (deferred only when all the vectors go to autograd as they are: wanted, and no cast of grad_res_bias)
            ok = ctx.needs_input_grad[2] and ctx.needs_input_grad[3] and (not want_rb or ctx.rb_dtype == torch.float32)
            gw, gb, grb = _layernorm_bwd_call(capi.lib(), s, gy, gres, w, stats, gx, want_rb, rows, C, defer_ok=ok)
        return gx, gx, gw, gb, None, None, None if grb is None else grb.to(ctx.rb_dtype)


def add_layer_norm(x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor,
                   eps: float = 1e-5, out_dtype: Optional[torch.dtype] = None,
                   residual_bias: Optional[torch.Tensor] = None):
    """Returns (x + residual, LayerNorm(x + residual)); both tensors have the shape of x.
    ``residual_bias``: see ``_AddLayerNormFn`` (gradient routing of a deferred Linear bias)."""
    return _AddLayerNormFn.apply(x, residual, weight, bias, eps, out_dtype or x.dtype, residual_bias)


class _RouteBiasGradFn(torch.autograd.Function):
    """Identity on `delta` that returns the column sums of its gradient to `bias`: the landing pad of a
    deferred Linear bias gradient whose delta did not end in a fused add + LayerNorm."""

    @staticmethod
    def forward(ctx, delta, bias):
        ctx.b_dtype = bias.dtype
        return delta.view_as(delta)

    @staticmethod
    def backward(ctx, g):
        n = g.shape[-1]
        g2 = g.reshape(-1, n)
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        if not ctx.needs_input_grad[1]:
            return g, None
        return g, _bias_grad(g2, n, ctx.b_dtype)


def route_bias_grad(delta: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    return delta if bias is None else _RouteBiasGradFn.apply(delta, bias)


def layer_norm_supported(x: torch.Tensor) -> bool:
    return x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) and \
        bool(capi.lib().csb200_layernorm_supported(x.shape[-1], capi.dtype_code(x)))


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5,
               out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """LayerNorm over the last dimension in one HBM pass; ``out_dtype`` (default: x.dtype) lets the
    fp32 residual stream be normalised straight into the bf16 operand of the following GEMM."""
    return _LayerNormFn.apply(x, weight, bias, eps, out_dtype or x.dtype)


# ------------------------------------------------------------------------------------------------
# Linear with a one-pass bias gradient
# ------------------------------------------------------------------------------------------------
def column_sum(x2d: torch.Tensor, deferrable: bool = False) -> torch.Tensor:
    """fp32 column sums of a contiguous (rows, cols) CUDA matrix in one flat HBM pass.  ``deferrable``: inside a
    ``deferred_sums`` block the result may be filled at the end of the block (the caller must not read it)."""
    rows, cols = x2d.shape
    lib = capi.lib()
    out = torch.empty(cols, dtype=torch.float32, device=x2d.device)
    nws = lib.csb200_colsum_workspace_bytes(cols)
    wsp = torch.empty(nws, dtype=torch.uint8, device=x2d.device)
    with torch.cuda.device(x2d.device), _span("colsum", x2d.numel() * x2d.element_size(), 0, f"{rows}x{cols}"):
        if deferrable and _deferred is not None and rows > 0:
            pp, pr = ctypes.c_void_p(), ctypes.c_int32()
            capi.check(lib.csb200_colsum_partials(_ptr(x2d), _ptr(wsp), nws, rows, cols, capi.dtype_code(x2d),
                                                  ctypes.byref(pp), ctypes.byref(pr), _vp(capi.stream_of(x2d))),
                       "csb200_colsum_partials")
            out = _defer_sum(pp.value, pr.value, cols, out, wsp)
        else:
            capi.check(lib.csb200_colsum(_ptr(x2d), _ptr(out), _ptr(wsp), nws, rows, cols, capi.dtype_code(x2d),
                                         _vp(capi.stream_of(x2d))), "csb200_colsum")
    return out


# Low-precision shadows of the fp32 master parameters.  Under autocast every Linear used to cast its
# weight and bias on every call (250 cast launches per 512^2 step); TrainStep refreshes all shadows
# with ONE multi-tensor copy after the optimizer step instead.  A shadow is only trusted while the
# parameter's version counter still has the value it had when the shadow was written.
def shadow_params(params, dtype=torch.bfloat16):
    """(Re)build the shadows of `params`; returns (masters, shadows) for later ``refresh_shadows``."""
    masters = [p for p in params if p.is_cuda and p.dtype == torch.float32]
    shadows = []
    with torch.no_grad():
        for p in masters:
            old = getattr(p, "_csb_shadow", None)
            if old is not None and old[0].dtype == dtype and old[0].shape == p.shape and old[0].device == p.device:
                sh = old[0]  # keep the tensor: an optimizer kernel or a captured graph may write to it
                sh.copy_(p)
            else:
                sh = p.detach().to(dtype)
            p._csb_shadow = (sh, p._version)
            shadows.append(sh)
    return masters, shadows


def refresh_shadows(masters, shadows):
    """shadow <- master for every pair, as one multi-tensor launch; call after optimizer.step()."""
    with torch.no_grad():
        torch._foreach_copy_(shadows, masters)
    for p, sh in zip(masters, shadows):
        p._csb_shadow = (sh, p._version)


def restamp_shadows(masters, shadows):
    """Mark every shadow as current WITHOUT copying (the optimizer kernel has just written them)."""
    for p, sh in zip(masters, shadows):
        if p._csb_shadow[1] != p._version:
            p._csb_shadow = (sh, p._version)


def cast_param(p: Optional[torch.Tensor], dtype: torch.dtype) -> Optional[torch.Tensor]:
    if p is None or p.dtype == dtype:
        return p
    sh = getattr(p, "_csb_shadow", None)
    if sh is not None and sh[1] == p._version and sh[0].dtype == dtype:
        return sh[0]
    return p.to(dtype)


# Which Linear layers run on the csb200 tcgen05 GEMM (csb200_linear_fwd, K = C in {64, 128, 256}, bf16):
# "gelu": Mlp.fc1 + GELU with the activation in the epilogue; "plain": qkv / proj (bias-only epilogue).
# "wgrad": the weight (+ bias) gradients of those layers in one streaming tcgen05 pass (csb200_linear_wgrad).
TC_LINEAR = {"gelu": True, "plain": True, "wgrad": True}


def set_tc_linear(gelu: Optional[bool] = None, plain: Optional[bool] = None, wgrad: Optional[bool] = None):
    for key, val in (("gelu", gelu), ("plain", plain), ("wgrad", wgrad)):
        if val is not None:
            TC_LINEAR[key] = val if val == "always" else bool(val)  # wgrad="always": ignore the token-count policy


def _tc_linear_ok(x2: torch.Tensor, wc: torch.Tensor, bias: Optional[torch.Tensor]) -> bool:
    if x2.dtype != torch.bfloat16 or wc.dtype != torch.bfloat16 or not wc.is_contiguous() or x2.stride(1) != 1:
        return False
    if x2.data_ptr() % 16 or wc.data_ptr() % 16 or (x2.stride(0) * 2) % 16:
        return False
    return bool(capi.lib().csb200_linear_supported(x2.shape[0], wc.shape[0], wc.shape[1], capi.BF16))


def _tc_linear(x2: torch.Tensor, wc: torch.Tensor, bias: Optional[torch.Tensor], epilogue: int):
    """(y, pre_act) = csb200_linear_fwd on a (M, K) token matrix; bias is read in fp32."""
    M, K = x2.shape
    N = wc.shape[0]
    y = torch.empty((M, N), dtype=torch.bfloat16, device=x2.device)
    saves = epilogue in (capi.EPI_GELU_SAVE, capi.EPI_GELU_SAVE_DERIV)  # second output: h, or GELU'(h)
    h = torch.empty((M, N), dtype=torch.bfloat16, device=x2.device) if saves else None
    b32 = None if bias is None else bias.detach().float().contiguous()
    nbytes = 2 * (M * K + N * K + M * N * (2 if h is not None else 1))
    with torch.cuda.device(x2.device), _span("linear_tc", nbytes, 2 * M * N * K, f"M{M}xN{N}xK{K}e{epilogue}"):
        capi.check(capi.lib().csb200_linear_fwd(_ptr(x2), _ptr(wc), _ptr(b32), _ptr(y), _ptr(h), M, N, K, x2.stride(0),
                                                capi.BF16, epilogue, _vp(capi.stream_of(x2))), "csb200_linear_fwd")
    return y, h


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b.  GEMMs stay on cuBLAS (torch.mm) unless TC_LINEAR["plain"]; the bias gradient, which ATen computes with a
    generic strided reduction at ~1/9 of the HBM roofline, is one csb200_colsum pass."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, weight, bias, compute_dtype):
        xc = x if x.dtype == compute_dtype else x.to(compute_dtype)
        wc, bc = cast_param(weight, compute_dtype), cast_param(bias, compute_dtype)
        ctx.save_for_backward(xc, wc)
        ctx.meta = (x.dtype, weight.dtype, None if bias is None else bias.dtype)
        if TC_LINEAR["plain"] and compute_dtype == torch.bfloat16:
            x2 = xc.reshape(-1, xc.shape[-1])
            if _tc_linear_ok(x2, wc, bias):
                y, _ = _tc_linear(x2, wc, bias, capi.EPI_BIAS)
                return y.view(*xc.shape[:-1], wc.shape[0])
        return torch.nn.functional.linear(xc, wc, bc)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gy):
        xc, wc = ctx.saved_tensors
        x_dtype, w_dtype, b_dtype = ctx.meta
        n = wc.shape[0]
        g2 = gy.reshape(-1, n)
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        x2 = xc.reshape(-1, wc.shape[1])
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = torch.mm(g2, wc).reshape(xc.shape).to(x_dtype)
        want_b = b_dtype is not None and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1] and _tc_wgrad_ok(g2, x2, w_dtype, want_b):
            gw, gb = _tc_wgrad(g2, x2, want_b)  # the bias gradient rides in the same streaming pass
            gb = None if gb is None else gb.to(b_dtype)
        else:
            if ctx.needs_input_grad[1]:
                gw = _wgrad(g2, x2, w_dtype)
            if want_b:
                gb = _bias_grad(g2, n, b_dtype)
        return gx, gw, gb, None


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
           defer_bias_grad: bool = False) -> torch.Tensor:
    """Drop-in for ``F.linear`` on CUDA float32 / bfloat16 tensors (autocast aware).
    ``defer_bias_grad``: the bias enters detached (no column-sum pass in backward); the CALLER must hand
    the output and the bias to ``add_layer_norm(..., residual_bias=bias)`` or ``route_bias_grad``."""
    if defer_bias_grad and bias is not None:
        bias = bias.detach()
    if not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16):
        return torch.nn.functional.linear(x, weight, bias)
    dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
    if dt not in (torch.float32, torch.bfloat16):
        return torch.nn.functional.linear(x, weight, bias)
    return _LinearFn.apply(x, weight, bias, dt)


# Measured (profiles/r2_wgrad_bench.jsonl): csb200_linear_wgrad runs at 0.87-0.89 of the HBM roofline on the
# stage-1 token counts (= cuBLAS split-K) and always beats cuBLAS + a column-sum pass when the bias gradient is
# wanted (1.2-1.9x); on the short token counts of stages 2-4 cuBLAS alone is 10-25 % faster (fixed costs: two
# memsets, 24-deep vector reductions on an output of a few hundred KB), so those stay on cuBLAS.
WGRAD_MIN_TOKENS = 262144


def _tc_wgrad_ok(g2: torch.Tensor, x2: torch.Tensor, w_dtype: torch.dtype, want_bias: bool = True) -> bool:
    if not TC_LINEAR["wgrad"] or w_dtype != torch.float32 or g2.dtype != torch.bfloat16 or x2.dtype != torch.bfloat16:
        return False
    if not want_bias and g2.shape[0] < WGRAD_MIN_TOKENS and TC_LINEAR["wgrad"] != "always":
        return False
    if g2.stride(1) != 1 or x2.stride(1) != 1 or g2.data_ptr() % 16 or x2.data_ptr() % 16 \
            or (g2.stride(0) * 2) % 16 or (x2.stride(0) * 2) % 16:
        return False
    return bool(capi.lib().csb200_linear_wgrad_supported(g2.shape[0], g2.shape[1], x2.shape[1], capi.BF16))


def _tc_wgrad(g2: torch.Tensor, x2: torch.Tensor, want_bias: bool):
    """(grad_W fp32 [N][K], grad_b fp32 [N] or None) = csb200_linear_wgrad(g2 [M][N], x2 [M][K])."""
    M, N = g2.shape
    K = x2.shape[1]
    lib = capi.lib()
    arena = _zeroed(N * K + (N if want_bias else 0), g2.device)  # both outputs from one zero-filled slice, or None
    if arena is not None:
        gw, gb, fn, name = arena[:N * K].view(N, K), (arena[N * K:] if want_bias else None), lib.csb200_linear_wgrad_acc, \
            "csb200_linear_wgrad_acc"
    else:
        gw = torch.empty((N, K), dtype=torch.float32, device=g2.device)
        gb = torch.empty(N, dtype=torch.float32, device=g2.device) if want_bias else None
        fn, name = lib.csb200_linear_wgrad, "csb200_linear_wgrad"
    with torch.cuda.device(g2.device), _span("wgrad_tc", 2 * M * (N + K) + 4 * N * K, 2 * M * N * K, f"M{M}xN{N}xK{K}"):
        capi.check(fn(_ptr(g2), _ptr(x2), _ptr(gw), _ptr(gb), M, N, K, g2.stride(0), x2.stride(0), capi.BF16,
                      _vp(capi.stream_of(g2))), name)
    return gw, gb


def _wgrad(g2: torch.Tensor, x2: torch.Tensor, w_dtype: torch.dtype) -> torch.Tensor:
    """grad_W = g^T x with the output written in the master-weight dtype by the GEMM itself (fp32
    accumulators stored unrounded; no bf16 -> fp32 cast kernel per layer)."""
    if _tc_wgrad_ok(g2, x2, w_dtype, want_bias=False):
        return _tc_wgrad(g2, x2, False)[0]
    if g2.is_cuda and g2.dtype == torch.bfloat16 and w_dtype == torch.float32:
        return torch.mm(g2.t(), x2, out_dtype=torch.float32)
    return torch.mm(g2.t(), x2).to(w_dtype)


def _bias_grad(g2: torch.Tensor, n: int, final_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Column sums of a (rows, n) gradient, accumulated in fp32 and returned as ``final_dtype`` (None: fp32):
    csb200_colsum when the width tiles, else ATen.  A caller that passes ``final_dtype`` promises to hand the
    result to autograd unread, so the fp32 case may be deferred (``deferred_sums``)."""
    if final_dtype is not None:
        if final_dtype == torch.float32 and g2.is_contiguous() and g2.data_ptr() % 16 == 0 \
                and capi.lib().csb200_colsum_supported(n, capi.dtype_code(g2)):
            return column_sum(g2, deferrable=True)
        return _bias_grad(g2, n).to(final_dtype)
    if g2.is_contiguous() and g2.data_ptr() % 16 == 0:
        code, rows = capi.dtype_code(g2), g2.shape[0]
        if capi.lib().csb200_colsum_supported(n, code):
            return column_sum(g2)
        # widths that are not a multiple of the 16-byte vector (1 logit channel, 36 CARAFE taps): k rows
        # side by side form one row of k*n columns; fold the k column groups afterwards
        for k in (2, 4, 8):
            if rows % k == 0 and capi.lib().csb200_colsum_supported(n * k, code):
                return column_sum(g2.view(rows // k, n * k)).view(k, n).sum(0)
    return g2.sum(0, dtype=torch.float32)


class _LinearGeluFn(torch.autograd.Function):
    """a = GELU(x W^T + b) — fc1 + act of the Mlp (C:188-196).  The GEMMs stay on cuBLAS; the exact-erf
    GELU is one csb200 pass, and its backward pass also emits the bias gradient (csb200_gelu_bwd), so
    neither ATen's GeluBackward nor a separate column-sum pass over the 4C-wide tensor runs."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, weight, bias, compute_dtype):
        xc = x if x.dtype == compute_dtype else x.to(compute_dtype)
        wc = cast_param(weight, compute_dtype)
        n = wc.shape[0]
        ctx.meta = (x.dtype, weight.dtype, bias.dtype)
        x2 = xc.reshape(-1, xc.shape[-1])
        if TC_LINEAR["gelu"] and compute_dtype == torch.bfloat16 and _tc_linear_ok(x2, wc, bias):
            # one tcgen05 GEMM whose epilogue adds the bias, applies GELU and stores both a and h (C:188-190)
            a, h = _tc_linear(x2, wc, bias, capi.EPI_GELU_SAVE)
            ctx.save_for_backward(xc, wc, h.view(*xc.shape[:-1], n))
            return a.view(*xc.shape[:-1], n)
        h = torch.nn.functional.linear(xc, wc, cast_param(bias, compute_dtype))
        h2 = h.reshape(-1, n)
        a = torch.empty_like(h)
        lib = capi.lib()
        with torch.cuda.device(h.device), _span("gelu_fwd", 2 * h.numel() * h.element_size()):
            capi.check(lib.csb200_gelu_fwd(_ptr(h2), _ptr(a), h2.shape[0], n, capi.dtype_code(h),
                                           _vp(capi.stream_of(h))), "csb200_gelu_fwd")
        ctx.save_for_backward(xc, wc, h)
        return a

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, ga):
        xc, wc, h = ctx.saved_tensors
        x_dtype, w_dtype, b_dtype = ctx.meta
        n = wc.shape[0]
        g2 = ga.reshape(-1, n)
        if not g2.is_contiguous() or g2.dtype != h.dtype:
            g2 = g2.to(h.dtype).contiguous()
        rows = g2.shape[0]
        dh = torch.empty_like(g2)
        gb = torch.empty(n, dtype=torch.float32, device=h.device)
        lib = capi.lib()
        nws = lib.csb200_gelu_bwd_workspace_bytes(n)
        wsp = torch.empty(nws, dtype=torch.uint8, device=h.device)
        with torch.cuda.device(h.device), _span("gelu_bwd", 3 * h.numel() * h.element_size()):
            capi.check(lib.csb200_gelu_bwd(_ptr(g2), _ptr(h), _ptr(dh), _ptr(gb), _ptr(wsp), nws, rows, n,
                                           capi.dtype_code(h), _vp(capi.stream_of(h))), "csb200_gelu_bwd")
        x2 = xc.reshape(-1, wc.shape[1])
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = torch.mm(dh, wc).reshape(xc.shape).to(x_dtype)
        if ctx.needs_input_grad[1]:
            gw = _wgrad(dh, x2, w_dtype)
        return gx, gw, (gb.to(b_dtype) if ctx.needs_input_grad[2] else None), None


def _tc_dgelu(g2: torch.Tensor, w2c: torch.Tensor, h2: torch.Tensor, deriv: bool = False, defer_bias: bool = False):
    """(grad_h, grad_bias) = csb200_linear_dgelu_bwd: grad_h = (g W2) * GELU'(h) in one tcgen05 GEMM.
    ``deriv``: ``h2`` already holds GELU'(h) (forward epilogue EPI_GELU_SAVE_DERIV) -> csb200_linear_dact_bwd,
    whose epilogue is one multiplication per element."""
    M, K = g2.shape
    N = w2c.shape[1]
    lib = capi.lib()
    dh = torch.empty((M, N), dtype=torch.bfloat16, device=g2.device)
    gb = torch.empty(N, dtype=torch.float32, device=g2.device)
    nws = lib.csb200_linear_dgelu_workspace_bytes(N)
    wsp = torch.empty(nws, dtype=torch.uint8, device=g2.device)
    nbytes = 2 * (M * K + N * K + 2 * M * N)
    fn = lib.csb200_linear_dact_bwd if deriv else lib.csb200_linear_dgelu_bwd
    with torch.cuda.device(g2.device), _span("linear_tc", nbytes, 2 * M * N * K,
                                             f"M{M}xN{N}xK{K}{'dact' if deriv else 'dgelu'}"):
        if defer_bias and _deferred is not None and M > 0:  # gb is filled when the deferred_sums block exits
            pp, pr = ctypes.c_void_p(), ctypes.c_int32()
            capi.check(lib.csb200_linear_dact_bwd_partials(_ptr(g2), _ptr(w2c), _ptr(h2), _ptr(dh), _ptr(wsp), nws, M, N,
                                                           K, g2.stride(0), capi.BF16, int(deriv), ctypes.byref(pp),
                                                           ctypes.byref(pr), _vp(capi.stream_of(g2))),
                       "csb200_linear_dact_bwd_partials")
            gb = _defer_sum(pp.value, pr.value, N, gb, wsp)
        else:
            capi.check(fn(_ptr(g2), _ptr(w2c), _ptr(h2), _ptr(dh), _ptr(gb), _ptr(wsp), nws, M, N, K,
                          g2.stride(0), capi.BF16, _vp(capi.stream_of(g2))),
                       "csb200_linear_dact_bwd" if deriv else "csb200_linear_dgelu_bwd")
    return dh, gb


# fused Mlp: save GELU'(h) in forward (True) or h (False: the backward epilogue re-evaluates erf and exp)
MLP_SAVE_DERIV = True


class _MlpFn(torch.autograd.Function):
    """y = fc2(GELU(fc1(x))) — Mlp.forward, C:188-196 (dropout 0) — with the activation inside the GEMMs that
    surround it: forward = tcgen05 fc1 whose epilogue stores GELU(h) and GELU'(h), then fc2 on cuBLAS; backward =
    ONE tcgen05 GEMM for grad_h = (grad_y W2) * GELU'(h) that also emits the fc1 bias gradient, then the
    weight gradients and the fc1 input gradient on cuBLAS.  Neither flat GELU pass runs."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, w1, b1, w2, b2):
        xc = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
        w1c, w2c = cast_param(w1, torch.bfloat16), cast_param(w2, torch.bfloat16)
        x2 = xc.reshape(-1, xc.shape[-1])
        # the second output is GELU'(h), not h: backward multiplies instead of evaluating erf + exp again
        a, h = _tc_linear(x2, w1c, b1, capi.EPI_GELU_SAVE_DERIV if MLP_SAVE_DERIV else capi.EPI_GELU_SAVE)
        y = torch.nn.functional.linear(a, w2c, cast_param(b2, torch.bfloat16))
        ctx.save_for_backward(x2, w1c, w2c, h, a)
        ctx.deriv = MLP_SAVE_DERIV
        ctx.meta = (x.dtype, x.shape, w1.dtype, b1.dtype, w2.dtype, None if b2 is None else b2.dtype)
        return y.view(*x.shape[:-1], w2c.shape[0])

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gy):
        x2, w1c, w2c, h, a = ctx.saved_tensors
        x_dtype, x_shape, w1_dtype, b1_dtype, w2_dtype, b2_dtype = ctx.meta
        g2 = gy.reshape(-1, w2c.shape[0])
        if g2.dtype != torch.bfloat16 or not g2.is_contiguous():
            g2 = g2.to(torch.bfloat16).contiguous()
        # (the fc1 bias gradient may be deferred when it goes to autograd as it is: fp32 parameter)
        dh, gb1 = _tc_dgelu(g2, w2c, h, ctx.deriv, defer_bias=b1_dtype == torch.float32 and ctx.needs_input_grad[2])
        gx = gw1 = gw2 = gb2 = None
        if ctx.needs_input_grad[0]:
            gx = torch.mm(dh, w1c).view(x_shape).to(x_dtype)
        if ctx.needs_input_grad[1]:
            gw1 = _wgrad(dh, x2, w1_dtype)
        want_b2 = b2_dtype is not None and ctx.needs_input_grad[4]
        if ctx.needs_input_grad[3] and _tc_wgrad_ok(g2, a, w2_dtype, want_b2):
            gw2, gb2 = _tc_wgrad(g2, a, want_b2)
            gb2 = None if gb2 is None else gb2.to(b2_dtype)
        else:
            if ctx.needs_input_grad[3]:
                gw2 = _wgrad(g2, a, w2_dtype)
            if want_b2:
                gb2 = _bias_grad(g2, g2.shape[1], b2_dtype)
        return gx, gw1, (gb1.to(b1_dtype) if ctx.needs_input_grad[2] else None), gw2, gb2


def mlp_fused_supported(x: torch.Tensor, w1: torch.Tensor, b1, w2: torch.Tensor) -> bool:
    """bf16 autocast (or bf16 input), K = C in {64, 128, 256}, hidden width a multiple of 64, and the fused path
    switched on (TC_LINEAR["gelu"])."""
    if not TC_LINEAR["gelu"] or b1 is None or not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16):
        return False
    dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
    if dt != torch.bfloat16 or w2.shape[1] != w1.shape[0] or w2.shape[0] != w1.shape[1]:
        return False
    M = x.numel() // x.shape[-1]
    N, K = w1.shape
    lib = capi.lib()
    return bool(lib.csb200_linear_supported(M, N, K, capi.BF16)) and bool(lib.csb200_linear_dgelu_supported(M, N, K, capi.BF16))


def mlp_fused(x, w1, b1, w2, b2, defer_fc2_bias_grad: bool = False) -> torch.Tensor:
    """fc2(GELU(fc1(x))) through _MlpFn.  ``defer_fc2_bias_grad``: as in ``linear``."""
    if defer_fc2_bias_grad and b2 is not None:
        b2 = b2.detach()
    return _MlpFn.apply(x, w1, b1, w2, b2)


def linear_gelu_supported(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> bool:
    if bias is None or not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16):
        return False
    dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
    return dt in (torch.float32, torch.bfloat16) and \
        bool(capi.lib().csb200_gelu_supported(weight.shape[0], capi._DTYPES[dt]))


def linear_gelu(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """GELU(F.linear(x, weight, bias)) with the fused csb200 GELU passes (exact erf form)."""
    dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
    return _LinearGeluFn.apply(x, weight, bias, dt)


class _Conv2dFn(torch.autograd.Function):
    """F.conv2d on cuDNN with the bias gradient taken out of ATen's strided reduce_kernel: on a
    channels-last gradient it is a column sum of the (B*H*W, C) matrix (csb200_colsum)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, weight, bias, stride, padding, compute_dtype):
        xc = x if x.dtype == compute_dtype else x.to(compute_dtype)
        wc, bc = cast_param(weight, compute_dtype), cast_param(bias, compute_dtype)
        ctx.save_for_backward(xc, wc)
        ctx.cfg = (stride, padding, x.dtype, weight.dtype, None if bias is None else bias.dtype)
        if bias is None:
            return torch.nn.functional.conv2d(xc, wc, None, stride, padding)
        y = torch.nn.functional.conv2d(xc, wc, None, stride, padding)
        return add_row_bias(y, bias, inplace=True)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gy):
        xc, wc = ctx.saved_tensors
        stride, padding, x_dtype, w_dtype, b_dtype = ctx.cfg
        gy = gy.to(xc.dtype)
        gx, gw, _ = torch.ops.aten.convolution_backward(
            gy, xc, wc, None, list(stride), list(padding), [1, 1], False, [0, 0], 1,
            [ctx.needs_input_grad[0], ctx.needs_input_grad[1], False])
        gb = None
        if b_dtype is not None and ctx.needs_input_grad[2]:
            gb = channel_sum(gy, b_dtype)
        return (None if gx is None else gx.to(x_dtype)), (None if gw is None else gw.to(w_dtype)), gb, None, None, None


def channel_sum(g: torch.Tensor, final_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """fp32 sum over (B, H, W) of a (B, C, H, W) tensor; one flat csb200 pass when it is channels-last.
    ``final_dtype``: as in ``_bias_grad`` (the result goes to autograd unread)."""
    B, C, H, W = g.shape
    if g.dtype in (torch.float32, torch.bfloat16) and g.is_cuda:
        if not g.is_contiguous(memory_format=torch.channels_last) and g.stride(1) == 1:
            # a channel slice of a wider channels-last tensor (the backward of the decoder's torch.cat,
            # C:657): one compacting copy + a flat column sum beats ATen's strided reduction 5x
            g = g.contiguous(memory_format=torch.channels_last)
        if g.is_contiguous(memory_format=torch.channels_last):
            return _bias_grad(g.permute(0, 2, 3, 1).reshape(B * H * W, C), C, final_dtype)
    out = g.sum((0, 2, 3), dtype=torch.float32)
    return out if final_dtype is None else out.to(final_dtype)


def conv2d(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], stride=1, padding=0) -> torch.Tensor:
    """Drop-in for ``F.conv2d`` (groups = 1, dilation = 1) on CUDA float32 / bfloat16 tensors."""
    pair = lambda v: (v, v) if isinstance(v, int) else tuple(v)
    if not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16):
        return torch.nn.functional.conv2d(x, weight, bias, stride, padding)
    dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
    if dt not in (torch.float32, torch.bfloat16):
        return torch.nn.functional.conv2d(x, weight, bias, stride, padding)
    return _Conv2dFn.apply(x, weight, bias, pair(stride), pair(padding), dt)


def add_row_bias(x: torch.Tensor, bias: torch.Tensor, inplace: bool = False) -> torch.Tensor:
    """x + bias[None, :, None, None] for a (B, C, H, W) tensor.  Channels-last tensors of a tiled width
    take one flat csb200 pass (cuDNN leaves the bias of a convolution to a broadcasting add_ that ATen
    runs through its non-vectorised elementwise kernel); everything else takes the ATen add."""
    B, C, H, W = x.shape
    if x.is_cuda and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous() \
            and x.dtype in (torch.float32, torch.bfloat16) and x.data_ptr() % 16 == 0 \
            and capi.lib().csb200_colsum_supported(C, capi.dtype_code(x)):
        y = x if inplace else torch.empty_like(x)  # preserves the channels-last strides
        b32 = bias.detach().float().contiguous()
        with torch.cuda.device(x.device), _span("row_bias", 2 * x.numel() * x.element_size()):
            capi.check(capi.lib().csb200_add_row_bias(_ptr(x), _ptr(b32), _ptr(y), B * H * W, C, capi.dtype_code(x),
                                                      _vp(capi.stream_of(x))), "csb200_add_row_bias")
        return y
    b = bias.to(x.dtype).view(1, -1, 1, 1)
    return x.add_(b) if inplace else x + b


class _ChannelBiasFn(torch.autograd.Function):
    """x + bias[None, :, None, None] whose bias gradient is a csb200 column sum."""

    @staticmethod
    def forward(ctx, x, bias):
        ctx.b_dtype = bias.dtype
        return add_row_bias(x, bias)

    @staticmethod
    def backward(ctx, g):
        return g, (channel_sum(g, ctx.b_dtype) if ctx.needs_input_grad[1] else None)


def add_channel_bias(x: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    return _ChannelBiasFn.apply(x, bias) if x.is_cuda else x + bias.to(x.dtype).view(1, -1, 1, 1)


# ------------------------------------------------------------------------------------------------
# CARAFE reassembly
# ------------------------------------------------------------------------------------------------
class _CarafeFn(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, low, enc, up):
        capi.require_cuda(low, enc)
        CL = torch.channels_last
        low = low.contiguous(memory_format=CL)
        enc = enc.to(low.dtype).contiguous(memory_format=CL)
        B, C, H, W = low.shape
        out = torch.empty((B, C, H * up, W * up), dtype=low.dtype, device=low.device, memory_format=CL)
        wt = torch.empty((B, H * up, W * up, 9), dtype=low.dtype, device=low.device)
        nbytes = (low.numel() + enc.numel() + out.numel() + wt.numel()) * low.element_size()
        with torch.cuda.device(low.device), _span("carafe_fwd", nbytes):
            capi.check(capi.lib().csb200_carafe_fwd(_ptr(low), _ptr(enc), _ptr(out), _ptr(wt), B, H, W, C, up,
                                                    capi.dtype_code(low), _vp(capi.stream_of(low))),
                       "csb200_carafe_fwd")
        ctx.save_for_backward(low, wt)
        ctx.cfg = (up, enc.shape)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        low, wt = ctx.saved_tensors
        up, enc_shape = ctx.cfg
        CL = torch.channels_last
        B, C, H, W = low.shape
        gout = gout.to(low.dtype).contiguous(memory_format=CL)
        d_low = torch.empty_like(low, memory_format=CL)
        d_enc = torch.empty(enc_shape, dtype=low.dtype, device=low.device, memory_format=CL)
        nbytes = (2 * low.numel() + d_enc.numel() + 2 * gout.numel() + 2 * wt.numel()) * low.element_size()
        with torch.cuda.device(low.device), _span("carafe_bwd", nbytes):
            capi.check(capi.lib().csb200_carafe_bwd(_ptr(low), _ptr(wt), _ptr(gout), _ptr(d_low), _ptr(d_enc), B, H, W,
                                                    C, up, capi.dtype_code(low), _vp(capi.stream_of(low))),
                       "csb200_carafe_bwd")
        return d_low, d_enc, None


def carafe_supported(low: torch.Tensor) -> bool:
    return low.is_cuda and low.dtype in (torch.float32, torch.bfloat16) and \
        bool(capi.lib().csb200_carafe_supported(low.shape[1]))


def carafe_reassemble(low: torch.Tensor, enc: torch.Tensor, up: int) -> torch.Tensor:
    """Fused kernel-softmax + content-aware reassembly (C:408-431).  low: (B, C, H, W) features,
    enc: (B, 9*up*up, H, W) raw encoder logits; returns (B, C, H*up, W*up), channels-last."""
    return _CarafeFn.apply(low, enc, up)


# ------------------------------------------------------------------------------------------------
# stripe attention
# ------------------------------------------------------------------------------------------------
class Branch:
    """Geometry of one LePEAttention branch inside a (B, L, C_total) channel layout."""
    __slots__ = ("h_sp", "w_sp", "heads", "chan0", "chans")

    def __init__(self, h_sp: int, w_sp: int, heads: int, chan0: int, chans: int):
        self.h_sp, self.w_sp, self.heads, self.chan0, self.chans = h_sp, w_sp, heads, chan0, chans


# ---- attention dropout (`attn = self.attn_drop(attn)`, C:290) -------------------------------------------------
# The keep decisions are generated INSIDE the forward kernels (Philox4x32-10) from a device-resident
# (seed, call counter) pair, so a captured CUDA graph draws a fresh mask on every replay; the forward kernel
# writes the mask as bits and the backward kernels read it back (include/csb200.h, csb200_stripe_desc.drop_p).
_drop_state = {}
KEEP_LAST_DROP_MASKS = False   # tests: keep the masks of the last forward call in `last_drop_masks`
last_drop_masks = None


def attention_dropout_state(device) -> torch.Tensor:
    """int64 [2] on `device`: seed, number of attention calls drawn so far."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _drop_state.get(idx)
    if st is None:
        st = _drop_state[idx] = torch.tensor([torch.initial_seed() & (2 ** 62 - 1), 0], dtype=torch.int64,
                                             device=torch.device("cuda", idx))
    return st


def seed_attention_dropout(seed: int, device=None, counter: int = 0) -> None:
    """(Re)seed the attention-dropout generator of `device` (default: the current CUDA device)."""
    st = attention_dropout_state(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    st.copy_(torch.tensor([seed & (2 ** 62 - 1), counter], dtype=torch.int64))


def _desc(dtype_code, B, H, W, br: Branch, scale, engine, qkv_strides, o_strides, g_strides=None,
          drop=None):
    d = capi.StripeDesc()
    d.dtype, d.batch, d.height, d.width = dtype_code, B, H, W
    d.h_sp, d.w_sp, d.heads, d.head_dim = br.h_sp, br.w_sp, br.heads, br.chans // br.heads
    d.scale, d.engine = float(scale), engine
    (d.q_sb, d.q_sl), (d.k_sb, d.k_sl), (d.v_sb, d.v_sl) = qkv_strides
    d.o_sb, d.o_sl = o_strides
    if g_strides is not None:
        (d.dq_sb, d.dq_sl), (d.dk_sb, d.dk_sl), (d.dv_sb, d.dv_sl) = g_strides
    if drop is not None:  # (p, salt, rng state tensor or None, mask tensor)
        d.drop_p, d.drop_salt = float(drop[0]), int(drop[1])
        d.rng_state = None if drop[2] is None else drop[2].data_ptr()
        d.drop_mask = drop[3].data_ptr()
    return d


def _attn_work(B, L, br: Branch, itemsize, backward):
    """Algorithmic (bytes, flops) of one branch call: q, k, v, out cross HBM once
    (backward: + grad_out in, 3 gradients out); 4 N^2 hd FLOPs per (stripe, head) forward, 10 backward."""
    N, hd = br.h_sp * br.w_sp, br.chans // br.heads
    tensors = 8 if backward else 4
    return tensors * B * L * br.chans * itemsize, (10 if backward else 4) * N * hd * B * L * br.heads


def _attn_tag(B, L, C, branches) -> str:
    return f"B{B}xL{L}xC{C}xN{branches[0].h_sp * branches[0].w_sp}x{len(branches)}br"


_ENGINE = {"auto": capi.ENGINE_AUTO, "simt": capi.ENGINE_SIMT, "tcgen05": capi.ENGINE_TCGEN05}


def stripe_engine(dtype: torch.dtype, B, H, W, br: Branch, backward=False, engine="auto") -> str:
    """Which engine csb200_stripe_attn_fwd / _bwd runs for this shape (csb200_stripe_attn_engine): "simt" or
    "tcgen05".  Raises like the launch would for an invalid / untileable request."""
    C = br.chans
    d = _desc(0 if dtype == torch.float32 else 1, B, H, W, br, 1.0, _ENGINE[engine], [(H * W * 3 * C, 3 * C)] * 3,
              (H * W * C, C), [(H * W * 3 * C, 3 * C)] * 3)
    rc = capi.lib().csb200_stripe_attn_engine(ctypes.byref(d), int(backward))
    if rc < 0:
        raise RuntimeError(capi.last_error())
    return {capi.ENGINE_SIMT: "simt", capi.ENGINE_TCGEN05: "tcgen05"}[rc]


class _CrossStripeFn(torch.autograd.Function):
    """All branches of one CSWinBlock on the packed (B, L, 3C) qkv buffer -> (B, L, C).

    Fuses what the reference does with slicing, 12 copies per branch and a torch.cat (C:358-363):
    every branch reads q/k/v in place through strides and writes its channel range of ``out``;
    backward writes one packed grad_qkv, so the qkv Linear sees a single contiguous gradient.
    """

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, qkv, H, W, branches, scale, engine, drop_p, *wb):
        global last_drop_masks
        capi.require_cuda(qkv)
        qkv = qkv.contiguous()
        B, L, C3 = qkv.shape
        C = C3 // 3
        if L != H * W:
            raise AssertionError("flatten img_tokens has wrong size")  # C:281, C:356
        code = capi.dtype_code(qkv)
        out = torch.empty((B, L, C), dtype=qkv.dtype, device=qkv.device)
        lses = [torch.empty((B, b.heads, L), dtype=torch.float32, device=qkv.device) for b in branches]
        ws = [w.detach().float().contiguous() for w in wb]
        lib = capi.lib()
        st = _vp(capi.stream_of(qkv))
        n = len(branches)
        descs = (capi.StripeDesc * n)()
        ios = (capi.BranchIO * n)()
        nbytes = flops = 0
        masks, used = [], None
        if drop_p > 0:  # one call counter per attention call, advanced on the device (graph-capturable)
            state = attention_dropout_state(qkv.device)
            used = state.clone()
            state[1:].add_(1)
            masks = [torch.empty((B, b.heads, L, (b.h_sp * b.w_sp + 31) // 32), dtype=torch.int32, device=qkv.device)
                     for b in branches]
            if KEEP_LAST_DROP_MASKS:
                last_drop_masks = masks
        for i, br in enumerate(branches):
            descs[i] = _desc(code, B, H, W, br, scale, engine, [(L * C3, C3)] * 3, (L * C, C),
                             drop=(drop_p, i, used, masks[i]) if drop_p > 0 else None)
            io = ios[i]
            io.q, io.k, io.v = (_ptr(qkv, br.chan0).value, _ptr(qkv, C + br.chan0).value,
                                _ptr(qkv, 2 * C + br.chan0).value)
            io.lepe_w, io.lepe_b = ws[2 * i].data_ptr(), ws[2 * i + 1].data_ptr()
            io.out, io.lse = _ptr(out, br.chan0).value, lses[i].data_ptr()
            w_ = _attn_work(B, L, br, qkv.element_size(), False)
            nbytes, flops = nbytes + w_[0], flops + w_[1]
        with torch.cuda.device(qkv.device), _span("attn_fwd", nbytes, flops, _attn_tag(B, L, C, branches)):
            capi.check(lib.csb200_cross_stripe_attn_fwd(n, descs, ios, st), "csb200_cross_stripe_attn_fwd")
        ctx.save_for_backward(qkv, out, *lses, *ws, *masks)
        ctx.cfg = (H, W, branches, scale, engine, drop_p)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        H, W, branches, scale, engine, drop_p = ctx.cfg
        qkv, out, *rest = ctx.saved_tensors
        nb = len(branches)
        lses, ws, masks = rest[:nb], rest[nb:3 * nb], rest[3 * nb:]
        B, L, C3 = qkv.shape
        C = C3 // 3
        code = capi.dtype_code(qkv)
        gout = gout.contiguous()
        if gout.dtype != qkv.dtype:
            gout = gout.to(qkv.dtype)
        gqkv = torch.empty_like(qkv)
        lib = capi.lib()
        st = _vp(capi.stream_of(qkv))
        grads, keep = [], []
        n = len(branches)
        descs = (capi.StripeDesc * n)()
        ios = (capi.BranchIO * n)()
        nbytes = flops = 0
        for i, br in enumerate(branches):
            s3 = [(L * C3, C3)] * 3
            descs[i] = _desc(code, B, H, W, br, scale, engine, s3, (L * C, C), s3,
                             drop=(drop_p, i, None, masks[i]) if drop_p > 0 else None)
            nws = lib.csb200_stripe_attn_bwd_workspace_bytes(ctypes.byref(descs[i]))
            wsp = torch.empty(max(nws, 16), dtype=torch.uint8, device=qkv.device)
            gw, gb = torch.empty_like(ws[2 * i]), torch.empty_like(ws[2 * i + 1])
            keep.append(wsp)
            io = ios[i]
            io.q, io.k, io.v = (_ptr(qkv, br.chan0).value, _ptr(qkv, C + br.chan0).value,
                                _ptr(qkv, 2 * C + br.chan0).value)
            io.lepe_w, io.lepe_b = ws[2 * i].data_ptr(), ws[2 * i + 1].data_ptr()
            io.out, io.lse = _ptr(out, br.chan0).value, lses[i].data_ptr()
            io.grad_out = _ptr(gout, br.chan0).value
            io.dq, io.dk, io.dv = (_ptr(gqkv, br.chan0).value, _ptr(gqkv, C + br.chan0).value,
                                   _ptr(gqkv, 2 * C + br.chan0).value)
            io.grad_lepe_w, io.grad_lepe_b = gw.data_ptr(), gb.data_ptr()
            io.workspace, io.workspace_bytes = wsp.data_ptr(), nws
            grads += [gw, gb]
            w_ = _attn_work(B, L, br, qkv.element_size(), True)
            nbytes, flops = nbytes + w_[0], flops + w_[1]
        with torch.cuda.device(qkv.device), _span("attn_bwd", nbytes, flops, _attn_tag(B, L, C, branches)):
            capi.check(lib.csb200_cross_stripe_attn_bwd(n, descs, ios, st), "csb200_cross_stripe_attn_bwd")
        return (gqkv, None, None, None, None, None, None, *grads)


def cross_stripe_attention(qkv: torch.Tensor, H: int, W: int, branches: Sequence[Branch], scale: float,
                           weights_and_biases: Sequence[torch.Tensor], engine: str = "auto",
                           drop_p: float = 0.0) -> torch.Tensor:
    """qkv: (B, L, 3C) packed as [q | k | v] along channels (the output of CSWinBlock.qkv, C:358).

    ``branches`` partition the C channels; ``weights_and_biases`` = [w0, b0, w1, b1, ...] are the
    get_v parameters ((C',1,3,3), (C',)) of each branch.  ``drop_p``: attention dropout on the softmax
    probabilities (C:290), generated inside the kernels (quantised to k/256).  Returns (B, L, C).
    """
    if not 0.0 <= drop_p < 1.0:
        raise ValueError(f"dropout probability has to be in [0, 1), got {drop_p}")
    out = _CrossStripeFn.apply(qkv, H, W, tuple(branches), scale, _ENGINE[engine], float(drop_p), *weights_and_biases)
    return out


def stripe_attention(q, k, v, lepe_w, lepe_b, H, W, h_sp, w_sp, heads, scale=None, engine="auto", drop_p=0.0):
    """One branch on separate (B, L, C') q, k, v (any strides) — LePEAttention.forward, C:271-298."""
    Cb = q.shape[-1]
    scale = (Cb // heads) ** -0.5 if scale is None else scale
    qkv = torch.cat([q, k, v], dim=-1)  # generic-stride entry: one packing copy, then the fused path
    return cross_stripe_attention(qkv, H, W, [Branch(h_sp, w_sp, heads, 0, Cb)], scale, [lepe_w, lepe_b], engine,
                                  drop_p)
