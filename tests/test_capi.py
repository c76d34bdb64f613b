"""CPU: the C-ABI library loads, exports every symbol include/csb200.h declares, and validates its
arguments without touching a GPU (no compute calls here)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from cswin_simam_unet_b200 import capi, functional as F_


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "csb200.h")).read()
    return sorted(set(re.findall(r"CSB200_API[^;(]*?\b(csb200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    declared = _declared_symbols()
    assert len(declared) >= 9
    assert set(declared) == set(capi.EXPORTS)
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert capi.lib().csb200_abi_version() == 2


def test_descriptor_layout_matches_header():
    # 10 x 4-byte fields, 14 x int64 (the first int64 is 8-byte aligned: 40 bytes in), then the dropout block:
    # float + int32 + two pointers
    assert ctypes.sizeof(capi.StripeDesc) == 40 + 14 * 8 + 8 + 16
    assert capi.StripeDesc.q_sb.offset == 40
    assert capi.StripeDesc.drop_p.offset == 152 and capi.StripeDesc.rng_state.offset == 160


def test_adam_tensor_layout_and_cpu_refusal():
    from cswin_simam_unet_b200 import optim
    # csb200_adam_tensor: 5 pointers, int64 numel, 2 x int32
    assert ctypes.sizeof(optim._AdamTensor) == 5 * 8 + 8 + 8 and optim._AdamTensor.numel.offset == 40
    assert capi.lib().csb200_adam_chunk_elems() > 0
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    with pytest.raises(RuntimeError, match="no CPU"):
        optim.FusedAdamW([p]).step()
    o = optim.fused_adam([p], lr=1e-3, weight_decay=1e-4)
    assert o.param_groups[0]["decoupled"] is False and o.param_groups[0]["betas"] == (0.9, 0.999)
    # checkpoints interchange with torch.optim.AdamW: per-parameter "step" entries collapse to ONE counter
    q = [torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2, 2))]
    t = torch.optim.AdamW(q, lr=1e-3)
    for v in q:
        v.grad = torch.ones_like(v)
    t.step(); t.step()
    f = optim.FusedAdamW(q, lr=1e-3)
    f.load_state_dict(t.state_dict())
    assert float(f._steps) == 2.0 and all(f.state[v]["step"] is f._steps for v in q)
    assert torch.equal(f.state[q[1]]["exp_avg"], t.state[q[1]]["exp_avg"])
    back = torch.optim.AdamW(q, lr=1e-3)
    back.load_state_dict(f.state_dict())
    assert float(back.state[q[0]]["step"]) == 2.0


def _desc(**kw):
    d = capi.StripeDesc()
    base = dict(dtype=capi.F32, batch=1, height=8, width=8, h_sp=8, w_sp=2, heads=1, head_dim=32, scale=0.1,
                engine=capi.ENGINE_AUTO, q_sb=8 * 8 * 96, q_sl=96, k_sb=8 * 8 * 96, k_sl=96, v_sb=8 * 8 * 96,
                v_sl=96, o_sb=8 * 8 * 32, o_sl=32)
    base.update(kw)
    for k, v in base.items():
        setattr(d, k, v)
    return d


def test_shape_validation_mirrors_reference_failures():
    lib = capi.lib()
    assert lib.csb200_stripe_attn_engine(ctypes.byref(_desc()), 0) == capi.ENGINE_SIMT
    # stripe does not divide the grid: the reference raises RuntimeError from view() (C:204)
    rc = lib.csb200_stripe_attn_engine(ctypes.byref(_desc(w_sp=7)), 0)
    assert rc == -capi.ERR_INVALID and "not divisible" in capi.last_error()
    assert lib.csb200_stripe_attn_engine(ctypes.byref(_desc(head_dim=64)), 0) == -capi.ERR_UNSUPPORTED
    assert lib.csb200_stripe_attn_engine(ctypes.byref(_desc(dtype=7)), 0) == -capi.ERR_INVALID
    assert lib.csb200_stripe_attn_engine(ctypes.byref(_desc(q_sl=97)), 0) == -capi.ERR_INVALID
    assert lib.csb200_stripe_attn_engine(None, 0) == -capi.ERR_INVALID
    # forced tcgen05 on a shape it cannot tile is refused, not silently rerouted
    assert lib.csb200_stripe_attn_engine(ctypes.byref(_desc(engine=capi.ENGINE_TCGEN05)), 0) == -capi.ERR_UNSUPPORTED
    assert lib.csb200_stripe_attn_bwd_workspace_bytes(ctypes.byref(_desc())) >= 8 * 8 * 4
    with pytest.raises(RuntimeError, match="not divisible"):
        capi.check(-lib.csb200_stripe_attn_engine(ctypes.byref(_desc(w_sp=7)), 0), "engine")


def test_engine_selection():
    lib = capi.lib()
    bf = dict(dtype=capi.BF16)
    # config-3 stripes (N = 128 / 256, width dividing 128) go to the tcgen05 engine in forward
    for kw in (dict(height=128, width=128, h_sp=128, w_sp=1), dict(height=64, width=64, h_sp=2, w_sp=64),
               dict(height=32, width=32, h_sp=32, w_sp=8), dict(height=16, width=16, h_sp=16, w_sp=16)):
        L = kw["height"] * kw["width"]
        d = _desc(q_sb=L * 96, k_sb=L * 96, v_sb=L * 96, o_sb=L * 32, **bf, **kw)
        assert lib.csb200_stripe_attn_engine(ctypes.byref(d), 0) == capi.ENGINE_TCGEN05, kw
        d.engine = capi.ENGINE_SIMT
        assert lib.csb200_stripe_attn_engine(ctypes.byref(d), 0) == capi.ENGINE_SIMT
    # config-1 stripes (N = 49, 56, 98) and fp32 stay on the CUDA-core engine
    d = _desc(height=14, width=14, h_sp=14, w_sp=7, q_sb=196 * 96, k_sb=196 * 96, v_sb=196 * 96, o_sb=196 * 32, **bf)
    assert lib.csb200_stripe_attn_engine(ctypes.byref(d), 0) == capi.ENGINE_SIMT
    d = _desc(height=16, width=16, h_sp=16, w_sp=16, q_sb=256 * 96, k_sb=256 * 96, v_sb=256 * 96, o_sb=256 * 32)
    assert lib.csb200_stripe_attn_engine(ctypes.byref(d), 0) == capi.ENGINE_SIMT


def test_engine_selection_for_config5_stripes():
    """BASELINE config 5 (1024^2 inference, C:232-242 geometry): stripes of 64 tokens (two per tile) and of
    512 / 1024 / 2048 tokens (key/value-tiled kernel) take the tcgen05 engine in FORWARD only; so does width 7."""
    lib = capi.lib()
    cases = (dict(height=64, width=64, h_sp=64, w_sp=1),      # split 1, stage 3: N = 64
             dict(height=64, width=64, h_sp=1, w_sp=64),
             dict(height=256, width=256, h_sp=256, w_sp=2),   # split 2, stage 1: N = 512
             dict(height=256, width=256, h_sp=8, w_sp=256),   # split 8, stage 1: N = 2048
             dict(height=32, width=32, h_sp=32, w_sp=32))     # last stage, full window: N = 1024
    for kw in cases:
        L = kw["height"] * kw["width"]
        d = _desc(dtype=capi.BF16, q_sb=L * 96, k_sb=L * 96, v_sb=L * 96, o_sb=L * 32, **kw)
        assert lib.csb200_stripe_attn_engine(ctypes.byref(d), 0) == capi.ENGINE_TCGEN05, kw
        assert lib.csb200_stripe_attn_engine(ctypes.byref(d), 1) == capi.ENGINE_SIMT, kw   # backward: CUDA cores
        d.engine = capi.ENGINE_TCGEN05
        assert lib.csb200_stripe_attn_engine(ctypes.byref(d), 1) == -capi.ERR_UNSUPPORTED  # forced: refused
        d.dtype = capi.F32
        d.engine = capi.ENGINE_AUTO
        assert lib.csb200_stripe_attn_engine(ctypes.byref(d), 0) == capi.ENGINE_SIMT
    L = 224 * 224   # 896^2, split 7: N = 1568 — ragged tiles, masked: still the tcgen05 engine in forward
    d = _desc(dtype=capi.BF16, height=224, width=224, h_sp=224, w_sp=7, q_sb=L * 96, k_sb=L * 96, v_sb=L * 96, o_sb=L * 32)
    assert lib.csb200_stripe_attn_engine(ctypes.byref(d), 0) == capi.ENGINE_TCGEN05
    assert lib.csb200_stripe_attn_engine(ctypes.byref(d), 1) == capi.ENGINE_SIMT
    # config 1 (224^2, split [1, 2, 7, 7]): stripes of <= 128 tokens that are not 64 or 128 stay on the CUDA cores
    d = _desc(dtype=capi.BF16, height=14, width=14, h_sp=14, w_sp=7, q_sb=196 * 96, k_sb=196 * 96, v_sb=196 * 96, o_sb=196 * 32)
    assert lib.csb200_stripe_attn_engine(ctypes.byref(d), 0) == capi.ENGINE_SIMT
    # the SimAM workspace query never fails: 0 = "the grid-resident kernels do not apply" (here: no GPU, or NCHW)
    assert lib.csb200_simam_workspace_bytes(32, 64, 16384, capi.NCHW, capi.BF16) == 0


def test_simam_argument_validation():
    lib = capi.lib()
    one = ctypes.c_void_p(16)
    assert lib.csb200_simam_fwd(one, one, None, 1, 1, 4, 9, capi.F32, 1e-4, None) == capi.ERR_INVALID  # layout
    assert lib.csb200_simam_fwd(one, one, None, 1, 1, 4, capi.NCHW, 5, 1e-4, None) == capi.ERR_INVALID  # dtype
    assert lib.csb200_simam_fwd(None, one, None, 1, 1, 4, capi.NCHW, capi.F32, 1e-4, None) == capi.ERR_INVALID
    assert lib.csb200_simam_bwd(one, None, None, one, 1, 1, 4, capi.NCHW, capi.F32, 1e-4, None) == capi.ERR_INVALID
    # empty tensors are a no-op, not an error (and launch nothing)
    n0 = capi.launch_count()
    assert lib.csb200_simam_fwd(None, None, None, 0, 4, 16, capi.NCHW, capi.F32, 1e-4, None) == capi.OK
    assert capi.launch_count() == n0


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        F_.simam(torch.randn(1, 2, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        F_.cross_stripe_attention(torch.randn(1, 16, 96), 4, 4, [F_.Branch(4, 4, 1, 0, 32)], 0.1,
                                  [torch.randn(32, 1, 3, 3), torch.randn(32)])
    with pytest.raises(TypeError):
        capi.dtype_code(torch.zeros(1, dtype=torch.float16))


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "cswin-simam-unet_b200")
    pat = re.compile(r"^\s*(from|import)\s+\.*oracle\b|import_module\([\"']oracle|#include\s+[<\"].*oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_deferred_sum_recorder_is_host_side_bookkeeping():
    """csb200_sum_rows_deferred only RECORDS (no device work until the flush): arguments are validated, the list
    is process-wide and can be dropped; functional.deferred_sums refuses to nest and drops its records when the
    block raises (no flush, hence nothing a CPU-only box could not do)."""
    lib = capi.lib()
    assert lib.csb200_sum_rows_discard() == capi.OK and lib.csb200_sum_rows_pending() == 0
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    assert lib.csb200_sum_rows_deferred(p, 4, 8, 8, p) == capi.OK
    assert lib.csb200_sum_rows_deferred(p, 2, 8, 16, p) == capi.OK
    assert lib.csb200_sum_rows_pending() == 2
    for bad in ((None, 4, 8, 8, p), (p, 4, 8, 8, None), (p, 0, 8, 8, p), (p, 4, 0, 8, p), (p, 4, 8, 4, p)):
        assert lib.csb200_sum_rows_deferred(*bad) == capi.ERR_INVALID
        assert "csb200_sum_rows_deferred" in capi.last_error()
    assert lib.csb200_sum_rows_pending() == 2
    assert lib.csb200_sum_rows_discard() == capi.OK and lib.csb200_sum_rows_pending() == 0
    with pytest.raises(ZeroDivisionError):
        with F_.deferred_sums("cpu") as block:
            assert lib.csb200_sum_rows_deferred(p, 4, 8, 8, p) == capi.OK
            with pytest.raises(RuntimeError, match="nest"):
                F_.deferred_sums("cpu").__enter__()
            assert F_._zeroed(100, "cpu") is None and block.zero_arena_numel == 0  # no arena: callers zero their own
            1 / 0
    assert lib.csb200_sum_rows_pending() == 0 and F_._deferred is None and F_._zero_arena is None
    assert block.arena_demand == 128  # what the block was asked for, rounded to 256-byte slices
