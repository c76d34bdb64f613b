"""ctypes binding of libcsb200.so (include/csb200.h) — the only door between PyTorch host code and
the sm_100a kernels.  There is NO CPU fallback: a missing library or a CPU tensor raises."""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcsb200.so")

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE = 0, 1, 2, 3, 4
F32, BF16 = 0, 1
NCHW, NLC = 0, 1
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1, 2

_DTYPES = {torch.float32: F32, torch.bfloat16: BF16}


class StripeDesc(ctypes.Structure):
    """Mirror of csb200_stripe_desc."""
    _fields_ = [("dtype", ctypes.c_int32), ("batch", ctypes.c_int32), ("height", ctypes.c_int32),
                ("width", ctypes.c_int32), ("h_sp", ctypes.c_int32), ("w_sp", ctypes.c_int32),
                ("heads", ctypes.c_int32), ("head_dim", ctypes.c_int32), ("scale", ctypes.c_float),
                ("engine", ctypes.c_int32)] + [
        (n, ctypes.c_int64) for n in ("q_sb", "q_sl", "k_sb", "k_sl", "v_sb", "v_sl", "o_sb", "o_sl",
                                      "dq_sb", "dq_sl", "dk_sb", "dk_sl", "dv_sb", "dv_sl")] + [
        ("drop_p", ctypes.c_float), ("drop_salt", ctypes.c_int32), ("rng_state", ctypes.c_void_p),
        ("drop_mask", ctypes.c_void_p)]


class BranchIO(ctypes.Structure):
    """Mirror of csb200_branch_io."""
    _fields_ = [(n, ctypes.c_void_p) for n in ("q", "k", "v", "lepe_w", "lepe_b", "out", "lse", "grad_out", "dq",
                                               "dk", "dv", "grad_lepe_w", "grad_lepe_b", "workspace")] + \
               [("workspace_bytes", ctypes.c_size_t)]


_lib = None
_lock = threading.Lock()

EXPORTS = ("csb200_abi_version", "csb200_last_error_string", "csb200_launch_count", "csb200_simam_fwd",
           "csb200_simam_bwd", "csb200_simam_workspace_bytes", "csb200_simam_fwd_ws", "csb200_simam_bwd_ws",
           "csb200_layernorm_supported", "csb200_layernorm_fwd",
           "csb200_layernorm_bwd_workspace_bytes", "csb200_layernorm_bwd", "csb200_add_layernorm_fwd",
           "csb200_add_layernorm_bwd", "csb200_add_layernorm_bwd_rb", "csb200_colsum_supported",
           "csb200_colsum_workspace_bytes", "csb200_colsum", "csb200_add_row_bias", "csb200_gelu_supported", "csb200_gelu_fwd",
           "csb200_gelu_bwd_workspace_bytes", "csb200_gelu_bwd", "csb200_carafe_supported", "csb200_carafe_fwd",
           "csb200_carafe_bwd", "csb200_stripe_attn_engine", "csb200_stripe_attn_fwd",
           "csb200_stripe_attn_bwd_workspace_bytes", "csb200_stripe_attn_bwd", "csb200_cross_stripe_attn_fwd",
           "csb200_cross_stripe_attn_bwd", "csb200_adam_chunk_elems", "csb200_adam_step",
           "csb200_linear_supported", "csb200_linear_fwd", "csb200_linear_dgelu_supported",
           "csb200_linear_dgelu_workspace_bytes", "csb200_linear_dgelu_bwd", "csb200_linear_dact_bwd",
           "csb200_linear_wgrad_supported",
           "csb200_linear_wgrad", "csb200_linear_wgrad_acc", "csb200_layernorm_bwd_partials", "csb200_colsum_partials",
           "csb200_linear_dact_bwd_partials", "csb200_sum_rows_deferred", "csb200_sum_rows_pending",
           "csb200_sum_rows_flush", "csb200_sum_rows_discard")
EPI_BIAS, EPI_GELU, EPI_GELU_SAVE, EPI_GELU_SAVE_DERIV = 0, 1, 2, 3


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C cswin-simam-unet_b200/csrc`). There is no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        vp, i64, f32p = ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p
        L.csb200_abi_version.restype = ctypes.c_int
        L.csb200_last_error_string.restype = ctypes.c_char_p
        L.csb200_launch_count.restype = ctypes.c_uint64
        L.csb200_simam_fwd.argtypes = [vp, vp, f32p, i64, i64, i64, ctypes.c_int, ctypes.c_int, ctypes.c_float, vp]
        L.csb200_simam_bwd.argtypes = [vp, vp, f32p, vp, i64, i64, i64, ctypes.c_int, ctypes.c_int, ctypes.c_float, vp]
        L.csb200_simam_workspace_bytes.argtypes = [i64, i64, i64, ctypes.c_int, ctypes.c_int]
        L.csb200_simam_workspace_bytes.restype = ctypes.c_size_t
        L.csb200_simam_fwd_ws.argtypes = [vp, vp, f32p, i64, i64, i64, ctypes.c_int, ctypes.c_int, ctypes.c_float, vp,
                                          ctypes.c_size_t, vp]
        L.csb200_simam_bwd_ws.argtypes = [vp, vp, f32p, vp, i64, i64, i64, ctypes.c_int, ctypes.c_int, ctypes.c_float, vp,
                                          ctypes.c_size_t, vp]
        L.csb200_simam_fwd_ws.restype = L.csb200_simam_bwd_ws.restype = ctypes.c_int
        L.csb200_layernorm_supported.argtypes = [i64, ctypes.c_int]
        L.csb200_layernorm_supported.restype = ctypes.c_int
        L.csb200_layernorm_fwd.argtypes = [vp, vp, vp, vp, vp, i64, i64, ctypes.c_int, ctypes.c_int, ctypes.c_float, vp]
        L.csb200_layernorm_fwd.restype = ctypes.c_int
        L.csb200_layernorm_bwd_workspace_bytes.argtypes = [i64, i64]
        L.csb200_layernorm_bwd_workspace_bytes.restype = ctypes.c_size_t
        L.csb200_layernorm_bwd.argtypes = [vp] * 8 + [ctypes.c_size_t, i64, i64, ctypes.c_int, ctypes.c_int, vp]
        L.csb200_layernorm_bwd.restype = ctypes.c_int
        L.csb200_add_layernorm_fwd.argtypes = [vp] * 7 + [i64, i64, ctypes.c_int, ctypes.c_int, ctypes.c_float, vp]
        L.csb200_add_layernorm_fwd.restype = ctypes.c_int
        L.csb200_add_layernorm_bwd.argtypes = [vp] * 9 + [ctypes.c_size_t, i64, i64, ctypes.c_int, ctypes.c_int, vp]
        L.csb200_add_layernorm_bwd.restype = ctypes.c_int
        L.csb200_add_layernorm_bwd_rb.argtypes = [vp] * 10 + [ctypes.c_size_t, i64, i64, ctypes.c_int, ctypes.c_int, vp]
        L.csb200_add_layernorm_bwd_rb.restype = ctypes.c_int
        L.csb200_adam_chunk_elems.restype = ctypes.c_int64
        L.csb200_adam_step.argtypes = [vp, vp, i64, vp, vp, vp]
        L.csb200_adam_step.restype = ctypes.c_int
        L.csb200_colsum_supported.argtypes = [i64, ctypes.c_int]
        L.csb200_colsum_supported.restype = ctypes.c_int
        L.csb200_colsum_workspace_bytes.argtypes = [i64]
        L.csb200_colsum_workspace_bytes.restype = ctypes.c_size_t
        L.csb200_colsum.argtypes = [vp, vp, vp, ctypes.c_size_t, i64, i64, ctypes.c_int, vp]
        L.csb200_colsum.restype = ctypes.c_int
        L.csb200_add_row_bias.argtypes = [vp, vp, vp, i64, i64, ctypes.c_int, vp]
        L.csb200_add_row_bias.restype = ctypes.c_int
        L.csb200_gelu_supported.argtypes = [i64, ctypes.c_int]
        L.csb200_gelu_supported.restype = ctypes.c_int
        L.csb200_gelu_fwd.argtypes = [vp, vp, i64, i64, ctypes.c_int, vp]
        L.csb200_gelu_fwd.restype = ctypes.c_int
        L.csb200_gelu_bwd_workspace_bytes.argtypes = [i64]
        L.csb200_gelu_bwd_workspace_bytes.restype = ctypes.c_size_t
        L.csb200_gelu_bwd.argtypes = [vp, vp, vp, vp, vp, ctypes.c_size_t, i64, i64, ctypes.c_int, vp]
        L.csb200_gelu_bwd.restype = ctypes.c_int
        L.csb200_linear_supported.argtypes = [i64, i64, i64, ctypes.c_int]
        L.csb200_linear_supported.restype = ctypes.c_int
        L.csb200_linear_fwd.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i64, ctypes.c_int, ctypes.c_int, vp]
        L.csb200_linear_fwd.restype = ctypes.c_int
        L.csb200_linear_dgelu_supported.argtypes = [i64, i64, i64, ctypes.c_int]
        L.csb200_linear_dgelu_supported.restype = ctypes.c_int
        L.csb200_linear_dgelu_workspace_bytes.argtypes = [i64]
        L.csb200_linear_dgelu_workspace_bytes.restype = ctypes.c_size_t
        L.csb200_linear_dgelu_bwd.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_size_t, i64, i64, i64, i64, ctypes.c_int, vp]
        L.csb200_linear_dgelu_bwd.restype = ctypes.c_int
        L.csb200_linear_dact_bwd.argtypes = L.csb200_linear_dgelu_bwd.argtypes
        L.csb200_linear_dact_bwd.restype = ctypes.c_int
        L.csb200_linear_wgrad_supported.argtypes = [i64, i64, i64, ctypes.c_int]
        L.csb200_linear_wgrad_supported.restype = ctypes.c_int
        L.csb200_linear_wgrad.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, i64, ctypes.c_int, vp]
        L.csb200_linear_wgrad.restype = ctypes.c_int
        L.csb200_linear_wgrad_acc.argtypes = L.csb200_linear_wgrad.argtypes
        L.csb200_linear_wgrad_acc.restype = ctypes.c_int
        pp, ip = ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int32)
        L.csb200_layernorm_bwd_partials.argtypes = [vp] * 6 + [ctypes.c_int, vp, ctypes.c_size_t, i64, i64, ctypes.c_int,
                                                    ctypes.c_int, pp, ip, vp]
        L.csb200_colsum_partials.argtypes = [vp, vp, ctypes.c_size_t, i64, i64, ctypes.c_int, pp, ip, vp]
        L.csb200_linear_dact_bwd_partials.argtypes = [vp, vp, vp, vp, vp, ctypes.c_size_t, i64, i64, i64, i64,
                                                      ctypes.c_int, ctypes.c_int, pp, ip, vp]
        L.csb200_sum_rows_deferred.argtypes = [vp, i64, i64, i64, vp]
        L.csb200_sum_rows_flush.argtypes = [vp]
        L.csb200_sum_rows_pending.restype = ctypes.c_int64
        for fn in ("csb200_layernorm_bwd_partials", "csb200_colsum_partials", "csb200_linear_dact_bwd_partials",
                   "csb200_sum_rows_deferred", "csb200_sum_rows_flush", "csb200_sum_rows_discard"):
            getattr(L, fn).restype = ctypes.c_int
        L.csb200_carafe_supported.argtypes = [i64]
        L.csb200_carafe_supported.restype = ctypes.c_int
        L.csb200_carafe_fwd.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, ctypes.c_int, ctypes.c_int, vp]
        L.csb200_carafe_fwd.restype = ctypes.c_int
        L.csb200_carafe_bwd.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i64, ctypes.c_int, ctypes.c_int, vp]
        L.csb200_carafe_bwd.restype = ctypes.c_int
        dp = ctypes.POINTER(StripeDesc)
        L.csb200_stripe_attn_engine.argtypes = [dp, ctypes.c_int]
        L.csb200_stripe_attn_fwd.argtypes = [dp, vp, vp, vp, f32p, f32p, vp, f32p, vp]
        L.csb200_stripe_attn_bwd_workspace_bytes.argtypes = [dp]
        L.csb200_stripe_attn_bwd_workspace_bytes.restype = ctypes.c_size_t
        L.csb200_stripe_attn_bwd.argtypes = [dp] + [vp] * 14 + [ctypes.c_size_t, vp]
        bp = ctypes.POINTER(BranchIO)
        L.csb200_cross_stripe_attn_fwd.argtypes = [ctypes.c_int, dp, bp, vp]
        L.csb200_cross_stripe_attn_bwd.argtypes = [ctypes.c_int, dp, bp, vp]
        L.csb200_cross_stripe_attn_fwd.restype = L.csb200_cross_stripe_attn_bwd.restype = ctypes.c_int
        for fn in ("csb200_simam_fwd", "csb200_simam_bwd", "csb200_stripe_attn_engine",
                   "csb200_stripe_attn_fwd", "csb200_stripe_attn_bwd"):
            getattr(L, fn).restype = ctypes.c_int
        if L.csb200_abi_version() != 2:
            raise RuntimeError("libcsb200.so ABI version mismatch; rebuild")
        _lib = L
    return _lib


def last_error() -> str:
    return lib().csb200_last_error_string().decode()


def launch_count() -> int:
    return int(lib().csb200_launch_count())


def check(rc: int, what: str):
    """Shape errors surface as RuntimeError, like the view() failure of the reference (C:204)."""
    if rc != OK:
        raise RuntimeError(f"{what} failed (csb200 status {rc}): {last_error()}")


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"csb200 kernels take float32 or bfloat16 tensors, got {t.dtype}") from None


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("csb200 kernels run on sm_100a only: got a CPU tensor and there is no CPU fallback")


def stream_of(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream
