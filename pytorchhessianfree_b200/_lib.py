"""ctypes binding of ``csrc/libhf_b200.so`` (the C ABI declared in ``include/hf_b200.h``).

There is no CPU or PyTorch fallback: if the library is missing, or a call is made
with tensors that are not on a CUDA device, the error is raised here.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libhf_b200.so")

ABI_VERSION = 2
HF_F32, HF_F64 = 0, 1
LAYER_KIND = {"linear": 0, "conv2d": 1, "avgpool": 2}
ACT = {"none": 0, "relu": 1, "sigmoid": 2, "tanh": 3}
LOSS = {"mse": 0, "ce": 1, "bce": 2}
REDUCTION = {"mean": 0, "sum": 1}
PCG_ALPHA, PCG_BETA, PCG_FUSED = 1, 2, 3
LIN_HESSIAN, LIN_LOSS_ONLY = 1, 2
REASONS = {
    1: "Convergence (Martens)",
    2: "Number of iterations",
    3: "Divergence",
    4: "Convergence (tolerances)",
}


class PcgStatus(C.Structure):
    """Mirror of ``hf_pcg_status``."""

    _fields_ = [
        ("iter", C.c_int32), ("reason", C.c_int32), ("nonpos_iter", C.c_int32), ("pad_", C.c_int32),
        ("nonpos_pAp", C.c_double), ("ry", C.c_double), ("pAp", C.c_double), ("alpha", C.c_double),
        ("beta", C.c_double), ("rnorm", C.c_double), ("m", C.c_double), ("res_bound", C.c_double),
    ]


class LayerDesc(C.Structure):
    """Mirror of ``hf_layer_desc``."""

    _fields_ = [
        ("in_features", C.c_int32), ("out_features", C.c_int32), ("act", C.c_int32), ("has_bias", C.c_int32),
        ("w_offset", C.c_int64), ("b_offset", C.c_int64), ("d_w_frozen", C.c_void_p), ("d_b_frozen", C.c_void_p),
        ("kind", C.c_int32), ("c_in", C.c_int32), ("h_in", C.c_int32), ("w_in", C.c_int32),
        ("k_h", C.c_int32), ("k_w", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
        ("h_out", C.c_int32), ("w_out", C.c_int32),
    ]


class Operand(C.Structure):
    """Mirror of ``hf_operand``."""

    _fields_ = [("d_ptr", C.c_void_p), ("stride_mn", C.c_int64), ("stride_k", C.c_int64)]


_vp, _i32, _i64, _dbl, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_size_t

# name -> (restype, argtypes); every symbol of include/hf_b200.h appears here (tests check that)
SIGNATURES = {
    "hf_abi_version": (C.c_int, []),
    "hf_last_error_string": (C.c_char_p, []),
    "hf_device_sm_count": (C.c_int, []),
    "hf_debug_launch_count": (C.c_longlong, []),
    "hf_debug_pcg_trace": (C.c_int, [_vp]),
    "hf_debug_tc_trace": (C.c_int, [_vp]),
    "hf_debug_tc2_trace": (C.c_int, [_vp]),
    "hf_debug_tc_trace_iters": (C.c_int, [_vp]),
    "hf_pcg_state_bytes": (_sz, [_i64]),
    "hf_pcg_m_iters_offset": (_sz, []),
    "hf_pcg_set_progress": (C.c_int, [_vp, _vp, _vp]),
    "hf_pcg_init": (C.c_int, [C.c_int, _i64, _vp, _sz, _vp, _vp, _vp, _vp, _dbl, _dbl, _dbl, _i64, C.c_int, C.c_int,
                              _vp, _vp, _vp, _vp]),
    "hf_pcg_iter": (C.c_int, [C.c_int, _i64, _vp, C.c_int, _vp, _vp, _vp, _vp, _dbl, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hf_precond_power": (C.c_int, [C.c_int, _i64, _vp, _dbl, _dbl, _vp, _vp]),
    "hf_axpy_out": (C.c_int, [C.c_int, _i64, _vp, _dbl, _vp, _vp, _vp]),
    "hf_net_create": (C.c_int, [C.POINTER(LayerDesc), _i32, _i32, _i32, _i64, C.POINTER(_vp)]),
    "hf_net_destroy": (None, [_vp]),
    "hf_net_set_engine": (C.c_int, [_vp, _i32]),
    "hf_lin_workspace_bytes": (_sz, [_vp, _i64, _i32]),
    "hf_lin_create": (C.c_int, [_vp, _i64, _i32, _vp, _sz, C.POINTER(_vp)]),
    "hf_lin_destroy": (None, [_vp]),
    "hf_lin_forward": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "hf_lin_gradient": (C.c_int, [_vp, _vp, _vp, _i32, _vp]),
    "hf_ggn_matvec": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "hf_hessian_matvec": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "hf_fisher_diag": (C.c_int, [_vp, _vp, _vp, _i32, _vp]),
    "hf_matvec_phase": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _i32]),
    "hf_net_first_layer_span": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "hf_lin_logits": (_vp, [_vp]),
    "hf_allreduce_multimem": (C.c_int, [_vp, _vp, _i32, _i32, _i64, _i64, _i32, _vp, _vp]),
    "hf_debug_allreduce_variant": (None, [_i32]),
    "hf_contract_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32]),
    "hf_contract": (C.c_int, [_i32, _i64, _i64, _i64, _i32, C.POINTER(Operand), C.POINTER(Operand), _vp, _i64, _vp, _sz,
                              _vp]),
}

_ERRORS = {-1: ValueError, -2: NotImplementedError, -3: ValueError, -4: RuntimeError}
_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("HF_B200_LIB", LIB_PATH)  # override: A/B builds of the same ABI
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: the sm_100a kernels are not built. Run "
            "`python -m pytorchhessianfree_b200.build` (there is no CPU or PyTorch fallback)."
        )
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.hf_abi_version() != ABI_VERSION:
        raise RuntimeError(f"{path}: ABI version {lib.hf_abi_version()} != {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc):
    """Turn a negative ``hf_status`` into the exception type the reference would raise."""
    if rc != 0:
        msg = load().hf_last_error_string().decode(errors="replace")
        raise _ERRORS.get(rc, RuntimeError)(f"hf_b200: {msg}")


def dtype_code(t):
    if t.dtype == torch.float32:
        return HF_F32
    if t.dtype == torch.float64:
        return HF_F64
    raise TypeError(f"hf_b200 kernels take float32 or float64 vectors, got {t.dtype}")


def require_cuda(t, what):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{what} must be a torch.Tensor, not {type(t)}")
    if not t.is_cuda:
        raise RuntimeError(
            f"{what} lives on {t.device}: pytorchhessianfree_b200 runs on a CUDA (sm_100a) device only; "
            "there is no CPU path"
        )


def vec(t):
    """Contiguous, 16-byte aligned view or copy of a 1-D tensor (what the fused kernels need)."""
    if t.dim() != 1:
        t = t.reshape(-1)
    if not t.is_contiguous() or t.data_ptr() % 16:
        t = t.clone(memory_format=torch.contiguous_format)
        if t.data_ptr() % 16:  # cannot happen with the caching allocator; be loud if it does
            raise RuntimeError("could not obtain a 16-byte aligned copy")
    return t


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream
