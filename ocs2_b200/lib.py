"""ctypes binding of include/ocs2_ddp_cuda.h. Fails loudly when the CUDA library is missing (no fallback of any kind)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libocs2_ddp_cuda.so"

ALG_ILQR, ALG_SLQ = 0, 1
FORM_FULL, FORM_REDUCED = 0, 1
STRATEGY_LINE_SEARCH, STRATEGY_LEVENBERG_MARQUARDT = 0, 1
HC_DIAGONAL_SHIFT, HC_CHOLESKY_MODIFICATION, HC_EIGENVALUE_MODIFICATION, HC_GERSHGORIN_MODIFICATION = 0, 1, 2, 3
STATUS_CHOL_NOT_PD, STATUS_NONFINITE, STATUS_CONSTRAINT_RANK, STATUS_NOT_PSD = 1, 2, 4, 8
LQ_SYMMETRIC_PACKED = 1

_ERR_NAMES = {1: "INVALID_ARGUMENT", 2: "UNSUPPORTED", 3: "CUDA", 4: "OUT_OF_MEMORY", 5: "NOT_READY"}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class O2cError(RuntimeError):
    """Raised for every non-zero o2c_error (the reference throws std::runtime_error on this path)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"o2c error {code} ({_ERR_NAMES.get(code, '?')}): {message}")
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("nu", C.c_int32), ("nc_max", C.c_int32), ("num_stages", C.c_int32), ("batch", C.c_int32),
        ("algorithm", C.c_int32), ("riccati_form", C.c_int32), ("strategy", C.c_int32), ("hessian_correction", C.c_int32),
        ("device", C.c_int32), ("max_alphas", C.c_int32), ("has_nominal", C.c_int32),
        ("hessian_multiple", C.c_double), ("lm_riccati_multiple", C.c_double), ("time_step", C.c_double),
    ]


class LineSearchSettings(C.Structure):
    _fields_ = [("min_step_length", C.c_double), ("max_step_length", C.c_double), ("contraction_rate", C.c_double),
                ("armijo_coefficient", C.c_double)]


class Field(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("problem_stride", C.c_int64), ("node_stride", C.c_int64)]


class LqView(C.Structure):
    _fields_ = [(n, Field) for n in ("A", "B", "Hv", "Q", "P", "R", "q", "r", "c", "C", "D", "e")] + [
        ("nc", C.c_void_p), ("nc_problem_stride", C.c_int64), ("nc_node_stride", C.c_int64)] + [
        (n, Field) for n in ("Qf", "qf", "cf", "x_nom", "u_nom", "x0")] + [("time", C.c_void_p)] + [
        ("event", C.c_void_p), ("event_problem_stride", C.c_int64), ("event_node_stride", C.c_int64)] + [
        (n, Field) for n in ("jump_A", "jump_Hv", "jump_Q", "jump_q", "jump_c")] + [("flags", C.c_int32)]


class DiscretizationView(C.Structure):
    _fields_ = [("dfdx", Field * 4), ("dfdu", Field * 4), ("dt", C.c_void_p), ("stages", C.c_int32)]


class SolutionView(C.Structure):
    _fields_ = [(n, Field) for n in ("K", "dbias", "bias", "Sm", "Sv", "s", "x", "u")] + [
        ("x_alpha_stride", C.c_int64), ("u_alpha_stride", C.c_int64), ("status", C.c_void_p)]


EXPORTED_SYMBOLS = [
    "o2c_abi_version", "o2c_last_error", "o2c_create", "o2c_destroy", "o2c_get_config", "o2c_sync", "o2c_compute_stream",
    "o2c_device_lq_view", "o2c_device_solution_view", "o2c_rollout_num_nodes", "o2c_rollout_times", "o2c_upload", "o2c_import_device",
    "o2c_download", "o2c_set_time", "o2c_backward", "o2c_rollout", "o2c_solve", "o2c_launch_count", "o2c_kernel_variant",
    "o2c_solve_host", "o2c_generate_synthetic", "o2c_host_alloc", "o2c_host_free", "o2c_line_search", "o2c_line_search_result",
    "o2c_download_flattened_controller", "o2c_discretize", "o2c_check_numerical_stability",
    "o2c_device_count", "o2c_set_lm_riccati_multiple",
]


def library_path() -> str:
    return os.path.join(_HERE, _LIB_NAME)


_lib = None


def load_library():
    """Loads the in-tree libocs2_ddp_cuda.so. Raises if it has not been built: there is no CPU path to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -m ocs2_b200.build` (needs nvcc); ocs2_b200 has no CPU fallback")
    lib = C.CDLL(path)
    lib.o2c_abi_version.restype = C.c_int
    lib.o2c_last_error.restype = C.c_char_p
    lib.o2c_kernel_variant.restype = C.c_char_p
    lib.o2c_kernel_variant.argtypes = [C.c_void_p]
    hp = C.c_void_p
    sig = {
        "o2c_create": [C.POINTER(Config), C.POINTER(hp)],
        "o2c_destroy": [hp],
        "o2c_get_config": [hp, C.POINTER(Config)],
        "o2c_sync": [hp],
        "o2c_compute_stream": [hp, C.POINTER(C.c_void_p)],
        "o2c_device_lq_view": [hp, C.POINTER(LqView)],
        "o2c_device_solution_view": [hp, C.POINTER(SolutionView)],
        "o2c_rollout_num_nodes": [hp, C.POINTER(C.c_int32)],
        "o2c_rollout_times": [hp, _dp],
        "o2c_upload": [hp, C.POINTER(LqView), C.c_int32, C.c_int32],
        "o2c_import_device": [hp, C.POINTER(LqView), C.c_int32, C.c_int32],
        "o2c_download": [hp, C.POINTER(SolutionView), C.c_int32, C.c_int32, C.c_int32],
        "o2c_set_time": [hp, _dp],
        "o2c_backward": [hp, C.c_int32, C.c_int32],
        "o2c_rollout": [hp, _dp, C.c_int32, C.c_int32, C.c_int32],
        "o2c_solve": [hp, C.c_double, C.c_int32, C.c_int32],
        "o2c_launch_count": [hp, C.POINTER(C.c_int64)],
        "o2c_solve_host": [hp, C.POINTER(LqView), C.POINTER(SolutionView), C.c_double, C.c_int32, C.c_int32],
        "o2c_generate_synthetic": [hp, C.c_uint64, C.c_int64, C.c_double],
        "o2c_host_alloc": [C.POINTER(C.c_void_p), C.c_uint64],
        "o2c_host_free": [C.c_void_p],
        "o2c_line_search": [hp, C.POINTER(LineSearchSettings), _dp, C.c_int32, C.c_int32],
        "o2c_line_search_result": [hp, _dp, _ip, _dp, _dp, _dp, _dp, _ip, C.c_int32, C.c_int32],
        "o2c_download_flattened_controller": [hp, C.POINTER(C.c_float), C.c_double, C.c_int32, C.c_int32],
        "o2c_discretize": [hp, C.POINTER(DiscretizationView), C.c_int32, C.c_int32, C.c_int32],
        "o2c_check_numerical_stability": [hp, C.c_int32, C.c_int32],
        "o2c_device_count": [_ip],
        "o2c_set_lm_riccati_multiple": [hp, C.c_double],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    if lib.o2c_abi_version() != 4:
        raise RuntimeError("libocs2_ddp_cuda.so ABI version mismatch")
    _lib = lib
    return lib


def check(code: int):
    if code != 0:
        raise O2cError(code, load_library().o2c_last_error().decode("utf-8", "replace"))
