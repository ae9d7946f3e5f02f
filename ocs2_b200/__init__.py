"""ocs2_b200 — B200-native batched LQ solver for the data-parallel core of OCS2's DDP inner loop.

The product is the C-ABI shared library ``libocs2_ddp_cuda.so`` (sources in ``ocs2_b200/csrc``, header ``include/ocs2_ddp_cuda.h``).
This package is the thin Python host side over that ABI (ctypes), mirroring the reference's names for the path:
``ddp.Settings``, ``solveSequentialRiccatiEquations``, ``calculateController``, ``rolloutTrajectory`` and the ``LinearController``
arrays. There is no CPU fallback: importing works anywhere, but creating a solver without the built library or without a CUDA
device raises.
"""
from .lib import (ALG_ILQR, ALG_SLQ, FORM_FULL, FORM_REDUCED, HC_CHOLESKY_MODIFICATION, HC_DIAGONAL_SHIFT, HC_EIGENVALUE_MODIFICATION,
                  HC_GERSHGORIN_MODIFICATION, LQ_SYMMETRIC_PACKED, STATUS_CHOL_NOT_PD, STATUS_CONSTRAINT_RANK, STATUS_NONFINITE,
                  STATUS_NOT_PSD, STRATEGY_LEVENBERG_MARQUARDT,
                  STRATEGY_LINE_SEARCH, O2cError, library_path, load_library)
from .solver import BatchedLqSolver, LinearController, LineSearchResult, LineSearchSettings, LqBatch, Settings, Solution, pack_upper

__all__ = [
    "ALG_ILQR", "ALG_SLQ", "FORM_FULL", "FORM_REDUCED", "HC_DIAGONAL_SHIFT", "HC_CHOLESKY_MODIFICATION", "HC_EIGENVALUE_MODIFICATION",
    "HC_GERSHGORIN_MODIFICATION", "STRATEGY_LINE_SEARCH", "STRATEGY_LEVENBERG_MARQUARDT", "STATUS_CHOL_NOT_PD", "STATUS_NONFINITE",
    "STATUS_CONSTRAINT_RANK", "STATUS_NOT_PSD", "LQ_SYMMETRIC_PACKED", "O2cError", "library_path", "load_library", "BatchedLqSolver", "LinearController", "LqBatch", "Settings",
    "Solution", "LineSearchSettings", "LineSearchResult", "pack_upper",
]
