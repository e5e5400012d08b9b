"""B200-native Hessian-free inner solve: drop-in for ltatzel/PyTorchHessianFree's hot path.

    from pytorchhessianfree_b200 import HessianFree        # reference: hessianfree.optimizer.HessianFree
    from pytorchhessianfree_b200.cg import cg              # reference: hessianfree.cg.cg
    from pytorchhessianfree_b200.preconditioners import diag_EF_preconditioner

Everything numeric runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/hf_b200.h`` (``csrc/libhf_b200.so``).  There is no CPU or PyTorch fallback: importing the
package is cheap, the first kernel call loads the library and fails loudly if it is not built.
"""
from .cg import DiagonalPreconditioner, cg, cg_storing_grid, pcg_device  # noqa: F401
from .cg_backtracking import cg_backtracking, cg_efficient_backtracking  # noqa: F401
from .linesearch import simple_linesearch  # noqa: F401
from .optimizer import HessianFree  # noqa: F401
from .preconditioners import (  # noqa: F401
    diag_EF_autograd, diag_EF_backpack, diag_EF_preconditioner, diag_to_preconditioner)
from .utils import vector_to_parameter_list, vector_to_trainparams  # noqa: F401

__all__ = ["HessianFree", "cg", "pcg_device", "cg_storing_grid", "DiagonalPreconditioner", "cg_backtracking",
           "cg_efficient_backtracking", "simple_linesearch", "diag_EF_autograd", "diag_EF_backpack",
           "diag_EF_preconditioner", "diag_to_preconditioner", "vector_to_parameter_list", "vector_to_trainparams"]
