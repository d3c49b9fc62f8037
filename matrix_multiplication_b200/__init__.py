"""matrix_multiplication_b200 — B200-native (sm_100a) fp32 CSR SpMM engine behind the surface of
smoorjani/matrix-multiplication: the `custom_mm` extension, the `matmuls` autograd wrappers.

  matrix_multiplication_b200.capi        ctypes binding of the C ABI (include/spmm_b200.h)
  matrix_multiplication_b200.custom_mm   pybind11 module with the reference's entry points
  matrix_multiplication_b200.matmuls     cusparseMM / naiveSpMM / cublasMM ... wrappers
  matrix_multiplication_b200.layers      cusparseLinear (persistent CSR weight + cached transpose)
  matrix_multiplication_b200.sharding    nnz-balanced row shards for multi-GPU runs

The native libraries are built in-tree by `python -m matrix_multiplication_b200.build`
(`__graft_entry__.build()`); nothing here falls back to a CPU or torch implementation.
"""
from . import _capi as capi  # noqa: F401  (light: ctypes + numpy only; the library loads on first use)

__version__ = "0.1.0"


def __getattr__(name):
    # custom_mm / matmuls / layers import torch and the pybind module; load them on demand
    if name in ("custom_mm", "matmuls", "layers", "sharding", "synthetic"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
