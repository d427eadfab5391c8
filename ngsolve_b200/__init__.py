"""ngsolve_b200 -- NGSolve's assembled-system solve hot path (CSR SpMV + BaseVector updates and
reductions + Jacobi, i.e. one CGSolver / GMRes iteration) as hand-written CUDA for B200 (sm_100a)
behind a C ABI (include/ngsb200.h, ngsolve_b200/lib/libngsb200.so).

`ngsolve_b200.la` mirrors the part of `ngsolve.la` / `ngsolve.ngscuda` a solve script touches (ctypes, for boxes without
NGSolve; the deployed boundary is the compiled adapter integration/ngsb200_ngla.cpp); `ngsolve_b200.parallel` is the
one-GPU-per-process ParallelDofs / ParallelMatrix layer.  Nothing here computes on the CPU.
"""
from . import _capi  # noqa: F401
from ._capi import NgsbError  # noqa: F401
