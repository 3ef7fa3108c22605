"""B200-native Monte Carlo evaluation of EnhancedBayesianNetworks.jl functional nodes.

The directory name carries a dot, so it is imported through the ``ebn_b200`` shim
at the repository root (``import ebn_b200``).

Layout
    csrc/         hand-written sm_100a kernels + the C ABI (libebncuda.so)
    _lib.py       ctypes binding of include/ebncuda.h
    workloads.py  the BASELINE.json configurations as C-ABI jobs
    julia/        the `ccall` wrapper a maintainer drops into the reference
"""
from . import _lib  # noqa: F401
from ._lib import (EbnError, Job, INPUT_DTYPE, make_inputs, pf_batch, hist_batch, eval_matrix,  # noqa: F401
                   sample_matrix, peak_fp64, last_stats, init, shutdown, device_count)
