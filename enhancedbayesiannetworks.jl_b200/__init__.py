"""B200-native Monte Carlo evaluation of EnhancedBayesianNetworks.jl functional nodes.

The directory name carries a dot, so it is imported through the ``ebn_b200`` shim
at the repository root (``import ebn_b200``).

Layout
    csrc/         hand-written sm_100a kernels + the C ABI (libebncuda.so)
    _lib.py       ctypes binding of include/ebncuda.h
    expr.py       model formulas as text (Julia or Python syntax) -> expression program of the C ABI
    workloads.py  the BASELINE.json configurations as C-ABI jobs
    julia/        the `ccall` wrapper a maintainer drops into the reference
    distributions.py, nodes.py, network.py, uq.py, evaluation.py
                  host-side mirror of the reference's node / network API for this path
                  (DiscreteNode, ContinuousNode, functional nodes, EnhancedBayesianNetwork,
                  evaluate, reduce, dispatch) so the parity tests read like the reference's
"""
from . import _lib  # noqa: F401
from ._lib import (EbnError, Job, INPUT_DTYPE, make_inputs, pf_batch, pf_batch_resumable, hist_batch, eval_matrix,  # noqa: F401
                   sample_matrix, peak_fp64, last_stats, init, shutdown, device_count)
from .distributions import (Normal, LogNormal, Uniform, Gumbel, Gamma, Exponential, Truncated, truncated,  # noqa: F401
                            Tabulated, EmpiricalDistribution, distribution_parameters, Parameter,
                            RandomVariable, Interval, ProbabilityBox, GaussianCopula, JointDistribution)
from .nodes import (DiscreteConditionalProbabilityTable, ContinuousConditionalProbabilityTable,  # noqa: F401
                    ExactDiscretization, ApproximatedDiscretization, UnamedProbabilityBox, DiscreteNode,
                    ContinuousNode, DiscreteFunctionalNode, ContinuousFunctionalNode, Model, Poly, Catalogue, Expression,
                    CudaMonteCarlo, MonteCarlo, CudaSobolSampling, SobolSampling, CudaDoubleLoop, DoubleLoop, CudaRandomSlicing, RandomSlicing, FORM)
from .network import (EnhancedBayesianNetwork, BayesianNetwork, CredalNetwork, add_child, order, parents,  # noqa: F401
                      children, discrete_ancestors, markov_blanket, markov_envelope)
from .uq import probability_of_failure, UnsupportedModel  # noqa: F401
from .evaluation import evaluate, reduce, dispatch, evaluate_with_envelopes  # noqa: F401
