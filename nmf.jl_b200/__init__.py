"""nmf.jl_b200 -- B200-native accelerator for the iteration hot path of JuliaStats/NMF.jl.

The directory name contains a dot, so import it through the shim module at the repo root:
    import nmf_jl_b200 as NMF
    NMF.nnmf(X, k, alg="multmse", init="random")
"""
from .api import (  # noqa: F401
    ALSPGrad,
    ArgumentError,
    CoordinateDescent,
    DimensionMismatch,
    GreedyCD,
    MultUpdate,
    NmfB200Error,
    NumericalError,
    ProjectedALS,
    Result,
    Session,
    SPA,
    nndsvd,
    nnmf,
    randinit,
    rsvd,
    solve,
    solve_replicates,
)
from . import _lib, build, dist  # noqa: F401

__version__ = "0.1.0"
