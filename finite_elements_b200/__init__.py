"""finite_elements_b200 -- B200 (sm_100a) implementation of the hot path of
Dessia-tech/finite_elements behind the reference's own class surface.

    import finite_elements_b200 as fe
    fe.analysis.FiniteElementAnalysis(mesh, ...).solve()    # -> fe.results.Result

The compute path is libfe_b200.so (hand-written CUDA, C ABI in include/fe_b200.h).  Importing
the package loads it; there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError when libfe_b200.so has not been built)
from . import core, mesh, elements, loads, conditions, results, analysis  # noqa: F401

__version__ = "0.1.0"
