"""lmcma_path_planner_b200 — B200-native LM-CMA trajectory optimiser (hot path of
behnamasadi/lmcma_path_planner) behind a C ABI (include/lmcma_b200.h).

The package holds only what the hot path needs: csrc/ (CUDA kernels + the C ABI + the C++ facade),
the ctypes binding, the Python mirror of the reference-facing interface, and workload generators."""
from . import _capi  # noqa: F401
from .optimizer import LMCMA, LONGSAFE, SHORTRISKY, CostMap, Optimizer  # noqa: F401
from . import maps  # noqa: F401

__all__ = ["LMCMA", "Optimizer", "CostMap", "LONGSAFE", "SHORTRISKY", "maps"]
