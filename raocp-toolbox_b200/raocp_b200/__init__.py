"""raocp_b200 -- B200-native drop-in for the Chambolle-Pock hot path of raocp-toolbox.

Usage mirrors the reference (`/root/reference/raocp/__init__.py:1`)::

    import raocp_b200 as r
    tree = r.core.MarkovChainScenarioTreeFactory(P, v, N, tau).create()
    problem = r.core.RAOCP(tree).with_markovian_dynamics(...) ...
    solver = r.core.Solver(problem)
    status = solver.chock(x0, max_iters=2000, tol=1e-3)

All arithmetic of Cache / Operator / Solver / cone projections runs in hand-written
sm_100a CUDA kernels behind the C-ABI declared in include/raocp_b200.h.  There is no
CPU fallback: importing works anywhere, computing needs the built library and a GPU.
"""
from . import core  # noqa: F401

__version__ = "0.1.0"
