"""Import the UNMODIFIED reference package from /root/reference (this container only).

TEST INFRASTRUCTURE.  The reference imports turtle, matplotlib.pyplot and tikzplotlib at module level
(reference raocp/core/scenario_tree.py:4, raocp/core/solver.py:8-9); none is installed here and none is used by
the hot path, so empty stub modules are injected for those names before the import.  The GPU box has no
/root/reference: `available()` is False there and every caller must skip.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RAOCP_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "raocp", "core", "solver.py"))


def load():
    """Returns the reference's top-level `raocp` module (with .core loaded)."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    for name in ("turtle", "matplotlib", "matplotlib.pyplot", "tikzplotlib"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import raocp  # noqa: E402  (the reference, not raocp_b200)
    import raocp.core.dynamics  # noqa: F401  (not star-imported by the reference's core/__init__)
    import raocp.core.constraints.rectangle  # noqa: F401
    return raocp


class RefApi:
    """Namespace adapter so oracle.problems.build() can target the reference's classes."""

    def __init__(self):
        r = load()
        import raocp.core.nodes as nodes
        import raocp.core.dynamics as dynamics
        import raocp.core.costs as costs
        import raocp.core.risks as risks
        import raocp.core.constraints.rectangle as rectangle
        self.MarkovChainScenarioTreeFactory = r.core.MarkovChainScenarioTreeFactory
        self.RAOCP = r.core.RAOCP
        self.Nonleaf, self.Leaf = nodes.Nonleaf, nodes.Leaf
        self.Dynamics = dynamics.Dynamics
        self.Quadratic = costs.Quadratic
        self.AVaR = risks.AVaR
        self.Rectangle = rectangle.Rectangle
        self.Solver = r.core.Solver
        self.Cache = r.core.Cache
        self.Operator = r.core.Operator
