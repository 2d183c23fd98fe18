"""Import the UNMODIFIED reference package: /root/reference (this container) or baseline/_ref (the offline pip install
`python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`, git-ignored,
travels to the GPU box with the snapshot; used by `bench.py --impl reference` and the cpu_baseline leg only).

TEST INFRASTRUCTURE.  The reference imports turtle, matplotlib.pyplot and tikzplotlib at module level
(reference raocp/core/scenario_tree.py:4, raocp/core/solver.py:8-9); none is installed here and none is used by
the hot path, so empty stub modules are injected for those names before the import.  The GPU box has no
/root/reference; tests marked `reference` skip there, the bench takes baseline/_ref.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _has(root):
    return bool(root) and os.path.isfile(os.path.join(root, "raocp", "core", "solver.py"))


def _find():
    for root in (os.environ.get("RAOCP_REFERENCE_ROOT"), "/root/reference",
                 os.path.join(os.path.dirname(_HERE), "baseline", "_ref")):
        if _has(root):
            return root
    return os.environ.get("RAOCP_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find()


def available():
    return _has(REFERENCE_ROOT)


def source():
    """where the reference was found: 'source tree' (/root/reference) or 'baseline/_ref' (pip --target install)"""
    return "baseline/_ref (pip --target install of the unmodified reference)" if "baseline" in REFERENCE_ROOT \
        else f"{REFERENCE_ROOT} (source tree)"


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
