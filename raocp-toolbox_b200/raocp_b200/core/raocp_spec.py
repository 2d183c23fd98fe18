"""Problem container and fluent builder (API parity with reference raocp/core/raocp_spec.py:6-198).

Per-node lists of dynamics (indexed by the CHILD node), nonleaf costs (child), leaf costs (leaf), constraints
(nonleaf / leaf node) and risks (nonleaf node).  The reference deep-copies one object per node
(raocp_spec.py:127,139,147,161,171,182); here nodes share the per-mode objects (they are read-only to the solver)
so that building a 10^5-node problem is O(n) pointer stores, and the flattening pass (flatten.py) can
de-duplicate the matrices by object identity before uploading them.
"""
import copy

from . import constraints as core_constraints
from . import scenario_tree as core_tree


class RAOCP:
    def __init__(self, scenario_tree: core_tree.ScenarioTree):
        self._tree = scenario_tree
        n = self._tree.num_nodes
        m = self._tree.num_nonleaf_nodes
        self._num_nodes = n
        self._num_nonleaf_nodes = m
        self._dynamics = [None] * n
        self._nonleaf_costs = [None] * n
        self._leaf_costs = [None] * n
        no_constraint = core_constraints.No()
        self._nonleaf_constraints = [no_constraint] * m + [None] * (n - m)
        self._leaf_constraints = [None] * m + [no_constraint] * (n - m)
        self._risks = [None] * m

    # -- accessors -----------------------------------------------------------------------------------------------
    @property
    def tree(self):
        return self._tree

    @property
    def list_of_dynamics(self):
        return self._dynamics

    @property
    def list_of_nonleaf_costs(self):
        return self._nonleaf_costs

    @property
    def list_of_leaf_costs(self):
        return self._leaf_costs

    @property
    def list_of_nonleaf_constraints(self):
        return self._nonleaf_constraints

    @property
    def list_of_leaf_constraints(self):
        return self._leaf_constraints

    @property
    def list_of_risks(self):
        return self._risks

    def state_dynamics_at_node(self, idx):
        return self._dynamics[idx].state_dynamics

    def control_dynamics_at_node(self, idx):
        return self._dynamics[idx].control_dynamics

    def nonleaf_cost_at_node(self, idx):
        return self._nonleaf_costs[idx]

    def leaf_cost_at_node(self, idx):
        return self._leaf_costs[idx]

    def nonleaf_constraint_at_node(self, idx):
        return self._nonleaf_constraints[idx]

    def leaf_constraint_at_node(self, idx):
        return self._leaf_constraints[idx]

    def risk_at_node(self, idx):
        return self._risks[idx]

    # -- builder -------------------------------------------------------------------------------------------------
    def _require_markovian(self, what):
        if not self._tree.is_markovian:
            raise TypeError(f"{what} provided as Markovian, scenario tree provided is not Markovian")

    def _modes(self):
        return self._tree.values_array[1:].tolist()

    def with_markovian_dynamics(self, ordered_list_of_dynamics):
        first = ordered_list_of_dynamics[0]
        for dyn in ordered_list_of_dynamics:
            if dyn.state_dynamics.shape != first.state_dynamics.shape:
                raise ValueError("Markovian state dynamics matrices are different shapes")
            if dyn.control_dynamics.shape != first.control_dynamics.shape:
                raise ValueError("Markovian control dynamics matrices are different shapes")
        self._require_markovian("dynamics")
        self._dynamics[1:] = [ordered_list_of_dynamics[w] for w in self._modes()]
        return self

    def with_markovian_nonleaf_costs(self, ordered_list_of_costs):
        for cost in ordered_list_of_costs:
            if not cost.node_type.is_nonleaf:
                raise Exception("Markovian costs provided are not nonleaf")
        self._require_markovian("costs")
        self._nonleaf_costs[1:] = [ordered_list_of_costs[w] for w in self._modes()]
        return self

    def with_all_nonleaf_costs(self, cost):
        if not cost.node_type.is_nonleaf:
            raise Exception("Nonleaf cost provided is not nonleaf")
        self._nonleaf_costs[1:] = [cost] * (self._num_nodes - 1)
        return self

    def with_all_leaf_costs(self, cost):
        if not cost.node_type.is_leaf:
            raise Exception("Leaf cost provided is not leaf")
        m = self._num_nonleaf_nodes
        self._leaf_costs[m:] = [cost] * (self._num_nodes - m)
        return self

    # -- per-node data (beyond the reference's builder, SURVEY 8f rank 3: non-Markovian dynamics / node-wise costs and risks).  The
    #    device path needs nothing new: dynamics, costs and risk levels are tables indexed per node, de-duplicated by identity --
    def with_nodewise_dynamics(self, dynamics_of_nodes):
        """one Dynamics per node 1 .. n-1 (entry j-1 -> the edge into node j); shapes must agree"""
        dyn = list(dynamics_of_nodes)
        if len(dyn) != self._num_nodes - 1:
            raise ValueError(f"expected {self._num_nodes - 1} dynamics (nodes 1 .. n-1), got {len(dyn)}")
        for d in dyn:
            if d.state_dynamics.shape != dyn[0].state_dynamics.shape or d.control_dynamics.shape != dyn[0].control_dynamics.shape:
                raise ValueError("node-wise dynamics matrices are different shapes")
        self._dynamics[1:] = dyn
        return self

    def with_nodewise_nonleaf_costs(self, costs_of_nodes):
        """one nonleaf cost per node 1 .. n-1 (indexed by the child, like the Markovian builder)"""
        costs = list(costs_of_nodes)
        if len(costs) != self._num_nodes - 1:
            raise ValueError(f"expected {self._num_nodes - 1} nonleaf costs (nodes 1 .. n-1), got {len(costs)}")
        for cost in costs:
            if not cost.node_type.is_nonleaf:
                raise Exception("Nonleaf cost provided is not nonleaf")
        self._nonleaf_costs[1:] = costs
        return self

    def with_nodewise_leaf_costs(self, costs_of_leaves):
        costs = list(costs_of_leaves)
        m = self._num_nonleaf_nodes
        if len(costs) != self._num_nodes - m:
            raise ValueError(f"expected {self._num_nodes - m} leaf costs, got {len(costs)}")
        for cost in costs:
            if not cost.node_type.is_leaf:
                raise Exception("Leaf cost provided is not leaf")
        self._leaf_costs[m:] = costs
        return self

    def with_nodewise_risks(self, risks_of_nonleaf_nodes):
        """one risk measure per nonleaf node (e.g. AVaR with a node-dependent level)"""
        risks = list(risks_of_nonleaf_nodes)
        if len(risks) != self._num_nonleaf_nodes:
            raise ValueError(f"expected {self._num_nonleaf_nodes} risks, got {len(risks)}")
        out = []
        for i, risk in enumerate(risks):
            if not risk.is_risk:
                raise Exception("Risk provided is not of risk type")
            risk_i = copy.copy(risk)
            risk_i.probs = self._tree.conditional_probabilities_of_children(i)
            out.append(risk_i)
        self._risks = out
        return self

    def _check_dynamics_before_constraints(self):
        if self._num_nodes < 2 or self._dynamics[1] is None:
            raise Exception("Constraints provided before dynamics - dynamics must be provided first")

    def with_all_nonleaf_constraints(self, nonleaf_constraint):
        self._check_dynamics_before_constraints()
        if not nonleaf_constraint.node_type.is_nonleaf:
            raise Exception("Nonleaf constraint provided is not nonleaf")
        nonleaf_constraint.state_size = self._dynamics[-1].state_dynamics.shape[1]
        nonleaf_constraint.control_size = self._dynamics[-1].control_dynamics.shape[1]
        shared = copy.deepcopy(nonleaf_constraint)
        self._nonleaf_constraints[: self._num_nonleaf_nodes] = [shared] * self._num_nonleaf_nodes
        return self

    def with_all_leaf_constraints(self, leaf_constraint):
        self._check_dynamics_before_constraints()
        if not leaf_constraint.node_type.is_leaf:
            raise Exception("Leaf constraint provided is not leaf")
        leaf_constraint.state_size = self._dynamics[-1].state_dynamics.shape[1]
        m = self._num_nonleaf_nodes
        shared = copy.deepcopy(leaf_constraint)
        self._leaf_constraints[m:] = [shared] * (self._num_nodes - m)
        return self

    def with_all_risks(self, risk):
        if not risk.is_risk:
            raise Exception("Risk provided is not of risk type")
        tree = self._tree
        risks = []
        for i in range(self._num_nonleaf_nodes):
            risk_i = copy.copy(risk)  # shallow: (E, F, cone, b) are rebuilt lazily from the node's own probabilities
            risk_i.probs = tree.conditional_probabilities_of_children(i)
            risks.append(risk_i)
        self._risks = risks
        return self

    def __str__(self):
        return f"RAOCP\n+ Nodes: {self._tree.num_nodes}\n" \
               f"+ {self._nonleaf_costs[0]}\n" \
               f"+ {self._risks[0]}"

    def __repr__(self):
        return f"RAOCP with {self._tree.num_nodes} nodes, " \
               f"with root cost: {type(self._nonleaf_costs[0]).__name__}, " \
               f"with root risk: {type(self._risks[0]).__name__}."
