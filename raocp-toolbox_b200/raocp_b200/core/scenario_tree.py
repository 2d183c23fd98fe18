"""Scenario tree container and stopped-Markov-chain factory.

API parity with reference raocp/core/scenario_tree.py:22-163 (ScenarioTree) and :243-351
(MarkovChainScenarioTreeFactory); the turtle "bulls-eye" plot (:165-240) is out of scope.

The node numbering is the reference's (breadth first: stage by stage, children of a node contiguous and in the order
of the non-zero columns of the transition row), because the device layout relies on it (stage ranges and children
ranges are contiguous; tests/test_scenario_tree.py:88-98 of the reference pins the stage ranges).  Unlike the
reference (O(n^2) np.concatenate / np.where loops) everything here is built with O(n) vectorised NumPy so that
10^5-node trees take milliseconds.
"""
import numpy as np


def _check_probability_vector(p):
    if abs(sum(p) - 1) >= 1e-10:
        raise ValueError("probability vector does not sum up to 1")
    if any(pi <= -1e-16 for pi in p):
        raise ValueError("probability vector contains negative entries")
    return True


def _check_stopping_time(n, t):
    if t > n:
        raise ValueError("stopping time greater than number of stages")
    return True


class ScenarioTree:
    def __init__(self, stages, ancestors, probability, w_values=None, is_markovian=False):
        """
        :param stages: array, stage of every node (node 0 is the root at stage 0)
        :param ancestors: array, parent of every node (-1 for the root)
        :param probability: array, probability of visiting every node
        :param w_values: array, value of the disturbance w at every node
        """
        self._is_markovian = is_markovian
        self._stages = np.asarray(stages)
        self._ancestors = np.asarray(ancestors)
        self._probability = np.asarray(probability)
        self._w = w_values
        self._num_nodes = len(self._ancestors)
        self._num_stages = int(self._stages[-1]) + 1
        self._num_nonleaf = int(np.sum(self._stages < (self._num_stages - 1)))
        # children in CSR form: a stable sort of the node ids by parent keeps each child list increasing
        anc = self._ancestors[1:]
        order = np.argsort(anc, kind="stable") + 1
        counts = np.bincount(anc, minlength=self._num_nodes)[: self._num_nonleaf]
        self._child_ptr = np.concatenate(([0], np.cumsum(counts)))
        self._child_idx = order
        self._data = np.empty(shape=(self._num_nodes,), dtype=dict)

    # -- data slots ----------------------------------------------------------------------------------------------
    def get_data_at_node(self, node_idx):
        return self._data[node_idx]

    def set_data_at_node(self, node_idx, data_dict: dict):
        self._data[node_idx] = data_dict

    # -- sizes ---------------------------------------------------------------------------------------------------
    @property
    def is_markovian(self):
        return self._is_markovian

    @property
    def num_nonleaf_nodes(self):
        return self._num_nonleaf

    @property
    def num_nodes(self):
        return self._num_nodes

    @property
    def num_stages(self):
        """number of stages including stage zero"""
        return self._num_stages

    # -- queries -------------------------------------------------------------------------------------------------
    def ancestor_of(self, node_idx):
        return self._ancestors[node_idx]

    def children_of(self, node_idx):
        return self._child_idx[self._child_ptr[node_idx]: self._child_ptr[node_idx + 1]]

    def stage_of(self, node_idx):
        if node_idx < 0:
            raise ValueError("node_idx cannot be <0")
        return self._stages[node_idx]

    def value_at_node(self, node_idx):
        return self._w[node_idx]

    def nodes_at_stage(self, stage_idx):
        return np.where(self._stages == stage_idx)[0]

    def probability_of_node(self, node_idx):
        return self._probability[node_idx]

    def siblings_of_node(self, node_idx):
        if node_idx == 0:
            return [0]
        return self.children_of(self.ancestor_of(node_idx))

    def conditional_probabilities_of_children(self, node_idx):
        return self._probability[self.children_of(node_idx)] / self._probability[node_idx]

    # -- whole-array views used by the flattening pass (raocp_b200 extension, not in the reference) ----------------
    @property
    def stages_array(self):
        return self._stages

    @property
    def ancestors_array(self):
        return self._ancestors

    @property
    def probability_array(self):
        return self._probability

    @property
    def values_array(self):
        return self._w

    def bulls_eye_plot(self, *args, **kwargs):
        raise NotImplementedError("plotting is outside the scope of raocp_b200 (reference scenario_tree.py:165-240)")

    def __str__(self):
        return f"Scenario Tree\n+ Nodes: {self.num_nodes}\n+ Stages: {self.num_stages}\n" \
               f"+ Scenarios: {len(self.nodes_at_stage(self.num_stages - 1))}\n" \
               f"+ Data: {self._data is not None}"

    def __repr__(self):
        return f"Scenario tree with {self.num_nodes} nodes, {self.num_stages} stages " \
               f"and {len(self.nodes_at_stage(self.num_stages - 1))} scenarios"


class MarkovChainScenarioTreeFactory:
    """Scenario tree of a Markov chain that stops branching at `stopping_time`."""

    def __init__(self, transition_prob, initial_distribution, num_stages, stopping_time=None):
        if stopping_time is None:
            stopping_time = num_stages
        else:
            _check_stopping_time(num_stages, stopping_time)
        self._p = np.asarray(transition_prob)
        self._v = np.asarray(initial_distribution)
        self._horizon = num_stages
        self._tau = stopping_time
        for row in self._p:
            _check_probability_vector(row)
        _check_probability_vector(self._v)

    def create(self):
        p, v = self._p, self._v
        num_modes = p.shape[0]
        # cover[w] = modes reachable from w, padded into a rectangular table
        cover_len = np.count_nonzero(p, axis=1)
        cover_pad = np.zeros((num_modes, max(int(cover_len.max()), 1)), dtype=int)
        for w in range(num_modes):
            nz = np.flatnonzero(p[w])
            cover_pad[w, : nz.size] = nz

        first = np.flatnonzero(v > 0)
        anc_parts = [np.array([-1]), np.zeros(first.size, dtype=int)]
        val_parts = [np.array([-1]), first]
        stg_parts = [np.array([0]), np.ones(first.size, dtype=int)]
        prb_parts = [np.array([1.0]), v[first].astype(float)]

        start = 1  # id of the first node of the newest stage
        vals, probs = first, v[first].astype(float)
        for stage in range(1, self._tau):  # branching stages: newest stage is `stage`, we add stage+1
            ids = np.arange(start, start + vals.size)
            cnt = cover_len[vals]
            parent = np.repeat(ids, cnt)
            parent_val = np.repeat(vals, cnt)
            within = np.arange(parent.size) - np.repeat(np.cumsum(cnt) - cnt, cnt)
            child_val = cover_pad[parent_val, within]
            child_prob = np.repeat(probs, cnt) * p[parent_val, child_val]
            anc_parts.append(parent)
            val_parts.append(child_val)
            stg_parts.append(np.full(parent.size, stage + 1, dtype=int))
            prb_parts.append(child_prob)
            start += vals.size
            vals, probs = child_val, child_prob
        for stage in range(self._tau, self._horizon):  # after the stopping time: one child, same mode, same prob
            ids = np.arange(start, start + vals.size)
            anc_parts.append(ids)
            val_parts.append(vals)
            stg_parts.append(np.full(vals.size, stage + 1, dtype=int))
            prb_parts.append(probs)
            start += vals.size

        ancestors = np.concatenate(anc_parts)
        values = np.concatenate(val_parts)
        stages = np.concatenate(stg_parts)
        probability = np.concatenate(prb_parts)
        return ScenarioTree(stages, ancestors, probability, values, is_markovian=True)
