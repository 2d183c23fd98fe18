"""Quadratic stage / terminal cost (API parity with reference raocp/core/costs.py:9-57).

The matrix square roots are host-side setup data: they are computed once here (scipy.linalg.sqrtm, as
the reference does at costs.py:21,26) and uploaded to the device when a Cache is created.
"""
from scipy.linalg import sqrtm


class Quadratic:
    def __init__(self, node_type, state_weights, control_weights=None):
        self._node_type = node_type
        if node_type.is_nonleaf and control_weights is None:
            raise Exception("No control weights provided for a nonleaf node")
        if node_type.is_leaf and control_weights is not None:
            raise Exception("Control weights provided for a leaf node")
        if state_weights.shape[0] != state_weights.shape[1]:
            raise Exception("Quadratic cost state weight matrix is not square")
        self._q = state_weights
        self._sqrt_q = sqrtm(state_weights)
        self._r = None
        self._sqrt_r = None
        if control_weights is not None and control_weights.shape[0] != control_weights.shape[1]:
            raise Exception("Quadratic cost control weight matrix is not square")
        if node_type.is_nonleaf:
            self._r = control_weights
            self._sqrt_r = sqrtm(control_weights)
        elif not node_type.is_leaf:
            raise Exception("Control weights error in cost")

    @property
    def node_type(self):
        return self._node_type

    @property
    def state_weights(self):
        return self._q

    @property
    def control_weights(self):
        return self._r

    @property
    def sqrt_state_weights(self):
        return self._sqrt_q

    @property
    def sqrt_control_weights(self):
        return self._sqrt_r

    def __repr__(self):
        return f"Cost item; type: {type(self).__name__}"

    __str__ = __repr__
