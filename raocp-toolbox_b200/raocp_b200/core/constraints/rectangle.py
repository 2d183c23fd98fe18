"""Box constraint lo <= [x; u] <= hi (nonleaf) / lo <= x <= hi (leaf).

API parity with reference raocp/core/constraints/rectangle.py:5-69.  `project` clips on the device
(kernel k_box in csrc/ops.cu, via the C-ABI entry rb_box_project); a NaN entry raises ValueError like
the reference's `_constrain` (rectangle.py:50-59).
"""
import numpy as np
from . import base_constraint as bc


class Rectangle(bc.Constraint):
    def __init__(self, node_type, _min, _max):
        super().__init__(node_type)
        self._check_constraints(_min, _max)
        self._lo = _min
        self._hi = _max

    @property
    def is_active(self):
        return True

    @property
    def lower(self):
        return self._lo

    @property
    def upper(self):
        return self._hi

    def _set_matrices(self):
        nx, nu = self.state_size, self.control_size
        self.state_matrix = np.vstack((np.eye(nx), np.zeros((nu, nx))))
        if self.node_type.is_nonleaf:
            self.control_matrix = np.vstack((np.zeros((nx, nu)), np.eye(nu)))

    def project(self, vector):
        self._check_input(vector)
        from ... import _lib
        return _lib.box_project(vector, self._lo, self._hi)

    @staticmethod
    def _check_constraints(_min, _max):
        if _min.size != _max.size:
            raise Exception("Rectangle constraint - min and max vectors sizes are not equal")
        for i in range(_min.size):
            if _min[i] is None and _max[i] is None:
                raise Exception("Rectangle constraint - both min and max constraints cannot be None")
            if _min[i] is None or _max[i] is None:
                continue
            if _min[i] > _max[i]:
                raise Exception("Rectangle constraint - min greater than max")

    def _check_input(self, vector):
        if vector.size != self.state_matrix.shape[0]:
            raise Exception("Rectangle constraint - input vector does not equal expected size")

    def __repr__(self):
        return f"Constraint; type: {type(self).__name__}"

    __str__ = __repr__
