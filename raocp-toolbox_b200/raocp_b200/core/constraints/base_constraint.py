"""Constraint base class: sizes -> selector matrices -> transposes
(API parity with reference raocp/core/constraints/base_constraint.py:4-118)."""
import numpy as np


class Constraint:
    def __init__(self, node_type):
        self._node_type = node_type
        self._nx = None
        self._nu = None
        self._gamma_x = None
        self._gamma_u = None
        self._gamma_x_t = None
        self._gamma_u_t = None

    def project(self, vector):
        pass

    @property
    def is_active(self):
        raise Exception("Base constraint accessed - actual constraint must not be setup")

    @property
    def node_type(self):
        return self._node_type

    @property
    def state_size(self):
        return self._nx

    @property
    def control_size(self):
        return self._nu

    @property
    def state_matrix(self):
        return self._gamma_x

    @property
    def control_matrix(self):
        return self._gamma_u

    @property
    def state_matrix_transposed(self):
        if self._gamma_x_t is None:
            raise Exception("Constraint state matrix transpose called but is None")
        return self._gamma_x_t

    @property
    def control_matrix_transposed(self):
        if self._gamma_u_t is None:
            raise Exception("Constraint control matrix transpose called but is None")
        return self._gamma_u_t

    def _finish(self):
        self._set_matrices()
        self._get_transpose()

    @state_size.setter
    def state_size(self, size):
        self._nx = size
        if self._node_type.is_nonleaf:
            if self._nu is not None:
                self._finish()
        elif self._node_type.is_leaf:
            self._nu = 0
            self._finish()
        else:
            raise Exception("Node type missing")

    @control_size.setter
    def control_size(self, size):
        self._nu = size
        if self._node_type.is_nonleaf:
            if self._nx is not None:
                self._finish()
        elif self._node_type.is_leaf:
            raise Exception("Attempt to set control size on leaf node")
        else:
            raise Exception("Node type missing")

    def _set_matrices(self):
        pass

    def _get_transpose(self):
        if self._node_type.is_nonleaf:
            self._gamma_x_t = np.transpose(self.state_matrix)
            self._gamma_u_t = np.transpose(self.control_matrix)
        elif self._node_type.is_leaf:
            self._gamma_x_t = np.transpose(self.state_matrix)
        else:
            raise Exception("Node type missing")

    @state_matrix.setter
    def state_matrix(self, matrix):
        self._gamma_x = matrix

    @control_matrix.setter
    def control_matrix(self, matrix):
        if self._node_type.is_nonleaf:
            self._gamma_u = matrix
        elif self._node_type.is_leaf:
            raise Exception("Attempt to set control constraint matrix of leaf node")
        else:
            raise Exception("Node type missing")

    def __repr__(self):
        return "Base constraint"

    __str__ = __repr__
