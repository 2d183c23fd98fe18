"""AVaR risk measure as a conic ambiguity set (E, F, cone, b) (API parity with reference raocp/core/risks.py:5-82).

For a node with c children and conditional child probabilities pi:
    E = [alpha*I; -I; 1^T]  ((2c+1) x c),   F is (2c+1) x 0,
    cone = R_+^{2c} x {0},   b = [pi; 0_c; 1].
"""
import numpy as np
from .constraints import cones as core_cones


class AVaR:
    def __init__(self, alpha):
        if not (0 <= alpha <= 1):
            raise ValueError("alpha value '%d' not supported" % alpha)
        self._alpha = alpha
        self._probs = None
        self._e = self._f = self._cone = self._b = None

    @property
    def is_risk(self):
        return True

    @property
    def alpha(self):
        return self._alpha

    def _build(self):
        # (E, F, cone, b) of reference risks.py:28-35, built on first use: a 10^5-node problem never needs the
        # dense E on the host, the device kernels use the closed forms documented in DESIGN.md
        if self._e is None and self._probs is not None:
            c = self._probs.size
            ident = np.eye(c)
            self._e = np.vstack((self._alpha * ident, -ident, np.ones((1, c))))
            self._f = np.zeros((2 * c + 1, 0))
            self._cone = core_cones.Cartesian([core_cones.NonnegativeOrthant(dimension=2 * c),
                                               core_cones.Zero(dimension=1)])
            self._b = np.concatenate((self._probs.reshape(-1), np.zeros(c), [1.0])).reshape(-1, 1)

    @property
    def matrix_e(self):
        self._build()
        return self._e

    @property
    def matrix_f(self):
        self._build()
        return self._f

    @property
    def cone(self):
        self._build()
        return self._cone

    @property
    def vector_b(self):
        self._build()
        return self._b

    @property
    def probs(self):
        return self._probs

    @probs.setter
    def probs(self, vector):
        self._probs = np.asarray(vector)
        self._e = self._f = self._cone = self._b = None

    def __repr__(self):
        return f"Risk item; type: {type(self).__name__}, alpha: {self._alpha}; cone: {self.cone.types}"

    __str__ = __repr__
