"""Convex cones and their Euclidean projections.

API parity with reference raocp/core/constraints/cones.py:4-230 (Real, Zero, NonnegativeOrthant, SecondOrderCone,
Cartesian; `project` / `project_onto_dual`; same dimension checks and error messages).  The arithmetic runs on the
device through the C-ABI entry rb_cone_project (kernel k_cone in csrc/ops.cu) -- these classes are the
host-side mirror only; inside the solver the same device functions are fused into the dual pass.
"""
import numpy as np

# cone type codes shared with csrc/common.cuh
_REAL, _ZERO, _NONNEG, _SOC = 0, 1, 2, 3


def _check_dimension(cone_type, cone_dimension, vector):
    vector_dimension = vector.size
    if cone_dimension is None:
        cone_dimension = vector_dimension
    if cone_dimension != vector_dimension:
        raise ValueError('%s cone dimension error: cone dimension = %d, input vector dimension = %d'
                         % (cone_type, cone_dimension, vector_dimension))
    return vector_dimension


def _device_project(code, vector):
    from ... import _lib
    return _lib.cone_project(code, vector)


class _Cone:
    _code = None
    _dual_code = None

    def __init__(self, dimension=None):
        self._dimension = dimension
        self._shape = None

    def project(self, vector):
        self._dimension = _check_dimension(type(self), self._dimension, vector)
        self._shape = vector.shape
        return _device_project(self._code, vector)

    def project_onto_dual(self, vector):
        self._dimension = _check_dimension(type(self), self._dimension, vector)
        self._shape = vector.shape
        return _device_project(self._dual_code, vector)

    @property
    def dimension(self):
        return self._dimension


class Real(_Cone):
    """R^n; its dual is {0}"""
    _code, _dual_code = _REAL, _ZERO


class Zero(_Cone):
    """{0}; its dual is R^n"""
    _code, _dual_code = _ZERO, _REAL


class NonnegativeOrthant(_Cone):
    """R^n_+ (self dual)"""
    _code, _dual_code = _NONNEG, _NONNEG


class SecondOrderCone(_Cone):
    """{(z, t): ||z||_2 <= t} (self dual); the LAST entry of the vector is t"""
    _code, _dual_code = _SOC, _SOC

    def project(self, vector):
        self._dimension = _check_dimension(type(self), self._dimension, vector)
        if self._dimension < 3:
            raise Exception("Attempt to project a vector of size < 3 onto second order cone")
        self._shape = vector.shape
        return _device_project(_SOC, vector)

    def project_onto_dual(self, vector):
        return SecondOrderCone.project(self, vector)


class Cartesian:
    """Cartesian product of cones"""

    def __init__(self, cones):
        self._cones = cones
        self._num_cones = len(cones)
        self._dimension = 0
        for c in cones:
            if c.dimension is None:
                self._dimension = None
                break
            self._dimension += c.dimension
        self._dimensions = [None] * self._num_cones

    def _run(self, list_of_vectors, method):
        parts = self._check_list_of_vectors(list_of_vectors)
        out = []
        for i, cone in enumerate(self._cones):
            self._dimensions[i] = _check_dimension(type(cone), cone.dimension, parts[i])
            out.append(getattr(cone, method)(parts[i]))
        self._dimension = sum(self._dimensions)
        return np.vstack(out) if len(list_of_vectors) == 1 else out

    def project(self, list_of_vectors):
        return self._run(list_of_vectors, "project")

    def project_onto_dual(self, list_of_vectors):
        return self._run(list_of_vectors, "project_onto_dual")

    def _check_list_of_vectors(self, list_of_vectors):
        # a single stacked vector is split by the declared cone dimensions
        if len(list_of_vectors) != 1:
            return list_of_vectors
        parts, cursor = [], 0
        for cone in self._cones:
            parts.append(list_of_vectors[0][cursor: cursor + cone.dimension])
            cursor += cone.dimension
        return parts

    @property
    def types(self):
        return " x ".join(type(c).__name__ for c in self._cones)

    @property
    def dimension(self):
        return self._dimension

    @property
    def dimensions(self):
        return self._dimensions

    @property
    def num_cones(self):
        return self._num_cones

    @property
    def cones(self):
        return self._cones
