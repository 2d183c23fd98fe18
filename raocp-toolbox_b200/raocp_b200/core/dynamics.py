"""Pair (A, B) of a linear system x+ = A x + B u (API parity with reference raocp/core/dynamics.py:8-25)."""


class Dynamics:
    def __init__(self, state_dynamics, control_dynamics):
        if state_dynamics.shape[0] != control_dynamics.shape[0]:
            raise ValueError("Dynamics matrices rows are different sizes")
        self._a = state_dynamics
        self._b = control_dynamics

    @property
    def state_dynamics(self):
        return self._a

    @property
    def control_dynamics(self):
        return self._b
