"""Inactive constraint (API parity with reference raocp/core/constraints/no_constraint.py:4-13)."""
from . import base_constraint as bc


class No(bc.Constraint):
    def __init__(self, node_type=None):
        super().__init__(node_type)

    @property
    def is_active(self):
        return False
