"""Operator: the linear map L and its adjoint L* on the device.

Drop-in for the reference's raocp/core/operators.py:5-120: `ell` / `ell_transpose` take the reference's block lists and
write into the caller's output list, `linop_ell` / `linop_ell_transpose` take and return the flat np.vstack form
(placeholders included).  The arithmetic is kernels k_l_axpby / k_lt_axpby (csrc/ops.cu) via rb_apply_L / rb_apply_Lt.
"""
import numpy as np

from . import cache as core_cache


class Operator:
    def __init__(self, cache: core_cache.Cache):
        self.__cache = cache
        self.__raocp = cache.get_raocp()
        self.__flat = cache.flat_problem
        self.__dev = cache.device_solver
        self.__segment_p = cache.get_primal_segments()
        self.__segment_d = cache.get_dual_segments()
        mp = self.__flat.maps()
        self.__d_real = np.flatnonzero(mp["d_real"])
        self.__d_total = mp["d_total"]
        self.__d_gather = mp["d_gather"]

    def ell(self, input_primal, output_dual):
        compact = self.__dev.apply_L(self.__flat.primal_from_blocks(input_primal))[0]
        blocks = self.__flat.dual_to_blocks(compact)
        for k in self.__d_real:          # placeholders of the caller's list are left alone, like the reference
            output_dual[k] = blocks[k]

    def ell_transpose(self, input_dual, output_primal):
        compact = self.__dev.apply_Lt(self.__flat.dual_from_blocks(input_dual))[0]
        blocks = self.__flat.primal_to_blocks(compact)
        tau0 = self.__segment_p[4]       # never written by the reference (operators.py:55-94)
        for k, block in enumerate(blocks):
            if k != tau0:
                output_primal[k] = block

    def linop_ell(self, flat_primal):
        compact = self.__dev.apply_L(np.asarray(flat_primal, dtype=np.float64).reshape(-1))[0]
        flat = np.zeros(self.__d_total)
        flat[self.__d_gather] = compact
        return flat.reshape(-1, 1)

    def linop_ell_transpose(self, flat_dual):
        flat = np.asarray(flat_dual, dtype=np.float64).reshape(-1)
        out = self.__dev.apply_Lt(flat[self.__d_gather])[0]
        return out.reshape(-1, 1)
