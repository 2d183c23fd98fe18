"""Solver: Chambolle-Pock driver with the whole loop on the device.

Drop-in for the reference's raocp/core/solver.py:12-171 (plotting / printing helpers :173-253 are out of scope):
same constructor, the four half-step methods, `_calculate_chock_errors`, and
`chock(initial_state, max_iters=10, tol=1e-5) -> 0 | 1` with the reference's stopping rule (max_iters+1 iterations
when it does not converge, solver.py:156-161).  The loop body runs as fused sm_100a kernels (csrc/fused.cu); only
the six residual norms per iteration come back to the host.

Step size: alpha_1 = alpha_2 = 0.999 / lambda_max(L* L) like solver.py:105-118, but lambda_max is computed on the
device as the maximum over the diagonal blocks of L* L (rb_lambda_max) instead of ARPACK; pass `alpha=` to chock to
impose a value (parity tests pass the oracle's alpha, whose last bits depend on ARPACK's random start vector).
"""
import time

import numpy as np

from . import cache as cache
from . import operators as ops
from . import raocp_spec as spec


class Solver:
    def __init__(self, problem_spec: spec.RAOCP, batch=1, dedup=True, device=0, keep_history=False, verbose=True,
                 shard=None, sweep_cuts=None):
        self.__raocp = problem_spec
        self.__cache = cache.Cache(self.__raocp, batch=batch, dedup=dedup, device=device, keep_history=keep_history,
                                   shard=shard, sweep_cuts=sweep_cuts)
        self.__operator = ops.Operator(self.__cache)
        self.__dev = self.__cache.device_solver
        self.__initial_state = None
        self.__parameter_1 = None
        self.__parameter_2 = None
        self.__error = [np.zeros(1)] * 3
        self.__delta_error = [np.zeros(1)] * 3
        self.__error_cache = None
        self.__delta_error_cache = None
        self.__verbose = verbose
        self.iterations = 0

    # -- raocp_b200 extensions ---------------------------------------------------------------------------------------
    @property
    def cache(self):
        return self.__cache

    @property
    def operator(self):
        return self.__operator

    @property
    def step_size(self):
        return self.__parameter_1

    def set_step_size(self, alpha):
        self.__parameter_1 = self.__parameter_2 = float(alpha)

    @property
    def residual_history(self):
        """(xi, delta): arrays (iterations, 3) like the reference's __error_cache / __delta_error_cache"""
        return self.__error_cache, self.__delta_error_cache

    # -- the four half steps (solver.py:27-61) -----------------------------------------------------------------------
    def primal_k_plus_half(self):
        self.__dev.primal_half(self.__parameter_1)

    def primal_k_plus_one(self):
        self.__cache.proximal_of_f(self.__parameter_1)

    def dual_k_plus_half(self):
        self.__dev.dual_half(self.__parameter_2)

    def dual_k_plus_one(self):
        self.__cache.proximal_of_g_conjugate(self.__parameter_2)

    def _calculate_chock_errors(self):
        """xi_0, xi_1, xi_2, delta_0, delta_1, delta_2 as block lists (solver.py:63-95)"""
        f = self.__cache.flat_problem
        _, vec = self.__dev.residuals(self.__parameter_1, vectors=True)
        v = vec[0]
        np_, nd_ = f.np_, f.nd_
        cuts = np.cumsum([0, np_, np_, nd_, np_, np_, nd_])
        parts = [v[cuts[i]: cuts[i + 1]] for i in range(6)]
        to_p, to_d = f.primal_to_blocks, f.dual_to_blocks
        return to_p(parts[0]), to_p(parts[1]), to_d(parts[2]), to_p(parts[3]), to_p(parts[4]), to_d(parts[5])

    def compute_step_size(self):
        """0.999 / lambda_max(L* L) (solver.py:105-118), lambda_max from the device"""
        return 0.999 / self.__dev.lambda_max()

    def chock(self, initial_state, max_iters=10, tol=1e-5, alpha=None):
        """Chambolle-Pock algorithm (solver.py:97-171).  Returns 0 (converged) or 1 (not converged)."""
        self.__initial_state = initial_state
        self.__cache.cache_initial_state(self.__initial_state)
        self.set_step_size(self.compute_step_size() if alpha is None else alpha)

        if self.__verbose:
            print("timer started")
        tick = time.perf_counter()
        status, iters, xi, delta = self.__dev.iterate(self.__parameter_1, max_iters, tol, history=True)
        tock = time.perf_counter()
        if self.__verbose:
            print(f"timer stopped in {tock - tick:0.4f} seconds")
        self.iterations = iters
        batch = xi.shape[1]
        self.__error_cache = xi[:, 0, :] if batch == 1 else xi
        self.__delta_error_cache = delta[:, 0, :] if batch == 1 else delta
        self.__error = list(xi[-1, 0])
        self.__delta_error = list(delta[-1, 0])
        if status not in (0, 1):
            raise Exception("Iteration error in solver")
        return status

    # -- presentation helpers of the reference (solver.py:173-253) are out of scope ------------------------------------
    def print_states(self):
        primal, _ = self.__cache.get_primal()
        seg_p = self.__cache.get_primal_segments()
        print("states =\n")
        for i in range(seg_p[1], seg_p[2]):
            print(f"{primal[i]}\n")

    def print_inputs(self):
        primal, _ = self.__cache.get_primal()
        seg_p = self.__cache.get_primal_segments()
        print("inputs =\n")
        for i in range(seg_p[2], seg_p[3]):
            print(f"{primal[i]}\n")

    def plot_residuals(self):
        raise NotImplementedError("plotting is outside the scope of raocp_b200 (reference solver.py:187-200)")

    def plot_solution(self):
        raise NotImplementedError("plotting is outside the scope of raocp_b200 (reference solver.py:202-253)")
