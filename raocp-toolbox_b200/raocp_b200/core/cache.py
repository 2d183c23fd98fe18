"""Cache: iterate storage + offline factorisation + prox_f + prox_g* -- device resident.

Drop-in for the reference's raocp/core/cache.py:8-393 (same constructor and method names, same list-of-(k,1)-blocks
exchange format with (1,1) placeholders, same error messages); every method forwards to a hand-written sm_100a kernel
through the C-ABI (include/raocp_b200.h).  The iterates never leave HBM unless a getter is called.

Differences that are deliberate (DESIGN.md "Boundary"):
  * the full iterate history the reference appends on every update_cache() (cache.py:186-196) is opt-in
    (keep_history=True); by default only the current and the old iterate exist, on the device;
  * get_nullspace_matrices() builds the null-space bases lazily on the host for inspection only -- the device
    projects onto the kernel in closed form.
"""
import numpy as np

from . import raocp_spec as ps
from .device import DeviceSolver
from .flatten import FlatProblem


class Cache:
    def __init__(self, problem_spec: ps.RAOCP, batch=1, dedup=True, device=0, keep_history=False, shard=None,
                 sweep_cuts=None):
        self.__raocp = problem_spec
        self.__flat = FlatProblem(problem_spec, batch=batch, dedup=dedup, device=device, shard=shard,
                                  sweep_cuts=sweep_cuts)
        f = self.__flat
        self.__num_nodes = f.n
        self.__num_nonleaf_nodes = f.m
        self.__state_size = f.nx
        self.__control_size = f.nu
        self.__keep_history = keep_history
        self.__primal_cache = []
        self.__dual_cache = []
        self.__initial_state = None
        n, m = f.n, f.m
        # segment start tables of the block lists (cache.py:127-132, 142-155)
        self.__segment_p = [None, 0, n, n + m, n + 2 * m, 2 * n + 2 * m, 3 * n + 2 * m]
        self.__segment_d = [None, 0, n, 2 * n, 3 * n, 4 * n, 5 * n, 6 * n, 7 * n, None, None,
                            7 * n, 8 * n, 9 * n, 10 * n, 11 * n]
        mp = f.maps()
        self.__p_sizes, self.__d_sizes = mp["p_sizes"], mp["d_sizes"]
        self.__dev = DeviceSolver(f)
        self._offline()
        self.update_cache()

    # -- raocp_b200 extensions ---------------------------------------------------------------------------------------
    @property
    def device_solver(self):
        return self.__dev

    @property
    def flat_problem(self):
        return self.__flat

    # -- getters (cache.py:56-75) --------------------------------------------------------------------------------------
    def get_raocp(self):
        return self.__raocp

    def get_primal(self, instance=0):
        f = self.__flat
        return (f.primal_to_blocks(self.__dev.get_primal(0)[instance]),
                f.primal_to_blocks(self.__dev.get_primal(1)[instance]))

    def get_primal_segments(self):
        return self.__segment_p.copy()

    def get_dual(self, instance=0):
        f = self.__flat
        return (f.dual_to_blocks(self.__dev.get_dual(0)[instance]),
                f.dual_to_blocks(self.__dev.get_dual(1)[instance]))

    def get_dual_segments(self):
        return self.__segment_d.copy()

    def get_kernel_constraint_matrices(self):
        out = []
        for i in range(self.__num_nonleaf_nodes):
            risk = self.__raocp.risk_at_node(i)
            eye = np.eye(len(self.__raocp.tree.children_of(i)))
            zeros = np.zeros((risk.matrix_f.shape[1], eye.shape[0]))
            out.append(np.vstack((np.hstack((risk.matrix_e.T, -eye, -eye)), np.hstack((risk.matrix_f.T, zeros, zeros)))))
        return out

    def get_nullspace_matrices(self):
        import scipy.linalg
        return [scipy.linalg.null_space(mat) for mat in self.get_kernel_constraint_matrices()]

    # -- setters (cache.py:79-122) -------------------------------------------------------------------------------------
    def cache_initial_state(self, state):
        self.__initial_state = state
        self.__dev.set_initial_state(state)
        if self.__keep_history and self.__primal_cache:
            self.__primal_cache[0][0] = state

    def _check_blocks(self, candidate, sizes, what, segments, seg_ids):
        if len(candidate) != len(sizes):
            raise Exception(f"Candidate {what} list is wrong length")
        for i, block in enumerate(candidate):
            if np.asarray(block).shape != (sizes[i], 1):
                for s in reversed(seg_ids):
                    if i >= segments[s]:
                        segment, node = s, i - segments[s]
                        break
                raise Exception(f"Candidate {what} array shape error in segment {segment} at node {node},\n"
                                f"candidate shape: {np.asarray(block).shape},\n"
                                f"current shape: {(int(sizes[i]), 1)}")

    def set_primal(self, candidate_primal):
        self._check_blocks(candidate_primal, self.__p_sizes, "primal", self.__segment_p, range(1, 6))
        self.__dev.set_primal(0, self.__flat.primal_from_blocks(candidate_primal))

    def set_dual(self, candidate_dual):
        self._check_blocks(candidate_dual, self.__d_sizes, "dual", self.__segment_d,
                           [s for s in range(1, 15) if s not in (8, 9, 10)])
        self.__dev.set_dual(0, self.__flat.dual_from_blocks(candidate_dual))

    # -- cache (cache.py:186-196) --------------------------------------------------------------------------------------
    def update_cache(self):
        """old <- current on the device; the host-side history of the reference is kept only if keep_history=True"""
        self.__dev.update_cache()
        if self.__keep_history:
            self.__primal_cache.append(self.get_primal()[0])
            self.__dual_cache.append(self.get_dual()[0])

    # -- offline (cache.py:200-242) ------------------------------------------------------------------------------------
    def _offline(self):
        self.offline_projection_dynamics()
        self.offline_projection_kernel()

    def offline_projection_dynamics(self):
        self.__dev.offline()

    def offline_projection_kernel(self):
        """nothing to precompute: the AVaR kernel projector is closed-form on the device"""

    # -- proximal of f (cache.py:248-317) ------------------------------------------------------------------------------
    def proximal_of_f(self, solver_parameter):
        self.__dev.prox_f(solver_parameter)

    def proximal_of_relaxation_s_at_stage_zero(self, solver_parameter):
        self.__dev.s0_shift(float(np.asarray(solver_parameter).reshape(-1)[0]))

    def project_on_dynamics(self):
        self.__dev.project_dynamics()

    def project_on_kernel(self):
        self.__dev.project_kernel()

    # -- proximal of g conjugate (cache.py:321-393) --------------------------------------------------------------------
    def proximal_of_g_conjugate(self, solver_parameter):
        self.__dev.prox_g_conj(solver_parameter)

    def modify_dual(self, solver_parameter):
        self.__dev.modify_dual(solver_parameter)

    def add_halves(self):
        self.__dev.add_halves()

    def project_on_constraints_nonleaf(self):
        self.__dev.project_nonleaf()

    def project_on_constraints_leaf(self):
        self.__dev.project_leaf()

    def modify_projection(self, solver_parameter, modified_dual):
        self.__dev.modify_projection(solver_parameter, self.__flat.dual_from_blocks(modified_dual))
