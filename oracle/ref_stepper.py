"""Drive the UNMODIFIED reference `Solver` one Chambolle-Pock iteration at a time with bounded memory.

TEST / BENCH INFRASTRUCTURE (never imported by the product).  `Solver.chock` (reference raocp/core/solver.py:97-171)
cannot be used at the sizes BASELINE.json names: it keeps every iterate (`cache.py:186-196`, +0.8 GB per iteration at
cfg3) and computes the step size with ARPACK on Python callbacks.  This harness calls exactly the methods the loop
body of `chock` calls, in the same order (solver.py:124-141):

    primal_k_plus_half -> primal_k_plus_one -> dual_k_plus_half -> dual_k_plus_one -> _calculate_chock_errors
    -> the six inf-norms (solver.py:137-141) -> Cache.update_cache

and then drops all but the newest entry of the Cache's history lists (they are only ever read at [-1]).
The step size is passed in (the closed-form block value, equal to ARPACK's to 1e-15) or computed by the
reference's own ARPACK call when `alpha=None`.
"""
import time

import numpy as np


class RefStepper:
    def __init__(self, api, problem, x0, alpha=None):
        """api: oracle.ref_loader.RefApi; problem: a reference RAOCP; x0: (nx, 1)"""
        t0 = time.perf_counter()
        self.solver = api.Solver(problem)             # runs Cache._offline (cache.py:200-205)
        self.setup_s = time.perf_counter() - t0
        self.cache = self.solver._Solver__cache
        self.cache.cache_initial_state(np.asarray(x0, dtype=float).reshape(-1, 1))
        if alpha is None:
            alpha = self.arpack_step_size()
        self.alpha = float(alpha)
        self.solver._Solver__parameter_1 = self.alpha
        self.solver._Solver__parameter_2 = self.alpha
        self.iterations = 0
        self.xi, self.delta = [], []

    def arpack_step_size(self):
        """solver.py:105-118, verbatim calls"""
        from scipy.sparse.linalg import LinearOperator, eigs
        op = self.solver._Solver__operator
        _, prim = self.cache.get_primal()
        _, dual = self.cache.get_dual()
        sp, sd = np.vstack(prim).size, np.vstack(dual).size
        ell = LinearOperator(dtype=None, shape=(sd, sp), matvec=op.linop_ell)
        ell_t = LinearOperator(dtype=None, shape=(sp, sd), matvec=op.linop_ell_transpose)
        eigens, _ = eigs(ell_t * ell)
        return 0.999 / np.real(max(eigens))

    def step(self, norms=True):
        """one iteration of the loop body; returns (xi[3], delta[3]) or None when norms=False"""
        s = self.solver
        s.primal_k_plus_half()
        s.primal_k_plus_one()
        s.dual_k_plus_half()
        s.dual_k_plus_one()
        out = None
        if norms:
            res = s._calculate_chock_errors()
            vals = [np.linalg.norm([np.linalg.norm(a, ord=np.inf) for a in vec], np.inf) for vec in res]
            out = (np.array(vals[:3]), np.array(vals[3:]))
            self.xi.append(out[0])
            self.delta.append(out[1])
        self.cache.update_cache()
        del self.cache._Cache__primal_cache[:-1]      # history is only read at [-1] (cache.py:191-196)
        del self.cache._Cache__dual_cache[:-1]
        self.iterations += 1
        return out

    def primal(self):
        """raw flat primal (np.vstack of the block list) of the newest iterate"""
        return np.vstack(self.cache._Cache__primal_cache[-1]).reshape(-1)

    def dual(self):
        """raw flat dual, (1,1) placeholders included (cache.py:140-170)"""
        return np.vstack(self.cache._Cache__dual_cache[-1]).reshape(-1)

    def segment_lengths(self):
        """(primal, dual) flat lengths per reference list segment: the segment tables count BLOCKS
        (cache.py:127-132,142-155), so the flat lengths are sums of block sizes"""
        seg_p, seg_d = self.cache.get_primal_segments(), self.cache.get_dual_segments()
        pb = [b.size for b in self.cache._Cache__primal_cache[-1]]
        db = [b.size for b in self.cache._Cache__dual_cache[-1]]
        p_edges = [e for e in seg_p if e is not None]
        d_edges = [seg_d[k] for k in (1, 2, 3, 4, 5, 6, 7, 11, 12, 13, 14)] + [len(db)]
        cp, cd = np.concatenate(([0], np.cumsum(pb))), np.concatenate(([0], np.cumsum(db)))
        return cp[p_edges], cd[d_edges]
