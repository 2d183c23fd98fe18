"""Vectorised (stage-synchronous, structure-of-arrays) CPU restatement of the reference's Chambolle-Pock
iteration.  TEST INFRASTRUCTURE ONLY -- the "debugging twin" of the CUDA kernels.

Same arithmetic as oracle/cp_node_oracle.py (which is pinned against the unmodified reference), but all nodes of a
stage are processed with batched NumPy calls, so the 10^4..10^5-node configurations finish in seconds and can be used
for full-size parity checks.  PARITY PINNED: tests/test_oracle_pinning.py checks it against the node oracle, against
iterates recorded from the unmodified reference (tests/golden/*.npz) and against 4-3-residuals.tex.

COMPACT LAYOUT (also the C-ABI exchange format of include/raocp_b200.h; no placeholders):
  primal = [x (n*nx) | u (m*nu) | y (sum_i 2c_i+1, node i's block = [y_a(c_i); y_b(c_i); y_last]) | tau (n) | s (n)]
  dual   = [d1 (like y) | d2 (m) | d3 ((n-1)*nx, row j-1 = edge into node j) | d4 ((n-1)*nu) | d5 (n-1) | d6 (n-1)
            | d7 (m*(nx+nu), only if nonleaf rectangles) | d11 (L*nx) | d12 (L) | d13 (L) | d14 (L*nx, only if leaf
            rectangles)],   L = n - m leaves.
Reference segment numbering: cache.py:126-170.
"""
import numpy as np


class FlatProblem:
    """Arrays extracted from a problem object exposing the RAOCP accessor API (raocp_spec.py:56-75)."""

    def __init__(self, problem):
        tree = problem.tree
        n, m = int(tree.num_nodes), int(tree.num_nonleaf_nodes)
        self.n, self.m, self.nleaf = n, m, n - m
        self.nx = nx = problem.state_dynamics_at_node(1).shape[1]
        self.nu = nu = problem.control_dynamics_at_node(1).shape[1]
        self.parent = np.array([int(tree.ancestor_of(i)) for i in range(n)])
        self.stage = np.array([int(tree.stage_of(i)) for i in range(n)])
        self.num_stages = int(tree.num_stages)
        self.child_first = np.zeros(m, dtype=np.int64)
        self.child_count = np.zeros(m, dtype=np.int64)
        for i in range(m):
            ch = np.asarray(tree.children_of(i))
            assert np.array_equal(ch, np.arange(ch[0], ch[0] + ch.size)), "children must be a contiguous range"
            self.child_first[i], self.child_count[i] = ch[0], ch.size
        assert np.all(np.diff(self.stage) >= 0), "nodes must be numbered stage by stage"
        self.stage_off = np.searchsorted(self.stage, np.arange(self.num_stages + 1))

        def table(objs, get):
            """de-duplicate matrices by object identity: returns (stack of unique matrices, per-object index)"""
            ids, mats, idx = {}, [], np.zeros(len(objs), dtype=np.int64)
            for k, o in enumerate(objs):
                key = id(o)
                if key not in ids:
                    ids[key] = len(mats)
                    mats.append(np.real(np.asarray(get(o), dtype=complex)).astype(float))
                idx[k] = ids[key]
            return np.stack(mats), idx

        dyn = problem.list_of_dynamics[1:]
        self.A_tab, dyn_idx = table(dyn, lambda o: o.state_dynamics)
        self.B_tab, _ = table(dyn, lambda o: o.control_dynamics)
        self.dyn_idx = np.concatenate(([0], dyn_idx))  # per node j (entry 0 unused)
        costs = problem.list_of_nonleaf_costs[1:]
        self.sq_tab, cost_idx = table(costs, lambda o: o.sqrt_state_weights)
        self.sr_tab, _ = table(costs, lambda o: o.sqrt_control_weights)
        self.cost_idx = np.concatenate(([0], cost_idx))
        self.sqf_tab, self.leafcost_idx = table(problem.list_of_leaf_costs[m:], lambda o: o.sqrt_state_weights)
        # constraints: all-or-nothing rectangles
        nl_active = [bool(problem.nonleaf_constraint_at_node(i).is_active) for i in range(m)]
        lf_active = [bool(problem.leaf_constraint_at_node(i).is_active) for i in range(m, n)]
        assert all(nl_active) or not any(nl_active)
        assert all(lf_active) or not any(lf_active)
        self.nl_rect, self.leaf_rect = all(nl_active), all(lf_active)

        def bounds(con):
            lo = con._Rectangle__min if hasattr(con, "_Rectangle__min") else con.lower
            hi = con._Rectangle__max if hasattr(con, "_Rectangle__max") else con.upper
            return np.asarray(lo, dtype=float).reshape(-1), np.asarray(hi, dtype=float).reshape(-1)

        if self.nl_rect:
            lo_hi = [bounds(problem.nonleaf_constraint_at_node(i)) for i in range(m)]
            self.nl_lo = np.stack([b[0] for b in lo_hi])
            self.nl_hi = np.stack([b[1] for b in lo_hi])
        if self.leaf_rect:
            lo_hi = [bounds(problem.leaf_constraint_at_node(i)) for i in range(m, n)]
            self.leaf_lo = np.stack([b[0] for b in lo_hi])
            self.leaf_hi = np.stack([b[1] for b in lo_hi])
        # risks: AVaR only (cache.py:172-178)
        self.risk_alpha = np.array([problem.risk_at_node(i).alpha for i in range(m)], dtype=float)
        self.pi = np.zeros(n)
        for i in range(m):
            self.pi[self.child_first[i]: self.child_first[i] + self.child_count[i]] = \
                np.asarray(problem.risk_at_node(i).probs, dtype=float).reshape(-1)
        # ragged y layout
        self.ysize = 2 * self.child_count + 1
        self.yoff = np.concatenate(([0], np.cumsum(self.ysize)))
        kpos = np.arange(n) - np.concatenate(([0], self.child_first[self.parent[1:]]))  # position among siblings
        par = self.parent.copy()
        par[0] = 0
        self.ya_idx = self.yoff[par] + kpos            # per edge j >= 1 (entry 0 unused)
        self.yb_idx = self.ya_idx + self.child_count[par]
        self.ylast_idx = self.yoff[:-1] + 2 * self.child_count
        # sizes
        self.ysz = int(self.yoff[-1])
        self.np_ = n * nx + m * nu + self.ysz + 2 * n
        self.nd_ = self.ysz + m + (n - 1) * (nx + nu + 2) + (m * (nx + nu) if self.nl_rect else 0) \
            + self.nleaf * (nx + 2) + (self.nleaf * nx if self.leaf_rect else 0)

    # segment reductions over the (contiguous) children of the nonleaf nodes lo..hi-1
    def child_sum(self, vals_per_edge, lo, hi):
        """vals_per_edge is indexed by edge j-1 for all edges; returns sum over children for nodes lo..hi-1."""
        first = self.child_first[lo:hi] - 1
        return self.child_sum_local(vals_per_edge[first[0]: first[-1] + self.child_count[hi - 1]], lo, hi)

    def child_sum_local(self, vals, lo, hi):
        """vals holds exactly the children of nodes lo..hi-1 (one contiguous node range), in order."""
        first = self.child_first[lo:hi]
        return np.add.reduceat(vals, first - first[0], axis=0)


class FlatOracle:
    def __init__(self, problem):
        self.fp = fp = problem if isinstance(problem, FlatProblem) else FlatProblem(problem)
        self.alpha = None
        self.x0 = None
        self.p = self.zero_primal()
        self.d = self.zero_dual()
        self.p_old = self.zero_primal()
        self.d_old = self.zero_dual()
        self._offline()

    # ------------------------------------------------------------------------------------------------ storage
    def zero_primal(self):
        fp = self.fp
        return dict(x=np.zeros((fp.n, fp.nx)), u=np.zeros((fp.m, fp.nu)), y=np.zeros(fp.ysz),
                    tau=np.zeros(fp.n), s=np.zeros(fp.n))

    def zero_dual(self):
        fp = self.fp
        d = {1: np.zeros(fp.ysz), 2: np.zeros(fp.m), 3: np.zeros((fp.n - 1, fp.nx)), 4: np.zeros((fp.n - 1, fp.nu)),
             5: np.zeros(fp.n - 1), 6: np.zeros(fp.n - 1), 11: np.zeros((fp.nleaf, fp.nx)), 12: np.zeros(fp.nleaf),
             13: np.zeros(fp.nleaf)}
        if fp.nl_rect:
            d[7] = np.zeros((fp.m, fp.nx + fp.nu))
        if fp.leaf_rect:
            d[14] = np.zeros((fp.nleaf, fp.nx))
        return d

    P_KEYS = ("x", "u", "y", "tau", "s")
    D_KEYS = (1, 2, 3, 4, 5, 6, 7, 11, 12, 13, 14)

    @classmethod
    def flat_primal(cls, p):
        return np.concatenate([p[k].reshape(-1) for k in cls.P_KEYS])

    @classmethod
    def flat_dual(cls, d):
        return np.concatenate([d[k].reshape(-1) for k in cls.D_KEYS if k in d])

    def unflat_primal(self, vec):
        out, cur = {}, 0
        for k, a in self.zero_primal().items():
            out[k] = np.array(vec[cur: cur + a.size], dtype=float).reshape(a.shape)
            cur += a.size
        return out

    def unflat_dual(self, vec):
        out, cur = {}, 0
        tmpl = self.zero_dual()
        for k in self.D_KEYS:
            if k in tmpl:
                a = tmpl[k]
                out[k] = np.array(vec[cur: cur + a.size], dtype=float).reshape(a.shape)
                cur += a.size
        return out

    # conversions from the reference's block lists (with placeholders) to the compact flat vectors
    def primal_from_blocks(self, blocks):
        return np.concatenate([np.asarray(b, dtype=float).reshape(-1) for b in blocks])

    def dual_from_blocks(self, blocks):
        fp, n, m = self.fp, self.fp.n, self.fp.m
        seg = lambda k: blocks[k * n: (k + 1) * n]  # noqa: E731  (11 segments of n blocks, cache.py:142-156)
        cat = lambda bl: np.concatenate([np.asarray(b, dtype=float).reshape(-1) for b in bl]) if len(bl) else np.zeros(0)  # noqa: E731
        parts = [cat(seg(0)[:m]), cat(seg(1)[:m]), cat(seg(2)[1:]), cat(seg(3)[1:]), cat(seg(4)[1:]), cat(seg(5)[1:])]
        if fp.nl_rect:
            parts.append(cat(seg(6)[:m]))
        parts += [cat(seg(7)[m:]), cat(seg(8)[m:]), cat(seg(9)[m:])]
        if fp.leaf_rect:
            parts.append(cat(seg(10)[m:]))
        return np.concatenate(parts)

    # ------------------------------------------------------------------------------------------------ offline
    def _offline(self):
        """cache.py:207-233, one batched step per stage (nodes of a stage are independent)."""
        fp = self.fp
        n, m, nx, nu = fp.n, fp.m, fp.nx, fp.nu
        self.P = np.zeros((n, nx, nx))
        self.P[m:] = np.eye(nx)
        self.K = np.zeros((m, nu, nx))
        self.Rt = np.zeros((m, nu, nu))
        self.Abar = np.zeros((n, nx, nx))
        A = lambda j: fp.A_tab[fp.dyn_idx[j]]  # noqa: E731
        B = lambda j: fp.B_tab[fp.dyn_idx[j]]  # noqa: E731
        for t in reversed(range(fp.num_stages - 1)):
            lo, hi = fp.stage_off[t], fp.stage_off[t + 1]
            ch = np.arange(fp.child_first[lo], fp.child_first[hi - 1] + fp.child_count[hi - 1])
            Bj, Aj, Pj = B(ch), A(ch), self.P[ch]
            BtP = np.einsum("jab,jac->jbc", Bj, Pj)                      # B' P
            Rt = np.eye(nu) + fp.child_sum_local(BtP @ Bj, lo, hi)
            S = fp.child_sum_local(BtP @ Aj, lo, hi)
            K = -np.linalg.solve(Rt, S)
            self.Rt[lo:hi], self.K[lo:hi] = Rt, K
            Ab = Aj + Bj @ K[fp.parent[ch] - lo]
            self.Abar[ch] = Ab
            edge_p = np.einsum("jab,jac->jbc", Ab, Pj @ Ab)
            self.P[lo:hi] = np.eye(nx) + np.einsum("iab,iac->ibc", K, K) + fp.child_sum_local(edge_p, lo, hi)
        self.PB = np.zeros((n, nx, nu))
        self.PB[1:] = self.P[1:] @ B(np.arange(1, n))

    # ------------------------------------------------------------------------------------------------ operators
    def ell(self, p):
        """L, operators.py:19-53."""
        fp = self.fp
        n, m = fp.n, fp.m
        par = fp.parent[1:]
        d = self.zero_dual()
        d[1] = p["y"].copy()
        edge_b = fp.pi[1:] * p["y"][fp.ya_idx[1:]]
        d[2] = p["s"][:m] - (fp.child_sum(edge_b, 0, m) + p["y"][fp.ylast_idx])
        d[3] = np.einsum("jab,jb->ja", fp.sq_tab[fp.cost_idx[1:]], p["x"][par])
        d[4] = np.einsum("jab,jb->ja", fp.sr_tab[fp.cost_idx[1:]], p["u"][par])
        d[5] = 0.5 * p["tau"][1:]
        d[6] = 0.5 * p["tau"][1:]
        if fp.nl_rect:
            d[7] = np.hstack((p["x"][:m], p["u"]))
        d[11] = np.einsum("iab,ib->ia", fp.sqf_tab[fp.leafcost_idx], p["x"][m:])
        d[12] = 0.5 * p["s"][m:]
        d[13] = 0.5 * p["s"][m:]
        if fp.leaf_rect:
            d[14] = p["x"][m:].copy()
        return d

    def ell_transpose(self, d):
        """L*, operators.py:55-94 (tau_0 stays 0)."""
        fp = self.fp
        n, m, nx = fp.n, fp.m, fp.nx
        par = fp.parent[1:]
        p = self.zero_primal()
        y = d[1].copy()
        y[fp.ya_idx[1:]] -= fp.pi[1:] * d[2][par]
        y[fp.ylast_idx] -= d[2]
        p["y"] = y
        p["s"][:m] = d[2]
        p["s"][m:] = 0.5 * (d[12] + d[13])
        ex = np.einsum("jab,jb->ja", fp.sq_tab[fp.cost_idx[1:]], d[3])
        eu = np.einsum("jab,jb->ja", fp.sr_tab[fp.cost_idx[1:]], d[4])
        p["x"][:m] = fp.child_sum(ex, 0, m)
        p["u"][:] = fp.child_sum(eu, 0, m)
        if fp.nl_rect:
            p["x"][:m] += d[7][:, :nx]
            p["u"] += d[7][:, nx:]
        p["x"][m:] = np.einsum("iab,ib->ia", fp.sqf_tab[fp.leafcost_idx], d[11])
        if fp.leaf_rect:
            p["x"][m:] += d[14]
        p["tau"][1:] = 0.5 * (d[5] + d[6])
        return p

    # ------------------------------------------------------------------------------------------------ prox_f
    def project_on_dynamics(self, p):
        """cache.py:259-288 with the reference's own formulas, one batched step per stage."""
        fp = self.fp
        n, m, nx, nu = fp.n, fp.m, fp.nx, fp.nu
        xb, ub = p["x"], p["u"]
        q = np.zeros((n, nx))
        q[m:] = -xb[m:]
        dv = np.zeros((m, nu))
        B = lambda j: fp.B_tab[fp.dyn_idx[j]]  # noqa: E731
        for t in reversed(range(fp.num_stages - 1)):
            lo, hi = fp.stage_off[t], fp.stage_off[t + 1]
            ch = np.arange(fp.child_first[lo], fp.child_first[hi - 1] + fp.child_count[hi - 1])
            edge = np.einsum("jab,ja->jb", B(ch), q[ch])
            rhs = ub[lo:hi] - fp.child_sum_local(edge, lo, hi)
            dv[lo:hi] = np.linalg.solve(self.Rt[lo:hi], rhs[..., None])[..., 0]
            inner = np.einsum("jab,jb->ja", self.PB[ch], dv[fp.parent[ch]]) + q[ch]
            edge_q = np.einsum("jab,ja->jb", self.Abar[ch], inner)
            q[lo:hi] = -xb[lo:hi] + np.einsum("iab,ia->ib", self.K[lo:hi], dv[lo:hi] - ub[lo:hi]) \
                + fp.child_sum_local(edge_q, lo, hi)
        x = np.zeros((n, nx))
        u = np.zeros((m, nu))
        x[0] = self.x0
        for t in range(fp.num_stages - 1):
            lo, hi = fp.stage_off[t], fp.stage_off[t + 1]
            u[lo:hi] = np.einsum("iab,ib->ia", self.K[lo:hi], x[lo:hi]) + dv[lo:hi]
            ch = np.arange(fp.child_first[lo], fp.child_first[hi - 1] + fp.child_count[hi - 1])
            x[ch] = np.einsum("jab,jb->ja", self.Abar[ch], x[fp.parent[ch]]) \
                + np.einsum("jab,jb->ja", B(ch), dv[fp.parent[ch]])
        p["x"], p["u"] = x, u

    def project_on_kernel(self, p):
        """cache.py:290-317.  For AVaR M = [alpha I, -I, 1, -I, -I] and M M' = (alpha^2+3) I + 1 1', so the
        orthogonal projector onto ker M has the closed form used here (equal to N lstsq(N, v) of the reference;
        tests/test_oracle_pinning.py checks that against the node oracle)."""
        fp = self.fp
        m = fp.m
        a = fp.risk_alpha
        c = fp.child_count.astype(float)
        par = fp.parent[1:]
        ya, yb = p["y"][fp.ya_idx[1:]], p["y"][fp.yb_idx[1:]]
        r = a[par] * ya - yb + p["y"][fp.ylast_idx][par] - p["tau"][1:] - p["s"][1:]   # M v, per edge
        rsum = fp.child_sum(r, 0, m)
        den = a * a + 3.0
        w = (r - (rsum / (den + c))[par]) / den[par]                                      # (M M')^-1 M v
        p["y"][fp.ya_idx[1:]] = ya - a[par] * w
        p["y"][fp.yb_idx[1:]] = yb + w
        p["y"][fp.ylast_idx] -= fp.child_sum(w, 0, m)
        p["tau"][1:] += w
        p["s"][1:] += w

    def proximal_of_f(self, p, alpha):
        """cache.py:248-257."""
        p["s"][0] -= alpha
        self.project_on_dynamics(p)
        self.project_on_kernel(p)

    # ------------------------------------------------------------------------------------------------ prox_g*
    @staticmethod
    def soc_rows(v):
        """Row-wise cones.py:113-132 (last column is t), same branch order."""
        z = v[:, :-1]
        t = v[:, -1]
        r = np.sqrt(np.sum(z * z, axis=1))
        out = v.copy()
        zero = (r > t) & (r <= -t)
        mid = (r > t) & ~zero
        out[zero] = 0.0
        tn = (r[mid] + t[mid]) / 2
        out[mid, :-1] = tn[:, None] * (z[mid] / r[mid][:, None])
        out[mid, -1] = tn
        return out

    @staticmethod
    def box(v, lo, hi):
        """rectangle.py:50-59 (NaN raises)."""
        if np.isnan(v).any():
            raise ValueError("Rectangle constraint - 'nan' value cannot be constrained")
        return np.where(v <= lo, lo, np.where(v >= hi, hi, v))

    def proximal_of_g_conjugate(self, d, alpha):
        """cache.py:321-393."""
        fp = self.fp
        nx, nu = fp.nx, fp.nu
        w = {k: v / alpha for k, v in d.items()}
        w[5] = w[5] - 0.5
        w[6] = w[6] + 0.5
        w[12] = w[12] - 0.5
        w[13] = w[13] + 0.5
        z = {}
        z1 = np.maximum(0.0, w[1])
        z1[fp.ylast_idx] = w[1][fp.ylast_idx]
        z[1] = z1
        z[2] = np.maximum(0.0, w[2])
        soc = self.soc_rows(np.hstack((w[3], w[4], w[5][:, None], w[6][:, None])))
        z[3], z[4], z[5], z[6] = soc[:, :nx], soc[:, nx: nx + nu], soc[:, nx + nu], soc[:, nx + nu + 1]
        if fp.nl_rect:
            z[7] = self.box(w[7], fp.nl_lo, fp.nl_hi)
        soc = self.soc_rows(np.hstack((w[11], w[12][:, None], w[13][:, None])))
        z[11], z[12], z[13] = soc[:, :nx], soc[:, nx], soc[:, nx + 1]
        if fp.leaf_rect:
            z[14] = self.box(w[14], fp.leaf_lo, fp.leaf_hi)
        return {k: alpha * (w[k] - z[k]) for k in w}

    # ------------------------------------------------------------------------------------------------ driver
    def cache_initial_state(self, x0):
        self.x0 = np.asarray(x0, dtype=float).reshape(-1).copy()
        self.p_old["x"][0] = self.x0

    def lambda_max(self):
        """lambda_max(L* L).  L* L is block diagonal (one block per node / edge), so the largest eigenvalue is the
        maximum over small dense symmetric blocks (equal to ARPACK's answer to 1e-15):
          [x_i; u_i]   : sum_j Q_j^(1/2)' Q_j^(1/2) (+ I if rectangles)  and same with R      (nonleaf)
          [y_i; s_i]   : [[I + b b', -b], [-b', 1]]                                           (nonleaf)
          tau_j        : 1/2;    leaf x: Qf + (I);   leaf s: 1/2."""
        fp = self.fp
        m, nx, nu = fp.m, fp.nx, fp.nu
        sq, sr = fp.sq_tab[fp.cost_idx[1:]], fp.sr_tab[fp.cost_idx[1:]]
        gx = fp.child_sum(np.einsum("jab,jac->jbc", sq, sq), 0, m) + (np.eye(nx) if fp.nl_rect else 0)
        gu = fp.child_sum(np.einsum("jab,jac->jbc", sr, sr), 0, m) + (np.eye(nu) if fp.nl_rect else 0)
        best = max(np.linalg.eigvalsh(gx).max(), np.linalg.eigvalsh(gu).max(), 0.5)
        gl = np.einsum("kab,kac->kbc", fp.sqf_tab, fp.sqf_tab) + (np.eye(nx) if fp.leaf_rect else 0)
        best = max(best, np.linalg.eigvalsh(gl).max())
        # [[I + b b', -b], [-b', 1]] has eigenvalues 1 (multiplicity) and the roots of
        # l^2 - (2 + |b|^2) l + 1 = 0  ->  l_max = (2 + |b|^2 + sqrt(|b|^4 + 4 |b|^2)) / 2
        bb = fp.child_sum(fp.pi[1:] ** 2, 0, m) + 1.0
        best = max(best, ((2 + bb + np.sqrt(bb * bb + 4 * bb)) / 2).max())
        return best

    def step_size(self):
        return 0.999 / self.lambda_max()

    @staticmethod
    def _axpy(a, x, y):
        return {k: a * x[k] + y[k] for k in x}

    def iterate(self):
        """One loop body of Solver.chock, solver.py:124-143."""
        a = self.alpha
        lt = self.ell_transpose(self.d_old)
        p = self._axpy(-a, lt, self.p_old)
        self.proximal_of_f(p, a)
        lp = self.ell({k: 2 * p[k] - self.p_old[k] for k in p})
        d = self.proximal_of_g_conjugate(self._axpy(a, lp, self.d_old), a)
        self.p, self.d = p, d
        xi, delta = self.residuals()
        self.p_old = {k: v.copy() for k, v in p.items()}
        self.d_old = {k: v.copy() for k, v in d.items()}
        return xi, delta

    def residuals(self):
        """solver.py:63-95,137-141."""
        a = self.alpha
        p_new, p, d_new, d = self.p, self.p_old, self.d, self.d_old
        dp = {k: p[k] - p_new[k] for k in p}
        dd = {k: d[k] - d_new[k] for k in d}
        lt = self.ell_transpose(dd)
        xi1 = {k: dp[k] / a - lt[k] for k in dp}
        pn = {k: p_new[k] - p[k] for k in p}
        lp = self.ell(pn)
        xi2 = {k: dd[k] / a + lp[k] for k in dd}
        lt2 = self.ell_transpose(xi2)
        xi0 = {k: xi1[k] + lt2[k] for k in xi1}
        delta2 = {k: d_new[k] - d[k] for k in d}
        lt3 = self.ell_transpose(delta2)
        delta0 = {k: pn[k] - lt3[k] for k in pn}
        nrm = lambda seg: max((np.max(np.abs(v)) if v.size else 0.0) for v in seg.values())  # noqa: E731
        return [nrm(xi0), nrm(xi1), nrm(xi2)], [nrm(delta0), nrm(pn), nrm(delta2)]

    def chock(self, x0, max_iters=10, tol=1e-5, alpha=None):
        """Solver.chock, solver.py:97-171 (runs max_iters+1 iterations if it does not converge)."""
        self.cache_initial_state(x0)
        self.alpha = self.step_size() if alpha is None else alpha
        xi_hist, delta_hist = [], []
        k = 0
        while True:
            xi, delta = self.iterate()
            xi_hist.append(xi)
            delta_hist.append(delta)
            if k >= max_iters or max(xi) <= tol:
                break
            k += 1
        return (0 if k < max_iters else 1), np.array(xi_hist), np.array(delta_hist)
