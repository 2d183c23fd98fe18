"""RAOCP object -> flat structure-of-arrays problem description for the C-ABI (include/raocp_b200.h rb_problem).

Reads the problem once through the reference's accessor API (raocp_spec.py:56-75, scenario_tree.py:75-154):
topology to stage offsets + contiguous child ranges, per-node matrices de-duplicated into small tables by object
identity, AVaR parameters, rectangle bounds, and the factorisation classes (nodes that share (P, K, R~)).
Also holds the index maps between the reference's block lists (with (1,1) placeholders, cache.py:126-170) and the
compact exchange layout.
"""
import numpy as np

from . import risks as core_risks
from .constraints import rectangle as core_rectangle
from .constraints import no_constraint as core_no


def _real(mat):
    """sqrtm may hand back a complex dtype with zero imaginary part"""
    arr = np.asarray(mat)
    if np.iscomplexobj(arr):
        if np.max(np.abs(arr.imag)) > 1e-12 * max(1.0, np.max(np.abs(arr.real))):
            raise Exception("cost weight matrix has a complex square root")
        arr = arr.real
    return np.ascontiguousarray(arr, dtype=np.float64)


def _table(objs, getters):
    """de-duplicate by object identity: returns ([stacked unique matrices per getter], int32 index per object)"""
    seen, idx = {}, np.empty(len(objs), dtype=np.int32)
    uniq = []
    for k, o in enumerate(objs):
        key = id(o)
        pos = seen.get(key)
        if pos is None:
            pos = seen[key] = len(uniq)
            uniq.append(o)
        idx[k] = pos
    tabs = [np.ascontiguousarray(np.stack([_real(g(o)) for o in uniq])) for g in getters]
    return tabs, idx


class FlatProblem:
    def __init__(self, problem, batch=1, dedup=True, device=0, shard=None, sweep_cuts=None):
        tree = problem.tree
        self.problem = problem
        self.n = n = int(tree.num_nodes)
        self.m = m = int(tree.num_nonleaf_nodes)
        self.nleaf = n - m
        self.num_stages = int(tree.num_stages)
        self.batch = int(batch)
        self.device = int(device)
        self.shard_rank, self.shard_world = (int(shard[0]), int(shard[1])) if shard else (0, 1)
        self.sweep_cuts = (int(sweep_cuts[0]), int(sweep_cuts[1])) if sweep_cuts else (0, 0)   # rb_problem.sweep_cut*_min
        if problem.list_of_dynamics[1] is None:
            raise Exception("RAOCP has no dynamics")
        self.nx = int(problem.state_dynamics_at_node(1).shape[1])
        self.nu = int(problem.control_dynamics_at_node(1).shape[1])

        # ---- topology ------------------------------------------------------------------------------------------
        if hasattr(tree, "ancestors_array"):
            parent = np.asarray(tree.ancestors_array, dtype=np.int64)
            stages = np.asarray(tree.stages_array, dtype=np.int64)
        else:  # any object with the reference's query API
            parent = np.array([int(tree.ancestor_of(i)) for i in range(n)], dtype=np.int64)
            stages = np.array([int(tree.stage_of(i)) for i in range(n)], dtype=np.int64)
        if np.any(np.diff(stages) < 0) or np.any(np.diff(parent[1:]) < 0) or parent[0] != -1:
            raise Exception("raocp_b200 needs nodes numbered stage by stage with contiguous children "
                            "(the numbering MarkovChainScenarioTreeFactory produces)")
        self.parent = parent.astype(np.int32)
        self.stage_off = np.searchsorted(stages, np.arange(self.num_stages + 1)).astype(np.int32)
        counts = np.bincount(parent[1:], minlength=n)[:m]
        self.child_count = counts.astype(np.int32)
        self.child_first = (1 + np.concatenate(([0], np.cumsum(counts)[:-1]))).astype(np.int32)
        if counts.min() < 1 or counts.sum() != n - 1:
            raise Exception("every nonleaf node needs at least one child")

        # ---- matrices ------------------------------------------------------------------------------------------
        (self.A, self.B), dyn_idx = _table(problem.list_of_dynamics[1:],
                                           (lambda o: o.state_dynamics, lambda o: o.control_dynamics))
        self.dyn_idx = np.concatenate(([0], dyn_idx)).astype(np.int32)
        if None in problem.list_of_nonleaf_costs[1:] or None in problem.list_of_leaf_costs[m:]:
            raise Exception("RAOCP has no costs")
        (self.sqrtQ, self.sqrtR), cost_idx = _table(problem.list_of_nonleaf_costs[1:],
                                                    (lambda o: o.sqrt_state_weights, lambda o: o.sqrt_control_weights))
        self.cost_idx = np.concatenate(([0], cost_idx)).astype(np.int32)
        (self.sqrtQf,), self.leafcost_idx = _table(problem.list_of_leaf_costs[m:], (lambda o: o.sqrt_state_weights,))
        if self.A.shape[1:] != (self.nx, self.nx) or self.B.shape[1:] != (self.nx, self.nu) \
                or self.sqrtQ.shape[1:] != (self.nx, self.nx) or self.sqrtR.shape[1:] != (self.nu, self.nu) \
                or self.sqrtQf.shape[1:] != (self.nx, self.nx):
            raise Exception("dynamics / cost matrices have inconsistent shapes")

        # ---- constraints: Rectangle or No, the same kind on every node of a type -----------------------------------
        self.nl_rect = self._rectangles(problem.list_of_nonleaf_constraints[:m], self.nx + self.nu, "nonleaf")
        self.leaf_rect = self._rectangles(problem.list_of_leaf_constraints[m:], self.nx, "leaf")

        # ---- risks: AVaR only (cache.py:172-178) ------------------------------------------------------------------
        risk_alpha = np.empty(m)
        cond_prob = np.zeros(n)
        for i, risk in enumerate(problem.list_of_risks):
            if type(risk) is not core_risks.AVaR and type(risk).__name__ != "AVaR":
                raise Exception(f"Risk at node {i} not defined")
            risk_alpha[i] = risk.alpha
            c0 = self.child_first[i]
            cond_prob[c0: c0 + self.child_count[i]] = np.asarray(risk.probs, dtype=np.float64).reshape(-1)
        self.risk_alpha, self.cond_prob = risk_alpha, cond_prob

        # ---- factorisation classes ------------------------------------------------------------------------------
        self.dedup = bool(dedup)
        if dedup:
            cls = np.empty(m, dtype=np.int64)
            keys = {}
            dyn = self.dyn_idx.tolist()
            cf, cc = self.child_first.tolist(), self.child_count.tolist()
            cl = [-1] * n
            for i in range(m - 1, -1, -1):   # children first; ids assigned bottom-up, flipped below
                key = tuple((dyn[j], cl[j]) for j in range(cf[i], cf[i] + cc[i]))
                pos = keys.get(key)
                if pos is None:
                    pos = keys[key] = len(keys)
                cl[i] = pos
            self.num_cls = len(keys)
            cls[:] = cl[:m]
            self.cls = (self.num_cls - 1 - cls).astype(np.int32)   # parents get SMALLER ids than their children
        else:
            self.num_cls = m
            self.cls = np.arange(m, dtype=np.int32)

        # ---- ragged y layout and sizes ---------------------------------------------------------------------------
        self.ysize = 2 * counts + 1
        self.yoff = np.concatenate(([0], np.cumsum(self.ysize)))
        self.ysz = int(self.yoff[-1])
        nx, nu = self.nx, self.nu
        self.np_ = n * nx + m * nu + self.ysz + 2 * n
        self.nd_ = self.ysz + m + (n - 1) * (nx + nu + 2) + (m * (nx + nu) if self.nl_rect else 0) \
            + self.nleaf * (nx + 2) + (self.nleaf * nx if self.leaf_rect else 0)
        self._maps = None

    def _rectangles(self, cons, dim, what):
        active = [bool(c.is_active) for c in cons]
        if not any(active):
            setattr(self, f"{what}_lo", None)
            return False
        if not all(active):
            raise Exception(f"mixed active / inactive {what} constraints are not supported by raocp_b200")
        for c in cons:
            if not (isinstance(c, core_rectangle.Rectangle) or type(c).__name__ == "Rectangle"):
                raise Exception(f"{what} constraint type {type(c).__name__} is not supported by raocp_b200")

        def lo(c):
            return np.asarray(c.lower if hasattr(c, "lower") else c._Rectangle__min, dtype=np.float64).reshape(-1)

        def hi(c):
            return np.asarray(c.upper if hasattr(c, "upper") else c._Rectangle__max, dtype=np.float64).reshape(-1)

        (lo_tab, hi_tab), idx = _table(cons, (lo, hi))
        if lo_tab.shape[1] != dim:
            raise Exception("Rectangle constraint - input vector does not equal expected size")
        setattr(self, f"{what}_lo", lo_tab)
        setattr(self, f"{what}_hi", hi_tab)
        setattr(self, f"{what}_rect_idx", idx.astype(np.int32))
        return True

    # ---- reference block lists <-> compact vectors -----------------------------------------------------------------
    def block_sizes(self):
        """(primal block sizes, dual block sizes, dual 'real block' mask) of the reference's lists (cache.py:126-170)"""
        n, m, nx, nu = self.n, self.m, self.nx, self.nu
        ps = np.concatenate((np.full(n, nx), np.full(m, nu), self.ysize, np.ones(n, int), np.ones(n, int)))
        node = np.arange(n)
        nonleaf, leaf, edge = node < m, node >= m, node > 0
        ys = np.ones(n, int)
        ys[:m] = self.ysize
        segs = [
            (ys, nonleaf),                                                       # 1
            (np.ones(n, int), nonleaf),                                          # 2
            (np.where(edge, nx, 1), edge),                                       # 3
            (np.where(edge, nu, 1), edge),                                       # 4
            (np.ones(n, int), edge),                                             # 5
            (np.ones(n, int), edge),                                             # 6
            (np.where(nonleaf & self.nl_rect, nx + nu, 1), nonleaf & self.nl_rect),   # 7
            (np.where(leaf, nx, 1), leaf),                                       # 11
            (np.ones(n, int), leaf),                                             # 12
            (np.ones(n, int), leaf),                                             # 13
            (np.where(leaf & self.leaf_rect, nx, 1), leaf & self.leaf_rect),     # 14
        ]
        ds = np.concatenate([s for s, _ in segs])
        real = np.concatenate([r for _, r in segs])
        return ps, ds, real

    def maps(self):
        """index arrays: compact_dual = ref_flat_dual[dual_gather]; primal needs none (same order, no placeholders)"""
        if self._maps is None:
            ps, ds, real = self.block_sizes()
            starts = np.concatenate(([0], np.cumsum(ds)))
            keep = np.flatnonzero(real)
            gather = np.concatenate([np.arange(starts[b], starts[b + 1]) for b in keep]) if keep.size else np.zeros(0, int)
            assert gather.size == self.nd_
            self._maps = dict(p_sizes=ps, d_sizes=ds, d_real=real, d_gather=gather, d_total=int(starts[-1]),
                              p_starts=np.concatenate(([0], np.cumsum(ps))), d_starts=starts)
        return self._maps

    def primal_from_blocks(self, blocks):
        return np.ascontiguousarray(np.concatenate([np.asarray(b, dtype=np.float64).reshape(-1) for b in blocks]))

    def dual_from_blocks(self, blocks):
        flat = np.concatenate([np.asarray(b, dtype=np.float64).reshape(-1) for b in blocks])
        return np.ascontiguousarray(flat[self.maps()["d_gather"]])

    def primal_to_blocks(self, compact):
        st = self.maps()["p_starts"]
        return [compact[st[k]: st[k + 1]].reshape(-1, 1).copy() for k in range(len(st) - 1)]

    def dual_to_blocks(self, compact):
        mp = self.maps()
        flat = np.zeros(mp["d_total"])
        flat[mp["d_gather"]] = compact
        st = mp["d_starts"]
        return [flat[st[k]: st[k + 1]].reshape(-1, 1).copy() for k in range(len(st) - 1)]

    # ---- subtree sharding: ownership of the compact entries, given the cut stage THE DEVICE chose ---------------------------
    # The cut rule lives in rb_create only (csrc/api.cu: balanced first cut of the sweep plan); DeviceSolver stores the
    # answer of rb_shard_info() in `shard_cut` right after rb_create.  Host-only tests pass a cut stage explicitly.
    shard_cut = None

    def shard_cut_stage(self):
        if self.shard_cut is None:
            raise Exception("the cut stage is decided by the device (rb_shard_info): create the DeviceSolver first "
                            "or set FlatProblem.shard_cut")
        return int(self.shard_cut)

    def shard_owned_nodes(self, rank, world):
        """bool[n]: nodes whose node-indexed quantities rank `rank` owns (its subtrees below the cut stage; rank 0 also
        answers for the replicated top of the tree)"""
        t_c = self.shard_cut_stage()
        first, count = int(self.stage_off[t_c]), int(self.stage_off[t_c + 1] - self.stage_off[t_c])
        if count < world:
            raise Exception("fewer cut-stage subtrees than ranks")
        a, b = first + rank * count // world, first + (rank + 1) * count // world
        own = np.zeros(self.n, dtype=bool)
        if rank == 0:
            own[:first] = True
        for t in range(t_c, self.num_stages):
            own[a:b] = True
            if t + 1 < self.num_stages:
                a, b = int(self.child_first[a]), int(self.child_first[b - 1] + self.child_count[b - 1])
        return own

    def shard_masks(self, rank, world):
        """(primal mask, dual mask) over the compact layouts: entries for which rank `rank` holds the authoritative
        value after a sharded solve.  Node-indexed entries follow the node, edge-indexed entries (tau_j, s_j, d3..d6 of
        edge j) follow the PARENT (they are computed by the parent's thread); the masks of all ranks partition both
        vectors."""
        own = self.shard_owned_nodes(rank, world)
        n, m, nx, nu = self.n, self.m, self.nx, self.nu
        par = self.parent.copy()
        par[0] = 0
        edge_own = own[par]              # entry j: owner of node j's parent (entry 0: the root itself)
        edge_own[0] = own[0]
        ymask = np.repeat(own[:m], self.ysize)
        pm = np.concatenate((np.repeat(own, nx), np.repeat(own[:m], nu), ymask, edge_own, edge_own))
        parts = [ymask, own[:m], np.repeat(edge_own[1:], nx), np.repeat(edge_own[1:], nu), edge_own[1:], edge_own[1:]]
        if self.nl_rect:
            parts.append(np.repeat(own[:m], nx + nu))
        parts += [np.repeat(own[m:], nx), own[m:], own[m:]]
        if self.leaf_rect:
            parts.append(np.repeat(own[m:], nx))
        dm = np.concatenate(parts)
        assert pm.size == self.np_ and dm.size == self.nd_
        return pm, dm
