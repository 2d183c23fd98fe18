"""DP sweeps (cache.py:259-288) on trees big / chain-like enough to reach every kernel of the sweep plan: the subtree
levels out of shared memory (tree_sweeps.cu) or global memory (sweeps.cu), the one-warp-per-chain walker (sweeps.cu) and
the eight-chains-per-warp tensor-core walker (chain_mma.cu), each
against the NumPy oracle (pinned to the reference by tests/test_oracle_pinning.py), through the C-ABI."""
import numpy as np
import pytest

from helpers import seg_rel_err

pytestmark = pytest.mark.gpu

# (problem, sweep_cuts): cut stage minima so that small trees still get a top part, a branching level and a chain level
CASES = [
    ("chain21", (2, 8)),        # nx=2, nu=1 (scalar loads), 8 chains = exactly one tile
    ("chain32", (3, 9)),        # odd nx, 9 chains: a full tile and a padded one
    ("mini2", (9, 27)),         # nx=4, nu=2, level 0 of depth 1
    ("chain63", (16, 64)),      # sizes without an instantiation: run-time-size walker
    ("shard", (64, 200)),       # nx=8, nu=4
    ("chain105", (64, 200)),    # nx=10, nu=5 (odd nu)
    ("chain2010", None),        # default plan: cut at 64 nodes, 256 chains
    ("chain6432", (9, 27)),     # nx=64, nu=32 (cfg5's sizes): tensor-core walker with the fragments in shared memory
    ("cfg2", None),             # default plan: one tree level + 243 chains
    ("wide", (2, 4)),           # nx=40, nu=36: rows wider than a warp, run-time-size kernels
    ("cfg1", None),             # 31 nodes: the whole tree is the "top"
]


MODES = ["mma", "mma_fused_check", "mma_launch_overlap", "mma_four_warps", "mma_one_warp_wide_rows", "warp_per_chain", "tree_per_level",
         "global_stage_kernels"]


def _pair(name, cuts, mode, batch=1):
    import raocp_b200 as r
    from oracle import problems
    from oracle.cp_flat_oracle import FlatOracle
    s = problems.spec(name, batch=batch)
    problem = problems.build(s, r.core)
    solver = r.core.Solver(problem, verbose=False, sweep_cuts=cuts, batch=batch)
    # chain_mma.cu (one warp per tile; four where instantiated: nx=20, nu=10) vs the sweeps.cu chain walker
    solver.cache.device_solver.use_mma_sweeps({"mma": 1, "mma_fused_check": 1, "mma_launch_overlap": 1, "mma_four_warps": 2, "mma_one_warp_wide_rows": 3}.get(mode, 0))
    solver.cache.device_solver.use_fused_check(mode == "mma_fused_check")   # stopping test by the last CTA of the dual passes (ablation)
    solver.cache.device_solver.use_launch_overlap(mode == "mma_launch_overlap")   # programmatic dependent launch (ablation)
    # tree_sweeps.cu fused with the top (default) / one launch per level, or the sweeps.cu stage kernels
    solver.cache.device_solver.use_tree_kernels({"global_stage_kernels": 0, "tree_per_level": 1}.get(mode, 2))
    return s, solver, FlatOracle(problem)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name,cuts", CASES, ids=[c[0] for c in CASES])
def test_projection_on_dynamics(name, cuts, mode):
    s, solver, oracle = _pair(name, cuts, mode)
    cache, dev, flat = solver.cache, solver.cache.device_solver, solver.cache.flat_problem
    x0 = s["x0"][:, :1]
    rng = np.random.default_rng(5)
    vec = rng.standard_normal(flat.np_)
    cache.cache_initial_state(x0)
    dev.set_primal(0, vec)
    cache.project_on_dynamics()
    oracle.cache_initial_state(x0)
    p = oracle.unflat_primal(vec)
    oracle.project_on_dynamics(p)
    assert seg_rel_err(flat, dev.get_primal(0)[0], oracle.flat_primal(p), dual=False) < 1e-11


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name,cuts", CASES, ids=[c[0] for c in CASES])
def test_iterates(name, cuts, mode):
    """30 Chambolle-Pock iterations, iterates within 1e-9 (north star) of the oracle"""
    s, solver, oracle = _pair(name, cuts, mode)
    flat, dev = solver.cache.flat_problem, solver.cache.device_solver
    x0 = s["x0"][:, :1]
    alpha = oracle.step_size()
    assert abs(solver.compute_step_size() - alpha) <= 1e-11 * alpha
    assert solver.chock(x0, max_iters=29, tol=0.0, alpha=alpha) == 1 and solver.iterations == 30
    oracle.cache_initial_state(x0)
    oracle.alpha = alpha
    for _ in range(30):
        xi, delta = oracle.iterate()
    assert seg_rel_err(flat, dev.get_primal(0)[0], oracle.flat_primal(oracle.p), dual=False) < 1e-9
    assert seg_rel_err(flat, dev.get_dual(0)[0], oracle.flat_dual(oracle.d), dual=True) < 1e-9
    got = np.concatenate((solver.residual_history[0][-1], solver.residual_history[1][-1]))
    want = np.concatenate((np.array(xi), np.array(delta)))
    assert np.max(np.abs(got - want) / want) < 1e-6


def test_batched_instances_through_chain_tiles():
    """batch > 1: every instance walks the same tiles on its own rows"""
    from oracle.cp_flat_oracle import FlatOracle
    s, solver, oracle = _pair("chain2010", None, "mma", batch=3)
    flat, dev = solver.cache.flat_problem, solver.cache.device_solver
    alpha = oracle.step_size()
    x0 = s["x0"][:, :3]
    assert solver.chock(x0, max_iters=9, tol=0.0, alpha=alpha) == 1
    for b in range(3):
        oracle = FlatOracle(flat.problem)       # fresh zero iterates
        oracle.cache_initial_state(x0[:, b:b + 1])
        oracle.alpha = alpha
        for _ in range(10):
            oracle.iterate()
        assert seg_rel_err(flat, dev.get_primal(0)[b], oracle.flat_primal(oracle.p), dual=False) < 1e-9
        assert seg_rel_err(flat, dev.get_dual(0)[b], oracle.flat_dual(oracle.d), dual=True) < 1e-9


@pytest.mark.parametrize("graphs", [True, False])
@pytest.mark.parametrize("name", ["chain2010", "shard", "mini2", "wide", "cfg1"])
def test_pipelined_loop_equals_unpipelined(name, graphs):
    """the pipelined loop (dual pass hands pbar to the next iteration; branching / chain / leaf dual kernels) against
    the loop with a primal pass per iteration: same iterates to rounding, same residual history, and both 1e-9 from
    the oracle; a stopping tolerance is met at the same iteration"""
    import raocp_b200 as r
    from oracle import problems
    from oracle.cp_flat_oracle import FlatOracle
    s = problems.spec(name)
    problem = problems.build(s, r.core)
    x0 = s["x0"][:, :1]
    oracle = FlatOracle(problem)
    alpha = oracle.step_size()
    out = {}
    for pipe in (True, False, 3, 4):   # 3: forward chain walk in two pieces; 4: risk block inside the chain dual pass
        solver = r.core.Solver(problem, verbose=False)
        dev = solver.cache.device_solver
        dev.use_pipeline(pipe)
        dev.use_graphs(graphs)
        assert solver.chock(x0, max_iters=40, tol=0.0, alpha=alpha) == 1 and solver.iterations == 41
        out[pipe] = (dev.get_primal(0)[0], dev.get_dual(0)[0], solver.residual_history[0].copy(), solver)
    flat = out[True][3].cache.flat_problem
    for other in (False, 3, 4):
        assert seg_rel_err(flat, out[True][0], out[other][0], dual=False) < 1e-12
        assert seg_rel_err(flat, out[True][1], out[other][1], dual=True) < 1e-12
        assert np.max(np.abs(out[True][2] - out[other][2]) / out[other][2]) < 1e-9
    oracle.cache_initial_state(x0)
    oracle.alpha = alpha
    for _ in range(41):
        oracle.iterate()
    assert seg_rel_err(flat, out[True][0], oracle.flat_primal(oracle.p), dual=False) < 1e-9
    assert seg_rel_err(flat, out[True][1], oracle.flat_dual(oracle.d), dual=True) < 1e-9
    # stopping test: a tolerance first met inside the recorded history stops both loops at the same iteration
    tol = float(np.sort(out[False][2].max(axis=1))[3]) * (1 + 1e-9)
    its = []
    for pipe in (True, False):
        solver = r.core.Solver(problem, verbose=False)
        solver.cache.device_solver.use_pipeline(pipe)
        solver.cache.device_solver.use_graphs(graphs)
        assert solver.chock(x0, max_iters=200, tol=tol, alpha=alpha) == 0
        its.append(solver.iterations)
    assert its[0] == its[1]


def test_pipelined_batch_and_second_solve():
    """batch > 1 through the pipelined kernels, and a second chock() on the same solver (the pbar hand-over must not
    leak from one loop into the next)"""
    from oracle.cp_flat_oracle import FlatOracle
    s, solver, oracle = _pair("chain2010", None, "mma", batch=2)
    flat, dev = solver.cache.flat_problem, solver.cache.device_solver
    alpha = oracle.step_size()
    oracles = [FlatOracle(flat.problem) for _ in range(2)]
    for x0 in (s["x0"][:, :2], 0.5 - s["x0"][:, :2]):   # the second solve warm-starts from the first one's iterate
        assert solver.chock(x0, max_iters=11, tol=0.0, alpha=alpha) == 1
        for b in range(2):
            oracles[b].cache_initial_state(x0[:, b:b + 1])
            oracles[b].alpha = alpha
            for _ in range(12):
                oracles[b].iterate()
            assert seg_rel_err(flat, dev.get_primal(0)[b], oracles[b].flat_primal(oracles[b].p), dual=False) < 1e-9
            assert seg_rel_err(flat, dev.get_dual(0)[b], oracles[b].flat_dual(oracles[b].d), dual=True) < 1e-9


def test_profile_hook_advances_like_an_iteration():
    """rb_profile_iteration (plain launches with events) is one ordinary iteration of the loop"""
    s, solver, oracle = _pair("chain2010", None, "mma")
    flat, dev = solver.cache.flat_problem, solver.cache.device_solver
    alpha = oracle.step_size()
    x0 = s["x0"][:, :1]
    solver.cache.cache_initial_state(x0)
    dev.loop_begin(alpha, 1 << 30, -1.0, 0)
    dev.loop_enqueue(3)
    phases, parts = dev.profile_iteration_full()
    assert len(phases) >= 3 and len(parts) in (1, 3) and all(v > 0 for v in phases + parts)
    dev.loop_enqueue(2)
    dev.profile_iteration()
    dev.loop_end()
    oracle.cache_initial_state(x0)
    oracle.alpha = alpha
    for _ in range(7):
        oracle.iterate()
    assert seg_rel_err(flat, dev.get_primal(0)[0], oracle.flat_primal(oracle.p), dual=False) < 1e-9
    assert seg_rel_err(flat, dev.get_dual(0)[0], oracle.flat_dual(oracle.d), dual=True) < 1e-9


@pytest.mark.parametrize("pipe", [True, False])
def test_step_with_a_new_initial_state_mid_loop(pipe):
    """rb_step (the end-to-end call of bench.py: x0 from host memory, one iteration, six norms back) with x0 CHANGED in the
    middle of a running loop -- the receding-horizon use.  In the pipelined loop the half step of the next iteration has
    already been written when the new x0 arrives; it must not matter (xbar_0 is never used by the projection)."""
    import torch
    s, solver, oracle = _pair("chain2010", None, "mma")
    flat, dev = solver.cache.flat_problem, solver.cache.device_solver
    dev.use_pipeline(pipe)
    alpha = oracle.step_size()
    xa = s["x0"][:, :1]
    xb = 0.25 - 0.5 * xa
    solver.cache.cache_initial_state(xa)
    dev.loop_begin(alpha, 1 << 30, -1.0, 0)
    x_host = torch.zeros(1, flat.nx, dtype=torch.float64).pin_memory()
    n_host = torch.zeros(1, 6, dtype=torch.float64).pin_memory()
    oracle.alpha = alpha
    for x0, steps in ((xa, 4), (xb, 5), (xa, 3)):
        x_host[0] = torch.from_numpy(x0[:, 0].copy())
        oracle.cache_initial_state(x0)
        for _ in range(steps):
            dev.step(x_host.data_ptr(), n_host.data_ptr())
            xi, delta = oracle.iterate()
            want = np.concatenate((np.array(xi), np.array(delta)))
            assert np.max(np.abs(n_host.numpy()[0] - want) / want) < 1e-6
    dev.loop_end()
    assert seg_rel_err(flat, dev.get_primal(0)[0], oracle.flat_primal(oracle.p), dual=False) < 1e-9
    assert seg_rel_err(flat, dev.get_dual(0)[0], oracle.flat_dual(oracle.d), dual=True) < 1e-9


@pytest.mark.parametrize("graphs", [True, False])
@pytest.mark.parametrize("name", ["cfg1", "chain2010"])
def test_fused_check_stops_like_the_check_launch(name, graphs):
    """rb_use_fused_check(1): the stopping test of solver.py:156-161 run by the last CTA of the dual passes -- same stopping
    iteration, same status, same residual history and iterates as with the k_check launch, with a tolerance, with the iteration cap,
    across a second chock() (warm start) and through rb_step"""
    import raocp_b200 as r
    import torch
    from oracle import problems
    s = problems.spec(name)
    problem = problems.build(s, r.core)
    x0 = s["x0"][:, :1]
    out = {}
    for fused in (False, True):
        solver = r.core.Solver(problem, verbose=False)
        dev = solver.cache.device_solver
        dev.use_fused_check(fused)
        dev.use_graphs(graphs)
        alpha = solver.compute_step_size()
        st_cap = solver.chock(x0, max_iters=40, tol=0.0, alpha=alpha)            # the cap fires: M + 1 iterations
        hist_cap = solver.residual_history[0].copy()
        n_cap = solver.iterations
        st_warm = solver.chock(x0, max_iters=9, tol=0.0, alpha=alpha)            # warm start from iterate 41
        hist_warm = solver.residual_history[0].copy()
        # a tolerance first met inside the recorded window, fresh solver
        tol = float(np.sort(hist_cap.max(axis=1))[1]) * (1 + 1e-9)
        fresh = r.core.Solver(problem, verbose=False)
        fresh.cache.device_solver.use_fused_check(fused)
        fresh.cache.device_solver.use_graphs(graphs)
        st_tol = fresh.chock(x0, max_iters=400, tol=tol, alpha=alpha)
        n_tol = fresh.iterations
        # rb_step: norms of every step on the host
        dev.loop_begin(alpha, 1 << 30, -1.0, 0)
        xh = torch.from_numpy(np.ascontiguousarray(x0.reshape(1, -1))).pin_memory()
        nh = torch.zeros(1, 6, dtype=torch.float64).pin_memory()
        steps = []
        for _ in range(5):
            dev.step(xh.data_ptr(), nh.data_ptr())
            steps.append(nh.numpy().copy())
        dev.loop_end()
        out[fused] = (st_cap, n_cap, hist_cap, st_warm, hist_warm, st_tol, n_tol, fresh.residual_history[0].copy(),
                      dev.get_primal(0)[0], dev.get_dual(0)[0], np.array(steps))
    a, b = out[False], out[True]
    assert (a[0], a[1]) == (b[0], b[1]) == (1, 41)
    assert np.array_equal(a[2], b[2])
    assert a[3] == b[3] == 1 and a[4].shape[0] == 10 and np.array_equal(a[4], b[4])
    assert (a[5], a[6]) == (b[5], b[6]) and a[5] == 0 and 1 < a[6] <= 41
    assert np.array_equal(a[7], b[7]) and np.array_equal(a[8], b[8]) and np.array_equal(a[9], b[9])
    assert np.array_equal(a[10], b[10]) and np.all(a[10] > 0)


@pytest.mark.parametrize("n_gpus,horizon,tau", [(2, 8, 4), (8, 7, 5)])
def test_widened_tree_sibling_matches_oracle(n_gpus, horizon, tau):
    """small siblings of the trees `bench.py --gpus N` shards (oracle/problems.py wide_spec: block-diagonal transition matrix, so the
    root has eight children and every other node four): default kernels, 30 iterations, 1e-9 from the oracle"""
    import raocp_b200 as r
    from oracle import problems
    from oracle.cp_flat_oracle import FlatOracle
    s = problems.wide_spec(n_gpus, horizon=horizon, tau=tau)
    problem = problems.build(s, r.core)
    x0 = s["x0"][:, :1]
    oracle = FlatOracle(problem)
    alpha = oracle.step_size()
    solver = r.core.Solver(problem, verbose=False)
    flat, dev = solver.cache.flat_problem, solver.cache.device_solver
    assert int(np.max(flat.child_count)) == 8 and int(flat.child_count[1]) == 4
    assert abs(solver.compute_step_size() - alpha) <= 1e-11 * alpha
    assert solver.chock(x0, max_iters=29, tol=0.0, alpha=alpha) == 1 and solver.iterations == 30
    oracle.cache_initial_state(x0)
    oracle.alpha = alpha
    for _ in range(30):
        xi, delta = oracle.iterate()
    assert seg_rel_err(flat, dev.get_primal(0)[0], oracle.flat_primal(oracle.p), dual=False) < 1e-9
    assert seg_rel_err(flat, dev.get_dual(0)[0], oracle.flat_dual(oracle.d), dual=True) < 1e-9
    got = np.concatenate((solver.residual_history[0][-1], solver.residual_history[1][-1]))
    want = np.concatenate((np.array(xi), np.array(delta)))
    assert np.max(np.abs(got - want) / want) < 1e-6
