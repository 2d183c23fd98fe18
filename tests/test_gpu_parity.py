"""Parity of the CUDA path (through the C-ABI) against the oracles and the committed reference fixtures.

Bars (BASELINE.json north_star): iterates within 1e-9 relative (per segment, relative to the segment's inf-norm)
of the reference's float64 solver for the first 100 iterations; same iteration count +-1 to a tolerance.
"""
import numpy as np
import pytest

from helpers import golden, rel_err, seg_rel_err, spec_for

pytestmark = pytest.mark.gpu

TOL_ITERATE = 1e-9      # north_star: 1e-9 relative per iterate
TOL_OPERATOR = 1e-12    # single operator applications


def _build(name, **kw):
    import raocp_b200 as r
    from oracle import problems
    s = spec_for(name)
    return s, problems.build(s, r.core), r


@pytest.fixture(scope="module", params=["demo", "cfg1", "mini2", "mini3", "mini5", "dense", "wide"])
def case(request):
    s, problem, r = _build(request.param)
    cache = r.core.Cache(problem)
    return dict(name=request.param, spec=s, problem=problem, r=r, cache=cache, flat=cache.flat_problem,
                dev=cache.device_solver)


def test_offline_matches_reference(case):
    """P, K of the offline factorisation (cache.py:207-233) against the reference's own Cache"""
    g = golden(f"{case['name']}_iterates.npz")
    flat, dev = case["flat"], case["dev"]
    P, K, Rinv = dev.get_offline()
    cls = flat.cls
    for i in range(flat.m):
        assert rel_err(P[cls[i]], g["P"][i]) < 1e-12
        assert rel_err(K[cls[i]], g["K"][i]) < 1e-12


def test_offline_without_dedup_matches(case):
    r = case["r"]
    cache = r.core.Cache(case["problem"], dedup=False)
    g = golden(f"{case['name']}_iterates.npz")
    P, K, _ = cache.device_solver.get_offline()
    assert rel_err(P, g["P"][: cache.flat_problem.m]) < 1e-12
    assert rel_err(K, g["K"]) < 1e-12


def test_operators_match_reference(case):
    """L and L* on random vectors (operators.py:19-94) against the reference's outputs"""
    g = golden(f"{case['name']}_ops.npz")
    flat, dev = case["flat"], case["dev"]
    gather = flat.maps()["d_gather"]
    got = dev.apply_L(g["rand_p"])[0]
    assert seg_rel_err(flat, got, g["ell"][gather], dual=True) < TOL_OPERATOR
    got = dev.apply_Lt(g["rand_d"][gather])[0]
    assert seg_rel_err(flat, got, g["ell_t"], dual=False) < TOL_OPERATOR


def test_adjointness(case):
    """<L p, d> == <p, L* d> (reference tests/test_operators.py:118-335)"""
    flat, dev = case["flat"], case["dev"]
    rng = np.random.default_rng(5)
    p = rng.standard_normal(flat.np_)
    p[flat.n * flat.nx + flat.m * flat.nu + flat.ysz] = 0.0   # tau_0 is not a variable
    d = rng.standard_normal(flat.nd_)
    lhs = float(dev.apply_L(p)[0] @ d)
    rhs = float(p @ dev.apply_Lt(d)[0])
    assert abs(lhs - rhs) <= 1e-10 * max(1.0, abs(lhs))


def test_prox_f_pieces_match_reference(case):
    g = golden(f"{case['name']}_ops.npz")
    flat, dev, cache = case["flat"], case["dev"], case["cache"]
    cache.cache_initial_state(g["x0"])
    dev.set_primal(0, g["rand_p"])
    cache.project_on_dynamics()
    assert seg_rel_err(flat, dev.get_primal(0)[0], g["dyn"], dual=False) < 1e-11
    dev.set_primal(0, g["rand_p"])
    cache.project_on_kernel()
    assert seg_rel_err(flat, dev.get_primal(0)[0], g["ker"], dual=False) < 1e-12
    dev.set_primal(0, g["rand_p"])
    cache.proximal_of_f(0.37)
    assert seg_rel_err(flat, dev.get_primal(0)[0], g["proxf"], dual=False) < 1e-11


def test_prox_g_matches_reference(case):
    g = golden(f"{case['name']}_ops.npz")
    flat, dev, cache = case["flat"], case["dev"], case["cache"]
    gather = flat.maps()["d_gather"]
    dev.set_dual(0, g["big_d"][gather])
    cache.proximal_of_g_conjugate(0.37)
    assert seg_rel_err(flat, dev.get_dual(0)[0], g["proxg"][gather], dual=True) < 1e-12


def test_prox_g_piecewise_equals_fused(case):
    """modify_dual + add_halves + projections + modify_projection == proximal_of_g_conjugate (cache.py:321-327)"""
    g = golden(f"{case['name']}_ops.npz")
    flat, dev, cache = case["flat"], case["dev"], case["cache"]
    gather = flat.maps()["d_gather"]
    dev.set_dual(0, g["big_d"][gather])
    cache.modify_dual(0.37)
    cache.add_halves()
    modified = dev.get_dual(0)[0].copy()
    cache.project_on_constraints_nonleaf()
    cache.project_on_constraints_leaf()
    dev.modify_projection(0.37, modified)
    assert seg_rel_err(flat, dev.get_dual(0)[0], g["proxg"][gather], dual=True) < 1e-12


@pytest.mark.parametrize("mode", ["fused", "fused_unpipelined", "fused_tile_kernels", "fused_dense_costs", "fused_no_graph",
                                  "stepwise"])
def test_iterates_match_reference(case, mode):
    """first 100 iterates against the unmodified reference (same alpha, same x0), 1e-9 relative per segment"""
    g = golden(f"{case['name']}_iterates.npz")
    r, problem = case["r"], case["problem"]
    solver = r.core.Solver(problem, verbose=False)
    flat, dev = solver.cache.flat_problem, solver.cache.device_solver
    gather = flat.maps()["d_gather"]
    alpha = float(g["alpha"])
    keep = [int(k) for k in g["keep"]]
    worst = 0.0
    if mode.startswith("fused"):
        for k in keep:
            fresh = r.core.Solver(problem, verbose=False)
            if mode == "fused_dense_costs":
                fresh.cache.device_solver.force_dense_costs(True)
            if mode == "fused_tile_kernels":
                fresh.cache.device_solver.use_lane_kernels(False)
            if mode == "fused_no_graph":
                fresh.cache.device_solver.use_graphs(False)
            if mode == "fused_unpipelined":   # primal pass + one dual pass per iteration (no pbar hand-over)
                fresh.cache.device_solver.use_pipeline(False)
            status = fresh.chock(g["x0"], max_iters=k - 1, tol=0.0, alpha=alpha)
            assert status == 1 and fresh.iterations == k
            d2 = fresh.cache.device_solver
            worst = max(worst, seg_rel_err(flat, d2.get_primal(0)[0], g[f"p{k}"], dual=False),
                        seg_rel_err(flat, d2.get_dual(0)[0], g[f"d{k}"][gather], dual=True))
            xi, delta = fresh.residual_history
            assert rel_err(xi, g["xi"][:k]) < 1e-7 and rel_err(delta, g["delta"][:k]) < 1e-7
            assert np.max(np.abs(xi - g["xi"][:k]) / g["xi"][:k]) < 1e-6
    else:
        solver.cache.cache_initial_state(g["x0"])
        solver.set_step_size(alpha)
        for k in range(1, max(keep) + 1):
            solver.primal_k_plus_half()
            solver.primal_k_plus_one()
            solver.dual_k_plus_half()
            solver.dual_k_plus_one()
            norms, _ = dev.residuals(alpha)
            assert np.max(np.abs(norms[0, :3] - g["xi"][k - 1]) / g["xi"][k - 1]) < 1e-6
            if k in keep:
                worst = max(worst, seg_rel_err(flat, dev.get_primal(0)[0], g[f"p{k}"], dual=False),
                            seg_rel_err(flat, dev.get_dual(0)[0], g[f"d{k}"][gather], dual=True))
            solver.cache.update_cache()
    assert worst < TOL_ITERATE, worst


def test_demo_reproduces_golden_residual_history():
    """the reference's own artefact 4-3-residuals.tex: 937 iterations to max(xi) <= 1e-3 (main.py:80)"""
    s, problem, r = _build("demo")
    gold = golden("demo_residuals.npz")["xi"]
    solver = r.core.Solver(problem, verbose=False)
    status = solver.chock(s["x0"], max_iters=2000, tol=1e-3)   # own step size (device lambda_max)
    xi, _ = solver.residual_history
    assert status == 0
    assert abs(xi.shape[0] - gold.shape[0]) <= 1
    k = min(xi.shape[0], gold.shape[0])
    assert np.max(np.abs(xi[:k] - gold[:k]) / gold[:k]) < 1e-6


# iterations of the NumPy oracle (oracle/cp_flat_oracle.py, pinned to the reference) until max(xi) <= tol on cfg1, seed 0,
# alpha = 0.24574029027120411 -- recorded with the loop in the docstring below (21 s of CPU for the full 12 229 iterations)
CFG1_ITERATIONS_TO_TOL = {1e-3: 5133, 1e-4: 7011, 1e-5: 9672, 1e-6: 12229}


@pytest.mark.parametrize("tol", [1e-3, 1e-6])
def test_cfg1_converges_at_the_oracle_iteration(tol):
    """north star: convergence at the same iteration count +-1.  Recorded with
    `o = FlatOracle(problem); o.cache_initial_state(x0); o.alpha = o.step_size(); k = first k with max(o.iterate()[0]) <= tol`;
    the second solve checks that Solver.chock continues from the current iterate like the reference's (cache.py:79-82)."""
    s, problem, r = _build("cfg1")
    solver = r.core.Solver(problem, verbose=False)
    assert solver.chock(s["x0"][:, :1], max_iters=20000, tol=tol, alpha=0.24574029027120411) == 0
    assert abs(solver.iterations - CFG1_ITERATIONS_TO_TOL[tol]) <= 1
    assert np.max(solver.residual_history[0][-1]) <= tol < np.max(solver.residual_history[0][-2])
    assert solver.chock(s["x0"][:, :1], max_iters=20000, tol=tol, alpha=0.24574029027120411) == 0
    assert solver.iterations <= 50   # already converged: warm start


def test_step_size_matches_oracle(case):
    from oracle.cp_flat_oracle import FlatOracle
    lam = case["dev"].lambda_max()
    assert abs(0.999 / lam - float(golden(f"{case['name']}_iterates.npz")["alpha"])) < 1e-12
    assert abs(lam - FlatOracle(case["problem"]).lambda_max()) < 1e-12 * lam


def test_calculate_chock_errors_lists(case):
    """Solver._calculate_chock_errors returns the six block lists of solver.py:63-95"""
    g = golden(f"{case['name']}_iterates.npz")
    r = case["r"]
    solver = r.core.Solver(case["problem"], verbose=False)
    solver.cache.cache_initial_state(g["x0"])
    solver.set_step_size(float(g["alpha"]))
    solver.primal_k_plus_half()
    solver.primal_k_plus_one()
    solver.dual_k_plus_half()
    solver.dual_k_plus_one()
    lists = solver._calculate_chock_errors()
    norms = [max(np.max(np.abs(b)) for b in lst) for lst in lists]
    want = np.concatenate((g["xi"][0], g["delta"][0]))
    assert np.max(np.abs(np.array(norms) - want) / want) < 1e-9
    p_len, d_len = len(solver.cache.get_primal()[0]), len(solver.cache.get_dual()[0])
    assert [len(lst) for lst in lists] == [p_len, p_len, d_len, p_len, p_len, d_len]


@pytest.mark.parametrize("rectangles", [True, False])
def test_nan_in_the_initial_state_is_reported(rectangles):
    """a NaN must not be lost by the running maxima of the residual norms: with rectangles the reference raises
    ValueError in the projection (rectangle.py:58-59); without, the loop stops on a non-finite residual"""
    import raocp_b200 as r
    from oracle import problems
    s = dict(problems.spec("mini2"))
    s["rectangles"] = rectangles
    problem = problems.build(s, r.core)
    for lane_kernels in (True, False):
        solver = r.core.Solver(problem, verbose=False)
        solver.cache.device_solver.use_lane_kernels(lane_kernels)
        x0 = s["x0"][:, :1].copy()
        x0[0, 0] = np.nan
        with pytest.raises(ValueError if rectangles else Exception):
            solver.chock(x0, max_iters=5, tol=1e-3)
