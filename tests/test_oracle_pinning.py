"""Pin the oracles (oracle/cp_node_oracle.py, oracle/cp_flat_oracle.py) to the reference:
  * the reference's own golden artefact 4-3-residuals.tex (committed as tests/golden/demo_residuals.npz),
  * iterates / operator outputs recorded from the unmodified reference (oracle/make_golden.py),
  * and, when /root/reference is present (build container only), the live reference itself.
CPU only."""
import numpy as np
import pytest

from helpers import golden, rel_err, spec_for
from oracle import problems, ref_loader
from oracle.cp_flat_oracle import FlatOracle
from oracle.cp_node_oracle import NodeOracle

import raocp_b200 as r

NAMES = ["demo", "cfg1", "mini2", "mini3", "mini5", "dense", "wide"]


def _problem(name):
    s = spec_for(name)
    return s, problems.build(s, r.core)


@pytest.mark.parametrize("name", NAMES)
def test_flat_oracle_iterates_match_reference_fixture(name):
    s, problem = _problem(name)
    g = golden(f"{name}_iterates.npz")
    orc = FlatOracle(problem)
    orc.cache_initial_state(g["x0"])
    orc.alpha = float(g["alpha"])
    keep = set(int(k) for k in g["keep"])
    for k in range(1, 101):
        xi, delta = orc.iterate()
        assert np.max(np.abs(np.array(xi) - g["xi"][k - 1]) / g["xi"][k - 1]) < 1e-9
        assert np.max(np.abs(np.array(delta) - g["delta"][k - 1]) / g["delta"][k - 1]) < 1e-9
        if k in keep:
            # raw reference vectors: primal has no placeholders; dual blocks are re-ordered by the oracle itself
            assert rel_err(orc.flat_primal(orc.p), g[f"p{k}"]) < 1e-12
            blocks = _raw_dual_to_blocks(orc, g[f"d{k}"])
            assert rel_err(orc.flat_dual(orc.d), orc.dual_from_blocks(blocks)) < 1e-12


def _raw_dual_to_blocks(orc, raw):
    """split the reference's flat dual (np.vstack of its block list, cache.py:140-170) back into blocks"""
    fp = orc.fp
    n, m, nx, nu = fp.n, fp.m, fp.nx, fp.nu
    sizes = []
    for part in (1, 2, 3, 4, 5, 6, 7, 11, 12, 13, 14):
        for i in range(n):
            if part == 1:
                sizes.append(2 * fp.child_count[i] + 1 if i < m else 1)
            elif part == 3:
                sizes.append(nx if i > 0 else 1)
            elif part == 4:
                sizes.append(nu if i > 0 else 1)
            elif part == 7:
                sizes.append(nx + nu if (i < m and fp.nl_rect) else 1)
            elif part == 11:
                sizes.append(nx if i >= m else 1)
            elif part == 14:
                sizes.append(nx if (i >= m and fp.leaf_rect) else 1)
            else:
                sizes.append(1)
    cuts = np.concatenate(([0], np.cumsum(sizes)))
    assert cuts[-1] == raw.size
    return [raw[cuts[k]: cuts[k + 1]] for k in range(len(sizes))]


@pytest.mark.parametrize("name", ["demo", "cfg1", "mini2"])
def test_node_oracle_iterates_match_reference_fixture(name):
    s, problem = _problem(name)
    g = golden(f"{name}_iterates.npz")
    orc = NodeOracle(problem)
    orc.cache_initial_state(g["x0"])
    orc.alpha = float(g["alpha"])
    keep = set(int(k) for k in g["keep"])
    for k in range(1, 21):
        xi, delta = orc.iterate()
        assert np.max(np.abs(np.array(xi) - g["xi"][k - 1]) / g["xi"][k - 1]) < 1e-9
        if k in keep:
            assert rel_err(np.vstack(orc.primal_blocks()).reshape(-1), g[f"p{k}"]) < 1e-12
            assert rel_err(np.vstack(orc.dual_blocks()).reshape(-1), g[f"d{k}"]) < 1e-12


@pytest.mark.parametrize("name", ["demo", "mini2"])
def test_offline_matches_reference_fixture(name):
    s, problem = _problem(name)
    g = golden(f"{name}_iterates.npz")
    flat, node = FlatOracle(problem), NodeOracle(problem)
    assert rel_err(flat.P, g["P"]) < 1e-12 and rel_err(flat.K, g["K"]) < 1e-12
    assert rel_err(flat.Abar[1:], g["Abar"]) < 1e-12
    assert rel_err(np.stack(node.P), g["P"]) < 1e-13 and rel_err(np.stack(node.K), g["K"]) < 1e-13


def test_demo_golden_residual_history_and_iteration_count():
    """4-3-residuals.tex: the flat oracle reproduces all 937 x 3 numbers and stops at the same iteration"""
    s, problem = _problem("demo")
    gold = golden("demo_residuals.npz")["xi"]
    orc = FlatOracle(problem)
    status, xi, _ = orc.chock(s["x0"], max_iters=2000, tol=1e-3)
    assert status == 0 and xi.shape == gold.shape
    assert np.max(np.abs(xi - gold) / gold) < 1e-9


def test_closed_form_step_size_matches_arpack_fixture():
    for name in NAMES:
        s, problem = _problem(name)
        g = golden(f"{name}_iterates.npz")
        assert abs(FlatOracle(problem).step_size() - float(g["alpha"])) < 1e-13


@pytest.mark.parametrize("name", ["demo", "mini3"])
def test_operator_fixtures(name):
    s, problem = _problem(name)
    g = golden(f"{name}_ops.npz")
    orc = FlatOracle(problem)
    p = orc.unflat_primal(g["rand_p"])
    d_blocks = _raw_dual_to_blocks(orc, g["rand_d"])
    d = orc.unflat_dual(orc.dual_from_blocks(d_blocks))
    want = orc.dual_from_blocks(_raw_dual_to_blocks(orc, g["ell"]))
    assert rel_err(orc.flat_dual(orc.ell(p)), want) < 1e-13
    assert rel_err(orc.flat_primal(orc.ell_transpose(d)), g["ell_t"]) < 1e-13
    orc.x0 = g["x0"].reshape(-1)
    q = orc.unflat_primal(g["rand_p"])
    orc.proximal_of_f(q, 0.37)
    assert rel_err(orc.flat_primal(q), g["proxf"]) < 1e-12
    big = orc.unflat_dual(orc.dual_from_blocks(_raw_dual_to_blocks(orc, g["big_d"])))
    want = orc.dual_from_blocks(_raw_dual_to_blocks(orc, g["proxg"]))
    assert rel_err(orc.flat_dual(orc.proximal_of_g_conjugate(big, 0.37)), want) < 1e-13


def test_kernel_projection_properties():
    """restatement of reference tests/test_cache.py:161-209 without cvxpy: result lies in ker M, the projection is
    idempotent and the residual v - proj is orthogonal to ker M (so it IS the least-squares projection)"""
    s, problem = _problem("mini2")
    node = NodeOracle(problem)
    flat = FlatOracle(problem)
    rng = np.random.default_rng(3)
    p = flat.unflat_primal(rng.standard_normal(flat.fp.np_))
    before = {k: v.copy() for k, v in p.items()}
    flat.project_on_kernel(p)
    again = {k: v.copy() for k, v in p.items()}
    flat.project_on_kernel(again)
    for k in p:
        assert np.allclose(p[k], again[k], atol=1e-13)
    for i in range(flat.fp.m):
        ch = np.arange(flat.fp.child_first[i], flat.fp.child_first[i] + flat.fp.child_count[i])
        ysl = slice(flat.fp.yoff[i], flat.fp.yoff[i + 1])
        v0 = np.concatenate((before["y"][ysl], before["tau"][ch], before["s"][ch]))
        v1 = np.concatenate((p["y"][ysl], p["tau"][ch], p["s"][ch]))
        assert np.max(np.abs(node.M[i] @ v1)) < 1e-12
        assert np.max(np.abs(node.N[i].T @ (v0 - v1))) < 1e-12


def test_dynamics_projection_kkt():
    """restatement of reference tests/test_cache.py:111-159 without cvxpy: the DP result satisfies the dynamics and the
    residual (xbar - x, ubar - u) is orthogonal to every feasible direction (KKT of the equality-constrained LS)"""
    s, problem = _problem("mini2")
    flat = FlatOracle(problem)
    fp = flat.fp
    rng = np.random.default_rng(4)
    p = flat.unflat_primal(rng.standard_normal(fp.np_))
    flat.x0 = p["x"][0].copy()
    xbar, ubar = p["x"].copy(), p["u"].copy()
    flat.project_on_dynamics(p)
    x, u = p["x"], p["u"]
    assert np.allclose(x[0], xbar[0])
    for j in range(1, fp.n):
        i = fp.parent[j]
        assert np.allclose(x[j], fp.A_tab[fp.dyn_idx[j]] @ x[i] + fp.B_tab[fp.dyn_idx[j]] @ u[i], atol=1e-12)
    for _ in range(5):  # random feasible direction: dx0 = 0, du random, dx by the dynamics
        du = rng.standard_normal(u.shape)
        dx = np.zeros_like(x)
        for j in range(1, fp.n):
            i = fp.parent[j]
            dx[j] = fp.A_tab[fp.dyn_idx[j]] @ dx[i] + fp.B_tab[fp.dyn_idx[j]] @ du[i]
        inner = np.sum((xbar - x) * dx) + np.sum((ubar - u) * du)
        assert abs(inner) < 1e-10


def test_cfg1_convergence_fixture_iteration_count():
    """time-to-1e-6 configuration: the flat oracle needs the reference's iteration count (+-1)"""
    try:
        g = golden("cfg1_convergence.npz")
    except FileNotFoundError:
        pytest.skip("cfg1_convergence.npz not generated (python -m oracle.make_golden --long)")
    s, problem = _problem("cfg1")
    orc = FlatOracle(problem)
    status, xi, _ = orc.chock(s["x0"], max_iters=20000, tol=1e-6, alpha=float(g["alpha"]))
    assert status == 0 and abs(xi.shape[0] - int(g["iterations"])) <= 1


@pytest.mark.reference
@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_live_reference_agrees_with_both_oracles():
    api = ref_loader.RefApi()
    s = problems.spec("mini2", seed=7)
    prob = problems.build(s, api)
    solver = api.Solver(prob)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        solver.chock(s["x0"][:, :1], max_iters=14, tol=0.0)
    alpha = solver._Solver__parameter_1
    cache = solver._Solver__cache
    node, flat = NodeOracle(prob), FlatOracle(prob)
    for orc in (node, flat):
        orc.cache_initial_state(s["x0"][:, :1])
        orc.alpha = alpha
    for k in range(1, 16):
        node.iterate()
        flat.iterate()
        ref_p = np.vstack(cache._Cache__primal_cache[k]).reshape(-1)
        ref_d = cache._Cache__dual_cache[k]
        assert rel_err(np.vstack(node.primal_blocks()).reshape(-1), ref_p) < 1e-13
        assert rel_err(np.vstack(node.dual_blocks()).reshape(-1), np.vstack(ref_d).reshape(-1)) < 1e-13
        assert rel_err(flat.flat_primal(flat.p), ref_p) < 1e-12
        assert rel_err(flat.flat_dual(flat.d), flat.dual_from_blocks(ref_d)) < 1e-12
