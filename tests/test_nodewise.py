"""Node-wise (non-Markovian) dynamics, costs and risk levels: `RAOCP.with_nodewise_*` (beyond the reference's
builder).  The device path needs no new kernel -- operators are tables indexed per node -- but with all matrices different
there is one factorisation class per node (K, R~ stream from HBM) and no chain tile shares its matrices (warp-per-chain
walker).  Checked: the two oracles agree (CPU), the unmodified reference agrees with them when the same per-node lists are
injected into its RAOCP (CPU, build container only), the CUDA path agrees with the oracle (GPU)."""
import numpy as np
import pytest

from helpers import seg_rel_err


def _nodewise(api, seed=3, horizon=5, tau=3, nx=4, nu=2, modes=3):
    rng = np.random.default_rng(seed)
    p = rng.uniform(0.1, 1.0, size=(modes, modes))
    p /= p.sum(axis=1, keepdims=True)
    v = rng.uniform(0.1, 1.0, size=modes)
    v /= v.sum()
    tree = api.MarkovChainScenarioTreeFactory(p, v, horizon, tau).create()
    n, m = tree.num_nodes, tree.num_nonleaf_nodes
    nl, lf = api.Nonleaf(), api.Leaf()
    dyn, costs = [], []
    for _ in range(n - 1):
        a = rng.standard_normal((nx, nx))
        a *= 0.9 / np.max(np.abs(np.linalg.eigvals(a)))
        dyn.append(api.Dynamics(a, rng.standard_normal((nx, nu)) / np.sqrt(nx)))
        costs.append(api.Quadratic(nl, np.diag(rng.uniform(0.5, 2.0, nx)), np.diag(rng.uniform(0.5, 2.0, nu))))
    leaf_costs = [api.Quadratic(lf, np.diag(rng.uniform(0.5, 2.0, nx))) for _ in range(n - m)]
    levels = rng.uniform(0.2, 0.9, size=m)
    x0 = rng.uniform(-1.0, 1.0, size=(nx, 1))
    hi_nl = np.vstack((5.0 * np.ones((nx, 1)), np.ones((nu, 1))))
    return dict(tree=tree, dyn=dyn, costs=costs, leaf_costs=leaf_costs, levels=levels, x0=x0, hi_nl=hi_nl,
                hi_l=5.0 * np.ones((nx, 1)), nl=nl, lf=lf)


def _ours(d, api):
    return api.RAOCP(d["tree"]).with_nodewise_dynamics(d["dyn"]).with_nodewise_nonleaf_costs(d["costs"]) \
        .with_nodewise_leaf_costs(d["leaf_costs"]).with_nodewise_risks([api.AVaR(a) for a in d["levels"]]) \
        .with_all_nonleaf_constraints(api.Rectangle(d["nl"], -d["hi_nl"], d["hi_nl"])) \
        .with_all_leaf_constraints(api.Rectangle(d["lf"], -d["hi_l"], d["hi_l"]))


def test_builder_checks():
    import raocp_b200 as r
    d = _nodewise(r.core)
    with pytest.raises(ValueError):
        r.core.RAOCP(d["tree"]).with_nodewise_dynamics(d["dyn"][:-1])
    with pytest.raises(Exception):
        r.core.RAOCP(d["tree"]).with_nodewise_nonleaf_costs(d["leaf_costs"] + d["costs"][: len(d["costs"]) - len(d["leaf_costs"])])
    problem = _ours(d, r.core)
    assert problem.risk_at_node(0).alpha == d["levels"][0] and problem.state_dynamics_at_node(7) is d["dyn"][6].state_dynamics


def test_oracles_agree_on_nodewise_data():
    import raocp_b200 as r
    from oracle.cp_flat_oracle import FlatOracle
    from oracle.cp_node_oracle import NodeOracle
    d = _nodewise(r.core)
    problem = _ours(d, r.core)
    fo, no = FlatOracle(problem), NodeOracle(problem)
    alpha = fo.step_size()
    fo.cache_initial_state(d["x0"])
    no.cache_initial_state(d["x0"])
    fo.alpha = no.alpha = alpha
    for _ in range(15):
        xi_f, _ = fo.iterate()
        xi_n, _ = no.iterate()
    assert np.max(np.abs(np.array(xi_f) - np.array(xi_n)) / np.array(xi_n)) < 1e-9
    pf = fo.flat_primal(fo.p)
    pn = fo.primal_from_blocks(no.primal_blocks())
    assert np.max(np.abs(pf - pn)) <= 1e-10 * max(1.0, np.max(np.abs(pn)))


@pytest.mark.reference
def test_reference_agrees_with_oracle_on_nodewise_data():
    """the unmodified reference with the same per-node lists injected into its RAOCP (it has no builder for them)"""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference not present")
    from oracle.cp_flat_oracle import FlatOracle
    from oracle.ref_stepper import RefStepper
    api = ref_loader.RefApi()
    d = _nodewise(api)
    base = api.RAOCP(d["tree"]).with_markovian_dynamics(d["dyn"][:3]).with_markovian_nonleaf_costs(d["costs"][:3]) \
        .with_all_leaf_costs(d["leaf_costs"][0]).with_all_risks(api.AVaR(0.5)) \
        .with_all_nonleaf_constraints(api.Rectangle(d["nl"], -d["hi_nl"], d["hi_nl"])) \
        .with_all_leaf_constraints(api.Rectangle(d["lf"], -d["hi_l"], d["hi_l"]))
    m = d["tree"].num_nonleaf_nodes
    base._RAOCP__list_of_dynamics[1:] = d["dyn"]
    base._RAOCP__list_of_nonleaf_costs[1:] = d["costs"]
    base._RAOCP__list_of_leaf_costs[m:] = d["leaf_costs"]
    for i in range(m):
        risk = api.AVaR(float(d["levels"][i]))
        risk.probs = d["tree"].conditional_probabilities_of_children(i)
        base._RAOCP__list_of_risks[i] = risk
    fo = FlatOracle(base)
    alpha = fo.step_size()
    st = RefStepper(api, base, d["x0"], alpha)
    fo.cache_initial_state(d["x0"])
    fo.alpha = alpha
    for _ in range(10):
        xi_r, _ = st.step()
        xi_f, _ = fo.iterate()
    assert np.max(np.abs(np.array(xi_f) - xi_r) / xi_r) < 1e-9
    assert np.max(np.abs(fo.flat_primal(fo.p) - st.primal())) <= 1e-10 * max(1.0, np.max(np.abs(st.primal())))


@pytest.mark.gpu
@pytest.mark.parametrize("dedup", [True, False])
@pytest.mark.parametrize("shape", [dict(horizon=5, tau=3, nx=4, nu=2), dict(horizon=8, tau=3, nx=6, nu=3, modes=2)])
def test_device_matches_oracle_on_nodewise_data(dedup, shape):
    import raocp_b200 as r
    from oracle.cp_flat_oracle import FlatOracle
    d = _nodewise(r.core, **shape)
    problem = _ours(d, r.core)
    fo = FlatOracle(problem)
    alpha = fo.step_size()
    solver = r.core.Solver(problem, dedup=dedup, verbose=False)
    flat, dev = solver.cache.flat_problem, solver.cache.device_solver
    assert flat.num_cls == flat.m            # every node its own factorisation class: all matrices differ
    assert abs(solver.compute_step_size() - alpha) <= 1e-11 * alpha
    assert solver.chock(d["x0"], max_iters=39, tol=0.0, alpha=alpha) == 1
    fo.cache_initial_state(d["x0"])
    fo.alpha = alpha
    for _ in range(40):
        xi, _ = fo.iterate()
    assert seg_rel_err(flat, dev.get_primal(0)[0], fo.flat_primal(fo.p), dual=False) < 1e-9
    assert seg_rel_err(flat, dev.get_dual(0)[0], fo.flat_dual(fo.d), dual=True) < 1e-9
    assert np.max(np.abs(solver.residual_history[0][-1] - np.array(xi)) / np.array(xi)) < 1e-6
