"""Device tests of the stand-alone cone / rectangle classes and of the list-form Operator wrappers.

Ports of the reference's own tests onto the CUDA path (everything below runs rb_cone_project / rb_box_project /
rb_apply_L / rb_apply_Lt through the C-ABI):
  * reference tests/test_cones.py:30-276 -- variational inequality <v - P(v), s - P(v)> <= 0 for samples s of the cone
    (and of the dual cone for project_onto_dual), all five classes incl. Cartesian, seeded instead of np.random;
  * reference tests/test_rectangle.py:68-103 -- clip into [lo, hi], NaN raises ValueError;
  * reference tests/test_operators.py:118-373 -- adjointness through the block-list API, linop_* == list form.
Each projection is also compared bit for bit with the oracle's restatement (cones.py:113-132, rectangle.py:50-59).
"""
import numpy as np
import pytest

from helpers import spec_for

pytestmark = pytest.mark.gpu

DIM, SAMPLES, REPEATS, MULT = 20, 100, 100, 10


def _cones():
    import raocp_b200.core.constraints.cones as cones
    return cones


def _samples(rng, kind, dim=DIM):
    """members of the cone `kind` (same generators as the reference's tests)"""
    if kind == "real":
        return [rng.integers(-100, 100, dim).astype(float) for _ in range(SAMPLES)]
    if kind == "zero":
        return [np.zeros(dim) for _ in range(SAMPLES)]
    if kind == "nonneg":
        return [rng.integers(0, 100, dim).astype(float) for _ in range(SAMPLES)]
    out = []
    for _ in range(SAMPLES):
        s = rng.standard_normal(dim - 1)
        out.append(np.hstack((s, np.linalg.norm(s))))
    return out


def _variational(vector, projection, samples):
    v, p = vector.reshape(-1), projection.reshape(-1)
    scale = max(1.0, float(np.abs(v) @ np.abs(v)))
    for s in samples:
        assert np.inner(v - p, s - p) <= 1e-12 * scale * max(1.0, np.max(np.abs(s)))


CASES = [("Real", "real", "zero"), ("Zero", "zero", "real"), ("NonnegativeOrthant", "nonneg", "nonneg"),
         ("SecondOrderCone", "soc", "soc")]


def _expected(name, dual, v):
    from oracle.cp_node_oracle import NodeOracle
    if name == "SecondOrderCone":
        return NodeOracle.soc_project(v.reshape(-1)).reshape(v.shape)
    kind = {"Real": "zero" if dual else "real", "Zero": "real" if dual else "zero",
            "NonnegativeOrthant": "nonneg"}[name]
    return {"real": v.copy(), "zero": np.zeros_like(v), "nonneg": np.maximum(v, 0.0)}[kind]


@pytest.mark.parametrize("name,kind,dual_kind", CASES)
def test_cone_project_variational(name, kind, dual_kind):
    cones = _cones()
    rng = np.random.default_rng(11)
    for rep in range(REPEATS if name == "SecondOrderCone" else 3):
        cone = getattr(cones, name)()
        assert type(cone).__name__ == name
        v = (MULT * rng.standard_normal(DIM)).reshape(DIM, 1)
        p = cone.project(v)
        assert p.shape == v.shape
        _variational(v, p, _samples(rng, kind))
        assert np.array_equal(p, _expected(name, False, v)) or np.max(np.abs(p - _expected(name, False, v))) < 1e-13
        q = cone.project_onto_dual(v)
        _variational(v, q, _samples(rng, dual_kind))
        assert np.max(np.abs(q - _expected(name, True, v))) < 1e-13


def test_soc_all_three_branches():
    """cones.py:120-132: inside -> v; polar -> 0; else the boundary formula"""
    cones = _cones()
    soc = cones.SecondOrderCone()
    z = np.array([3.0, -4.0, 0.0, 12.0])                       # ||z|| = 13
    inside = np.append(z, 20.0).reshape(-1, 1)
    polar = np.append(z, -20.0).reshape(-1, 1)
    between = np.append(z, 1.0).reshape(-1, 1)
    assert np.array_equal(soc.project(inside), inside)
    assert np.array_equal(cones.SecondOrderCone().project(polar), np.zeros_like(polar))
    out = cones.SecondOrderCone().project(between)
    assert abs(out[-1, 0] - 7.0) < 1e-14 and np.max(np.abs(out[:-1, 0] - 7.0 * z / 13.0)) < 1e-14
    with pytest.raises(Exception):
        cones.SecondOrderCone().project(np.ones((2, 1)))


def test_cone_dimension_errors():
    cones = _cones()
    cones._check_dimension("Real", 5, np.ones(5))
    with pytest.raises(ValueError):
        cones._check_dimension("Real", 5, np.ones(6))
    with pytest.raises(ValueError):
        cones.NonnegativeOrthant(4).project(np.ones((5, 1)))


@pytest.mark.parametrize("dual", [False, True])
def test_cartesian_project(dual):
    cones = _cones()
    rng = np.random.default_rng(12)
    cart = cones.Cartesian([cones.Real(), cones.Zero(), cones.NonnegativeOrthant(), cones.SecondOrderCone()])
    assert type(cart).__name__ == "Cartesian" and cart.num_cones == 4
    assert cart.types == "Real x Zero x NonnegativeOrthant x SecondOrderCone"
    vec = [(MULT * rng.standard_normal(DIM)).reshape(DIM, 1) for _ in range(4)]
    kinds = ["zero", "real", "nonneg", "soc"] if dual else ["real", "zero", "nonneg", "soc"]
    out = cart.project_onto_dual(vec) if dual else cart.project(vec)
    for i in range(4):
        _variational(vec[i], out[i], _samples(rng, kinds[i]))
    assert cart.dimension == 4 * DIM and cart.dimensions == [DIM] * 4


def test_cartesian_single_stacked_vector():
    """cones.py:164-206 / risks.py:32-33: one stacked vector is split by the declared dimensions (the AVaR cone)"""
    cones = _cones()
    rng = np.random.default_rng(13)
    cart = cones.Cartesian([cones.NonnegativeOrthant(6), cones.Zero(1)])
    v = rng.standard_normal((7, 1))
    out = cart.project_onto_dual([v])
    assert out.shape == (7, 1)
    assert np.array_equal(out[:6], np.maximum(v[:6], 0.0)) and np.array_equal(out[6:], v[6:])
    out = cart.project([v])
    assert np.array_equal(out[:6], np.maximum(v[:6], 0.0)) and np.array_equal(out[6:], np.zeros((1, 1)))


def test_rectangle_project_on_device():
    import raocp_b200.core.constraints.rectangle as rect
    import raocp_b200.core.nodes as nodes
    from oracle.cp_node_oracle import NodeOracle
    rng = np.random.default_rng(14)
    nx, nu = 3, 2
    lo, hi = 4 * np.ones((nx + nu, 1)), 5 * np.ones((nx + nu, 1))
    con = rect.Rectangle(nodes.Nonleaf(), lo, hi)
    assert con.is_active is True
    con.state_size, con.control_size = nx, nu
    for _ in range(20):
        v = 10 * rng.standard_normal((nx + nu, 1))
        out = con.project(v)
        assert np.all(lo <= out) and np.all(out <= hi)
        assert np.array_equal(out.reshape(-1), NodeOracle.box_project(v.reshape(-1), lo.reshape(-1), hi.reshape(-1)))
    inside = 4.5 * np.ones((nx + nu, 1))
    assert np.array_equal(con.project(inside), inside)
    bad = inside.copy()
    bad[2, 0] = np.nan
    with pytest.raises(ValueError):
        con.project(bad)
    with pytest.raises(Exception):
        con.project(np.ones((nx + nu + 1, 1)))
    leaf = rect.Rectangle(nodes.Leaf(), lo[:nx], hi[:nx])
    leaf.state_size = nx
    out = leaf.project(10 * rng.standard_normal((nx, 1)))
    assert np.all(lo[:nx] <= out) and np.all(out <= hi[:nx])


# ---- Operator list wrappers (reference tests/test_operators.py) ---------------------------------------------------------
@pytest.fixture(scope="module", params=["demo", "mini2", "wide"])
def op_case(request):
    import raocp_b200 as r
    from oracle import problems
    s = spec_for(request.param)
    problem = problems.build(s, r.core)
    cache = r.core.Cache(problem)
    return dict(cache=cache, op=r.core.Operator(cache), problem=problem, r=r)


def _random_like(blocks, rng):
    return [rng.standard_normal(b.shape) for b in blocks]


def test_list_operators_adjoint(op_case):
    """<L p, d> == <p, L* d> through ell / ell_transpose on block lists (tests/test_operators.py:118-335)"""
    cache, op = op_case["cache"], op_case["op"]
    rng = np.random.default_rng(15)
    _, prim = cache.get_primal()
    _, dual = cache.get_dual()
    seg_p = cache.get_primal_segments()
    rand_p, rand_d = _random_like(prim, rng), _random_like(dual, rng)
    rand_p[seg_p[4]] = np.zeros((1, 1))               # tau_0 is not a variable (operators.py:55-94 never writes it)
    ell_p = [np.zeros_like(b) for b in dual]
    ell_t_d = [np.zeros_like(b) for b in prim]
    op.ell(rand_p, ell_p)
    op.ell_transpose(rand_d, ell_t_d)
    lhs = sum(float(a.T @ b) for a, b in zip(ell_p, rand_d))
    rhs = sum(float(a.T @ b) for a, b in zip(rand_p, ell_t_d))
    # placeholders of the dual: L writes nothing there (zeros stay), so they do not enter <L p, d>
    assert abs(lhs - rhs) <= 1e-10 * max(1.0, abs(lhs))


def test_list_operators_match_oracle(op_case):
    cache, op = op_case["cache"], op_case["op"]
    flat = cache.flat_problem
    rng = np.random.default_rng(16)
    _, prim = cache.get_primal()
    _, dual = cache.get_dual()
    rand_p, rand_d = _random_like(prim, rng), _random_like(dual, rng)
    ell_p = [b.copy() for b in dual]
    ell_t_d = [b.copy() for b in prim]
    op.ell(rand_p, ell_p)
    op.ell_transpose(rand_d, ell_t_d)
    from oracle.cp_flat_oracle import FlatOracle
    fo = FlatOracle(op_case["problem"])
    want_d = fo.flat_dual(fo.ell(fo.unflat_primal(flat.primal_from_blocks(rand_p))))
    want_p = fo.flat_primal(fo.ell_transpose(fo.unflat_dual(flat.dual_from_blocks(rand_d))))
    got_d = flat.dual_from_blocks(ell_p)
    got_p = flat.primal_from_blocks(ell_t_d)
    tau0 = flat.n * flat.nx + flat.m * flat.nu + flat.ysz
    want_p[tau0] = 0.0                                 # left as the caller supplied it (zero here), App. C.4
    assert np.max(np.abs(got_d - want_d)) <= 1e-12 * max(1.0, np.max(np.abs(want_d)))
    assert np.max(np.abs(got_p - want_p)) <= 1e-12 * max(1.0, np.max(np.abs(want_p)))


def test_linop_wrappers_equal_list_form(op_case):
    """tests/test_operators.py:337-373: linop_ell(vstack(p)) == vstack(ell(p)), same for the adjoint"""
    cache, op = op_case["cache"], op_case["op"]
    rng = np.random.default_rng(17)
    _, prim = cache.get_primal()
    _, dual = cache.get_dual()
    rand_p = _random_like(prim, rng)
    ell_p = [np.zeros_like(b) for b in dual]
    op.ell(rand_p, ell_p)
    assert np.array_equal(np.vstack(ell_p), op.linop_ell(np.vstack(rand_p)))
    rand_d = _random_like(dual, rng)
    ell_t_d = [np.zeros_like(b) for b in prim]
    op.ell_transpose(rand_d, ell_t_d)
    seg_p = cache.get_primal_segments()
    wrapped = op.linop_ell_transpose(np.vstack(rand_d))
    unwrapped = np.vstack(ell_t_d)
    keep = np.ones(unwrapped.size, dtype=bool)
    keep[int(np.sum([b.size for b in prim[: seg_p[4]]]))] = False      # tau_0: list form leaves the caller's value
    assert np.array_equal(unwrapped[keep], wrapped[keep])
