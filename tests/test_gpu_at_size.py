"""Parity AT THE SIZES BASELINE.json NAMES (cfg2 ... cfg5), default kernels (pipelined loop, CUDA graphs, tensor-core
chain walkers, fused tree kernel), through the C-ABI.

Two independent checkers:
  * the UNMODIFIED REFERENCE: tests/golden/<cfg>_at_size.npz (oracle/make_golden_at_size.py stepped the reference's own
    Solver methods for 100 iterations in the build container): values at 8 192 seeded positions of the raw flat primal and
    dual at iterations 1, 2, 3, 5, 10, 20, 50, 100, checksums (inf-norm, l1, l2, sum) of EVERY reference list segment at the
    same iterations, and the full 100 x 3 residual histories;
  * the vectorised oracle (oracle/cp_flat_oracle.py, itself pinned to the reference) run here on the host next to the GPU:
    every entry of the iterate for the first iterations.

Bars (north star): 1e-9 relative per segment (relative to the segment's inf-norm) over the first 100 iterations;
residual histories 1e-6 relative.  cfg4 runs the whole batch of 4096 initial states and checks instances 0, 1, 2047, 4095.
"""
import os

import numpy as np
import pytest

from helpers import GOLD, golden, seg_rel_err
from oracle.at_size_check import check_against_fixture as _check_against_fixture

pytestmark = pytest.mark.gpu

TOL = 1e-9


def _need(name):
    path = os.path.join(GOLD, f"{name}_at_size.npz")
    if not os.path.isfile(path):
        pytest.skip(f"{path} not generated")
    return golden(f"{name}_at_size.npz")


def _problem(name, batch=1):
    import raocp_b200 as r
    from oracle import problems
    s = problems.spec(name, batch=batch)
    return s, problems.build(s, r.core), r


# chock() continues from the current iterate like the reference's (cache.py:79-82), so the KEEP iterations are reached by
# successive calls: this also runs the first-iteration path of the pipelined loop (stand-alone primal pass) at every restart
def _advance(solver, x0, alpha, upto, done):
    solver.chock(x0, max_iters=upto - done - 1, tol=0.0, alpha=alpha)
    assert solver.iterations == upto - done
    return solver.residual_history


@pytest.mark.parametrize("name,kw", [("cfg2", {}), ("cfg2", {"dedup": False}), ("cfg5", {}), ("cfg3", {})],
                         ids=["cfg2", "cfg2_nodedup", "cfg5", "cfg3"])
def test_100_iterations_match_reference(name, kw):
    g = _need(name)
    s, problem, r = _problem(name)
    solver = r.core.Solver(problem, verbose=False, **kw)
    flat, dev = solver.cache.flat_problem, solver.cache.device_solver
    assert flat.n == int(g["n"]) and flat.m == int(g["m"])
    alpha, x0 = float(g["alpha"]), g["x0"]
    assert abs(solver.compute_step_size() - alpha) <= 1e-12 * alpha
    done, worst, xi_all, delta_all = 0, 0.0, [], []
    for k in [int(v) for v in g["keep"]]:
        xi, delta = _advance(solver, x0, alpha, k, done)
        xi_all.append(xi)
        delta_all.append(delta)
        done = k
        worst = max(worst, _check_against_fixture(flat, g, "", k, dev.get_primal(0)[0], dev.get_dual(0)[0]))
    xi, delta = np.vstack(xi_all), np.vstack(delta_all)
    assert xi.shape == g["xi"].shape
    assert np.max(np.abs(xi - g["xi"]) / g["xi"]) < 1e-6
    assert np.max(np.abs(delta - g["delta"]) / g["delta"]) < 1e-6
    assert worst < TOL, worst


def test_cfg3_uninterrupted_100_iterations_match_reference():
    """one chock() of 100 iterations (no restarts: the steady-state graph loop only), iterate 100 against the reference"""
    g = _need("cfg3")
    s, problem, r = _problem("cfg3")
    solver = r.core.Solver(problem, verbose=False)
    flat, dev = solver.cache.flat_problem, solver.cache.device_solver
    assert solver.chock(g["x0"], max_iters=99, tol=0.0, alpha=float(g["alpha"])) == 1 and solver.iterations == 100
    worst = _check_against_fixture(flat, g, "", 100, dev.get_primal(0)[0], dev.get_dual(0)[0])
    xi, delta = solver.residual_history
    assert np.max(np.abs(xi - g["xi"]) / g["xi"]) < 1e-6 and np.max(np.abs(delta - g["delta"]) / g["delta"]) < 1e-6
    assert worst < TOL, worst


def test_cfg4_batch_4096_matches_reference():
    g = _need("cfg4")
    batch = 4096
    s, problem, r = _problem("cfg4", batch=batch)
    solver = r.core.Solver(problem, batch=batch, verbose=False)
    flat, dev = solver.cache.flat_problem, solver.cache.device_solver
    inst = [int(i) for i in g["instances"]]
    alpha = float(g[f"alpha_i{inst[0]}"])
    for i in inst:
        assert np.array_equal(g[f"x0_i{i}"].reshape(-1), s["x0"][:, i])
    done, worst = 0, 0.0
    xi_all = []
    for k in [int(v) for v in g[f"keep_i{inst[0]}"]]:
        xi, _ = _advance(solver, s["x0"], alpha, k, done)
        xi_all.append(xi)
        done = k
        p, d = dev.get_primal(0), dev.get_dual(0)
        for i in inst:
            worst = max(worst, _check_against_fixture(flat, g, f"_i{i}", k, p[i], d[i]))
    xi = np.concatenate(xi_all, axis=0)                     # (100, batch, 3)
    for i in inst:
        want = g[f"xi_i{i}"]
        assert np.max(np.abs(xi[:, i, :] - want) / want) < 1e-6
    assert worst < TOL, worst


@pytest.mark.parametrize("name,iters", [("cfg3", 12), ("cfg5", 20)])
def test_every_entry_matches_flat_oracle(name, iters):
    """all Np + Nd entries (not a sample) against the vectorised oracle run on this box's host"""
    from oracle.cp_flat_oracle import FlatOracle
    s, problem, r = _problem(name)
    x0 = s["x0"][:, :1]
    oracle = FlatOracle(problem)
    alpha = oracle.step_size()
    solver = r.core.Solver(problem, verbose=False)
    flat, dev = solver.cache.flat_problem, solver.cache.device_solver
    oracle.cache_initial_state(x0)
    oracle.alpha = alpha
    done, worst = 0, 0.0
    for upto in (1, 2, iters):
        _advance(solver, x0, alpha, upto, done)
        while done < upto:
            oracle.iterate()
            done += 1
        worst = max(worst,
                    seg_rel_err(flat, dev.get_primal(0)[0], oracle.flat_primal(oracle.p), dual=False),
                    seg_rel_err(flat, dev.get_dual(0)[0], oracle.flat_dual(oracle.d), dual=True))
    assert worst < TOL, worst
