"""The batch-innermost ("panel") path of the fused loop (csrc/batch.cu): many initial states on one small tree, the 32 lanes
of a warp being 32 instances (BASELINE.json configs[3], SURVEY 8e "instance-parallel").

Checked against the per-instance NumPy oracle (1e-9 per segment, north star), against the instance-major kernels (same
iteration, different kernels) and through rb_step with a changing x0.  Batch sizes that do not fill the last panel are used on
purpose.  The full-size case (cfg4 x 4096 against the unmodified reference) is tests/test_gpu_at_size.py.
"""
import numpy as np
import pytest

from helpers import seg_rel_err

pytestmark = pytest.mark.gpu


def _setup(name, batch):
    import raocp_b200 as r
    from oracle import problems
    s = problems.spec(name, batch=batch)
    problem = problems.build(s, r.core)
    return s, problem, r


@pytest.mark.parametrize("name,batch,iters", [("cfg1", 70, 40), ("mini2", 64, 30), ("mini3", 97, 30), ("cfg4", 64, 12)])
def test_panel_loop_matches_oracle(name, batch, iters):
    from oracle.cp_flat_oracle import FlatOracle
    s, problem, r = _setup(name, batch)
    solver = r.core.Solver(problem, batch=batch, verbose=False)
    flat, dev = solver.cache.flat_problem, solver.cache.device_solver
    alpha = FlatOracle(problem).step_size()
    assert solver.chock(s["x0"], max_iters=iters - 1, tol=0.0, alpha=alpha) == 1 and solver.iterations == iters
    p, d = dev.get_primal(0), dev.get_dual(0)
    xi, delta = solver.residual_history          # (iters, batch, 3)
    for b in sorted({0, 31, 32, batch // 2, batch - 1}):
        orc = FlatOracle(problem)
        orc.cache_initial_state(s["x0"][:, b:b + 1])
        orc.alpha = alpha
        for k in range(iters):
            oxi, odelta = orc.iterate()
            assert np.max(np.abs(xi[k, b] - np.array(oxi)) / np.array(oxi)) < 1e-6
            assert np.max(np.abs(delta[k, b] - np.array(odelta)) / np.array(odelta)) < 1e-6
        assert seg_rel_err(flat, p[b], orc.flat_primal(orc.p), dual=False) < 1e-9
        assert seg_rel_err(flat, d[b], orc.flat_dual(orc.d), dual=True) < 1e-9


def test_panel_equals_instance_major_kernels():
    """same problem through the instance-major kernels (one grid row per instance): every instance, 1e-11"""
    s, problem, r = _setup("mini3", 70)
    out = []
    for panels in (True, False):
        solver = r.core.Solver(problem, batch=70, verbose=False)
        solver.cache.device_solver.use_batch_panels(panels)
        alpha = solver.compute_step_size()
        solver.chock(s["x0"], max_iters=24, tol=0.0, alpha=alpha)
        dev = solver.cache.device_solver
        out.append((dev.get_primal(0), dev.get_dual(0), solver.residual_history[0]))
    flat = solver.cache.flat_problem
    for b in range(70):
        assert seg_rel_err(flat, out[0][0][b], out[1][0][b], dual=False) < 1e-11
        assert seg_rel_err(flat, out[0][1][b], out[1][1][b], dual=True) < 1e-11
    assert np.max(np.abs(out[0][2] - out[1][2]) / out[1][2]) < 1e-9


def test_panel_second_solve_and_stopping():
    """a second chock() continues from the first one's iterate (conversion out of and back into the panels), and the stopping
    test waits for EVERY instance (solver.py:156-161 applied per instance)"""
    from oracle.cp_flat_oracle import FlatOracle
    s, problem, r = _setup("cfg1", 64)
    solver = r.core.Solver(problem, batch=64, verbose=False)
    flat, dev = solver.cache.flat_problem, solver.cache.device_solver
    alpha = FlatOracle(problem).step_size()
    for x0 in (s["x0"], 0.5 - s["x0"]):
        assert solver.chock(x0, max_iters=9, tol=0.0, alpha=alpha) == 1
    orc = FlatOracle(problem)
    orc.alpha = alpha
    for x0 in (s["x0"][:, 5:6], 0.5 - s["x0"][:, 5:6]):
        orc.cache_initial_state(x0)
        for _ in range(10):
            orc.iterate()
    assert seg_rel_err(flat, dev.get_primal(0)[5], orc.flat_primal(orc.p), dual=False) < 1e-9
    assert seg_rel_err(flat, dev.get_dual(0)[5], orc.flat_dual(orc.d), dual=True) < 1e-9
    fresh = r.core.Solver(problem, batch=64, verbose=False)
    assert fresh.chock(s["x0"], max_iters=20000, tol=1e-3, alpha=alpha) == 0
    xi = fresh.residual_history[0]
    assert np.all(np.max(xi[-1], axis=1) <= 1e-3) and np.any(np.max(xi[-2], axis=1) > 1e-3)


def test_panel_step_with_new_initial_states():
    """rb_step on a batch: x0 [batch][nx] from pinned host memory every step, norms [batch][6] back"""
    import torch
    from oracle.cp_flat_oracle import FlatOracle
    s, problem, r = _setup("mini2", 64)
    solver = r.core.Solver(problem, batch=64, verbose=False)
    flat, dev = solver.cache.flat_problem, solver.cache.device_solver
    alpha = FlatOracle(problem).step_size()
    xa = s["x0"]
    xb = 0.25 - 0.5 * xa
    solver.cache.cache_initial_state(xa)
    dev.loop_begin(alpha, 1 << 30, -1.0, 0)
    x_host = torch.zeros(64, flat.nx, dtype=torch.float64).pin_memory()
    n_host = torch.zeros(64, 6, dtype=torch.float64).pin_memory()
    orc = FlatOracle(problem)
    orc.alpha = alpha
    b = 37
    for x0, steps in ((xa, 3), (xb, 4), (xa, 2)):
        x_host.copy_(torch.from_numpy(np.ascontiguousarray(x0.T)))
        orc.cache_initial_state(x0[:, b:b + 1])
        for _ in range(steps):
            dev.step(x_host.data_ptr(), n_host.data_ptr())
            oxi, odelta = orc.iterate()
            want = np.concatenate((np.array(oxi), np.array(odelta)))
            assert np.max(np.abs(n_host.numpy()[b] - want) / want) < 1e-6
    dev.loop_end()
    assert seg_rel_err(flat, dev.get_primal(0)[b], orc.flat_primal(orc.p), dual=False) < 1e-9
    assert seg_rel_err(flat, dev.get_dual(0)[b], orc.flat_dual(orc.d), dual=True) < 1e-9
