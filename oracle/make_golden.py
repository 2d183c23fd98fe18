"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.  TEST INFRASTRUCTURE.

    python -m oracle.make_golden            # everything except the long cfg1 convergence run
    python -m oracle.make_golden --long     # also cfg1 to 1e-6 on the reference (about 5 minutes)

The reference (Python) cannot travel to the GPU box, so its outputs are committed as small fixtures:

  demo_residuals.npz   the 937 x 3 residual history parsed from the reference's own artefact 4-3-residuals.tex
                       (xi_0 :27-966, xi_1 :968-1907, xi_2 :1909-2848) -- the only golden numbers the reference ships
  <name>_iterates.npz  for the seeded problems of oracle/problems.py: alpha as computed by the reference (ARPACK),
                       raw flat iterates (np.vstack of the reference's block lists, placeholders included) after
                       selected iterations, the full residual histories, and offline data (P, K, A+BK of the Cache)
  <name>_ops.npz       random primal/dual vectors pushed through the reference's Operator.ell / ell_transpose and
                       Cache.project_on_dynamics / project_on_kernel / proximal_of_g_conjugate
  cfg1_convergence.npz iteration count and residual history of the reference solving cfg1 to 1e-6 (--long)
"""
import argparse
import os
import re
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import problems, ref_loader  # noqa: E402

KEEP = (1, 2, 3, 5, 10, 20, 50, 100)


def parse_tex():
    path = os.path.join(ref_loader.REFERENCE_ROOT, "4-3-residuals.tex")
    text = open(path).read()
    tables = re.findall(r"table \{%\n(.*?)\n\};", text, flags=re.S)
    assert len(tables) == 3
    cols = [np.array([float(line.split()[1]) for line in t.strip().splitlines()]) for t in tables]
    out = np.stack(cols, axis=1)
    assert out.shape == (937, 3)
    np.savez_compressed(os.path.join(GOLD, "demo_residuals.npz"), xi=out)
    print("demo_residuals", out.shape)


def raw(blocks):
    return np.vstack(blocks).reshape(-1)


def iterates(name, api, iters=100):
    s = problems.demo_spec() if name == "demo" else problems.spec(name)
    prob = problems.build(s, api)
    solver = api.Solver(prob)
    x0 = s["x0"][:, :1]
    t0 = time.time()
    solver.chock(x0, max_iters=iters - 1, tol=0.0)  # runs exactly `iters` iterations
    cache = solver._Solver__cache
    out = dict(alpha=solver._Solver__parameter_1, x0=x0,
               xi=solver._Solver__error_cache, delta=solver._Solver__delta_error_cache,
               n=prob.tree.num_nodes, m=prob.tree.num_nonleaf_nodes, keep=np.array(KEEP))
    for k in KEEP:
        out[f"p{k}"] = raw(cache._Cache__primal_cache[k])
        out[f"d{k}"] = raw(cache._Cache__dual_cache[k])
    out["P"] = np.stack(cache._Cache__P)
    out["K"] = np.stack(cache._Cache__K)
    out["Abar"] = np.stack(cache._Cache__sum_of_dynamics[1:])
    np.savez_compressed(os.path.join(GOLD, f"{name}_iterates.npz"), **out)
    print(name, "iterates", prob.tree.num_nodes, "nodes", f"{time.time() - t0:.1f}s")


def ops(name, api, seed=123):
    s = problems.demo_spec() if name == "demo" else problems.spec(name)
    prob = problems.build(s, api)
    rng = np.random.default_rng(seed)
    cache = api.Cache(prob)
    op = api.Operator(cache)
    _, tp = cache.get_primal()
    _, td = cache.get_dual()
    seg_p, seg_d = cache.get_primal_segments(), cache.get_dual_segments()
    rand_p = [rng.standard_normal(b.shape) for b in tp]
    rand_p[seg_p[4]] = np.zeros((1, 1))  # tau_0 is never used
    rand_d = [rng.standard_normal(b.shape) if _real_block(i, prob) else np.zeros((1, 1))
              for i, b in enumerate(td)]
    out = dict(rand_p=raw(rand_p), rand_d=raw(rand_d))
    lp = [b.copy() for b in td]
    op.ell(rand_p, lp)
    out["ell"] = raw(lp)
    lt = [b.copy() for b in tp]
    op.ell_transpose(rand_d, lt)
    out["ell_t"] = raw(lt)
    # prox_f pieces on the random primal
    x0 = rng.standard_normal((prob.state_dynamics_at_node(1).shape[1], 1))
    cache.cache_initial_state(x0)
    cache.set_primal(rand_p)
    cache.project_on_dynamics()
    out["x0"] = x0
    out["dyn"] = raw(cache.get_primal()[0])
    cache.set_primal(rand_p)
    cache.project_on_kernel()
    out["ker"] = raw(cache.get_primal()[0])
    cache.set_primal(rand_p)
    cache.proximal_of_f(0.37)
    out["proxf"] = raw(cache.get_primal()[0])
    # prox_g* on the random dual (scaled up so that every SOC / box branch is hit)
    big_d = [3.0 * b for b in rand_d]
    cache.set_dual(big_d)
    cache.proximal_of_g_conjugate(0.37)
    out["big_d"] = raw(big_d)
    out["proxg"] = raw(cache.get_dual()[0])
    np.savez_compressed(os.path.join(GOLD, f"{name}_ops.npz"), **out)
    print(name, "ops")


def _real_block(i, prob):
    """True if block i of the reference's dual list is a real (non-placeholder) block (cache.py:140-170)."""
    n, m = prob.tree.num_nodes, prob.tree.num_nonleaf_nodes
    seg, node = divmod(i, n)
    part = (1, 2, 3, 4, 5, 6, 7, 11, 12, 13, 14)[seg]
    if part in (1, 2):
        return node < m
    if part in (3, 4, 5, 6):
        return node > 0
    if part == 7:
        return node < m and prob.nonleaf_constraint_at_node(node).is_active
    if part in (11, 12, 13):
        return node >= m
    return node >= m and prob.leaf_constraint_at_node(node).is_active


def cfg1_convergence(api):
    s = problems.spec("cfg1")
    prob = problems.build(s, api)
    solver = api.Solver(prob)
    t0 = time.time()
    status = solver.chock(s["x0"][:, :1], max_iters=20000, tol=1e-6)
    err = solver._Solver__error_cache
    np.savez_compressed(os.path.join(GOLD, "cfg1_convergence.npz"), status=status, iterations=err.shape[0],
                        alpha=solver._Solver__parameter_1, xi=err, seconds=time.time() - t0)
    print("cfg1 convergence", status, err.shape, f"{time.time() - t0:.0f}s")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--long", action="store_true")
    ap.add_argument("--only", default=None, help="comma separated problem names")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    api = ref_loader.RefApi()
    parse_tex()
    names = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else \
        ("demo", "cfg1", "mini2", "mini3", "mini5", "dense", "wide")
    for nm in names:
        iterates(nm, api)
        ops(nm, api)
    if args.long:
        cfg1_convergence(api)
