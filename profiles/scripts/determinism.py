"""Run-to-run determinism as a race detector (compute-sanitizer is closed on this GPU pool: profiles/sanitizer/README.md).

Every kernel family switch is run REPEATS times from the same state; all data races the hand-off code could have (cp.async
rings, named barriers of the four-warp walkers, the counter / flag hand-off of the fused tree kernel, the pbar aliasing of the
dual passes, the peer-memory exchange) change the iterate in at least the last bits when they strike, while every intended
cross-thread combination in these kernels is order-independent (integer max of bit patterns).  So: bit-identical iterates and
residual histories across repeats, and agreement with the NumPy oracle to 1e-9.

    python profiles/scripts/determinism.py [repeats]        ->  one line per configuration, exit code 1 on any mismatch
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "raocp-toolbox_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

CASES = [  # name, iterations, pipeline, tree mode, mma mode, graphs, batch
    ("chain2010", 25, 1, 2, 1, 1, 1), ("chain2010", 25, 1, 2, 2, 1, 1), ("chain2010", 25, 3, 1, 1, 0, 1),
    ("chain2010", 25, 4, 2, 1, 1, 1), ("chain2010", 25, 0, 0, 0, 1, 1), ("chain6432", 15, 1, 2, 1, 1, 1),
    ("cfg2", 25, 1, 2, 1, 1, 1), ("cfg1", 40, 1, 2, 1, 1, 1), ("mini3", 20, 1, 2, 1, 1, 3), ("mini3", 20, 1, 2, 1, 1, 70),
    ("cfg3", 20, 1, 2, 1, 1, 1), ("cfg5", 10, 1, 2, 1, 1, 1),
]


def main():
    repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    import raocp_b200 as r
    from oracle import problems
    from oracle.cp_flat_oracle import FlatOracle
    bad = 0
    for name, iters, pipe, tree, mma, graphs, batch in CASES:
        s = problems.spec(name, batch=batch)
        problem = problems.build(s, r.core)
        x0 = s["x0"] if batch > 1 else s["x0"][:, :1]
        alpha = None
        ref = None
        diffs = 0
        for rep in range(repeats):
            sol = r.core.Solver(problem, batch=batch, verbose=False)
            dev = sol.cache.device_solver
            dev.use_graphs(bool(graphs)); dev.use_pipeline(pipe); dev.use_tree_kernels(tree); dev.use_mma_sweeps(mma)
            alpha = alpha or sol.compute_step_size()
            sol.chock(x0, max_iters=iters - 1, tol=0.0, alpha=alpha)
            out = (dev.get_primal(0).copy(), dev.get_dual(0).copy(), np.array(sol.residual_history[0]).copy())
            if ref is None:
                ref = out
            elif not all(np.array_equal(a, b) for a, b in zip(ref, out)):
                diffs += 1
        err = float("nan")
        if name not in ("cfg3", "cfg5"):
            orc = FlatOracle(problem)
            orc.cache_initial_state(s["x0"][:, :1])
            orc.alpha = alpha
            for _ in range(iters):
                orc.iterate()
            p, d = ref[0][0], ref[1][0]
            err = max(np.max(np.abs(p - orc.flat_primal(orc.p))) / max(1.0, np.max(np.abs(p))),
                      np.max(np.abs(d - orc.flat_dual(orc.d))) / max(1.0, np.max(np.abs(d))))
        ok = diffs == 0 and not (err >= 1e-9)
        bad += 0 if ok else 1
        print(f"{name:10s} iters={iters:3d} pipeline={pipe} tree={tree} mma={mma} graphs={graphs} batch={batch:3d}: {repeats} runs, "
              f"{diffs} differ from the first; vs oracle {err:.2e}  {'ok' if ok else 'FAIL'}", flush=True)
    print("DETERMINISM", "PASS" if bad == 0 else "FAIL")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
