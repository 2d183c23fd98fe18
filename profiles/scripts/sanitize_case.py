"""One short solve with a chosen kernel-family configuration, for compute-sanitizer (profiles/scripts/sanitize.sh).

    python profiles/scripts/sanitize_case.py <problem> <iters> <pipeline 0|1|3|4> <tree_mode 0|1|2> <mma 0|1> [batch]

Plain launches (no CUDA graph) so that a hazard is attributed to a kernel; the step-by-step API (one kernel per reference
method), the residual kernels, the offline factorisation and the step-size kernels run as well.  Exit code 0 = solve finished
and matches the NumPy oracle to 1e-9."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "raocp-toolbox_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402


def main():
    name, iters, pipe, tree, mma = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    batch = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    import raocp_b200 as r
    from oracle import problems
    from oracle.cp_flat_oracle import FlatOracle
    s = problems.spec(name, batch=batch)
    problem = problems.build(s, r.core)
    x0 = s["x0"] if batch > 1 else s["x0"][:, :1]
    orc = FlatOracle(problem)
    alpha = orc.step_size()
    sol = r.core.Solver(problem, batch=batch, verbose=False)
    dev = sol.cache.device_solver
    assert abs(sol.compute_step_size() - alpha) < 1e-12 * alpha
    dev.use_graphs(False)
    dev.use_pipeline(pipe)
    dev.use_tree_kernels(tree)
    dev.use_mma_sweeps(bool(mma))
    sol.chock(x0, max_iters=iters - 1, tol=0.0, alpha=alpha)
    orc.cache_initial_state(s["x0"][:, :1])
    orc.alpha = alpha
    for _ in range(iters):
        orc.iterate()
    p, d = dev.get_primal(0)[0], dev.get_dual(0)[0]
    err = max(np.max(np.abs(p - orc.flat_primal(orc.p))) / max(1.0, np.max(np.abs(p))),
              np.max(np.abs(d - orc.flat_dual(orc.d))) / max(1.0, np.max(np.abs(d))))
    # the step-by-step API and the stand-alone residual kernels, two iterations
    sol.set_step_size(alpha)
    for _ in range(2):
        sol.primal_k_plus_half(); sol.primal_k_plus_one(); sol.dual_k_plus_half(); sol.dual_k_plus_one()
        dev.residuals(alpha)
        sol.cache.update_cache()
    print(f"sanitize_case {name} iters={iters} pipeline={pipe} tree={tree} mma={mma} batch={batch}: err vs oracle {err:.2e}, "
          f"kernels launched {dev.launch_count()}")
    sys.exit(0 if err < 1e-9 else 1)


if __name__ == "__main__":
    main()
