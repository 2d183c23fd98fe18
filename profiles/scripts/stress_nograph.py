"""Stress check of the plain-launch (no CUDA graph) pipelined loop against the graph loop: same residual history?"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "raocp-toolbox_b200"), ROOT):
    sys.path.insert(0, p)
import raocp_b200 as r
from oracle import problems

name = sys.argv[1] if len(sys.argv) > 1 else "mini5"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
s = problems.spec(name)
problem = problems.build(s, r.core)
x0 = s["x0"][:, :1]
ref = r.core.Solver(problem, verbose=False)
alpha = ref.compute_step_size()
ref.chock(x0, max_iters=99, tol=0.0, alpha=alpha)
xi_ref = ref.residual_history[0].copy()
p_ref = ref.cache.device_solver.get_primal(0)[0].copy()
bad = 0
for rep in range(reps):
    sv = r.core.Solver(problem, verbose=False)
    sv.cache.device_solver.use_graphs(False)
    sv.chock(x0, max_iters=99, tol=0.0, alpha=alpha)
    xi = sv.residual_history[0]
    dp = np.max(np.abs(sv.cache.device_solver.get_primal(0)[0] - p_ref))
    rows = np.where(np.max(np.abs(xi - xi_ref) / xi_ref, axis=1) > 1e-9)[0]
    if len(rows) or dp > 1e-9:
        bad += 1
        print(f"rep {rep}: primal diff {dp:.3e}, bad rows {rows[:10]}", flush=True)
        for k in rows[:3]:
            print("   row", k, xi[k], xi_ref[k], flush=True)
print("bad reps", bad, "of", reps)
