import sys, time
sys.path.insert(0, "/root/repo/raocp-toolbox_b200"); sys.path.insert(0, "/root/repo")
import numpy as np
import raocp_b200 as r
from oracle import problems
s = problems.spec("cfg3"); problem = problems.build(s, r.core)
solver = r.core.Solver(problem, verbose=False)
dev = solver.cache.device_solver
x0 = s["x0"][:, :1]
alpha = solver.compute_step_size()
dev.set_initial_state(x0.reshape(-1))
for lane in (True, False):
    dev.use_lane_kernels(lane)
    dev.loop_begin(alpha, 1 << 30, -1.0, 0)
    dev.loop_enqueue(10)
    t = np.array([dev.profile_iteration() for _ in range(20)])
    dev.loop_end()
    print("lane" if lane else "tile", "us per launch:", np.round(np.median(t, axis=0) * 1e3, 1))
