"""Seeded synthetic RAOCP instances for the five BASELINE.json configurations plus the
reference's own demo problem (reference main.py:11-80).  TEST / BENCH INFRASTRUCTURE.

`spec(name)` returns plain arrays; `build(spec, api)` turns them into a problem with either the reference's classes
(oracle.ref_loader.RefApi) or raocp_b200.core -- both expose the same builder API -- so that the very same numbers
reach the reference, the oracles and the CUDA path.
"""
import numpy as np

# name: (modes, horizon N, stopping time tau, nx, nu)
SHAPES = {
    "cfg1": (2, 4, 4, 2, 1),       # 31 nodes      (BASELINE.json configs[0])
    "cfg2": (3, 10, 5, 10, 5),     # 1 579 nodes   (configs[1])
    "cfg3": (4, 20, 6, 20, 10),    # 62 805 nodes  (configs[2])
    "cfg4": (2, 9, 9, 8, 4),       # 1 023 nodes x 4096 initial states (configs[3])
    "cfg5": (3, 18, 6, 64, 32),    # 9 841 nodes   (configs[4])
    # scaled-down siblings used by parity tests so the reference / node oracle finish in seconds
    "mini2": (3, 5, 3, 4, 2),      # 3 modes, ragged-free, 94-ish nodes
    "mini3": (4, 7, 3, 6, 3),
    "mini5": (3, 6, 3, 16, 8),
    "dense": (3, 5, 3, 5, 3),      # like mini2 but with NON-diagonal SPD cost weights (general matvec path)
    "wide": (2, 3, 3, 40, 36),     # nx, nu > 32: more than one row per lane
    "shard": (3, 9, 5, 8, 4),      # 1 822 nodes, 81 subtrees below stage 4: multi-GPU subtree sharding checks
    # trees with a long chain part (stopping time << horizon): the cut-stage / chain kernels of the DP sweeps
    "chain21": (2, 8, 3, 2, 1),
    "chain32": (3, 7, 2, 3, 2),
    "chain63": (4, 7, 3, 6, 3),
    "chain105": (3, 9, 5, 10, 5),
    "chain2010": (4, 9, 4, 20, 10),  # 1 621 nodes, 256 chains of 6 nodes: a small cfg3
    "chain6432": (3, 6, 3, 64, 32),  # 121 nodes, 27 chains of 4 nodes, cfg5's sizes: fragments in shared memory (chain_mma BIG)
}


def spec(name, seed=0, batch=1):
    """Seeded random instance of one of SHAPES (P,v ~ normalised U(0.1,1); A_w ~ N(0,1) scaled to
    spectral radius 0.9; B_w ~ N(0,1)/sqrt(nx); Q_w,R_w,Qf = diag U(0.5,2); AVaR(0.5); |x|<=5, |u|<=1; x0 ~ U(-1,1))."""
    modes, horizon, tau, nx, nu = SHAPES[name]
    rng = np.random.default_rng(seed)
    p = rng.uniform(0.1, 1.0, size=(modes, modes))
    p /= p.sum(axis=1, keepdims=True)
    v = rng.uniform(0.1, 1.0, size=modes)
    v /= v.sum()
    a_list, b_list, q_list, r_list = [], [], [], []
    for _ in range(modes):
        a = rng.standard_normal((nx, nx))
        a *= 0.9 / np.max(np.abs(np.linalg.eigvals(a)))
        a_list.append(a)
        b_list.append(rng.standard_normal((nx, nu)) / np.sqrt(nx))
        q_list.append(np.diag(rng.uniform(0.5, 2.0, size=nx)))
        r_list.append(np.diag(rng.uniform(0.5, 2.0, size=nu)))
    qf = np.diag(rng.uniform(0.5, 2.0, size=nx))
    if name == "dense":   # random SPD weights with condition number <= 4

        def spd(k):
            g, _ = np.linalg.qr(rng.standard_normal((k, k)))
            return g @ np.diag(rng.uniform(0.5, 2.0, size=k)) @ g.T

        q_list = [spd(nx) for _ in range(modes)]
        r_list = [spd(nu) for _ in range(modes)]
        qf = spd(nx)
    x0 = rng.uniform(-1.0, 1.0, size=(nx, batch))
    return dict(name=name, seed=seed, p=p, v=v, horizon=horizon, tau=tau, nx=nx, nu=nu,
                a=a_list, b=b_list, q=q_list, r=r_list, qf=qf, avar=0.5,
                x_lim=5.0, u_lim=1.0, x0=x0, rectangles=True)


def wide_spec(n_gpus, seed=0, horizon=20, tau=None):
    """cfg3 made ~n_gpus times WIDER (the weak-scaling tree of the subtree-sharded bench): same nx = 20, nu = 10, horizon 20,
    AVaR(0.5), rectangles.  n_gpus = 2: eight modes in two blocks of four (block-diagonal transition matrix, so every node
    still has four children but the root has eight) -> 8 192 chains; 4: four modes with stopping time 7 -> 16 384 chains;
    8: both -> 32 768 chains.  At most 8 children per node (the limit of the lanes-per-node passes).
    horizon / tau: smaller siblings of the same shape for oracle checks (tests/test_gpu_sweeps.py)."""
    if n_gpus == 1:
        return spec("cfg3", seed=seed)
    blocks = 2 if n_gpus in (2, 8) else 1
    if tau is None:
        tau = 7 if n_gpus in (4, 8) else 6
    if n_gpus not in (2, 4, 8):
        raise ValueError("wide_spec: n_gpus must be 1, 2, 4 or 8")
    modes, nx, nu = 4 * blocks, 20, 10
    rng = np.random.default_rng(seed + 1000 * n_gpus)
    p = np.zeros((modes, modes))
    for b in range(blocks):
        blk = rng.uniform(0.1, 1.0, size=(4, 4))
        p[4 * b: 4 * b + 4, 4 * b: 4 * b + 4] = blk / blk.sum(axis=1, keepdims=True)
    v = rng.uniform(0.1, 1.0, size=modes)
    v /= v.sum()
    a_list, b_list, q_list, r_list = [], [], [], []
    for _ in range(modes):
        a = rng.standard_normal((nx, nx))
        a *= 0.9 / np.max(np.abs(np.linalg.eigvals(a)))
        a_list.append(a)
        b_list.append(rng.standard_normal((nx, nu)) / np.sqrt(nx))
        q_list.append(np.diag(rng.uniform(0.5, 2.0, size=nx)))
        r_list.append(np.diag(rng.uniform(0.5, 2.0, size=nu)))
    qf = np.diag(rng.uniform(0.5, 2.0, size=nx))
    x0 = rng.uniform(-1.0, 1.0, size=(nx, 1))
    return dict(name=f"cfg3x{n_gpus}", seed=seed, p=p, v=v, horizon=horizon, tau=tau, nx=nx, nu=nu, a=a_list, b=b_list, q=q_list,
                r=r_list, qf=qf, avar=0.5, x_lim=5.0, u_lim=1.0, x0=x0, rectangles=True)


def demo_spec():
    """The reference's demo problem, numbers from reference main.py:11-80 (43 nodes, nx=3, nu=2, AVaR 0.95)."""
    p = np.array([[0.1, 0.8, 0.1], [0.4, 0.6, 0.0], [0.0, 0.3, 0.7]])
    v = np.array([0.1, 0.6, 0.3])
    f = 0.1
    aw = f * np.array([[1, 2, 1], [1, 1, 2], [2, 1, 1]], dtype=float)
    bw = f * np.array([[1, 0], [1, 0], [0, 2]], dtype=float)
    q = 0.2 * f * np.eye(3)
    r = 0.2 * f * np.eye(2)
    return dict(name="demo", seed=None, p=p, v=v, horizon=4, tau=3, nx=3, nu=2,
                a=[0.5 * aw, aw, -0.5 * aw], b=[-0.5 * bw, bw, 0.5 * bw],
                q=[q, q, q], r=[r, r, r], qf=f * 0.1 * np.eye(3), avar=0.95,
                x_lim=7.0, u_lim=0.1, x0=np.array([[5.0], [-6.0], [-1.0]]), rectangles=True)


def build(s, api):
    """Problem object (RAOCP) for spec `s` using the classes of `api` (reference or raocp_b200.core)."""
    tree = api.MarkovChainScenarioTreeFactory(s["p"], s["v"], s["horizon"], s["tau"]).create()
    nl, lf = api.Nonleaf(), api.Leaf()
    nx, nu = s["nx"], s["nu"]
    dyn = [api.Dynamics(a, b) for a, b in zip(s["a"], s["b"])]
    nl_costs = [api.Quadratic(nl, q, r) for q, r in zip(s["q"], s["r"])]
    problem = api.RAOCP(tree).with_markovian_dynamics(dyn).with_markovian_nonleaf_costs(nl_costs) \
        .with_all_leaf_costs(api.Quadratic(lf, s["qf"])).with_all_risks(api.AVaR(s["avar"]))
    if s.get("rectangles", True):
        hi_nl = np.vstack((s["x_lim"] * np.ones((nx, 1)), s["u_lim"] * np.ones((nu, 1))))
        hi_l = s["x_lim"] * np.ones((nx, 1))
        problem = problem.with_all_nonleaf_constraints(api.Rectangle(nl, -hi_nl, hi_nl)) \
            .with_all_leaf_constraints(api.Rectangle(lf, -hi_l, hi_l))
    return problem
