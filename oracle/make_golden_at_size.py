"""Golden fixtures AT THE SIZES BASELINE.json NAMES, from the unmodified reference.  TEST INFRASTRUCTURE.

    python -m oracle.make_golden_at_size cfg2 cfg5 cfg4 cfg3        # build container only (needs /root/reference)

cfg3 takes ~55 s per iteration on the reference, cfg5 ~6.5 s: the reference is stepped manually with its
history pruned (oracle/ref_stepper.py) for 100 iterations and, because the iterates themselves are megabytes, each
fixture keeps for the iterations in KEEP

  * the values at SAMPLES seeded random positions of the raw flat primal and of the raw flat dual (placeholders
    included -- they must stay 0),
  * per reference list segment: inf-norm, l1, l2 and plain sum of the whole segment (a checksum over EVERY entry),

plus alpha, x0, the full residual histories (xi, delta: 100 x 3) and the reference's own wall time per iteration
(`sec_per_iteration`, the measured CPU figure that BASELINE.md / bench.py quote for this container).
cfg4 records instances 0, 1, 2047 and 4095 of the 4096 seeded initial states.

Fixtures: tests/golden/<name>_at_size.npz (cfg4: one file, arrays suffixed _i<instance>).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import problems, ref_loader  # noqa: E402
from oracle.cp_flat_oracle import FlatOracle  # noqa: E402
from oracle.ref_stepper import RefStepper  # noqa: E402

KEEP = (1, 2, 3, 5, 10, 20, 50, 100)
SAMPLES = 8192
CFG4_INSTANCES = (0, 1, 2047, 4095)


def seg_stats(vec, edges):
    out = np.empty((len(edges) - 1, 4))
    for k in range(len(edges) - 1):
        s = vec[edges[k]: edges[k + 1]]
        out[k] = (np.max(np.abs(s)), np.sum(np.abs(s)), np.sqrt(np.sum(s * s)), np.sum(s)) if s.size else 0.0
    return out


def run(name, api, x0, alpha, iters, log):
    s = problems.spec(name, batch=4096 if name == "cfg4" else 1)
    t0 = time.perf_counter()
    prob = problems.build(s, api)
    build_s = time.perf_counter() - t0
    st = RefStepper(api, prob, x0, alpha)
    pe, de = st.segment_lengths()
    rng = np.random.default_rng(2024)
    pidx = np.sort(rng.choice(int(pe[-1]), size=min(SAMPLES, int(pe[-1])), replace=False))
    didx = np.sort(rng.choice(int(de[-1]), size=min(SAMPLES, int(de[-1])), replace=False))
    out = dict(alpha=st.alpha, x0=np.asarray(x0), keep=np.array([k for k in KEEP if k <= iters]),
               n=prob.tree.num_nodes, m=prob.tree.num_nonleaf_nodes, pidx=pidx, didx=didx,
               p_edges=pe, d_edges=de, build_s=build_s, offline_s=st.setup_s)
    times = []
    for k in range(1, iters + 1):
        t0 = time.perf_counter()
        st.step()
        times.append(time.perf_counter() - t0)
        if k in KEEP:
            p, d = st.primal(), st.dual()
            out[f"p{k}"], out[f"d{k}"] = p[pidx], d[didx]
            out[f"ps{k}"], out[f"ds{k}"] = seg_stats(p, pe), seg_stats(d, de)
        log(f"{name} it {k} {times[-1]:.2f}s xi {st.xi[-1]}")
    out["xi"], out["delta"] = np.array(st.xi), np.array(st.delta)
    out["sec_per_iteration"] = np.array(times)
    return out


def main(argv):
    api = ref_loader.RefApi()
    names = argv or ["cfg2", "cfg5", "cfg4", "cfg3"]
    iters = int(os.environ.get("GOLDEN_ITERS", "100"))

    def log(msg):
        print(msg, flush=True)

    for name in names:
        s = problems.spec(name, batch=4096 if name == "cfg4" else 1)
        alpha = FlatOracle(problems.build(s, api)).step_size()
        if name == "cfg4":
            merged = {}
            for inst in CFG4_INSTANCES:
                o = run(name, api, s["x0"][:, inst: inst + 1], alpha, iters, log)
                for key, val in o.items():
                    merged[f"{key}_i{inst}"] = val
            merged["instances"] = np.array(CFG4_INSTANCES)
            np.savez_compressed(os.path.join(GOLD, f"{name}_at_size.npz"), **merged)
        else:
            o = run(name, api, s["x0"][:, :1], alpha, iters, log)
            np.savez_compressed(os.path.join(GOLD, f"{name}_at_size.npz"), **o)
        log(f"wrote {name}_at_size.npz")


if __name__ == "__main__":
    main(sys.argv[1:])
