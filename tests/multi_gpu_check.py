"""Subtree-sharded solve on W GPUs vs the oracle and vs the single-GPU solve (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py [--quick]

Every rank solves the whole problem on its own GPU (single-GPU path) and takes part in the sharded solve.  Checked per case:
  * the assembled sharded iterate against the NumPy oracle (1e-9 per segment, the north-star bar) -- small cases,
  * against the default single-GPU loop (1e-10: same arithmetic, the top of the tree replicated),
  * same residual history (1e-9) and the same stopping iteration at a tolerance,
for BOTH exchanges: device-initiated over NVLink peer memory (pipelined loop in one CUDA graph; default: the exchange inside the
fused tree kernel, RAOCP_SHARD_XCHG=kernel / split: its one-launch and three-launch forms) and the NCCL all-gather between plain
launches (RAOCP_SHARD_P2P=0).  tests/test_gpu_sharding.py runs this file at every world size the box
offers."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "raocp-toolbox_b200"), ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import raocp_b200 as r
    from oracle import problems
    from oracle.cp_flat_oracle import FlatOracle
    from helpers import seg_rel_err

    quick = "--quick" in sys.argv
    # (the lanes-per-node passes the sharded loop is built on need even nx, nu: cfg2's nu = 5 is out)
    cases = [("shard", 60, True), ("chain2010", 40, True)]
    if not quick:
        cases += [("cfg5", 12, False), ("cfg3", 25, False)]
    ok = True
    for name, iters, with_oracle in cases:
        s = problems.spec(name)
        problem = problems.build(s, r.core)
        x0 = s["x0"][:, :1]
        single = r.core.Solver(problem, device=local, verbose=False)
        alpha = single.compute_step_size()
        single.chock(x0, max_iters=iters - 1, tol=0.0, alpha=alpha)
        sd = single.cache.device_solver
        p1, d1 = sd.get_primal(0)[0], sd.get_dual(0)[0]
        xi1, _ = single.residual_history
        flat = single.cache.flat_problem
        cut = sd.shard_info()
        if cut[2] < world:
            if rank == 0:
                print(f"{name}: {cut[2]} cut-stage subtrees < {world} ranks, skipped", flush=True)
            continue
        orc = None
        if with_oracle:
            orc = FlatOracle(problem)
            orc.cache_initial_state(x0)
            orc.alpha = alpha
            for _ in range(iters):
                orc.iterate()
        variants = [("1", "", "peer-memory"), ("0", "", "nccl")]
        if name in (cases[0][0], "cfg3"):   # the ablation forms of the peer-memory exchange: one launch between the level kernels
            variants.insert(1, ("1", "kernel", "peer-memory one-launch"))   # (k_shard_xchg) / push, pull, check as three launches
            variants.insert(2, ("1", "split", "peer-memory split"))
        for p2p, xchg, label in variants:
            os.environ["RAOCP_SHARD_P2P"] = p2p
            os.environ["RAOCP_SHARD_XCHG"] = xchg
            sharded = r.core.Solver(problem, device=local, verbose=False, shard=(rank, world))
            dev = sharded.cache.device_solver
            dev.shard_init()
            assert sharded.cache.flat_problem.shard_cut == dev.shard_info()[0]
            sharded.chock(x0, max_iters=iters - 1, tol=0.0, alpha=alpha)
            p2, d2 = dev.gather_sharded(0)
            xi2, _ = sharded.residual_history
            e1 = max(seg_rel_err(flat, p2, p1, dual=False), seg_rel_err(flat, d2, d1, dual=True))
            er = float(np.max(np.abs(xi2 - xi1) / xi1))
            line = (f"[rank {rank}] {name} x{world} {label}: cut stage {cut[0]}, after {iters} iterations vs "
                    f"single GPU {e1:.2e}, residual history {er:.2e}")
            good = e1 < 1e-10 and er < 1e-9 and sharded.iterations == single.iterations
            if orc is not None:
                eo = max(seg_rel_err(flat, p2, orc.flat_primal(orc.p), dual=False),
                         seg_rel_err(flat, d2, orc.flat_dual(orc.d), dual=True))
                line += f", vs oracle {eo:.2e}"
                good = good and eo < 1e-9
            # stopping at a tolerance: same iteration count on both paths
            s1 = r.core.Solver(problem, device=local, verbose=False)
            s2 = r.core.Solver(problem, device=local, verbose=False, shard=(rank, world))
            s2.cache.device_solver.shard_init()
            tol = float(np.sort(xi1.max(axis=1))[1]) * (1 + 1e-9)   # first reached somewhere inside the recorded history
            st1 = s1.chock(x0, max_iters=400, tol=tol, alpha=alpha)
            st2 = s2.chock(x0, max_iters=400, tol=tol, alpha=alpha)
            line += f"; stop test {s1.iterations} / {s2.iterations} iterations (status {st1} / {st2})"
            good = good and st1 == st2 and s1.iterations == s2.iterations
            print(line + ("" if good else "   <-- FAIL"), flush=True)
            ok &= good
            del sharded, s2
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if int(t.item()) == 1 else "FAIL", flush=True)
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
