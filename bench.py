#!/usr/bin/env python
"""bench.py -- Chambolle-Pock iterations/sec of the raocp_b200 CUDA path (and of the CPU restatement of the reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3] [--batch B]

One "step" = one full CP iteration (primal half step, prox_f, dual half step, prox_g*, all six residual norms and the
stopping test, reference solver.py:124-161) over the synthetic, seeded problem of oracle/problems.py.
Default workload: BASELINE.json configs[2] ("cfg3": 62 805-node tree, nx=20, nu=10), the tree the north-star target is
quoted on; it fits one GPU.  Rank 0 prints ONE JSON line.

Timing (device, CUDA events on the launching stream, max over ranks):
  value    cold iterations/s: every timed iteration runs after an L2 flush (a 256 MiB buffer is overwritten), each
           iteration bracketed by its own event pair (whole iteration = one CUDA graph launch)
  roofline the dominant kernel (dual pass): algorithmic bytes per launch / its launch duration measured with CUDA events
           around plain launches of the same iteration, L2 flushed; the whole-iteration figure rides in roofline.iteration
  warm     the same K iterations back to back with no flush (what a real solve sees: the 50 MB of iterates of cfg3
           stay L2-resident on a B200), reported beside it
  e2e      through the host API with HOST buffers: every step copies x0 from pinned host memory to the device, runs
           one iteration and reads the six residual norms back (synchronised)
With N > 1 (torchrun) every rank solves its own instance(s) of the workload ("independent problem instances",
BASELINE.json north_star) -- no data-path collective, scaling "weak".
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "raocp-toolbox_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

# stdout carries exactly ONE JSON line: libraries that write to file descriptor 1 (NCCL prints its version banner there) are
# sent to stderr for the whole run, and the line itself goes to a duplicate of the original stdout
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


import numpy as np  # noqa: E402

METRIC = "CP iterations/sec"
# bounded CPU samples: same modes / nx / nu as the workload on a shorter tree, so the node-by-node CPU port (55 s per
# iteration on the full cfg3 tree) finishes in seconds; scaled to the full tree by node count (cost is linear in nodes)
CPU_SAMPLE = {"cfg1": None, "cfg2": (3, 7, 4), "cfg3": (4, 8, 4), "cfg4": None, "cfg5": (3, 6, 4)}  # (modes, N, tau)


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the roofline kernel, from the committed `ncu --set full`
# capture of the same command (profiles/): (workload, batch, dedup) -> bytes
TRAFFIC = {("cfg3", 1, True): 57.84e6 + 4.56e6}   # profiles/r1c_iter_raw.csv, k_dual_chain<20,10,4,3>


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        return json.load(open(path))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(flat, batch, dedup):
    """every iterate entry read once and written once, plus the node-specific DP operators (K twice,
    R~ factor once) when they are streamed per node; with (mode, stage) de-duplicated operators that term is dropped"""
    state = 8 * batch * 2 * (flat.np_ + flat.nd_)
    ops = 0 if dedup else 8 * flat.m * (2 * flat.nu * flat.nx + flat.nu * flat.nu)
    return state + ops


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons = [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1]))
                out["sm_max_mhz"] = float(f[2])
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def cpu_baseline(workload, budget_iters=4):
    """Times the node-by-node CPU port of the reference (oracle/cp_node_oracle.py, kind "port") on a bounded sample."""
    from oracle import problems
    from oracle.cp_node_oracle import NodeOracle
    import raocp_b200 as r
    full = problems.spec(workload)
    sample_shape = CPU_SAMPLE.get(workload)
    s = dict(full)
    if sample_shape is not None:
        modes, horizon, tau = sample_shape
        s["horizon"], s["tau"] = horizon, tau
    problem = problems.build(s, r.core)
    n_sample = problem.tree.num_nodes
    n_full = problems.build(full, r.core).tree.num_nodes if sample_shape is not None else n_sample
    t0 = time.perf_counter()
    orc = NodeOracle(problem)
    t_setup = time.perf_counter() - t0
    orc.cache_initial_state(full["x0"][:, :1])
    from oracle.cp_flat_oracle import FlatOracle
    orc.alpha = FlatOracle(problem).step_size()
    orc.iterate()  # warm-up
    t0 = time.perf_counter()
    for _ in range(budget_iters):
        orc.iterate()
    dt = (time.perf_counter() - t0) / budget_iters
    its_sample = 1.0 / dt
    value = its_sample * n_sample / n_full
    return {"value": value, "unit": "it/s", "cores": 1, "kind": "port",
            "sample": f"{budget_iters} iterations of oracle/cp_node_oracle.py (NumPy, node-by-node like the reference) on "
                      f"a {n_sample}-node tree with the workload's modes/nx/nu: {its_sample:.3f} it/s, scaled by "
                      f"{n_sample}/{n_full} nodes to the full tree; setup (offline + null spaces) {t_setup:.1f} s; "
                      f"host cores available: {os.cpu_count()}"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    t0 = time.perf_counter()
    base = cpu_baseline(args.workload, budget_iters=max(1, min(args.steps, 6)))
    line = {"metric": METRIC, "value": base["value"], "unit": "it/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / base["value"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": args.workload, "batch": 1, "note": "CPU port of the reference, bounded sample"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    emit(line)


def run_ours(args, rank, world, local_rank):
    import torch
    import raocp_b200 as r
    from oracle import problems

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (raocp_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    K, W, batch = args.steps, max(3, args.warmup), args.batch
    spec = problems.spec(args.workload, seed=0, batch=batch * world)
    t0 = time.perf_counter()
    problem = problems.build(spec, r.core)
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    cuts = tuple(int(v) for v in args.sweep_cuts.split(",")) if args.sweep_cuts else None
    solver = r.core.Solver(problem, batch=batch, dedup=not args.no_dedup, device=local_rank, verbose=False, sweep_cuts=cuts)
    solver.cache.device_solver.use_tree_kernels(args.tree_mode)
    solver.cache.device_solver.use_mma_sweeps(not args.no_mma)
    solver.cache.device_solver.use_pipeline(0 if args.no_pipeline else (3 if args.fwd_split else (4 if args.no_risk_split else 1)))
    dev = solver.cache.device_solver
    dev.synchronize()
    t_setup = time.perf_counter() - t0
    flat = solver.cache.flat_problem
    alpha = solver.compute_step_size()
    stream = torch.cuda.Stream()
    dev.set_stream(stream.cuda_stream)
    x0_host = torch.from_numpy(np.ascontiguousarray(spec["x0"][:, rank * batch:(rank + 1) * batch].T)).pin_memory()
    norms_host = torch.zeros(batch, 6, dtype=torch.float64).pin_memory()
    dev.set_initial_state(x0_host.numpy().T if batch > 1 else x0_host.numpy().reshape(-1))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        # ---- warm-up ---------------------------------------------------------------------------------------------
        dev.loop_begin(alpha, 1 << 30, -1.0, 0)
        dev.loop_enqueue(W)
        prof = [dev.profile_iteration_full() for _ in range(5)][2:]
        phases = np.array([q[0] for q in prof]).mean(axis=0)
        parts = np.array([q[1] for q in prof]).mean(axis=0)
        barrier()
        launches0 = dev.launch_count()
        sampler = ClockSampler(local_rank)
        sampler.start()
        # ---- cold: L2 flushed before every timed iteration --------------------------------------------------------
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        barrier()
        for k in range(K):
            flush.zero_()
            starts[k].record(stream)
            dev.loop_enqueue(1)
            ends[k].record(stream)
        barrier()
        cold_ms = np.array([s.elapsed_time(e) for s, e in zip(starts, ends)])
        # ---- warm: back to back ------------------------------------------------------------------------------------
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        dev.loop_enqueue(K)
        ev1.record(stream)
        barrier()
        warm_ms = ev0.elapsed_time(ev1)
        launches = dev.launch_count() - launches0
        # ---- per-launch durations (roofline of the dominant kernel): the same iteration as plain launches with a CUDA event
        #      after each one (rb_profile_iteration, on this stream), L2 flushed before every iteration like the timed region
        cold_phases, cold_parts = [], []
        for _ in range(min(K, 50)):
            flush.zero_()
            ph, pt = dev.profile_iteration_full()
            cold_phases.append(ph)
            cold_parts.append(pt)
        cold_phases = np.array(cold_phases).mean(axis=0)
        cold_parts = np.array(cold_parts).mean(axis=0)
        clocks = sampler.stop()
        # ---- end to end through the host API: pinned x0 -> device, one iteration, norms -> host, per step ----------------
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            dev.step(x0_host.data_ptr(), norms_host.data_ptr())
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        iters_done, _, last_norms = dev.loop_poll()
        dev.loop_end()
        # whole-solve call for context: Solver.chock(x0_host, K iterations) incl. x0 upload + residual history download
        t0 = time.perf_counter()
        solver.chock(spec["x0"][:, rank * batch:(rank + 1) * batch] if batch > 1 else spec["x0"][:, :1],
                     max_iters=K - 1, tol=0.0, alpha=alpha)
        solve_s = time.perf_counter() - t0
        # ---- time to the 1e-6 residual (the second half of BASELINE.json's metric): wall time of Solver.chock from the host
        #      x0 to the stop flag on the host, bounded by --ttt-iters iterations (rank 0, one instance per GPU only)
        ttt = None
        if rank == 0 and batch == 1 and args.ttt_iters > 0:
            # a FRESH solver: like the reference's, Solver.chock continues from the current iterate (cache.py:79-82 only
            # replaces x_0), and the legs above have already advanced `solver` by thousands of iterations
            cold = r.core.Solver(problem, batch=1, dedup=not args.no_dedup, device=local_rank, verbose=False)
            cold.cache.device_solver.set_stream(stream.cuda_stream)
            cold.cache.device_solver.synchronize()
            t0 = time.perf_counter()
            st_ttt = cold.chock(spec["x0"][:, :1], max_iters=args.ttt_iters, tol=1e-6, alpha=alpha)
            ttt_s = time.perf_counter() - t0
            ttt = {"seconds": ttt_s, "iterations": int(cold.iterations), "converged": st_ttt == 0, "tol": 1e-6,
                   "max_iters": args.ttt_iters, "final_residual": float(np.max(cold.residual_history[0][-1])),
                   "note": "fresh Solver, zero iterates: Solver.chock(x0, max_iters, tol=1e-6) -- x0 upload, device-side "
                           "stopping test after every iteration, host poll every 64 iterations, residual history download; "
                           "`converged` false means the cap was reached first and `seconds` is the time to `final_residual`"}
            del cold

    # ---- N > 1: the same single tree sharded by subtree over all ranks (one all-gather per iteration) -------------------
    shard_ms = None
    if dist is not None and batch == 1:
        try:
            sh = r.core.Solver(problem, device=local_rank, verbose=False, shard=(rank, world))
            sdev = sh.cache.device_solver
            sdev.shard_init()
            sdev.set_stream(stream.cuda_stream)
            sdev.set_initial_state(spec["x0"][:, :1].reshape(-1))
            with torch.cuda.stream(stream):
                sdev.loop_begin(alpha, 1 << 30, -1.0, 0)
                sdev.loop_enqueue(W)
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                sdev.loop_enqueue(K)
                e1.record(stream)
                barrier()
                shard_ms = e0.elapsed_time(e1)
                sdev.loop_end()
        except Exception as exc:   # e.g. the tree has fewer cut-stage subtrees than ranks
            shard_ms = None
            shard_err = str(exc)

    cold_total, e2e_t = float(cold_ms.sum()), e2e_s
    if dist is not None:
        t = torch.tensor([cold_total, warm_ms, e2e_t, solve_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cold_total, warm_ms, e2e_t, solve_s = t.tolist()
        if shard_ms is not None:
            t = torch.tensor([shard_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            shard_ms = float(t.item())
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    units = batch * world  # problem instances advanced by one iteration per step
    value = units * K / (cold_total * 1e-3)
    peak, peak_src = peaks()
    b_alg = algorithmic_bytes(flat, batch, not args.no_dedup)
    ms_iter = cold_total / K
    achieved = b_alg / (ms_iter * 1e-3) / 1e9
    # ---- roofline of the dominant HBM kernel -------------------------------------------------------------------------------
    # Pipelined loop: the chain dual pass k_dual_chain (nonleaf nodes with one child: L, dual half step, prox of g*, the
    # six residual norms, and pbar of the next iteration).  Algorithmic bytes per launch = SURVEY 8(d)'s figure -- every
    # iterate entry read once and written once -- restricted to the nodes of the launch: 16 B x (primal + dual doubles
    # of a chain node and of the edge to its child).  What the kernel has to move is more (it reads p, p+, d and writes d+
    # and pbar: 3 primal + 2 dual rows), reported as `bytes_moved_model`.
    n_chain = dev.chain_dual_nodes()
    pipelined = len(cold_parts) == 3 and n_chain > 0
    if pipelined:
        p_node = flat.nx + flat.nu + 3 + 2            # x_i, u_i, y_i (2c+1 = 3), tau_j, s_i
        d_node = 3 + 1 + 2 * (flat.nx + flat.nu) + 2  # d1_i, d2_i, d3_j, d4_j, d5_j, d6_j, d7_i
        if not args.no_risk_split:                    # y_i, s_i, d1_i, d2_i are k_dual_risk_chain's (under the sweeps)
            p_node -= 4
            d_node -= 4
        b_kernel = 16 * batch * n_chain * (p_node + d_node)
        b_moved = 8 * batch * n_chain * (3 * p_node + 2 * d_node)
        t_kernel = float(cold_parts[1]) * 1e-3
        kname = (f"chain dual pass k_dual_chain<{flat.nx},{flat.nu}> over {n_chain} of {flat.n} nodes (L, dual half step, "
                 "prox of g*, six residual norms, pbar of the next iteration"
                 + ("" if args.no_risk_split else "; the risk block d1, d2 of these nodes runs under the sweeps") + ")")
        traffic = TRAFFIC.get((args.workload, batch, not args.no_dedup))
    else:
        b_kernel = b_moved = 8 * batch * (2 * flat.np_ + 2 * flat.nd_)
        t_kernel = float(cold_phases[-1]) * 1e-3
        kname = "dual pass (L, dual half step, prox of g*, six residual norms), all nodes"
        traffic = None
    serial_ms = float(cold_phases.sum())
    roofline = {
        "bound": "hbm", "kernel": kname,
        "achieved": b_kernel / t_kernel / 1e9, "peak": peak, "unit": "GB/s", "frac": b_kernel / t_kernel / 1e9 / peak,
        "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes": b_kernel, "bytes_moved_model": b_moved,
        "launch_ms": t_kernel * 1e3, "share_of_step": float(t_kernel * 1e3 / serial_ms),
        "timing": "CUDA events around the plain launch on the bench stream, L2 flushed before every iteration, mean of "
                  f"{min(K, 50)} iterations (includes ~2 us of launch gap); share_of_step = launch_ms / sum of all launches "
                  "of the iteration run one after the other (the ncu launch list is serialised the same way)",
        "iteration": {"algorithmic_bytes": b_alg, "achieved": achieved, "frac": achieved / peak,
                      "note": "all launches of one CP iteration inside the CUDA graph, the timed region of `value`"},
        "launch_ms_all": {"primal_or_kernel_projection": float(cold_phases[0]),
                          "sweeps_in_launch_order": [float(v) for v in cold_phases[1:-1]],
                          "dual_and_check": float(cold_phases[-1]),
                          "dual_kernels_branching_chain_leaves": [float(v) for v in cold_parts]},
        "launch_ms_all_warm": {"primal_or_kernel_projection": float(phases[0]),
                               "sweeps_in_launch_order": [float(v) for v in phases[1:-1]],
                               "dual_and_check": float(phases[-1]),
                               "dual_kernels_branching_chain_leaves": [float(v) for v in parts]},
    }
    line = {
        "metric": METRIC, "value": value, "unit": "it/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_iter, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {flat.n}-node scenario tree ({flat.m} nonleaf), nx={flat.nx}, "
                               f"nu={flat.nu}, AVaR(0.5), rectangles, seed 0", "instances_per_gpu": batch,
                   "parallelism": "single GPU" if world == 1 else f"{world} x independent instances (no collective)",
                   "l2": "flushed (256 MiB overwrite) before every timed iteration", "dedup_operators": not args.no_dedup,
                   "residuals": "all six norms + stopping test every iteration", "alpha": alpha,
                   "loop": "pipelined (dual pass writes pbar of the next iteration)" if not args.no_pipeline else "primal pass + dual pass"},
        "warm": {"value": units * K / (warm_ms * 1e-3), "unit": "it/s", "ms_per_step": warm_ms / K,
                 "note": "same K iterations back to back, no L2 flush (iterates L2-resident when they fit)"},
        "e2e": {"value": units * K / e2e_t, "unit": "it/s", "h2d_bytes_per_step": int(x0_host.numel() * 8),
                "d2h_bytes_per_step": int(norms_host.numel() * 8),
                "note": "rb_step per iteration: x0 from pinned host memory, one iteration, residual norms to the host, "
                        "synchronised every step",
                "solve_call_it_s": units * K / solve_s},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "clocks": clocks,
        "setup": {"problem_build_s": t_build, "flatten_upload_offline_s": t_setup, "factorisation_classes": flat.num_cls},
        "residuals_last": [float(v) for v in last_norms[0]],
    }
    if world > 1:
        line["sharded_single_tree"] = (
            {"value": K / (shard_ms * 1e-3), "unit": "it/s", "ms_per_step": shard_ms / K, "scaling": "strong",
             "note": f"ONE {flat.n}-node tree sharded by subtree over {world} GPUs, one NCCL all-gather of the cut-stage "
                     "q_j, d2_j and residual maxima per iteration, K iterations back to back (warm), max over ranks; "
                     "compare with warm.value / n_gpus of the N=1 run"}
            if shard_ms is not None else {"unavailable": locals().get("shard_err", "batch > 1")})
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(args.workload)
    if ttt is not None:
        if "cpu_baseline" in line and line["cpu_baseline"].get("value"):
            ttt["cpu_seconds_extrapolated"] = ttt["iterations"] / line["cpu_baseline"]["value"]
        line["time_to_1e-6"] = ttt
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--batch", type=int, default=1, help="problem instances per GPU")
    ap.add_argument("--no-dedup", action="store_true", help="stream per-node K / R~ (one factorisation class per node)")
    ap.add_argument("--ttt-iters", type=int, default=50000,
                    help="iteration cap of the time-to-1e-6-residual leg (0: skip it)")
    ap.add_argument("--sweep-cuts", default="", help="ablation: 'a,b' = rb_problem.sweep_cut1_min, sweep_cut2_min")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--tree-mode", type=int, default=2, choices=[0, 1, 2],
                    help="ablation: branching sweep levels with sweeps.cu (0), tree_sweeps.cu per level (1), fused with the top (2)")
    ap.add_argument("--no-mma", action="store_true", help="ablation: chains with one warp per chain instead of chain_mma.cu")
    ap.add_argument("--no-risk-split", action="store_true",
                    help="ablation: risk block of the chain nodes inside the chain dual pass, not under the sweeps")
    ap.add_argument("--fwd-split", action="store_true",
                    help="ablation: forward chain walk in two pieces, the second overlapped with the dual pass of the first")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="ablation: primal pass + one dual pass per iteration (the loop before the pbar hand-over)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torchrun (python -m torch.distributed.run --nproc-per-node N ...)")
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
