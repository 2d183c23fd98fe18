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


def measured_traffic(workload, batch, dedup, kernel_tag):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the roofline kernel: profiles/traffic.json maps
    "<workload>/<batch>/<dedup>/<kernel tag>" to the bytes of a committed `ncu --set full` capture (file and commit named
    there).  No entry for the kernel this run timed -> null (never a stale literal)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        table = json.load(open(path))
    except Exception:
        return None, None
    ent = table.get(f"{workload}/{batch}/{'dedup' if dedup else 'nodedup'}/{kernel_tag}")
    if not ent:
        return None, None
    return float(ent["dram_bytes_per_launch"]), f"{ent['capture']} @ {ent['commit']}"


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
    """SM clock and throttle reasons DURING the timed region.  NVML is polled from a thread every millisecond (the default
    driver run times 20 iterations = 2 ms: `nvidia-smi -lms` never gets a sample in); the recipe's nvidia-smi query
    (B200_PROFILING.md) is the fallback when pynvml is missing."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index, pci_bus_id=None):
        self.index, self.pci = index, pci_bus_id
        self.proc = self.path = self.thread = None
        self.sm, self.mask, self.stop_flag, self.max_mhz, self.power = [], 0, False, None, 0.0

    def _poll(self, nv, h):
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                self.mask |= int(get(h))
                self.power = max(self.power, nv.nvmlDeviceGetPowerUsage(h) / 1e3)
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByPciBusId(self.pci.encode()) if self.pci else nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.sm:
                out.update(sm_mhz=float(np.median(self.sm)), sm_min_mhz=float(np.min(self.sm)), samples=len(self.sm),
                           power_w_max=self.power, source="NVML polled every 1 ms during the timed region")
            out["sm_max_mhz"] = self.max_mhz
            out["reasons"] = sorted(name for bit, name in self.BITS.items() if self.mask & bit)
            return out
        if self.proc is None:
            return out
        time.sleep(0.05)
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
            out["source"] = "nvidia-smi -lms 20"
        out["reasons"] = sorted(reasons)
        return out


def _x0_of(spec):
    return spec["x0"][:, :1]


def reference_timing(workload, step_budget_s, max_steps, warmup_steps=0):
    """The UNMODIFIED reference (oracle/ref_loader.py: /root/reference or baseline/_ref) on the ACTUAL workload, its own
    Solver methods called in the order of Solver.chock's loop body with the iterate history pruned (oracle/ref_stepper.py;
    BASELINE.md section 3).  Steps until `step_budget_s` seconds of stepping are used up (at least one timed iteration)."""
    from oracle import problems, ref_loader
    from oracle.cp_flat_oracle import FlatOracle
    from oracle.ref_stepper import RefStepper
    api = ref_loader.RefApi()
    spec = problems.spec(workload)
    t0 = time.perf_counter()
    problem = problems.build(spec, api)
    build_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    alpha = FlatOracle(problem).step_size()    # closed-form block lambda_max == ARPACK's to 1e-15; ARPACK itself
    alpha_s = time.perf_counter() - t0         # needs ~65 L / L* pairs through Python callbacks, minutes at cfg3
    st = RefStepper(api, problem, _x0_of(spec), alpha)
    for _ in range(warmup_steps):
        st.step()
    times = []
    while len(times) < max_steps and (not times or sum(times) + times[-1] <= step_budget_s):
        t0 = time.perf_counter()
        st.step()
        times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    return {"value": 1.0 / sec, "unit": "it/s", "cores": 1, "kind": "reference", "extrapolated": False,
            "seconds_per_iteration": sec, "iterations_timed": len(times), "warmup_iterations": warmup_steps,
            "offline_s": st.setup_s, "problem_build_s": build_s, "nodes": int(problem.tree.num_nodes),
            "residuals_last": [float(v) for v in st.xi[-1]],
            "sample": f"{len(times)} full Chambolle-Pock iteration(s) of the unmodified reference Solver on the whole "
                      f"{workload} problem ({int(problem.tree.num_nodes)} nodes; {ref_loader.source()}), manual stepping with "
                      f"pruned history, residual norms every iteration: {sec:.2f} s/iteration; Solver() incl. Cache._offline "
                      f"{st.setup_s:.1f} s, problem build {build_s:.1f} s; single Python process (NumPy on (k,1) blocks is "
                      f"single-threaded: 1 core used of {os.cpu_count()})"}


def port_timing(workload, step_budget_s, max_steps):
    """fallback when the reference is not installed: the node-by-node NumPy port (oracle/cp_node_oracle.py) on the full tree"""
    from oracle import problems
    from oracle.cp_node_oracle import NodeOracle
    from oracle.cp_flat_oracle import FlatOracle
    import raocp_b200 as r
    spec = problems.spec(workload)
    problem = problems.build(spec, r.core)
    t0 = time.perf_counter()
    orc = NodeOracle(problem)
    t_setup = time.perf_counter() - t0
    orc.cache_initial_state(_x0_of(spec))
    orc.alpha = FlatOracle(problem).step_size()
    times = []
    while len(times) < max_steps and (not times or sum(times) + times[-1] <= step_budget_s):
        t0 = time.perf_counter()
        orc.iterate()
        times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    return {"value": 1.0 / sec, "unit": "it/s", "cores": 1, "kind": "port", "extrapolated": False,
            "seconds_per_iteration": sec, "iterations_timed": len(times), "offline_s": t_setup,
            "nodes": int(problem.tree.num_nodes),
            "sample": f"{len(times)} iteration(s) of oracle/cp_node_oracle.py (NumPy, node by node like the reference) on the "
                      f"whole {workload} tree: {sec:.2f} s/iteration; the reference itself is not installed on this box"}


def cpu_baseline(workload, step_budget_s=20.0, max_steps=50, warmup_steps=0):
    from oracle import ref_loader
    if ref_loader.available():
        try:
            return reference_timing(workload, step_budget_s, max_steps, warmup_steps)
        except Exception as exc:   # e.g. a scipy / numpy combination the reference cannot run on
            why = f"{type(exc).__name__}: {exc}"
            out = port_timing(workload, step_budget_s, max_steps)
            out["sample"] += f" (reference failed: {why})"
            return out
    return port_timing(workload, step_budget_s, max_steps)


def run_reference(args, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path on this box's host cores, on the SAME workload
    (whole tree).  cfg3 costs ~40-55 s per iteration on the reference, so K and W are capped by a time budget; the line says
    how many iterations were timed."""
    if rank != 0:
        return
    t0 = time.perf_counter()
    base = cpu_baseline(args.workload, step_budget_s=args.ref_budget_s, max_steps=max(1, args.steps),
                        warmup_steps=1 if args.workload in ("cfg1", "cfg2", "cfg4") else 0)
    line = {"metric": METRIC, "value": base["value"], "unit": "it/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / base["value"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"{args.workload}: {base['nodes']}-node scenario tree, same seeded problem as the CUDA arm",
                       "batch": 1, "steps_timed": base["iterations_timed"], "warmup_done": base.get("warmup_iterations", 0),
                       "note": "the requested --steps / --warmup are capped by --ref-budget-s seconds of stepping"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "offline_s": base.get("offline_s"), "gpu_launches": 0,
            "wall_s": time.perf_counter() - t0}
    emit(line)


def parity_check(workload, r, local_rank):
    """pass / fail carried by every bench line (checker only, nothing here is timed): (1) 30 iterations of the default
    pipelined loop on chain2010 (a 1 621-node sibling of cfg3: same kernels) against the NumPy oracle; (2) when the at-size
    fixture of the workload is committed (tests/golden/<workload>_at_size.npz, recorded from the unmodified reference), 100
    iterations of a fresh solver on the WORKLOAD ITSELF against the reference's residual history, sampled entries and
    per-segment checksums."""
    from oracle import problems
    from oracle.cp_flat_oracle import FlatOracle
    out = {}
    s = problems.spec("chain2010")
    problem = problems.build(s, r.core)
    orc = FlatOracle(problem)
    alpha = orc.step_size()
    sol = r.core.Solver(problem, device=local_rank, verbose=False)
    sol.chock(s["x0"][:, :1], max_iters=29, tol=0.0, alpha=alpha)
    orc.cache_initial_state(s["x0"][:, :1])
    orc.alpha = alpha
    for _ in range(30):
        xi, _ = orc.iterate()
    dev = sol.cache.device_solver
    p, d = dev.get_primal(0)[0], dev.get_dual(0)[0]
    e_it = max(float(np.max(np.abs(p - orc.flat_primal(orc.p))) / max(1.0, np.max(np.abs(p)))),
               float(np.max(np.abs(d - orc.flat_dual(orc.d))) / max(1.0, np.max(np.abs(d)))))
    e_res = float(np.max(np.abs(sol.residual_history[0][-1] - np.array(xi)) / np.array(xi)))
    out["sibling_chain2010_30_iterations_vs_oracle"] = {"iterate_rel_err": e_it, "residual_rel_err": e_res,
                                                        "pass": bool(e_it < 1e-9 and e_res < 1e-6)}
    fixture = os.path.join(ROOT, "tests", "golden", f"{workload}_at_size.npz")
    if workload != "cfg4" and os.path.isfile(fixture):
        from oracle.at_size_check import check_against_fixture
        g = np.load(fixture)
        s = problems.spec(workload)
        sol = r.core.Solver(problems.build(s, r.core), device=local_rank, verbose=False)
        sol.chock(g["x0"], max_iters=99, tol=0.0, alpha=float(g["alpha"]))
        dev = sol.cache.device_solver
        worst = check_against_fixture(sol.cache.flat_problem, g, "", 100, dev.get_primal(0)[0], dev.get_dual(0)[0])
        e_res = float(np.max(np.abs(sol.residual_history[0] - g["xi"]) / g["xi"]))
        out[f"{workload}_100_iterations_vs_reference_fixture"] = {"iterate_rel_err": float(worst), "residual_rel_err": e_res,
                                                                   "pass": bool(worst < 1e-9 and e_res < 1e-6)}
    out["pass"] = all(v["pass"] for v in out.values())
    return out


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
    solver.cache.device_solver.use_batch_panels(not args.no_panels)
    solver.cache.device_solver.use_launch_overlap(args.overlap)
    solver.cache.device_solver.use_table_prefetch(not args.no_table_prefetch)
    solver.cache.device_solver.use_fused_check(args.fused_check)
    solver.cache.device_solver.use_tree_kernels(args.tree_mode)
    solver.cache.device_solver.use_mma_sweeps(0 if args.no_mma else (2 if args.mma_four_warps else (3 if args.mma_one_warp else 1)))
    solver.cache.device_solver.use_pipeline(0 if args.no_pipeline else (3 if args.fwd_split else (4 if args.no_risk_split else 1)))
    dev = solver.cache.device_solver
    dev.synchronize()
    t_setup = time.perf_counter() - t0
    # the offline factorisation on its own ("timed separately"): rb_offline re-run on the uploaded problem
    # (Riccati-like recursion per factorisation class, level by level, + the tensor-core fragment tables), wall time incl. sync
    off_times = []
    for _ in range(3):
        t0 = time.perf_counter()
        dev.offline()
        dev.synchronize()
        off_times.append(time.perf_counter() - t0)
    t_offline = float(min(off_times))
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
        props = torch.cuda.get_device_properties(local_rank)
        pci = None
        if all(hasattr(props, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            pci = f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        sampler = ClockSampler(local_rank, pci)
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
            hist = np.max(cold.residual_history[0], axis=1)            # max(xi0, xi1, xi2) per iteration
            per_it = ttt_s / max(1, int(cold.iterations))
            levels = {}
            for tol in (1e-3, 1e-4, 1e-5, 1e-6):
                hit = np.flatnonzero(hist <= tol)
                levels[f"{tol:.0e}"] = ({"converged": True, "iterations": int(hit[0]) + 1, "seconds": (int(hit[0]) + 1) * per_it}
                                        if hit.size else {"converged": False, "iterations": None, "seconds": None})
            ttt = {"seconds": ttt_s, "iterations": int(cold.iterations), "converged": st_ttt == 0, "tol": 1e-6,
                   "max_iters": args.ttt_iters, "final_residual": float(hist[-1]), "by_tolerance": levels,
                   "reference_demo_tolerance": "the reference's own main.py:80 stops at tol=1e-3",
                   "note": "fresh Solver, zero iterates: Solver.chock(x0, max_iters, tol=1e-6) -- x0 upload, device-side "
                           "stopping test after every iteration, host poll every 64 iterations, residual history download; "
                           "`converged` false means the cap was reached first and `seconds` is the time to `final_residual`; "
                           "by_tolerance: first iteration of THIS run whose max(xi) is below the tolerance, seconds = that "
                           "iteration count x the run's mean wall time per iteration"}
            del cold

    # ---- N > 1: ONE tree sharded by subtree over all ranks (one exchange of the cut-stage messages per iteration) -------------
    #      strong: the cfg3 tree itself; weak: a tree ~N times wider (oracle/problems.py wide_spec), so that every GPU owns about
    #      one cfg3 worth of nodes -- the headline at N > 1
    sharded = {}

    def _exchange_form():
        if os.environ.get("RAOCP_SHARD_P2P", "1") == "0":
            return "ncclAllGather between plain launches"
        return {"split": "NVLink peer memory: push / pull / stopping-test kernels inside the CUDA graph",
                "kernel": "NVLink peer memory: one exchange kernel (buffers + flags) inside the CUDA graph"}.get(
                    os.environ.get("RAOCP_SHARD_XCHG", ""),
                    "NVLink peer memory: 16-byte data+flag packets stored and polled by the kernel that sweeps the top of the tree "
                    "(no exchange launch), inside the CUDA graph")

    if dist is not None and batch == 1 and args.workload == "cfg3":
        def sharded_leg(prob, x0_col):
            sh = r.core.Solver(prob, device=local_rank, verbose=False, shard=(rank, world))
            sdev = sh.cache.device_solver
            sdev.shard_init()
            sdev.use_pipeline(0 if args.no_pipeline else 1)
            sdev.set_stream(stream.cuda_stream)
            a = sh.compute_step_size()
            sdev.set_initial_state(x0_col.reshape(-1))
            barrier()   # the first exchange waits at most ~2 s for a peer: all ranks enter the loop together
            with torch.cuda.stream(stream):
                sdev.loop_begin(a, 1 << 30, -1.0, 0)
                sdev.loop_enqueue(W)
                st_ = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
                en_ = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
                barrier()
                for k in range(K):
                    flush.zero_()
                    st_[k].record(stream)
                    sdev.loop_enqueue(1)
                    en_[k].record(stream)
                barrier()
                cold = float(sum(a_.elapsed_time(b_) for a_, b_ in zip(st_, en_)))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                e0.record(stream)
                sdev.loop_enqueue(K)
                e1.record(stream)
                barrier()
                warm = e0.elapsed_time(e1)
                # end to end: x0 from pinned host memory and the six norms back, every step, on every rank
                xh = torch.from_numpy(np.ascontiguousarray(x0_col.reshape(1, -1))).pin_memory()
                nh = torch.zeros(1, 6, dtype=torch.float64).pin_memory()
                barrier()
                t0_ = time.perf_counter()
                for _ in range(K):
                    sdev.step(xh.data_ptr(), nh.data_ptr())
                torch.cuda.synchronize()
                e2e = time.perf_counter() - t0_
                sdev.loop_end()
            t_ = torch.tensor([cold, warm, e2e], dtype=torch.float64, device="cuda")
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            cold, warm, e2e = t_.tolist()
            fl = sh.cache.flat_problem
            return {"nodes": int(fl.n), "cut_stage": int(sdev.shard_info()[0]), "cut_subtrees": int(sdev.shard_info()[2]),
                    "cold_ms_per_step": cold / K, "warm_ms_per_step": warm / K, "e2e_s_per_step": e2e / K,
                    "np": int(fl.np_), "nd": int(fl.nd_), "exchange": _exchange_form()}
        try:
            sharded["strong"] = sharded_leg(problem, spec["x0"][:, :1])
        except Exception as exc:   # e.g. the tree has fewer cut-stage subtrees than ranks
            sharded["strong"] = {"unavailable": f"{type(exc).__name__}: {exc}"}
        try:
            wide = problems.wide_spec(world)
            sharded["weak"] = sharded_leg(problems.build(wide, r.core), wide["x0"][:, :1])
        except Exception as exc:
            sharded["weak"] = {"unavailable": f"{type(exc).__name__}: {exc}"}

    cold_total, e2e_t = float(cold_ms.sum()), e2e_s
    if dist is not None:
        t = torch.tensor([cold_total, warm_ms, e2e_t, solve_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cold_total, warm_ms, e2e_t, solve_s = t.tolist()
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
        ktag = f"k_dual_chain<{flat.nx},{flat.nu}>" + ("" if not args.no_risk_split else "+risk")
        traffic, traffic_src = measured_traffic(args.workload, batch, not args.no_dedup, ktag)
    elif len(cold_parts) == 3 and batch >= 64 and not args.no_panels:
        # batch-innermost panel path (csrc/batch.cu): the dominant kernel is the x / u block of the nonleaf nodes' dual pass,
        # k_bp_dual_xu -- x_i, u_i, tau_j of the primal and d7_i, d3_j, d4_j, d5_j, d6_j of the dual, every entry read once and
        # written once (SURVEY 8d restricted to these entries); it also reads p+ and writes pbar (bytes_moved_model)
        e_p = flat.m * (flat.nx + flat.nu) + (flat.n - 1)
        e_d = flat.m * (flat.nx + flat.nu) + (flat.n - 1) * (flat.nx + flat.nu + 2)
        b_kernel = 16 * batch * (e_p + e_d)
        b_moved = 8 * batch * (3 * e_p + 2 * e_d)
        t_kernel = float(cold_parts[0]) * 1e-3
        kname = (f"k_bp_dual_xu<{flat.nx},{flat.nu}> (batch-innermost layout, lanes = instances): x / u block of the dual pass of all "
                 f"{flat.m} nonleaf nodes x {batch} instances (L, dual half step, rectangle and second-order-cone projections, "
                 "residual maxima, pbar of the next iteration)")
        traffic, traffic_src = measured_traffic(args.workload, batch, not args.no_dedup, f"k_bp_dual_xu<{flat.nx},{flat.nu}>")
    else:
        b_kernel = b_moved = 8 * batch * (2 * flat.np_ + 2 * flat.nd_)
        t_kernel = float(cold_phases[-1]) * 1e-3
        kname = "dual pass (L, dual half step, prox of g*, six residual norms), all nodes"
        traffic, traffic_src = None, None
    serial_ms = float(cold_phases.sum())
    roofline = {
        "bound": "hbm", "kernel": kname,
        "achieved": b_kernel / t_kernel / 1e9, "peak": peak, "unit": "GB/s", "frac": b_kernel / t_kernel / 1e9 / peak,
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes": b_kernel, "bytes_moved_model": b_moved,
        "launch_ms": t_kernel * 1e3, "share_of_step": float(t_kernel * 1e3 / serial_ms),
        "timing": "CUDA events around the plain launch on the bench stream, L2 flushed before every iteration, mean of "
                  f"{min(K, 50)} iterations (includes ~2 us of launch gap); share_of_step = launch_ms / sum of all launches "
                  "of the iteration run ONE AFTER THE OTHER (the ncu launch list is serialised the same way) -- not a share "
                  "of the overlapped graph iteration, where several of these launches run side by side",
        "iteration": {"algorithmic_bytes": b_alg, "achieved": achieved, "frac": achieved / peak,
                      "note": "all launches of one CP iteration inside the CUDA graph, the timed region of `value`"},
        "launch_ms_all": {"primal_or_kernel_projection": float(cold_phases[0]),
                          "sweeps_in_launch_order": [float(v) for v in cold_phases[1:-1]],
                          "dual_and_check": float(cold_phases[-1]),
                          "dual_kernels_branching_chain_leaves": [float(v) for v in cold_parts],
                          "note": "panel path (batch >= 64): kernel projection | backward stage launches | fused top + forward stage "
                                  "launches | dual + check; the three dual kernels are x/u block, risk block, leaves"
                          if batch >= 64 and not args.no_panels else "see DESIGN.md section 4"},
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
        "offline_s": {"value": t_offline, "what": "rb_offline (Cache._offline, cache.py:200-242) on the device, wall time incl. "
                      "synchronise, best of 3; the reference's time for the same call is cpu_baseline.offline_s",
                      "classes": flat.num_cls, "per_node_classes": bool(args.no_dedup)},
        "residuals_last": [float(v) for v in last_norms[0]],
    }
    if world > 1 and sharded:
        # headline at N > 1 for the tree workload: the WIDE tree sharded by subtree (weak scaling: one cfg3 worth of nodes per
        # GPU, one exchange per iteration).  value = cfg3-equivalents advanced by one iteration per second = (nodes / cfg3
        # nodes) x iterations/s; the N independent replicas measured above move to `replicas`.
        line["replicas"] = {"value": line["value"], "unit": "it/s", "ms_per_step": line["ms_per_step"],
                            "warm": line["warm"], "e2e": line["e2e"],
                            "note": f"{world} independent cfg3 instances, one per GPU, no collective (linear by construction)"}
        wk, sg = sharded.get("weak", {}), sharded.get("strong", {})
        if "cold_ms_per_step" in sg:
            line["sharded_cfg3_strong"] = {
                "value": 1e3 / sg["cold_ms_per_step"], "warm_value": 1e3 / sg["warm_ms_per_step"], "unit": "it/s", "scaling": "strong",
                "ms_per_step": sg["cold_ms_per_step"], "cut_stage": sg["cut_stage"], "exchange": sg["exchange"],
                "note": f"ONE {flat.n}-node cfg3 tree over {world} GPUs (latency-bound: sharding cuts the width of the sweeps, not "
                        "their depth); compare with the N=1 value"}
        else:
            line["sharded_cfg3_strong"] = sg
        if "cold_ms_per_step" in wk:
            equiv = wk["nodes"] / flat.n
            line.update({
                "value": equiv * 1e3 / wk["cold_ms_per_step"], "ms_per_step": wk["cold_ms_per_step"], "scaling": "weak",
                "warm": {"value": equiv * 1e3 / wk["warm_ms_per_step"], "unit": "it/s", "ms_per_step": wk["warm_ms_per_step"],
                         "note": "back to back, no L2 flush"},
                "e2e": {"value": equiv / wk["e2e_s_per_step"], "unit": "it/s", "h2d_bytes_per_step": int(world * flat.nx * 8),
                        "d2h_bytes_per_step": int(world * 48),
                        "note": "rb_step per iteration on every rank: x0 from pinned host memory, one sharded iteration, residual "
                                "norms to the host (those of the previous iteration: the stopping test of iteration k rides on the "
                                "exchange of iteration k + 1), synchronised every step"}})
            line["config"].update({
                "workload": f"cfg3 widened for {world} GPUs: ONE {wk['nodes']}-node scenario tree (= {equiv:.2f} x cfg3's {flat.n} nodes), "
                            f"nx={flat.nx}, nu={flat.nu}, AVaR(0.5), rectangles, seed 0, sharded by subtree below stage "
                            f"{wk['cut_stage']} ({wk['cut_subtrees']} subtrees), top of the tree replicated",
                "parallelism": f"subtree sharding over {world} GPUs, one exchange of the cut-stage q_j, sbar_j and residual maxima per "
                               f"iteration ({wk['exchange']})",
                "value_definition": "cfg3-equivalents x iterations/s = (tree nodes / 62 805) x CP iterations per second of the "
                                    "sharded tree; N = 1 is the plain cfg3 value"})
            b_w = 8 * 2 * (wk["np"] + wk["nd"])
            line["roofline"]["iteration"] = {"algorithmic_bytes": b_w, "achieved": b_w / (wk["cold_ms_per_step"] * 1e-3) / 1e9,
                                             "frac": b_w / (wk["cold_ms_per_step"] * 1e-3) / 1e9 / (world * peak),
                                             "note": f"sharded wide tree, all {world} GPUs: B_alg / t_iteration / (N x peak)"}
            line["roofline"]["note"] = "the per-kernel figures above are rank 0's replica leg (one cfg3 instance on one GPU)"
        else:
            line["sharded_weak"] = wk
    if not args.no_parity:
        try:
            line["parity_check"] = parity_check(args.workload, r, local_rank)
        except Exception as exc:
            line["parity_check"] = {"pass": False, "error": f"{type(exc).__name__}: {exc}"}
    if not args.no_cpu and world == 1:
        line["cpu_baseline"] = cpu_baseline(args.workload)
    if ttt is not None:
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
    ap.add_argument("--ref-budget-s", type=float, default=150.0,
                    help="--impl reference: seconds of stepping the reference may use (at least one iteration is timed)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity_check leg")
    ap.add_argument("--fused-check", action="store_true",
                    help="ablation: the stopping test by the last CTA of the dual passes instead of a k_check launch at the end of every "
                         "iteration (measured slower)")
    ap.add_argument("--no-table-prefetch", action="store_true",
                    help="ablation: no L2 prefetch of the operator tables at the head of the iteration")
    ap.add_argument("--overlap", action="store_true",
                    help="ablation: chain the walkers and the fused tree kernel by programmatic dependent launch")
    ap.add_argument("--no-panels", action="store_true",
                    help="ablation (batch >= 64): instance-major kernels instead of the batch-innermost panel path")
    ap.add_argument("--tree-mode", type=int, default=2, choices=[0, 1, 2],
                    help="ablation: branching sweep levels with sweeps.cu (0), tree_sweeps.cu per level (1), fused with the top (2)")
    ap.add_argument("--no-mma", action="store_true", help="ablation: chains with one warp per chain instead of chain_mma.cu")
    ap.add_argument("--mma-one-warp", action="store_true",
                    help="ablation (wide rows, nx=64 nu=32): one warp per chain tile instead of four")
    ap.add_argument("--mma-four-warps", action="store_true",
                    help="ablation: tensor-core chain walkers with four warps per tile (one output block each)")
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
