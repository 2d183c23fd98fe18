"""DeviceSolver: thin object wrapper over the C-ABI handle (include/raocp_b200.h) -- no arithmetic here."""
import ctypes as C

import numpy as np

from .. import _lib
from .flatten import FlatProblem


class DeviceSolver:
    def __init__(self, flat: FlatProblem):
        self.flat = flat
        self._lib = _lib.load()
        self._keep = []  # host arrays referenced by the rb_problem struct during rb_create

        def ip(arr):
            arr = np.ascontiguousarray(arr, dtype=np.int32)
            self._keep.append(arr)
            return _lib.iptr(arr)

        def dp(arr):
            arr = np.ascontiguousarray(arr, dtype=np.float64)
            self._keep.append(arr)
            return _lib.dptr(arr)

        f = flat
        pb = _lib.RbProblem()
        pb.n, pb.m, pb.nx, pb.nu, pb.num_stages, pb.batch = f.n, f.m, f.nx, f.nu, f.num_stages, f.batch
        pb.stage_off, pb.parent = ip(f.stage_off), ip(f.parent)
        pb.child_first, pb.child_count = ip(f.child_first), ip(f.child_count)
        pb.num_dyn, pb.dyn_idx, pb.A, pb.B = f.A.shape[0], ip(f.dyn_idx), dp(f.A), dp(f.B)
        pb.num_cost, pb.cost_idx, pb.sqrtQ, pb.sqrtR = f.sqrtQ.shape[0], ip(f.cost_idx), dp(f.sqrtQ), dp(f.sqrtR)
        pb.num_leafcost, pb.leafcost_idx, pb.sqrtQf = f.sqrtQf.shape[0], ip(f.leafcost_idx), dp(f.sqrtQf)
        if f.nl_rect:
            pb.num_nl_rect, pb.nl_rect_idx = f.nonleaf_lo.shape[0], ip(f.nonleaf_rect_idx)
            pb.nl_lo, pb.nl_hi = dp(f.nonleaf_lo), dp(f.nonleaf_hi)
        else:
            pb.num_nl_rect = 0
        if f.leaf_rect:
            pb.num_leaf_rect, pb.leaf_rect_idx = f.leaf_lo.shape[0], ip(f.leaf_rect_idx)
            pb.leaf_lo, pb.leaf_hi = dp(f.leaf_lo), dp(f.leaf_hi)
        else:
            pb.num_leaf_rect = 0
        pb.risk_alpha, pb.cond_prob = dp(f.risk_alpha), dp(f.cond_prob)
        pb.num_cls, pb.cls = f.num_cls, ip(f.cls)
        pb.device = f.device
        pb.shard_rank, pb.shard_world = f.shard_rank, f.shard_world
        pb.sweep_cut1_min, pb.sweep_cut2_min = f.sweep_cuts
        handle = C.c_void_p()
        _lib.check(self._lib.rb_create(C.byref(pb), C.byref(handle)))
        self._h = handle
        self._keep = []
        np_, nd_ = C.c_int64(), C.c_int64()
        self._call("rb_sizes", C.byref(np_), C.byref(nd_))
        assert (np_.value, nd_.value) == (f.np_, f.nd_), "host / device layout mismatch"
        self.np_, self.nd_, self.batch = f.np_, f.nd_, f.batch
        f.shard_cut = self.shard_info()[0]     # the device's cut stage is the only source of truth (flatten.shard_*)

    def _call(self, name, *args):
        _lib.check(getattr(self._lib, name)(self._h, *args), self._h)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing ------------------------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        self._call("rb_set_stream", C.c_void_p(cuda_stream_ptr))

    def synchronize(self):
        self._call("rb_synchronize")

    def launch_count(self):
        k = C.c_int64()
        self._call("rb_launch_count", C.byref(k))
        return k.value

    def offline(self):
        self._call("rb_offline")

    def get_offline(self):
        f = self.flat
        p = np.empty((f.num_cls, f.nx, f.nx))
        k = np.empty((f.num_cls, f.nu, f.nx))
        ri = np.empty((f.num_cls, f.nu, f.nu))
        self._call("rb_get_offline", _lib.dptr(p), _lib.dptr(k), _lib.dptr(ri))
        return p, k, ri

    def _batched(self, compact, size):
        arr = np.ascontiguousarray(np.asarray(compact, dtype=np.float64))
        if arr.size == size and self.batch > 1:
            arr = np.ascontiguousarray(np.broadcast_to(arr.reshape(1, size), (self.batch, size)))
        if arr.size != size * self.batch:
            raise Exception(f"expected {self.batch} x {size} doubles, got {arr.size}")
        return arr

    def set_primal(self, which, compact):
        arr = self._batched(compact, self.np_)
        self._call("rb_set_primal", which, _lib.dptr(arr))

    def get_primal(self, which):
        out = np.empty((self.batch, self.np_))
        self._call("rb_get_primal", which, _lib.dptr(out))
        return out

    def set_dual(self, which, compact):
        arr = self._batched(compact, self.nd_)
        self._call("rb_set_dual", which, _lib.dptr(arr))

    def get_dual(self, which):
        out = np.empty((self.batch, self.nd_))
        self._call("rb_get_dual", which, _lib.dptr(out))
        return out

    def set_initial_state(self, x0):
        """x0: (nx,), (nx,1) or (nx, batch) like the columns of the reference's initial_state"""
        f = self.flat
        arr = np.asarray(x0, dtype=np.float64)
        if arr.size == f.nx:
            arr = np.broadcast_to(arr.reshape(1, f.nx), (self.batch, f.nx))
        elif arr.shape == (f.nx, self.batch):
            arr = arr.T
        else:
            raise Exception(f"initial state must have {f.nx} entries (or shape ({f.nx}, {self.batch}))")
        arr = np.ascontiguousarray(arr)
        self._call("rb_set_initial_state", _lib.dptr(arr))

    def update_cache(self):
        self._call("rb_update_cache")

    def apply_L(self, primal_compact):
        arr = self._batched(primal_compact, self.np_)
        out = np.empty((self.batch, self.nd_))
        self._call("rb_apply_L", _lib.dptr(arr), _lib.dptr(out))
        return out

    def apply_Lt(self, dual_compact):
        arr = self._batched(dual_compact, self.nd_)
        out = np.empty((self.batch, self.np_))
        self._call("rb_apply_Lt", _lib.dptr(arr), _lib.dptr(out))
        return out

    def lambda_max(self):
        v = C.c_double()
        self._call("rb_lambda_max", C.byref(v))
        return v.value

    # ---- steps ---------------------------------------------------------------------------------------------------
    def primal_half(self, alpha):
        self._call("rb_primal_half", float(alpha))

    def prox_f(self, alpha):
        self._call("rb_prox_f", float(alpha))

    def s0_shift(self, alpha):
        self._call("rb_s0_shift", float(alpha))

    def project_dynamics(self):
        self._call("rb_project_dynamics")

    def project_kernel(self):
        self._call("rb_project_kernel")

    def dual_half(self, alpha):
        self._call("rb_dual_half", float(alpha))

    def prox_g_conj(self, alpha):
        self._call("rb_prox_g_conj", float(alpha))

    def modify_dual(self, alpha):
        self._call("rb_modify_dual", float(alpha))

    def add_halves(self):
        self._call("rb_add_halves")

    def project_nonleaf(self):
        self._call("rb_project_nonleaf")

    def project_leaf(self):
        self._call("rb_project_leaf")

    def modify_projection(self, alpha, modified_dual_compact):
        arr = self._batched(modified_dual_compact, self.nd_)
        self._call("rb_modify_projection", float(alpha), _lib.dptr(arr))

    def residuals(self, alpha, vectors=False):
        norms = np.empty((self.batch, 6))
        vec = np.empty((self.batch, 4 * self.np_ + 2 * self.nd_)) if vectors else None
        self._call("rb_residuals", float(alpha), _lib.dptr(norms), _lib.dptr(vec) if vectors else None)
        return norms, vec

    def _shard_rendezvous(self):
        """sharded solve with the peer-memory exchange: the device waits ~2 s for a peer's packets before it reports the peer as gone,
        so the ranks enter a loop together (torch.distributed barrier; a no-op for single-GPU solvers)"""
        if getattr(self, "_p2p_ready", False):
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                dist.barrier()

    def iterate(self, alpha, max_iters, tol, history=True):
        self._shard_rendezvous()
        cap = max_iters + 1
        xi = np.zeros((cap, self.batch, 3)) if history else None
        delta = np.zeros((cap, self.batch, 3)) if history else None
        iters, status = C.c_int32(), C.c_int32()
        self._call("rb_iterate", float(alpha), int(max_iters), float(tol), 1,
                   _lib.dptr(xi) if history else None, _lib.dptr(delta) if history else None, cap,
                   C.byref(iters), C.byref(status))
        if history:
            xi, delta = xi[: iters.value], delta[: iters.value]
        return status.value, iters.value, xi, delta

    def iterate_fixed(self, alpha, iters):
        self._shard_rendezvous()
        norms = np.empty((self.batch, 6))
        self._call("rb_iterate_fixed", float(alpha), int(iters), _lib.dptr(norms))
        return norms

    # ---- the fused loop in pieces (hosts that pipeline; bench.py) -----------------------------------------------------
    def use_graphs(self, enable=True):
        self._call("rb_use_graphs", 1 if enable else 0)

    def use_table_prefetch(self, enable=True):
        """pipelined loop: L2 prefetch of the operator tables at the head of every iteration (default on)"""
        self._call("rb_use_table_prefetch", 1 if enable else 0)

    def use_fused_check(self, enable=True):
        """pipelined loop, batch 1: stopping test by the last CTA of the dual passes (measured ablation) or by a k_check launch (default)"""
        self._call("rb_use_fused_check", 1 if enable else 0)

    def use_launch_overlap(self, enable=True):
        """pipelined loop: chain the walkers and the fused tree kernel by programmatic dependent launch (default) or not"""
        self._call("rb_use_launch_overlap", 1 if enable else 0)

    def use_batch_panels(self, enable=True):
        """batch >= 64: fused loop in the batch-innermost panel layout (default) or with the instance-major kernels"""
        self._call("rb_use_batch_panels", 1 if enable else 0)

    def use_pipeline(self, enable=True):
        """True / 1: pipelined loop (default); 3: additionally the forward chain walk in two pieces, overlapped with the
        dual pass (ablation); 4: the risk block of the chain nodes inside the chain dual pass instead of a kernel of its own
        under the sweeps (ablation); False / 0: primal pass + one dual pass per iteration"""
        self._call("rb_use_pipeline", int(enable))

    def pipeline_info(self):
        """(early, chain_first, chain_nodes): the node ranges of the pipelined dual pass (rb_pipeline_info)"""
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self._call("rb_pipeline_info", C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def chain_dual_nodes(self):
        return self.pipeline_info()[2]

    def force_dense_costs(self, enable=True):
        self._call("rb_force_dense_costs", 1 if enable else 0)

    def use_lane_kernels(self, enable=True):
        self._call("rb_use_lane_kernels", 1 if enable else 0)

    def use_mma_sweeps(self, enable=True):
        """False / 0: one warp per chain (sweeps.cu); True / 1 (default): tensor cores, one warp per tile; 2: tensor cores, four
        warps per tile where instantiated (ablation)"""
        self._call("rb_use_mma_sweeps", int(enable))

    def use_tree_kernels(self, mode=2):
        """0: sweeps.cu stage kernels; 1: tree_sweeps.cu per level; 2: level 0 + top fused (default)"""
        self._call("rb_use_tree_kernels", int(mode))

    def loop_begin(self, alpha, max_iters, tol=-1.0, hist_capacity=0):
        self._shard_rendezvous()
        self._call("rb_loop_begin", float(alpha), int(max_iters), float(tol), int(hist_capacity))

    def loop_enqueue(self, count=1):
        self._call("rb_loop_enqueue", int(count))

    def loop_poll(self):
        iters, done = C.c_int32(), C.c_int32()
        norms = np.empty((self.batch, 6))
        self._call("rb_loop_poll", C.byref(iters), C.byref(done), _lib.dptr(norms))
        return iters.value, bool(done.value), norms

    def step(self, x0_ptr, norms_ptr):
        """raw pointers (pinned host memory): x0 [batch][nx] in, norms [batch][6] out"""
        self._call("rb_step", C.cast(x0_ptr, _lib.c_double_p), C.cast(norms_ptr, _lib.c_double_p))

    def loop_end(self):
        iters, status = C.c_int32(), C.c_int32()
        self._call("rb_loop_end", None, None, C.byref(iters), C.byref(status))
        return iters.value, status.value

    def profile_iteration(self):
        """ms per launch of one iteration (CUDA events): primal (or kernel projection), sweep launches in order, dual +
        stopping test"""
        return self.profile_iteration_full()[0]

    def profile_iteration_full(self):
        """(phases, dual_parts): phases as profile_iteration(); dual_parts = the kernels of the pipelined dual pass on
        their own (branching nodes, chain nodes, leaves -- or one entry: all nodes); empty for the unpipelined loop"""
        ms = (C.c_float * 12)()
        self._call("rb_profile_iteration", ms)
        return [v for v in ms[:8] if v >= 0.0], [v for v in ms[8:] if v >= 0.0]

    # ---- subtree sharding over GPUs (one process per GPU) ------------------------------------------------------------------
    def shard_info(self):
        """(cut_stage, cut_first, num_cut, chain_stage) as chosen by rb_create (rb_shard_info)"""
        v = [C.c_int32() for _ in range(4)]
        self._call("rb_shard_info", *[C.byref(x) for x in v])
        return tuple(x.value for x in v)

    def shard_init(self, unique_id=None):
        """collective: rank 0 creates the NCCL id, torch.distributed carries it, every rank joins the communicator"""
        import torch
        import torch.distributed as dist
        buf = (C.c_char * 128)()
        if unique_id is None:
            if dist.get_rank() == 0:
                _lib.check(self._lib.rb_shard_unique_id(buf))
            t = torch.tensor(list(buf.raw), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.broadcast(t, src=0)
            unique_id = bytes(t.cpu().tolist())
        self._call("rb_shard_init", C.c_char_p(unique_id))
        import os
        if os.environ.get("RAOCP_SHARD_P2P", "1") != "0" and dist.get_world_size() <= 8:
            self.shard_p2p_init()

    def shard_p2p_init(self):
        """collective: exchange the CUDA IPC handles of the ranks' receive buffers / flags and map them (rb_shard_p2p_*): from
        then on the cut-stage exchange is device-initiated over NVLink peer memory and the sharded loop is pipelined + graphed"""
        import torch
        import torch.distributed as dist
        mine = (C.c_char * 128)()
        self._call("rb_shard_p2p_export", mine)
        t = torch.tensor(list(mine.raw), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        blob = b"".join(bytes(o.cpu().tolist()) for o in out)
        self._call("rb_shard_p2p_open", C.c_char_p(blob))
        self._p2p_ready = True

    def gather_sharded(self, which=0):
        """assemble the full compact iterates on every rank from the ranks' authoritative parts (torch.distributed)"""
        import torch
        import torch.distributed as dist
        f = self.flat
        pm, dm = f.shard_masks(f.shard_rank, f.shard_world)
        p = torch.from_numpy(np.where(pm, self.get_primal(which)[0], 0.0))
        d = torch.from_numpy(np.where(dm, self.get_dual(which)[0], 0.0))
        if dist.get_backend() == "nccl":
            p, d = p.cuda(), d.cuda()
        dist.all_reduce(p)
        dist.all_reduce(d)
        return p.cpu().numpy(), d.cpu().numpy()
