/*
 * raocp_b200.h -- C-ABI of the B200-native Chambolle-Pock path of raocp-toolbox.
 *
 * The reference (pure Python, /root/reference) has no FFI layer; its boundary for this path is the class API
 * raocp.core.Cache / Operator / Solver.  Each entry point below cites the reference interface it replaces
 * (paths relative to /root/reference).  Plain pointers and sizes only; all host buffers are caller-owned, all device
 * memory is library-owned for the lifetime of the handle.  One handle per host thread.  Every function returns
 * 0 on success or a negative rb_status; rb_last_error() gives the message.
 *
 * COMPACT EXCHANGE LAYOUT (float64, per problem instance; no placeholders -- see DESIGN.md "Data layout"):
 *   primal (Np doubles) = [ x (n*nx) | u (m*nu) | y (sum_i 2c_i+1; node i: [y_a(c_i); y_b(c_i); y_last]) | tau (n) | s (n) ]
 *   dual   (Nd doubles) = [ d1 (like y) | d2 (m) | d3 ((n-1)*nx, row j-1 = edge into node j) | d4 ((n-1)*nu) | d5 (n-1)
 *                           | d6 (n-1) | d7 (m*(nx+nu), iff nonleaf rectangles) | d11 (L*nx) | d12 (L) | d13 (L)
 *                           | d14 (L*nx, iff leaf rectangles) ],  L = n - m.
 *   Reference numbering of these segments: raocp/core/cache.py:126-170.  With batch > 1 the instances are
 *   concatenated (instance-major).
 */
#ifndef RAOCP_B200_H
#define RAOCP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rb_solver rb_solver; /* opaque handle */

typedef enum {
    RB_OK = 0,
    RB_ERR_INVALID = -1,      /* bad argument / inconsistent problem description */
    RB_ERR_CUDA = -2,         /* CUDA runtime error (message has the cudaError string) */
    RB_ERR_NO_DEVICE = -3,    /* no usable CUDA device: there is NO CPU fallback */
    RB_ERR_NUMERIC = -4,      /* NaN met in a rectangle projection (reference: ValueError, rectangle.py:58-59)
                                 or Cholesky breakdown in the offline factorisation */
    RB_ERR_STATE = -5         /* call order error (e.g. iterate before offline / initial state) */
} rb_status;

/* Problem description = what the reference's Cache reads through RAOCP accessors (raocp/core/raocp_spec.py:56-75)
 * and ScenarioTree queries (raocp/core/scenario_tree.py:75-154), flattened once on the host.
 * Nodes are numbered stage by stage, children of a node contiguous (reference factory order,
 * scenario_tree.py:273-315); nonleaf nodes are 0..m-1. */
typedef struct {
    int32_t n;                 /* nodes                     (tree.num_nodes) */
    int32_t m;                 /* nonleaf nodes             (tree.num_nonleaf_nodes) */
    int32_t nx, nu;            /* state / input sizes       (cache.py:19-20) */
    int32_t num_stages;        /* N+1                       (tree.num_stages) */
    int32_t batch;             /* independent initial states solved together (>=1; reference: 1) */
    const int32_t *stage_off;  /* [num_stages+1] first node of every stage (nodes_at_stage, scenario_tree.py:123) */
    const int32_t *parent;     /* [n]  ancestor_of, -1 at the root */
    const int32_t *child_first;/* [m]  children_of(i)[0] */
    const int32_t *child_count;/* [m]  len(children_of(i)) */
    /* dynamics, indexed by the CHILD node j>=1 through a table (raocp_spec.py:56-60) */
    int32_t num_dyn;
    const int32_t *dyn_idx;    /* [n]  table row of node j (entry 0 ignored) */
    const double *A;           /* [num_dyn][nx][nx] row-major */
    const double *B;           /* [num_dyn][nx][nu] row-major */
    /* nonleaf costs, indexed by the CHILD node (operators.py:33-36,84-87): matrix square roots as stored by the
     * host cost objects (costs.py:21,26) */
    int32_t num_cost;
    const int32_t *cost_idx;   /* [n] */
    const double *sqrtQ;       /* [num_cost][nx][nx] */
    const double *sqrtR;       /* [num_cost][nu][nu] */
    /* leaf costs (operators.py:46-47,89) */
    int32_t num_leafcost;
    const int32_t *leafcost_idx; /* [L] */
    const double *sqrtQf;      /* [num_leafcost][nx][nx] */
    /* rectangles (rectangle.py:24-35): bounds on [x;u] (nonleaf) and x (leaf); 0 tables = constraint inactive */
    int32_t num_nl_rect;
    const int32_t *nl_rect_idx;  /* [m] */
    const double *nl_lo, *nl_hi; /* [num_nl_rect][nx+nu] */
    int32_t num_leaf_rect;
    const int32_t *leaf_rect_idx;/* [L] */
    const double *leaf_lo, *leaf_hi; /* [num_leaf_rect][nx] */
    /* AVaR risks (risks.py:28-35): alpha per nonleaf node, conditional probability of every node given its parent */
    const double *risk_alpha;  /* [m] */
    const double *cond_prob;   /* [n]  (entry 0 ignored) = b_i[0:c_i] of the parent */
    /* factorisation classes: nodes with the same class share (P, K, R~^-1).  cls[i] in [0, num_cls); classes must be
     * ordered so that every child's class index is LARGER than its parent's (leaves have class -1 -> P = I).
     * The trivial choice cls[i] = i (one class per node) reproduces the reference's per-node loop (cache.py:207-233). */
    int32_t num_cls;
    const int32_t *cls;        /* [m] */
    int32_t device;            /* CUDA device ordinal */
    /* subtree sharding of ONE tree over the GPUs of a box: every rank passes the whole problem and its
     * (rank, world); world <= 1 = not sharded.  Rank r owns a contiguous block of the subtrees below the first stage
     * with >= 64 nodes and replicates the nodes above it; rb_shard_init must follow rb_create. */
    int32_t shard_rank, shard_world;
    /* tuning / test hook for the cut stages of the DP sweeps (sweeps.cu); 0 = defaults (64 nodes, 200 chains): the first cut
     * is the first stage with >= sweep_cut1_min nodes, the second the stage where the tree turns into chains if it has
     * >= sweep_cut2_min of them.  Must be 0 when shard_world > 1. */
    int32_t sweep_cut1_min, sweep_cut2_min;
} rb_problem;

/* -- lifetime ---------------------------------------------------------------------------------------------------- */
/* Cache.__init__ without the offline phase (cache.py:13-47): uploads the problem, allocates zero iterates. */
int rb_create(const rb_problem *problem, rb_solver **out);
void rb_destroy(rb_solver *s);
const char *rb_last_error(const rb_solver *s);       /* s may be NULL: error of the last failed rb_create */
int rb_set_stream(rb_solver *s, void *cuda_stream);  /* launch on this cudaStream_t (NULL = library's own stream) */
int rb_sizes(const rb_solver *s, int64_t *np, int64_t *nd); /* compact per-instance sizes Np, Nd */
int rb_synchronize(rb_solver *s);

/* -- offline ----------------------------------------------------------------------------------------------------- */
/* Cache._offline (cache.py:200-242): Riccati-like factorisation per class.  The null-space matrices of
 * offline_projection_kernel (cache.py:235-242) are not needed: the AVaR kernel projector is closed-form. */
int rb_offline(rb_solver *s);
/* test hook: per-class P [num_cls][nx][nx], K [num_cls][nu][nx], R~^-1 [num_cls][nu][nu] (any pointer may be NULL) */
int rb_get_offline(rb_solver *s, double *P, double *K, double *Rinv);

/* -- iterate access (cache.py:59-122) ---------------------------------------------------------------------------- */
/* which: 0 = current iterate (Cache.__primal / __dual), 1 = old iterate (__old_primal / __old_dual) */
int rb_set_primal(rb_solver *s, int which, const double *compact); /* Cache.set_primal, cache.py:84-101 */
int rb_get_primal(rb_solver *s, int which, double *compact);       /* Cache.get_primal, cache.py:59-60 */
int rb_set_dual(rb_solver *s, int which, const double *compact);   /* Cache.set_dual, cache.py:103-122 */
int rb_get_dual(rb_solver *s, int which, double *compact);         /* Cache.get_dual, cache.py:65-66 */
int rb_set_initial_state(rb_solver *s, const double *x0);          /* Cache.cache_initial_state, cache.py:79-82;
                                                                      x0 is [batch][nx] */
int rb_update_cache(rb_solver *s);                                 /* Cache.update_cache, cache.py:186-196: old <- current */

/* -- linear operator (operators.py:19-94), host vectors in / out (compact layout, batch instances) ---------------- */
int rb_apply_L(rb_solver *s, const double *primal_in, double *dual_out);   /* Operator.ell / linop_ell */
int rb_apply_Lt(rb_solver *s, const double *dual_in, double *primal_out);  /* Operator.ell_transpose / linop_ell_transpose */
/* lambda_max(L* L) as the maximum over the diagonal blocks of L* L (replaces ARPACK eigs, solver.py:105-118) */
int rb_lambda_max(rb_solver *s, double *lambda_max);

/* -- the four half steps on the device-resident iterates (solver.py:27-61) ----------------------------------------- */
int rb_primal_half(rb_solver *s, double alpha);   /* Solver.primal_k_plus_half: current_p = old_p - alpha L*(old_d) */
int rb_prox_f(rb_solver *s, double alpha);        /* Cache.proximal_of_f, cache.py:248-251 (on current_p) */
int rb_s0_shift(rb_solver *s, double alpha);      /* Cache.proximal_of_relaxation_s_at_stage_zero, cache.py:253-257 */
int rb_project_dynamics(rb_solver *s);            /* Cache.project_on_dynamics, cache.py:259-288 */
int rb_project_kernel(rb_solver *s);              /* Cache.project_on_kernel, cache.py:290-317 */
int rb_dual_half(rb_solver *s, double alpha);     /* Solver.dual_k_plus_half: current_d = old_d + alpha L(2 cur_p - old_p) */
int rb_prox_g_conj(rb_solver *s, double alpha);   /* Cache.proximal_of_g_conjugate, cache.py:321-327 (on current_d) */
int rb_modify_dual(rb_solver *s, double alpha);   /* Cache.modify_dual, cache.py:329-332 */
int rb_add_halves(rb_solver *s);                  /* Cache.add_halves, cache.py:334-347 */
int rb_project_nonleaf(rb_solver *s);             /* Cache.project_on_constraints_nonleaf, cache.py:349-371 */
int rb_project_leaf(rb_solver *s);                /* Cache.project_on_constraints_leaf, cache.py:373-390 */
int rb_modify_projection(rb_solver *s, double alpha, const double *modified_dual); /* cache.py:392-393 */

/* Solver._calculate_chock_errors + the inf-norms (solver.py:63-95,137-141) between old and current iterates.
 * norms: [batch][6] = xi0, xi1, xi2, delta0, delta1, delta2.  vectors (optional, may be NULL): six compact host
 * vectors [xi0 (Np) | xi1 (Np) | xi2 (Nd) | delta0 (Np) | delta1 (Np) | delta2 (Nd)] per instance. */
int rb_residuals(rb_solver *s, double alpha, double *norms, double *vectors);

/* -- the fused loop: Solver.chock's while-loop (solver.py:124-161) entirely on the device ---------------------------
 * Runs until iteration index >= max_iters or max(xi0,xi1,xi2) <= tol for every instance (so max_iters+1 iterations
 * when not converged, like the reference).  Only the six residual norms per iteration leave the device.
 * xi_hist / delta_hist: [capacity][batch][3] (may be NULL).  Returns the reference's status in *status
 * (0 converged, 1 not) and the number of iterations executed in *iters. */
int rb_iterate(rb_solver *s, double alpha, int32_t max_iters, double tol, int32_t check_every,
               double *xi_hist, double *delta_hist, int32_t hist_capacity, int32_t *iters, int32_t *status);
/* exactly `iters` fused iterations, no host synchronisation inside, residual norms of the last iteration in
 * norms[batch][6] (may be NULL). */
int rb_iterate_fixed(rb_solver *s, double alpha, int32_t iters, double *norms);
/* The same loop in pieces, for hosts that pipeline (rb_iterate is built from these):
 *   rb_loop_begin   arms the device-side stopping test (max_iters, tol as in chock; tol < 0 never stops early) and the
 *                   residual history (hist_capacity rows, 0 = none); iteration 0 starts from the OLD iterate like
 *                   Solver.chock (solver.py:29-37).
 *   rb_loop_enqueue enqueues `count` iterations (one CUDA-graph launch each) without synchronising; iterations after
 *                   the stopping test fired are no-ops.
 *   rb_loop_poll    synchronises and reports iterations executed, the stop flag and the last residual norms [batch][6].
 *   rb_step         one end-to-end step for a host caller: x0 (host, may be NULL = unchanged) -> device, one
 *                   iteration, the six residual norms -> host, synchronised.  In the pipelined loop this is ONE upload
 *                   (the kernel projection copies x0 into x_0 of the old iterate, cache.py:79-82), one graph launch and
 *                   no download: the stopping test writes the norms to mapped host memory as well.
 *   rb_loop_end     closes the loop: current == old == newest iterate (no copy), history to the host. */
int rb_loop_begin(rb_solver *s, double alpha, int32_t max_iters, double tol, int32_t hist_capacity);
int rb_loop_enqueue(rb_solver *s, int32_t count);
int rb_loop_poll(rb_solver *s, int32_t *iters, int32_t *done, double *last_norms);
int rb_step(rb_solver *s, const double *x0, double *norms);
int rb_loop_end(rb_solver *s, double *xi_hist, double *delta_hist, int32_t *iters, int32_t *status);
/* measurement hook: one iteration as plain launches on one stream with a CUDA event after every launch; ms[12]:
 * ms[0] primal pass (pipelined loop: the in-place kernel projection), then the sweep launches in order (backward levels
 * bottom-up, top, forward levels top-down), then the dual pass (+ stopping test) as a whole; ms[8..10]: the kernels of
 * the pipelined dual pass on their own (branching nodes, chain nodes, leaves; or ms[8] alone: all nodes); unused
 * entries -1.  Advances the loop by one iteration. */
int rb_profile_iteration(rb_solver *s, float *ms);
/* test / ablation hook: 0 = the loop of round-1 v11 (primal pass, sweeps, one dual pass per iteration); 1 (default) = the
 * pipelined loop: the dual pass of iteration k also writes pbar = p+ - alpha L* d+ (solver.py:27-39 of iteration k+1), so
 * iteration k+1 only runs the kernel projection (cache.py:290-317) next to the backward sweep, and the dual pass is split
 * into branching nodes (next to the forward chain sweep), chain nodes (k_dual_chain) and leaves; the forward chain walk
 * 3 = additionally the forward chain walk is cut in two pieces and the dual pass of the first piece's nodes runs next to
 * the second piece (measured slower on cfg3: the walker is latency-bound and loses more to the co-resident dual warps
 * than the overlap gains; kept as an ablation).  4 = the risk block (d1, d2) of the chain nodes inside the chain dual pass
 * instead of a kernel of its own under the sweeps (ablation).  Needs the lane passes; ignored otherwise and under subtree sharding. */
int rb_use_pipeline(rb_solver *s, int32_t enable);
/* how the pipelined dual pass is split: nodes [0, early) run next to the forward chain sweep, [chain_first, chain_first +
 * chain_nodes) through k_dual_chain; chain_nodes = 0 when the loop is not pipelined or has no chain kernel for (nx, nu) */
int rb_pipeline_info(const rb_solver *s, int32_t *early_nodes, int32_t *chain_first, int32_t *chain_nodes);
int rb_use_graphs(rb_solver *s, int32_t enable); /* 1 (default): one CUDA graph per iteration; 0: plain launches */
/* 1 (default): every pipelined iteration starts with an L2 prefetch of the read-only operator tables on the side stream
 * (k_prefetch_ranges); 0: off (ablation). */
int rb_use_table_prefetch(rb_solver *s, int32_t enable);
/* 1: in the pipelined single-GPU loop (batch 1) the stopping test of solver.py:137-161 is run by the last CTA of the iteration's
 * dual-pass kernels to finish (every CTA counts itself in after folding its residual maxima) instead of by a k_check launch at the
 * end of the iteration.  Same decision, same iteration count, same histories.  Measured ablation: the counted arrival keeps every
 * CTA of the bandwidth-bound dual pass resident for one more atomic round trip, which costs more than the launch it saves
 * (cfg3: 9 138 vs 9 525 it/s in one call); 0 (default) = the k_check launch. */
int rb_use_fused_check(rb_solver *s, int32_t enable);
/* 1: in the pipelined loop the backward chain walker, the fused tree kernel and the forward chain walker are chained by
 * programmatic dependent launch -- each starts while the one before it still runs, stages its tables and waits for the data
 * itself (csrc/chain_mma.cu "launch overlap").  Measured ablation, results identical, not faster on cfg3 (9 586 vs 9 750 it/s):
 * 0 (default) = plain stream-ordered launches. */
int rb_use_launch_overlap(rb_solver *s, int32_t enable);
/* batch >= 64 instances of one tree (instance-parallel mode): 1 (default) = the fused loop runs in the
 * batch-innermost "panel" layout (csrc/batch.cu: the 32 lanes of a warp are 32 instances; iterates are converted at
 * rb_loop_begin / rb_loop_end); 0 = the instance-major kernels of batch == 1, one grid row per instance (ablation). */
int rb_use_batch_panels(rb_solver *s, int32_t enable);
/* subtree sharding: rank 0 calls rb_shard_unique_id and distributes the 128 bytes (e.g. torch.distributed broadcast),
 * every rank then calls rb_shard_init (collective: ncclCommInitRank).  Afterwards rb_iterate / rb_loop_* run the
 * sharded loop: per iteration ONE all-gather of the cut-stage q_j, d2_j and the residual maxima; the iterates of a
 * rank are valid on its own nodes and on the replicated top of the tree. */
int rb_shard_unique_id(char *id128);
int rb_shard_init(rb_solver *s, const char *id128);
/* Device-initiated exchange over NVLink peer memory (optional, after rb_shard_init): every rank calls rb_shard_p2p_export
 * (128 bytes: two CUDA IPC handles), the host distributes them, every rank calls rb_shard_p2p_open with all world * 128 bytes.
 * Afterwards the cut-stage messages are written straight into the peers' buffers by a kernel and awaited by a kernel, the
 * whole sharded iteration is ONE CUDA graph, and the pipelined loop (rb_use_pipeline) runs under sharding as well.  Without
 * it the exchange is a host-launched ncclAllGather between plain launches of the unpipelined loop. */
int rb_shard_p2p_export(rb_solver *s, char *handles128);
int rb_shard_p2p_open(rb_solver *s, const char *all_handles);
/* the cut the device chose for the DP sweeps and for subtree sharding -- the ONE source of truth for it (the host
 * never re-derives the rule): cut_stage = first stage below the replicated top, cut_first = its first node, num_cut =
 * its width (rank r of W owns the cut nodes [r*num_cut/W, (r+1)*num_cut/W) and their descendants), chain_stage = the
 * stage from which the chain walkers take over (-1: none).  Any pointer may be NULL.  Valid after rb_create. */
int rb_shard_info(const rb_solver *s, int32_t *cut_stage, int32_t *cut_first, int32_t *num_cut, int32_t *chain_stage);
/* test hook: 0 = never use the lanes-per-node passes (lane.cu), always the warp-per-node tile kernels */
int rb_use_lane_kernels(rb_solver *s, int32_t enable);
/* test hook: 0 = walk chains with one warp per chain (sweeps.cu) instead of eight chains per tile on the FP64 tensor
 * cores (chain_mma.cu); 1 (default) = tensor cores: one warp per tile, except wide rows (nx = 64, nu = 32), which are walked
 * by four warps per tile, one tile per CTA; 2 = additionally four warps per tile for nx = 20, nu = 10 (a measured ablation:
 * not faster, see profiles/r2_kernel_evolution.md); 3 = one warp per tile everywhere (ablation for the wide rows).  All
 * implement cache.py:259-288. */
int rb_use_mma_sweeps(rb_solver *s, int32_t enable);
/* test hook for the branching levels and the top of the tree: 0 = global-memory stage kernels (sweeps.cu); 1 = the
 * shared-memory-resident subtree kernels (tree_sweeps.cu), one launch per level; 2 (default) = additionally the first
 * level and the top fused into one cooperative launch when all its CTAs can be co-resident */
int rb_use_tree_kernels(rb_solver *s, int32_t mode);
/* test hook: 1 = use the general dense-matrix cost path even if sqrtQ, sqrtR, sqrtQf are all diagonal */
int rb_force_dense_costs(rb_solver *s, int32_t enable);
int rb_launch_count(const rb_solver *s, int64_t *kernels_launched); /* kernels launched by this handle so far */

/* -- stand-alone projections (cones.py:30-132, rectangle.py:29-59): host vector in, host vector out ----------------
 * cone: 0 Real, 1 Zero, 2 NonnegativeOrthant, 3 SecondOrderCone (last entry is t). */
int rb_cone_project(int32_t cone, int32_t dim, const double *in, double *out);
int rb_box_project(int32_t dim, const double *in, const double *lo, const double *hi, double *out);

#ifdef __cplusplus
}
#endif
#endif /* RAOCP_B200_H */
