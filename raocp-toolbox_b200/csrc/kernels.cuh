// kernels.cuh -- declarations of all raocp_b200 kernels (definitions in ops.cu, fused.cu, offline.cu).
#pragma once
#include "common.cuh"

namespace rb {

// ---- ops.cu: one reference method per kernel ------------------------------------------------------------------
__global__ void k_lt_axpby(const __grid_constant__ Params P, const double *__restrict__ dual,
                           const double *__restrict__ base, double *__restrict__ out, double beta, double gamma);
__global__ void k_l_axpby(const __grid_constant__ Params P, const double *__restrict__ p1, const double *__restrict__ p2,
                          double c1, double c2, const double *__restrict__ base, double *__restrict__ out, double beta,
                          double gamma);
__global__ void k_kernel_proj(const __grid_constant__ Params P, double *__restrict__ prim);
__global__ void k_s0_shift(const __grid_constant__ Params P, double *__restrict__ prim, double alpha);
__global__ void k_prox_g(const __grid_constant__ Params P, double *__restrict__ dual, double alpha, int mode,
                         int *__restrict__ status);
__global__ void k_axpby(double *__restrict__ out, double a, const double *__restrict__ x, double b,
                        const double *__restrict__ y, long long count);
__global__ void k_div_add(double *__restrict__ out, const double *__restrict__ x, double a, double b,
                          const double *__restrict__ y, long long count);
__global__ void k_absmax(const double *__restrict__ x, long long stride, double *__restrict__ slot, int slot_stride,
                         int *__restrict__ status);
// L2 prefetch of read-only byte ranges (the operator tables), one 128-byte line per thread and step
struct PrefetchRange {
    const char *ptr;
    long long bytes;
};
struct Ctrl;
__global__ void k_prefetch_ranges(const Ctrl *__restrict__ ctrl, const PrefetchRange *__restrict__ ranges, int count);
__global__ void k_cone(int cone, int dim, const double *__restrict__ in, double *__restrict__ out);
__global__ void k_box(int dim, const double *__restrict__ in, const double *__restrict__ lo, const double *__restrict__ hi,
                      double *__restrict__ out, int *__restrict__ status);

// ---- fused.cu: Solver.chock's loop body --------------------------------------------------------------------------
// control block shared by the kernels of the fused loop
struct Ctrl {
    int done;        // set by k_check when the stopping test of solver.py:156-161 fires; later launches are no-ops
    int iters;       // iterations executed so far
    int status;      // bit 0: NaN met in a rectangle projection; bit 1: non-finite residual
    int max_iters;   // stop when the iteration index reaches this (solver.py:156-157)
    double tol;      // stop when max(xi0, xi1, xi2) <= tol for every instance (solver.py:158-159)
    double alpha;    // step size alpha_1 = alpha_2 (solver.py:116-118)
    double *hist;    // [hist_capacity][batch][6] residual history or null
    int hist_capacity;
    int pending;     // set by the dual pass: `slots` hold the residual maxima of an iteration k_check has not tested yet
    int mirror;      // k_check also writes the norms to mapped host memory (switched on by the first rb_step of a loop: the
    int pad;         // posted write to system memory costs the kernel ~1 us, which a loop that never reads them need not pay)
    int chk_fail, chk_nan, chk_arrived;   // scratch of k_check_wide (many CTAs): some instance above tol / NaN seen / CTAs done
    // stopping test without a launch of its own (batch 1, pipelined loop): every CTA of the dual-pass kernels counts itself in after
    // its maxima are folded into `slots`; the one that completes arr_expected runs k_check's body (lane.cu iteration_arrive)
    int arr_count, arr_expected;
    double *last, *host_last;             // k_check's outputs, for that CTA
};
// tiles of consecutive nodes for the node-parallel passes (fused.cu): tiles[t] = (first node, one past the last);
// a tile is either all nonleaf or all leaf nodes
struct TilePlan {
    const int2 *tiles;
    int num_tiles;
    int rowlen;      // doubles per warp-private scratch row (max(nx, nu) rounded up to even)
};
// host launchers of the tiled kernels (templates on DIAG = all cost square roots diagonal -> entrywise products
// instead of matvecs; defined in fused.cu next to the kernels)
cudaError_t tile_kernels_set_smem(size_t primal_bytes, size_t dual_bytes);
void launch_primal_tile(bool diag, dim3 grid, size_t smem, cudaStream_t st, const Params &P, const Ctrl *ctrl,
                        const TilePlan &plan, const double *p_old, const double *d_old, double *p_new);
void launch_dual_tile(bool diag, dim3 grid, size_t smem, cudaStream_t st, const Params &P, Ctrl *ctrl, const TilePlan &plan,
                      const double *p_old, const double *p_new, const double *d_old, double *d_new, double *slots);
constexpr int kDualRowsHost = 15;

// Subtree sharding: the nodes a rank owns in a run of equally wide stages (the chain part of the tree) are the columns
// [lo, lo + w) of every stage.  at(i) maps the i-th owned node to its offset from the first node of the run; w = 0: identity.
struct OwnMap {
    int w, lo, full;
    __host__ __device__ int at(int i) const { return w > 0 ? (i / w) * full + lo + i % w : i; }
};

// ---- lane.cu: eight lanes per node (diagonal cost square roots, even nx / nu, <= kLaneMaxChildren children per node) ---------------
constexpr int kLaneThreads = 256;
constexpr int kLaneOctet = 8;    // lanes that share one node
constexpr int kLaneMaxChildren = 8;
// node_list / count: the nodes to process (lane group t takes node_list[t]); null = all nodes 0..n-1.  The kernels are
// templates on the number of lanes per node (4 or 8); nodes_batch.y = batch
void launch_primal_lane(dim3 nodes_batch, cudaStream_t st, const Params &P, const Ctrl *ctrl, const double *p_old,
                        const double *d_old, double *p_new, const int *node_list, int count);
// first: node of lane group 0 when node_list is null.  pbar (may be null) = the p_old buffer: the pass also leaves
// there the half step pbar = p+ - alpha L* d+ of the NEXT iteration (then the next iteration needs no primal pass, only
// the kernel projection k_kproj_node in place)
// (the three dual-pass launchers return the number of CTAs launched; arrive: those CTAs count themselves into Ctrl::arr_count)
int launch_dual_lane(dim3 nodes_batch, cudaStream_t st, const Params &P, Ctrl *ctrl, const double *p_old,
                     const double *p_new, const double *d_old, double *d_new, double *slots, const int *node_list,
                     int first, int count, double *pbar, bool narrow = false, bool arrive = false);
// narrow: keep the default lanes per node for a small launch (fewest CTAs: it runs next to a latency-critical kernel)
// the same pass specialised for a run of nonleaf nodes with ONE child each (the chain part of the tree)
bool dual_chain_supported(int nx, int nu);
// recs[i]: packed topology of node first + i = (child, cost-table row of the child, offset of y_i, rectangle row);
// stride > 0: the child of node i is i + stride for the whole run; yo0 = offset of y_first
int launch_dual_chain(int batch, cudaStream_t st, const Params &P, Ctrl *ctrl, const double *p_old, const double *p_new,
                      const double *d_old, double *d_new, double *slots, const int4 *recs, int first, int count,
                      int stride, int yo0, double *pbar, int with_risk = 1, OwnMap own = OwnMap{0, 0, 0}, bool arrive = false);
// the risk block (d1, d2, ybar, sbar, y / s residual rows) of the same run of nodes on its own: needs y, s only, runs
// under the sweeps; the chain pass is then launched with with_risk = 0
int launch_dual_risk_chain(int batch, cudaStream_t st, const Params &P, Ctrl *ctrl, const double *p_old, const double *p_new,
                           const double *d_old, double *d_new, double *slots, int first, int count, int stride, int yo0,
                           double *pbar, OwnMap own = OwnMap{0, 0, 0}, bool arrive = false);
// x0 / p_old (may be null): also copy the initial state x0 [batch][nx] into x_0 of the OLD iterate (what
// cache_initial_state does, cache.py:79-82) -- rb_step then uploads x0 once
void launch_kproj(int batch, cudaStream_t st, const Params &P, const Ctrl *ctrl, double *prim, const double *x0 = nullptr,
                  double *p_old = nullptr, const int *node_list = nullptr, int count = 0);
__global__ void k_check(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl, double *__restrict__ slots,
                        double *__restrict__ last, double *__restrict__ host_last);
// the same stopping test for large batches (one thread per slot, the last CTA to finish closes the iteration); launch_check picks
void launch_check(cudaStream_t st, const Params &P, Ctrl *ctrl, double *slots, double *last, double *host_last);

// ---- sweeps.cu: the DP sweeps in a handful of launches ---------------------------------------------------------------
struct SweepLevel {
    int t_lo, depth;       // the level's subtrees are rooted at stage t_lo and cover stages [t_lo, t_lo + depth)
    int num_sub;           // nodes at stage t_lo
    int warps_per_sub;     // warps cooperating on one subtree (block barriers per stage if > 1)
    int subs_per_cta;      // subtrees packed into one CTA
    int chain;             // 1: every subtree is a chain down to the leaves (<= 64 nodes): prefetching chain walker
    int stage_cap;         // max over subtrees / stages of max(#nodes, #children) of one stage (group-shared buffer rows)
    const int *lo, *hi;    // [num_sub][depth] node range of the subtree at stage t_lo + d
    const int *tiles;      // chain levels: [num_tiles][8] chains with identical dynamics / class sequences (-1 = padding)
    int num_tiles;         //               0 = no tiling (chain_mma.cu is not used)
    const int *tile_meta;  // [num_tiles][depth * 10 + 8]: node ids [depth][8], dynamics row [depth], class [depth], valid [8]
};
struct SweepPlan {
    SweepLevel lv[2];
    int num_levels;        // 0, 1 or 2
    int t_top;             // stages [0, t_top) belong to the top kernel
    int top_cap;           // widest stage (nodes or children) handled by the top kernel
    const int *stage_off;  // [num_stages + 1]
};
// host launchers (templates on <NX, NU> are instantiated and dispatched in sweeps.cu)
cudaError_t sweep_kernels_set_smem(int bytes);
void launch_sweep_sub_bwd(dim3 grid, int threads, size_t smem, cudaStream_t st, const Params &P, const Ctrl *ctrl,
                          const SweepLevel &lv, const double *prim, double *q, double *r);
void launch_sweep_sub_fwd(dim3 grid, int threads, size_t smem, cudaStream_t st, const Params &P, const Ctrl *ctrl,
                          const SweepLevel &lv, double *prim, const double *r);
void launch_sweep_top(int grid, int threads, size_t smem, cudaStream_t st, const Params &P, const Ctrl *ctrl,
                      const SweepPlan &plan, double *prim, double *q, double *r, const double *x0);

// ---- tree_sweeps.cu: branching subtrees with everything resident in shared memory ---------------------------------------------
struct TreeLevel {
    int depth, num_sub;      // stages per subtree, subtrees (one CTA each)
    int desc_stride;         // ints per subtree descriptor
    int max_nodes, max_ext;  // largest subtree (nodes) / largest set of children below a subtree
    int max_row;             // widest stage or child set (rows of the ping-pong / contribution buffers)
    int num_dyn;
    int resident;            // the dynamics tables and the per-node K, [K R~^-1] are staged in shared memory (they fit)
    int warps;
    const int *desc;         // [num_sub][desc_stride]; null = level not served by tree_sweeps.cu
};
size_t tree_smem_bytes(const TreeLevel &lv, int nx, int nu, int warps, bool top);
cudaError_t tree_kernels_set_smem(int bytes);
void launch_tree_bwd(dim3 grid, int threads, size_t smem, cudaStream_t st, const Params &P, const Ctrl *ctrl, const TreeLevel &lv,
                     const double *prim, double *q, double *r);
void launch_tree_fwd(dim3 grid, int threads, size_t smem, cudaStream_t st, const Params &P, const Ctrl *ctrl, const TreeLevel &lv,
                     double *prim, const double *r);
cudaError_t launch_tree_fused(int batch, int threads, size_t smem, cudaStream_t st, const Params &P, const Ctrl *ctrl,
                              const TreeLevel &lvs, const TreeLevel &lvt, double *prim, double *q, double *r, const double *x0,
                              int *sync, int *walk_count = nullptr, int walk_tiles = 0, int *tree_done = nullptr);
bool tree_fused_fits(int nx, int nu, bool resident, int threads, size_t smem, int ctas);
void launch_tree_top(int grid, int threads, size_t smem, cudaStream_t st, const Params &P, const Ctrl *ctrl, const TreeLevel &lv,
                     double *prim, double *q, double *r, const double *x0);

// ---- chain_mma.cu: chain levels on the FP64 tensor cores, 8 chains per warp ---------------------------------------------------
bool chain_mma_supported(int nx, int nu);
bool chain_mma_w4(int nx, int nu);   // four warps per tile are instantiated for these sizes (k_chain_mma_*_w4)
bool chain_mma_wide(int nx, int nu); // wide rows: four warps per tile, one tile per CTA (k_chain_mma_*_wide), the default there
size_t chain_mma_smem_bytes(int nx, int nu, int depth, bool backward, bool w4 = false);
cudaError_t chain_mma_set_smem(int bytes);
void chain_mma_frag_counts(int nx, int nu, int *f_ab, int *f_abt, int *f_k, int *f_kr);
void launch_chain_mma_frags(cudaStream_t st, const Tabs &M, int nx, int nu, int num_dyn, int num_cls, bool dynamics,
                            bool classes);
// w4: walk a tile with four warps (one output block each) -- only where chain_mma_w4(nx, nu)
// walk_count / tree_done (null = off): the launch-overlap protocol with the fused tree kernel (chain_mma.cu); one warp per tile only
void launch_chain_mma_bwd(cudaStream_t st, const Params &P, const Ctrl *ctrl, const SweepLevel &lv, const double *prim,
                          double *q, double *r, bool w4 = false, int *walk_count = nullptr, int *tree_done = nullptr);
// d_begin, d_end: the steps of the walk to run (depth below the chain heads); d_end < 0 = to the leaves
void launch_chain_mma_fwd(cudaStream_t st, const Params &P, const Ctrl *ctrl, const SweepLevel &lv, double *prim,
                          const double *r, int d_begin = 0, int d_end = -1, bool w4 = false, const int *tree_done = nullptr,
                          int tree_ctas = 0);

// ---- batch.cu: many instances of one small tree, batch-innermost ("panel") layout: the lanes of a warp are 32 instances ------
bool batch_panel_supported(int nx, int nu);
size_t batch_panel_doubles(long long stride, int batch);   // doubles of a panel buffer with `stride` elements per instance
void launch_to_panels(cudaStream_t st, const double *src, double *dst, long long stride, int batch);
void launch_from_panels(cudaStream_t st, const double *src, double *dst, long long stride, int batch);
void launch_bp_primal(cudaStream_t st, const Params &P, const Ctrl *ctrl, const double *p_old, const double *d_old, double *p_out);
void launch_bp_kproj(cudaStream_t st, const Params &P, const Ctrl *ctrl, double *prim, const double *x0, double *p_old);
void launch_bp_bwd(cudaStream_t st, const Params &P, const Ctrl *ctrl, const double *prim, double *q, double *r, int first, int count);
void launch_bp_fwd(cudaStream_t st, const Params &P, const Ctrl *ctrl, double *prim, const double *r, const double *x0, int first,
                   int count);
// the three dual kernels (x / u block and risk block of the nonleaf nodes, leaves); pbar = the p_old buffer; c2: [m][nx + nu]
// diagonal of L* L on the x / u rows (launch_bp_c2, once)
void launch_bp_dual(cudaStream_t st, const Params &P, Ctrl *ctrl, const double *p_old, const double *p_new, const double *d_old,
                    double *d_new, double *slots, double *pbar, const double *c2, cudaEvent_t *evs = nullptr);   // evs: 2 events (after the first two kernels)
void launch_bp_c2(cudaStream_t st, const Params &P, double *c2);
// the stages [0, t_top) backward and forward in one launch (one CTA per panel, a warp per parent: stages of <= 32 parents)
void launch_bp_top(cudaStream_t st, const Params &P, const Ctrl *ctrl, double *prim, double *q, double *r, const double *x0,
                   const int *stage_off, int t_top);
int batch_panel_max_children();

// ---- shard.cu: one tree sharded by subtree over the GPUs of a box ----------------------------------------------------------
struct ShardPlan {
    int rank, world;
    int cut_first;          // first node of the cut stage (= first sweep level)
    int cut_lo, cut_hi;     // this rank's block of cut nodes, as offsets into the cut stage
    int cap;                // largest block of any rank (message rows)
    const int *cut_bounds;  // [world + 1] block boundaries of all ranks (device)
};
struct NcclId {
    char internal[128];
};
// peer-mapped exchange buffers (CUDA IPC; one box: <= 8 ranks)
constexpr int kMaxPeers = 8;
struct PeerXchg {
    double *recv[kMaxPeers];               // receive buffer of every rank: [2 parities][world][cap * (nx + 1) + 6]
    unsigned long long *flag[kMaxPeers];   // flags of every rank: [2 parities][world]
    unsigned long long *seq;               // local: [0] number of the next exchange (starts at 1), [1] CTAs of k_shard_xchg that arrived
    // packet area of every rank, behind its receive buffer: [2 parities][world][cap * (nx + 1) + 6] 16-byte packets
    // {low word, flag, high word, flag} (flag = low 32 bits of the sequence number): data and flag travel in one store, so the
    // exchange needs no fence and no separate flag -- one NVLink flight (the "LL" protocol of NCCL, on doubles)
    uint4 *ll[kMaxPeers];
};
// the exchange folded into the top-of-the-tree kernel (k_tree_top<.., SHARD = true>, tree_sweeps.cu): the one CTA that walks the
// replicated top stores this rank's q_j, aux_j and residual maxima into every peer's packet area, polls the peers' packets straight
// into its shared-memory row buffer and runs the stopping test, while its tables and rows are still being staged
struct ShardHand {
    ShardPlan sp;
    PeerXchg px;
    double *aux;                       // per-node scalar that travels with q_j (sbar_j of pbar)
    double *slots, *last, *host_last;  // residual maxima of the previous iteration / k_check's outputs
    int check;                         // an iteration is waiting for its stopping test
};
void launch_tree_top_sharded(int threads, size_t smem, cudaStream_t st, const Params &P, Ctrl *ctrl, const TreeLevel &lv, double *prim,
                             double *q, double *r, const double *x0, const ShardHand &sh);
void launch_shard_xchg(cudaStream_t st, const Params &P, Ctrl *ctrl, const ShardPlan &sp, double *q, double *aux, double *slots,
                       const PeerXchg &px, double *last, double *host_last, bool check);
void launch_shard_push(cudaStream_t st, const Params &P, const Ctrl *ctrl, const ShardPlan &sp, const double *q, const double *aux,
                       const double *slots, const PeerXchg &px);
void launch_shard_pull(cudaStream_t st, const Params &P, Ctrl *ctrl, const ShardPlan &sp, double *q, double *aux, double *slots,
                       const PeerXchg &px);
__global__ void k_shard_pack(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl, ShardPlan sp,
                             const double *__restrict__ q, const double *__restrict__ aux,
                             const double *__restrict__ slots, double *__restrict__ send);
__global__ void k_shard_unpack(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl, ShardPlan sp,
                               const double *__restrict__ recv, double *__restrict__ q, double *__restrict__ aux,
                               double *__restrict__ slots);
const char *nccl_load();   // nullptr on success, else the reason
int nccl_unique_id(NcclId *id);
int nccl_comm_init(void **comm, int world, const NcclId &id, int rank);
int nccl_all_gather_f64(const double *send, double *recv, size_t count, void *comm, cudaStream_t st);
void nccl_comm_destroy(void *comm);
const char *nccl_error(int code);

// ---- offline.cu --------------------------------------------------------------------------------------------------
struct ClassView {
    const int *child_ptr;   // [num_cls+1]
    const int *child_dyn;   // dyn table row of every child of the class representative
    const int *child_cls;   // class of that child, -1 = leaf (P = I)
    const int *level_list;  // classes of the level being processed
};
__global__ void k_offline_level(const __grid_constant__ Params P, ClassView cv, int level_begin, int level_count,
                                double *__restrict__ Ptab, double *__restrict__ Ktab, double *__restrict__ KRcatT,
                                int *__restrict__ status);
size_t offline_smem_bytes(int nx, int nu);

// lambda_max(L* L) pieces
__global__ void k_gram_eig(const __grid_constant__ Params P, const int *__restrict__ grp_ptr, const int *__restrict__ grp_idx,
                           int kind, int num_groups, double *__restrict__ work, double *__restrict__ out_max,
                           int *__restrict__ status);
__global__ void k_risk_block_eig(const __grid_constant__ Params P, double *__restrict__ out_max);

}  // namespace rb
