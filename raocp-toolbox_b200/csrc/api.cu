// api.cu -- host side of the C-ABI (include/raocp_b200.h): handle, device memory, launches, CUDA graphs.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../include/raocp_b200.h"
#include "kernels.cuh"

using namespace rb;

namespace {

thread_local std::string g_create_error;

}  // namespace

struct rb_solver {
    Params P{};
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    std::string err;
    std::vector<void *> allocs;
    // host copies of the topology
    std::vector<int> stage_off, parent, child_first, child_count, dyn_idx, cost_idx, cls;
    int num_cls = 0, num_dyn = 0, num_cost = 0, num_leafcost = 0;
    int64_t np = 0, nd = 0;  // compact sizes
    // compact segment tables: offset in compact vector, offset in padded buffer, length
    struct Seg {
        int64_t compact, padded, len;
    };
    std::vector<Seg> pseg, dseg;
    // iterates: two device copies.  cur_i / old_i say which copy plays "current" (Cache.__primal/__dual) and which
    // "old" (Cache.__old_primal/__old_dual).  After update_cache() or a fused loop both views show the same
    // iterate: instead of copying, `collapsed` marks that only buffer old_i is meaningful; the copy is made lazily by
    // materialise() when a step-by-step method needs two distinct buffers.
    double *prim[2] = {nullptr, nullptr};
    double *dual[2] = {nullptr, nullptr};
    int cur_i = 0, old_i = 1;
    bool collapsed = true;
    bool in_loop = false;
    int loop_old0 = 1;  // which buffer was "old" when the fused loop began
    double *h_pinned = nullptr;  // pinned staging for control block / residual read-back
    double *h_last = nullptr, *h_last_dev = nullptr;   // mapped pinned mirror of `last` (k_check writes it; rb_step reads it)
    bool mirror_on = false;                            // Ctrl::mirror has been switched on in the running loop
    double *q = nullptr, *r = nullptr, *x0 = nullptr;
    Ctrl *ctrl = nullptr;
    double *slots = nullptr, *last = nullptr, *hist = nullptr;
    int hist_capacity = 0;
    int *status = nullptr;
    bool have_x0 = false, have_offline = false;
    // offline
    std::vector<int> cls_child_ptr, cls_child_dyn, cls_child_cls, level_list, level_ptr;
    int *d_cls_child_ptr = nullptr, *d_cls_child_dyn = nullptr, *d_cls_child_cls = nullptr, *d_level_list = nullptr;
    double *Ptab = nullptr, *Ktab = nullptr, *KRcatT = nullptr;
    bool diag_costs = false;
    int max_children = 1;
    bool allow_lane = true;   // one-thread-per-node passes (lane.cu) when the problem qualifies
    // residual temporaries (allocated on first use)
    double *tp[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double *td[4] = {nullptr, nullptr, nullptr, nullptr};
    double *staging_p = nullptr, *staging_d = nullptr;  // device staging for host-vector operator calls
    // node tiles of the fused passes (fused.cu)
    TilePlan tiles{};
    size_t primal_smem = 0, dual_smem = 0;
    // DP sweeps in three launches (sweeps.cu)
    SweepPlan plan{};
    int top_warps = 16;
    size_t sweep_smem_max = 0;
    // subtree sharding over GPUs (shard.cu)
    ShardPlan shard{};
    bool sharded = false;
    void *nccl_comm = nullptr;
    int *owned_nodes = nullptr, *top_nodes = nullptr, *dual_nodes = nullptr;   // device node lists
    int n_owned = 0, n_top = 0;
    SweepLevel shard_lv[2]{};
    std::vector<int> chain_lo[2];   // host copy of lv.lo of chain levels (tile building)
    bool allow_mma = true;
    bool mma_wide = true;    // wide rows: four warps per tile, one tile per CTA (rb_use_mma_sweeps(3): the one-warp BIG kernels)
    bool mma_w4 = false;     // rb_use_mma_sweeps(2): four warps per chain tile where instantiated (measured ablation, chain_mma.cu)
    int tree_mode = 2;       // 0: sweeps.cu stage kernels; 1: tree_sweeps.cu, one launch per level; 2: + level 0 fused with the top
    bool fuse_ok = false;    // the fused launch is possible (co-residency, same residency mode, one smem footprint)
    size_t fuse_smem = 0;
    int fuse_threads = 0;
    int *tree_sync = nullptr;
    // read-only operator tables (dynamics, K, R~^-1, tensor-core fragments): prefetched into L2 at the head of every pipelined
    // iteration by k_prefetch_ranges (ops.cu) -- a few MB that every sweep kernel otherwise fetches through dependent misses
    std::vector<PrefetchRange> h_tables;
    PrefetchRange *d_tables = nullptr;
    bool allow_table_prefetch = true;
    int *overlap_sync = nullptr;   // [2 * batch]: walk_count, tree_done of the launch-overlap protocol (chain_mma.cu)
    bool allow_overlap = false;   // rb_use_launch_overlap(1): measured ablation (not faster on cfg3: 9 586 vs 9 750 it/s cold)
    TreeLevel tree_top{}, tree_lv[2]{}, shard_tree_lv[2]{};
    size_t tree_smem[3]{};   // dynamic shared memory of the top / level launches   // chain levels on the FP64 tensor cores when the level carries tiles (rb_use_mma_sweeps)
    double *xchg_send = nullptr, *xchg_recv = nullptr;
    size_t xchg_count = 0;
    // device-initiated exchange over peer memory (shard.cu k_shard_push / _pull) and the pipelined sharded loop it enables
    bool p2p = false;
    PeerXchg px{};
    double *p2p_recv = nullptr;
    unsigned long long *p2p_flag = nullptr;
    void *p2p_opened[2 * kMaxPeers] = {};
    int *own_nonleaf = nullptr, *top_nonleaf = nullptr, *own_lane = nullptr;   // node lists of the pipelined sharded loop
    int n_own_nonleaf = 0, n_top_nonleaf = 0, n_own_lane = 0;
    OwnMap own_chain{0, 0, 0};      // the rank's columns of the chain stages
    int n_own_chain = 0;
    bool fused_check = false;     // rb_use_fused_check(1): stopping test by the last CTA of the dual passes (measured ablation: slower)
    bool loop_armed = false;      // the running loop's control block carries arr_expected
    int arr_expected = 0;         // CTAs of the dual-pass launches of one pipelined iteration (0: not counted yet)
    int arr_counted = 0;          // ... as counted by the latest unarmed enqueue / capture
    bool xchg_in_top = true;     // ... and inside the top-of-the-tree kernel (k_tree_top<.., SHARD>) in the pipelined loop
    bool xchg_fused = true;       // peer-memory exchange as ONE launch (k_shard_xchg); RAOCP_SHARD_XCHG=split: push / pull / check
    bool shard_pending = false;   // an executed iteration whose residuals have not been gathered / tested yet
    // pipelined loop (lane passes only): the dual pass of iteration k also writes pbar of iteration k+1 into the old
    // primal buffer, so iteration k+1 has no primal pass -- only the kernel projection, in place, next to the backward
    // sweep -- and the dual pass is split into the branching part (runs next to the forward chain sweep), the chain part
    // (k_dual_chain) and the leaves
    bool allow_pipe = true;
    bool pipe_fwd_split = false;
    bool risk_split = true;         // the risk block of the chain nodes' dual pass as a kernel of its own under the sweeps    // cut the forward chain walk in two and overlap the second piece with the dual pass of the first
    bool pbar_ready = false;        // prim[cur_i] holds pbar of the next iteration
    int chain_first = 0;            // nodes [chain_first, m) are nonleaf nodes with exactly one child
    int4 *chain_recs = nullptr;     // their packed topology records (launch_dual_chain)
    int chain_stride = -1;          // > 0: child_first[i] = i + chain_stride on the whole run
    int chain_yo0 = 0;              // offset of y of node chain_first
    cudaStream_t side[2] = {nullptr, nullptr};
    cudaEvent_t pev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // batch-innermost ("panel") path of the fused loop (batch.cu): buffers allocated on first use; panel_live = the loop's
    // iterates are in the panel buffers (rb_loop_begin converted them, rb_loop_end converts the newest one back)
    bool allow_panel = true;
    bool panel_live = false;
    double *pprim[2] = {nullptr, nullptr}, *pdual[2] = {nullptr, nullptr}, *pq = nullptr, *pr = nullptr, *pc2 = nullptr;
    // fused loop
    cudaGraphExec_t graph[2] = {nullptr, nullptr};  // graph[src]: one iteration reading buffer src, writing 1-src
    bool use_graphs = true;
    int64_t launches = 0;
    int kernels_per_iter = 0;
};

namespace {

int fail(rb_solver *s, int code, const std::string &msg) {
    if (s) s->err = msg;
    else g_create_error = msg;
    return code;
}

#define RB_CUDA(s, call)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail((s), e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? RB_ERR_NO_DEVICE : RB_ERR_CUDA, \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                          \
    } while (0)

template <typename T>
int upload(rb_solver *s, const T *host, size_t count, T **dev) {
    *dev = nullptr;
    if (count == 0) count = 1;
    RB_CUDA(s, cudaMalloc((void **)dev, count * sizeof(T)));
    s->allocs.push_back(*dev);
    if (host) RB_CUDA(s, cudaMemcpy(*dev, host, count * sizeof(T), cudaMemcpyHostToDevice));
    else RB_CUDA(s, cudaMemset(*dev, 0, count * sizeof(T)));
    return RB_OK;
}

template <typename T>
int dev_zero(rb_solver *s, size_t count, T **dev) {
    return upload<T>(s, nullptr, count, dev);
}

std::vector<double> transpose_tab(const double *tab, int count, int rows, int cols) {
    std::vector<double> out((size_t)count * rows * cols);
    for (int t = 0; t < count; ++t)
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c)
                out[(size_t)t * rows * cols + (size_t)c * rows + r] = tab[(size_t)t * rows * cols + (size_t)r * cols + c];
    return out;
}

int is_diag(const double *tab, int count, int dim) {
    for (int t = 0; t < count; ++t)
        for (int r = 0; r < dim; ++r)
            for (int c = 0; c < dim; ++c)
                if (r != c && tab[(size_t)t * dim * dim + (size_t)r * dim + c] != 0.0) return 0;
    return 1;
}

inline int64_t round16(int64_t v) { return (v + 15) / 16 * 16; }

dim3 node_grid(const rb_solver *s, int nodes) {
    return dim3((unsigned)((nodes + kWarpsPerBlock - 1) / kWarpsPerBlock), (unsigned)s->P.L.batch);
}

int check_status(rb_solver *s) {
    int st = 0;
    RB_CUDA(s, cudaMemcpyAsync(&st, s->status, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    RB_CUDA(s, cudaStreamSynchronize(s->stream));
    if (st) {
        RB_CUDA(s, cudaMemsetAsync(s->status, 0, sizeof(int), s->stream));
        if (st & 1) return fail(s, RB_ERR_NUMERIC, "Rectangle constraint - 'nan' value cannot be constrained");
        if (st & 4) return fail(s, RB_ERR_NUMERIC, "offline factorisation: R~ is not positive definite");
        return fail(s, RB_ERR_NUMERIC, "non-finite value in the residuals");
    }
    return RB_OK;
}

// compact host vector <-> padded device buffer, all instances (2-D copies: one per segment)
int copy_segments(rb_solver *s, const std::vector<rb_solver::Seg> &segs, int64_t compact_stride, int64_t padded_stride,
                  double *dev, const double *host_in, double *host_out) {
    for (const auto &g : segs) {
        if (g.len == 0) continue;
        if (host_in)
            RB_CUDA(s, cudaMemcpy2DAsync(dev + g.padded, padded_stride * sizeof(double), host_in + g.compact,
                                         compact_stride * sizeof(double), g.len * sizeof(double), s->P.L.batch,
                                         cudaMemcpyHostToDevice, s->stream));
        else
            RB_CUDA(s, cudaMemcpy2DAsync(host_out + g.compact, compact_stride * sizeof(double), dev + g.padded,
                                         padded_stride * sizeof(double), g.len * sizeof(double), s->P.L.batch,
                                         cudaMemcpyDeviceToHost, s->stream));
    }
    RB_CUDA(s, cudaStreamSynchronize(s->stream));
    return RB_OK;
}

int launch_ok(rb_solver *s, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(s, RB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return RB_OK;
}

#define RB_LAUNCHED(s, what)                       \
    do {                                           \
        ++(s)->launches;                           \
        int rc_ = launch_ok((s), (what));          \
        if (rc_ != RB_OK) return rc_;              \
    } while (0)

// make "current" a distinct copy of "old" (see rb_solver::collapsed)
int materialise(rb_solver *s) {
    if (!s->collapsed) return RB_OK;
    const Layout &L = s->P.L;
    RB_CUDA(s, cudaMemcpyAsync(s->prim[s->cur_i], s->prim[s->old_i], (size_t)L.batch * L.np_pad * sizeof(double),
                               cudaMemcpyDeviceToDevice, s->stream));
    RB_CUDA(s, cudaMemcpyAsync(s->dual[s->cur_i], s->dual[s->old_i], (size_t)L.batch * L.nd_pad * sizeof(double),
                               cudaMemcpyDeviceToDevice, s->stream));
    s->collapsed = false;
    return RB_OK;
}
inline double *view_p(rb_solver *s, int which) { return s->prim[(which == 0 && !s->collapsed) ? s->cur_i : s->old_i]; }
inline double *view_d(rb_solver *s, int which) { return s->dual[(which == 0 && !s->collapsed) ? s->cur_i : s->old_i]; }
#define RB_MATERIALISE(s)               \
    do {                                \
        int rc_ = materialise(s);       \
        if (rc_ != RB_OK) return rc_;   \
    } while (0)

bool use_lane(const rb_solver *s) {
    return s->allow_lane && s->diag_costs && s->max_children <= kLaneMaxChildren && s->P.L.nx % 2 == 0 && s->P.L.nu % 2 == 0;
}

void launch_primal(rb_solver *s, cudaStream_t st, int src, int dst, const int *nodes = nullptr, int count = -1) {
    const Layout &L = s->P.L;
    if (count < 0) count = L.n;
    if (use_lane(s)) {
        launch_primal_lane(dim3(1, L.batch), st, s->P, s->ctrl, s->prim[src], s->dual[src], s->prim[dst], nodes, count);
    } else {
        launch_primal_tile(s->diag_costs, dim3(s->tiles.num_tiles, L.batch), s->primal_smem, st, s->P, s->ctrl, s->tiles,
                           s->prim[src], s->dual[src], s->prim[dst]);
    }
}

// four warps per chain tile: the default where rows are wide (chain_mma_wide: nx = 64, nu = 32; rb_use_mma_sweeps(3) = one warp,
// ablation), an ablation where four tiles per SM already fill the tensor pipes (chain_mma_w4: rb_use_mma_sweeps(2))
bool mma_four_warps(const rb_solver *s) {
    const Layout &L = s->P.L;
    return (s->mma_w4 && chain_mma_w4(L.nx, L.nu)) || (s->mma_wide && chain_mma_wide(L.nx, L.nu));
}

// the captured iterations are stale (a kernel-family switch, another stream, another layout): drop them, and with them the CTA
// count the fused stopping test was armed with
void drop_graphs(rb_solver *s) {
    for (int i = 0; i < 2; ++i)
        if (s->graph[i]) {
            cudaGraphExecDestroy(s->graph[i]);
            s->graph[i] = nullptr;
        }
    s->arr_expected = 0;
    s->arr_counted = 0;
}

bool use_pipe(const rb_solver *s) { return use_lane(s) && s->allow_pipe && !s->sharded; }

// many instances of a small tree: lanes = instances (batch.cu).  Needs diagonal cost square roots; the padding instances of
// the last panel cost as much as real ones, so a batch has to fill at least a few panels to pay
bool use_panel(const rb_solver *s) {
    const Layout &L = s->P.L;
    return s->allow_panel && !s->sharded && L.batch >= 64 && s->diag_costs && batch_panel_supported(L.nx, L.nu) &&
           s->max_children <= batch_panel_max_children();
}

void launch_dual(rb_solver *s, cudaStream_t st, int src, int dst, const int *nodes = nullptr, int count = -1) {
    const Layout &L = s->P.L;
    if (count < 0) count = L.n;
    if (use_lane(s)) {
        launch_dual_lane(dim3(1, L.batch), st, s->P, s->ctrl, s->prim[src], s->prim[dst], s->dual[src], s->dual[dst], s->slots,
                         nodes, 0, count, nullptr);
    } else {
        launch_dual_tile(s->diag_costs, dim3(s->tiles.num_tiles, L.batch), s->dual_smem, st, s->P, s->ctrl, s->tiles,
                         s->prim[src], s->prim[dst], s->dual[src], s->dual[dst], s->slots);
    }
}

// kernels launched per iteration of the fused loop: primal pass, sweeps, dual pass, stopping test
struct PipeSplit {
    int early;   // nodes [0, early): final once the top of the tree is done (dual pass runs next to the forward levels)
    int cf;      // nodes [early, cf): general lane pass after the sweeps; [cf, m): chain pass; [m, n): leaves
    int split;   // > 0: the last forward level (chain walker) runs as steps [0, split) and [split, depth), and the chain
    int mid;     //      dual pass of the nodes [cf, mid) -- final after the first piece -- runs next to the second piece
};
bool sweeps_fused(const rb_solver *s) {
    const SweepPlan &pl = s->plan;
    return s->tree_mode > 1 && s->fuse_ok && !(s->allow_mma && pl.lv[0].num_tiles > 0) && !s->sharded;
}
PipeSplit pipe_split(const rb_solver *s) {
    const Layout &L = s->P.L;
    const SweepPlan &pl = s->plan;
    PipeSplit ps;
    ps.cf = dual_chain_supported(L.nx, L.nu) ? std::max(1, std::min(s->chain_first, L.m)) : L.m;
    const int v0 = sweeps_fused(s) ? 1 : 0;   // first forward level launched after the top
    ps.early = v0 < pl.num_levels ? std::min(ps.cf, s->stage_off[pl.lv[v0].t_lo]) : 0;
    static const bool no_early = getenv("RB_NO_EARLY") != nullptr;   // ablation knob
    if (no_early) ps.early = 0;
    ps.split = 0;
    ps.mid = ps.cf;
    if (s->pipe_fwd_split && pl.num_levels > v0 && ps.cf < L.m) {
        const SweepLevel &lv = pl.lv[pl.num_levels - 1];
        if (s->allow_mma && lv.num_tiles > 0 && lv.depth >= 6 && s->stage_off[lv.t_lo] >= ps.cf) {
            ps.split = lv.depth / 2;
            ps.mid = std::min(L.m, s->stage_off[lv.t_lo + ps.split]);
            if (ps.mid <= ps.cf) ps.split = 0, ps.mid = ps.cf;
        }
    }
    return ps;
}
// kernels of the dual pass of the pipelined loop
int pipe_dual_launches(const rb_solver *s) {
    const Layout &L = s->P.L;
    const PipeSplit ps = pipe_split(s);
    if (ps.cf >= L.m) return (ps.early > 0 ? 1 : 0) + 1;
    return (ps.early > 0 ? 1 : 0) + 1 + (ps.cf > ps.early ? 1 : 0) + 1 + (ps.split > 0 ? 2 : 0);   // split: + walker piece + dual piece
}
// stages [0, panel_top) of the sweeps run fused in one launch (k_bp_top): the leading stages with <= 32 parents
int panel_top(const rb_solver *s) {
    const int stages = s->P.L.num_stages - 1;   // nonleaf stages
    int t = 0;
    while (t < stages && s->stage_off[t + 1] - s->stage_off[t] <= 32) ++t;
    return t >= 2 ? t : 0;
}
int iter_launches(const rb_solver *s) {
    const SweepPlan &pl = s->plan;
    if (use_panel(s)) {   // kproj | stages backward, (fused top), forward | dual x 3 | check
        const int top = panel_top(s);
        return 1 + 2 * (s->P.L.num_stages - 1 - top) + (top > 0 ? 1 : 0) + 3 + 1;
    }
    const int sweeps = 1 + 2 * pl.num_levels - (sweeps_fused(s) ? 2 : 0);
    if (use_pipe(s))
        return 1 + sweeps + pipe_dual_launches(s) + (s->loop_armed ? 0 : 1) + (s->risk_split && pipe_split(s).cf < s->P.L.m ? 1 : 0) +
               (s->allow_table_prefetch && !s->h_tables.empty() ? 1 : 0);   // (armed loop: no k_check launch)
    return 1 + sweeps + 1 + 1;
}

// the launches of the DP sweeps on `prim` (x, u rows hold xbar, ubar on entry and the projection on exit); evs / nev:
// optional event after every launch
int launch_sweeps(rb_solver *s, const Ctrl *ctrl, double *prim, cudaStream_t st, cudaEvent_t *evs = nullptr, int *nev = nullptr,
                  const std::function<void()> &after_top = nullptr, int fwd_split = 0,
                  const std::function<void()> &after_piece = nullptr) {
    int ne = 0;
    const SweepPlan &pl = s->plan;
    const Layout &L = s->P.L;
    const unsigned batch = (unsigned)L.batch;
    const size_t per_warp = (size_t)(2 * L.nxu + 32) * sizeof(double);
    auto grid = [&](const SweepLevel &lv) { return dim3((lv.num_sub + lv.subs_per_cta - 1) / lv.subs_per_cta, batch); };
    auto threads = [&](const SweepLevel &lv) { return 32 * lv.warps_per_sub * lv.subs_per_cta; };
    auto smem = [&](const SweepLevel &lv) {
        return per_warp * lv.warps_per_sub * lv.subs_per_cta + (size_t)lv.subs_per_cta * lv.stage_cap * L.nxu * sizeof(double);
    };
    const bool tree = s->tree_mode > 0;
    const bool fused = s->tree_mode > 1 && s->fuse_ok && !(s->allow_mma && pl.lv[0].num_tiles > 0);
    // launch overlap (programmatic dependent launch): backward walker -> fused tree kernel -> forward walker, inside the loop only
    // (ctrl carries the status word of the bounded waits), one-warp tensor-core walkers on the level right below the tree kernel
    const bool w4 = mma_four_warps(s);
    const bool overlap = s->allow_overlap && fused && ctrl && s->overlap_sync && !evs && fwd_split == 0 && pl.num_levels == 2 &&
                         s->allow_mma && pl.lv[1].num_tiles > 0 && !w4;
    int *walk_count = overlap ? s->overlap_sync : nullptr, *tree_done = overlap ? s->overlap_sync + L.batch : nullptr;
    auto bwd = [&](int v) {
        if (s->allow_mma && pl.lv[v].num_tiles > 0)
            launch_chain_mma_bwd(st, s->P, ctrl, pl.lv[v], prim, s->q, s->r, w4, walk_count, tree_done);
        else if (tree && s->tree_lv[v].desc)
            launch_tree_bwd(dim3(s->tree_lv[v].num_sub, batch), 32 * s->tree_lv[v].warps, s->tree_smem[1 + v], st, s->P, ctrl,
                            s->tree_lv[v], prim, s->q, s->r);
        else
            launch_sweep_sub_bwd(grid(pl.lv[v]), threads(pl.lv[v]), smem(pl.lv[v]), st, s->P, ctrl, pl.lv[v], prim, s->q, s->r);
        if (evs) cudaEventRecord(evs[ne++], st);
    };
    auto fwd = [&](int v) {
        if (s->allow_mma && pl.lv[v].num_tiles > 0 && fwd_split > 0 && v == pl.num_levels - 1) {
            launch_chain_mma_fwd(st, s->P, ctrl, pl.lv[v], prim, s->r, 0, fwd_split, mma_four_warps(s));
            if (after_piece) after_piece();
            launch_chain_mma_fwd(st, s->P, ctrl, pl.lv[v], prim, s->r, fwd_split, -1, mma_four_warps(s));
        } else if (s->allow_mma && pl.lv[v].num_tiles > 0)
            launch_chain_mma_fwd(st, s->P, ctrl, pl.lv[v], prim, s->r, 0, -1, w4, tree_done, s->tree_lv[0].num_sub);
        else if (tree && s->tree_lv[v].desc)
            launch_tree_fwd(dim3(s->tree_lv[v].num_sub, batch), 32 * s->tree_lv[v].warps, s->tree_smem[1 + v], st, s->P, ctrl,
                            s->tree_lv[v], prim, s->r);
        else
            launch_sweep_sub_fwd(grid(pl.lv[v]), threads(pl.lv[v]), smem(pl.lv[v]), st, s->P, ctrl, pl.lv[v], prim, s->r);
        if (evs) cudaEventRecord(evs[ne++], st);
    };
    for (int v = pl.num_levels - 1; v >= (fused ? 1 : 0); --v) bwd(v);
    if (fused) {
        cudaError_t e = launch_tree_fused((int)batch, s->fuse_threads, s->fuse_smem, st, s->P, ctrl, s->tree_lv[0], s->tree_top,
                                          prim, s->q, s->r, s->x0, s->tree_sync, walk_count, overlap ? pl.lv[1].num_tiles : 0,
                                          tree_done);
        if (e != cudaSuccess) return fail(s, RB_ERR_CUDA, std::string("fused tree launch: ") + cudaGetErrorString(e));
    } else if (tree && s->tree_top.desc) {
        launch_tree_top(batch, 32 * s->tree_top.warps, s->tree_smem[0], st, s->P, ctrl, s->tree_top, prim, s->q, s->r, s->x0);
    } else {
        launch_sweep_top(batch, 32 * s->top_warps, per_warp * s->top_warps + (size_t)pl.top_cap * L.nxu * sizeof(double), st,
                         s->P, ctrl, pl, prim, s->q, s->r, s->x0);
    }
    if (evs) cudaEventRecord(evs[ne++], st);
    if (after_top) after_top();
    for (int v = fused ? 1 : 0; v < pl.num_levels; ++v) fwd(v);
    if (nev) *nev = ne;
    return launch_ok(s, "DP sweeps");
}

// Chains whose nodes share the dynamics row and the factorisation class at every depth are grouped eight to a tile
// (chain_mma.cu).  `lo` = [num][depth] node ids.  The tiling is dropped (num_tiles = 0) when it would leave more than half
// of the tile columns empty -- then the one-warp-per-chain walker is the better kernel.
int build_chain_tiles(rb_solver *s, const std::vector<int> &lo, int depth, int first, int num, const int **tiles_out,
                      int *num_tiles_out, const int **meta_out) {
    *tiles_out = nullptr;
    *num_tiles_out = 0;
    *meta_out = nullptr;
    if (num <= 0) return RB_OK;
    const int m = s->P.L.m;
    std::map<std::vector<int>, std::vector<int>> groups;
    std::vector<int> sig((size_t)2 * depth);
    for (int c = first; c < first + num; ++c) {
        bool one_dyn = true;   // the walkers keep ONE dynamics row in registers per tile
        for (int d = 0; d < depth; ++d) {
            const int node = lo[(size_t)c * depth + d];
            sig[2 * d] = d == 0 ? 0 : s->dyn_idx[node];
            sig[2 * d + 1] = node < m ? s->cls[node] : -1;
            one_dyn = one_dyn && (d < 2 || sig[2 * d] == sig[2]);
        }
        if (!one_dyn) return RB_OK;
        groups[sig].push_back(c);
    }
    std::vector<int> tiles;
    for (auto &kv : groups) {
        const std::vector<int> &cs = kv.second;
        for (size_t i = 0; i < cs.size(); i += 8)
            for (size_t k = 0; k < 8; ++k) tiles.push_back(i + k < cs.size() ? cs[i + k] : -1);
    }
    const int num_tiles = (int)(tiles.size() / 8);
    if ((long long)num_tiles * 8 > 2LL * num) return RB_OK;
    int *d_tiles = nullptr;
    int rc = upload(s, tiles.data(), tiles.size(), &d_tiles);
    if (rc != RB_OK) return rc;
    // the image of the per-warp metadata the walkers keep in shared memory
    const int stride = depth * 10 + 8;
    std::vector<int> meta((size_t)num_tiles * stride);
    for (int t = 0; t < num_tiles; ++t) {
        int *mt = meta.data() + (size_t)t * stride;
        const int c0 = tiles[(size_t)t * 8];
        for (int g = 0; g < 8; ++g) {
            const int own = tiles[(size_t)t * 8 + g], c = own >= 0 ? own : c0;
            for (int d = 0; d < depth; ++d) mt[d * 8 + g] = lo[(size_t)c * depth + d];
            mt[depth * 10 + g] = own >= 0 ? 1 : 0;
        }
        for (int d = 0; d < depth; ++d) {
            const int node = lo[(size_t)c0 * depth + d];
            mt[depth * 8 + d] = s->dyn_idx[node];
            mt[depth * 9 + d] = node < m ? s->cls[node] : -1;
        }
    }
    int *d_meta = nullptr;
    rc = upload(s, meta.data(), meta.size(), &d_meta);
    if (rc != RB_OK) return rc;
    *tiles_out = d_tiles;
    *num_tiles_out = num_tiles;
    *meta_out = d_meta;
    return RB_OK;
}

// Pack the topology of every subtree of a level into one contiguous descriptor per subtree (tree_sweeps.cu) and decide what
// fits into shared memory.  lo / hi: [num_sub][depth] node ranges.  out->desc stays null if the level cannot be served.
int build_tree_level(rb_solver *s, const std::vector<int> &lo, const std::vector<int> &hi, int depth, int num_sub, bool top,
                     TreeLevel *out, size_t *smem_out) {
    *out = TreeLevel{};
    const Layout &L = s->P.L;
    const int m = L.m;
    if (depth <= 0 || num_sub <= 0) return RB_OK;
    TreeLevel lv{};
    lv.depth = depth;
    lv.num_sub = num_sub;
    lv.num_dyn = s->num_dyn;
    std::vector<int> ns(num_sub), ne(num_sub, 0), xf(num_sub, 0);
    for (int c = 0; c < num_sub; ++c) {
        int tot = 0;
        for (int d = 0; d < depth; ++d) {
            const int a = lo[(size_t)c * depth + d], b = hi[(size_t)c * depth + d];
            tot += b - a;
            lv.max_row = std::max(lv.max_row, b - a);
        }
        ns[c] = tot;
        const int a = lo[(size_t)c * depth + depth - 1], b = hi[(size_t)c * depth + depth - 1];
        if (a < m) {
            if (b > m) return RB_OK;   // a stage mixing leaves and nonleaf nodes: not handled here
            xf[c] = s->child_first[a];
            ne[c] = s->child_first[b - 1] + s->child_count[b - 1] - xf[c];
        }
        lv.max_nodes = std::max(lv.max_nodes, tot);
        lv.max_ext = std::max(lv.max_ext, ne[c]);
        lv.max_row = std::max(lv.max_row, ne[c]);
    }
    lv.desc_stride = 4 + 3 * depth + 5 * lv.max_nodes + 2 * lv.max_ext;
    lv.warps = std::min(16, std::max(1, lv.max_row));
    // what fits: operator tables resident in shared memory, or only the rows
    const size_t limit = 200 * 1024;
    bool fits = false;
    for (int opt = 0; opt < 2 && !fits; ++opt) {
        lv.resident = opt == 0;
        fits = tree_smem_bytes(lv, L.nx, L.nu, lv.warps, top) <= limit;
    }
    if (!fits) return RB_OK;
    std::vector<int> desc((size_t)num_sub * lv.desc_stride, 0);
    for (int c = 0; c < num_sub; ++c) {
        int *h = desc.data() + (size_t)c * lv.desc_stride;
        int *dlo = h + 4, *dw = dlo + depth, *doff = dw + depth, *dyn = doff + depth, *cls = dyn + lv.max_nodes,
            *cfirst = cls + lv.max_nodes, *ccount = cfirst + lv.max_nodes, *par = ccount + lv.max_nodes,
            *xdyn = par + lv.max_nodes, *xpar = xdyn + lv.max_ext;
        h[0] = ns[c];
        h[1] = ne[c];
        h[2] = xf[c];
        int off = 0;
        for (int d = 0; d < depth; ++d) {
            dlo[d] = lo[(size_t)c * depth + d];
            dw[d] = hi[(size_t)c * depth + d] - dlo[d];
            doff[d] = off;
            off += dw[d];
        }
        for (int d = 0; d < depth; ++d)
            for (int p = 0; p < dw[d]; ++p) {
                const int node = dlo[d] + p, i = doff[d] + p;
                dyn[i] = s->dyn_idx[node];
                cls[i] = node < m ? s->cls[node] : -1;
                par[i] = d > 0 ? doff[d - 1] + (s->parent[node] - dlo[d - 1]) : -1;
                if (node < m) {
                    cfirst[i] = s->child_first[node] - (d + 1 < depth ? dlo[d + 1] : xf[c]);
                    ccount[i] = s->child_count[node];
                }
            }
        for (int e = 0; e < ne[c]; ++e) {
            xdyn[e] = s->dyn_idx[xf[c] + e];
            xpar[e] = doff[depth - 1] + (s->parent[xf[c] + e] - dlo[depth - 1]);
        }
    }
    int *d_desc = nullptr;
    int rc = upload(s, desc.data(), desc.size(), &d_desc);
    if (rc != RB_OK) return rc;
    lv.desc = d_desc;
    *out = lv;
    *smem_out = tree_smem_bytes(lv, L.nx, L.nu, lv.warps, top);
    return RB_OK;
}

#define CTRY(x)                       \
    do {                              \
        int rc_ = (x);                \
        if (rc_ != RB_OK) return rc_; \
    } while (0)
#define CTRYC(call) RB_CUDA(s, call)

// rb_create, part: the topology the kernels rely on: stages partition the nodes, children are consecutive ranges in node order, class order
int create_validate(rb_solver *s, const rb_problem *pb) {
    const int n = pb->n, m = pb->m, nx = pb->nx, nu = pb->nu, nl = n - m;
    Layout &L = s->P.L;
    Topo &T = s->P.t;
    Tabs &M = s->P.m;
    (void)n; (void)m; (void)nx; (void)nu; (void)nl; (void)L; (void)T; (void)M;
    // ---- validate the topology the kernels rely on
    if (s->stage_off[0] != 0 || s->stage_off[pb->num_stages] != n || s->stage_off[pb->num_stages - 1] != m) {
        return fail(s, RB_ERR_INVALID, "stage_off must partition 0..n with the leaves as the last stage");
    }
    {
        int expect = 1;
        for (int i = 0; i < m; ++i) {
            if (s->child_count[i] < 1 || s->child_first[i] != expect) {
                return fail(s, RB_ERR_INVALID, "children of the nonleaf nodes must be consecutive ranges in node order (renumber the tree " "breadth first)");
            }
            for (int j = expect; j < expect + s->child_count[i]; ++j)
                if (j >= n || s->parent[j] != i) {
                    return fail(s, RB_ERR_INVALID, "parent[] and child ranges disagree");
                }
            expect += s->child_count[i];
        }
        if (expect != n) {
            return fail(s, RB_ERR_INVALID, "child ranges do not cover all nodes");
        }
        for (int i = 0; i < m; ++i) {
            if (s->cls[i] < 0 || s->cls[i] >= s->num_cls) {
                return fail(s, RB_ERR_INVALID, "cls[] out of range");
            }
            for (int j = s->child_first[i]; j < s->child_first[i] + s->child_count[i]; ++j)
                if (j < m && s->cls[j] <= s->cls[i]) {
                    return fail(s, RB_ERR_INVALID, "class of a child must be larger than the class of its parent");
                }
        }
    }
    return RB_OK;
}

// rb_create, part: compact / padded layout of the iterates, topology and operator tables on the device
int create_layout_and_tables(rb_solver *s, const rb_problem *pb) {
    const int n = pb->n, m = pb->m, nx = pb->nx, nu = pb->nu, nl = n - m;
    Layout &L = s->P.L;
    Topo &T = s->P.t;
    Tabs &M = s->P.m;
    (void)n; (void)m; (void)nx; (void)nu; (void)nl; (void)L; (void)T; (void)M;
    // ---- layout
    L.n = n; L.m = m; L.nleaf = nl; L.nx = nx; L.nu = nu; L.nxu = nx + nu;
    L.num_stages = pb->num_stages; L.batch = pb->batch;
    L.has_nl_rect = pb->num_nl_rect > 0; L.has_leaf_rect = pb->num_leaf_rect > 0;
    std::vector<int> yoff(m + 1, 0);
    for (int i = 0; i < m; ++i) yoff[i + 1] = yoff[i] + 2 * s->child_count[i] + 1;
    L.ysz = yoff[m];
    for (int i = 0; i < m; ++i) s->max_children = std::max(s->max_children, s->child_count[i]);
    s->chain_first = m;   // trailing run of nonleaf nodes with one child: the chain part of the tree (k_dual_chain)
    while (s->chain_first > 1 && s->child_count[s->chain_first - 1] == 1) --s->chain_first;
    {
        int64_t pc = 0, pp = 0;  // compact / padded cursors
        auto addp = [&](long long &field, int64_t len) {
            field = pp;
            s->pseg.push_back({pc, pp, len});
            pc += len;
            pp = round16(pp + len);
        };
        addp(L.px, (int64_t)n * nx); addp(L.pu, (int64_t)m * nu); addp(L.py, L.ysz); addp(L.ptau, n); addp(L.ps, n);
        s->np = pc; L.np_pad = pp;
        pc = 0; pp = 0;
        auto addd = [&](long long &field, int64_t len) {
            field = pp;
            s->dseg.push_back({pc, pp, len});
            pc += len;
            pp = round16(pp + len);
        };
        addd(L.d1, L.ysz); addd(L.d2, m); addd(L.d3, (int64_t)(n - 1) * nx); addd(L.d4, (int64_t)(n - 1) * nu);
        addd(L.d5, n - 1); addd(L.d6, n - 1); addd(L.d7, L.has_nl_rect ? (int64_t)m * (nx + nu) : 0);
        addd(L.d11, (int64_t)nl * nx); addd(L.d12, nl); addd(L.d13, nl); addd(L.d14, L.has_leaf_rect ? (int64_t)nl * nx : 0);
        s->nd = pc; L.nd_pad = pp;
    }
    // ---- topology and tables to the device
    int *tmp_i = nullptr;
    double *tmp_d = nullptr;
    CTRY(upload(s, pb->parent, n, &tmp_i)); T.parent = tmp_i;
    CTRY(upload(s, pb->child_first, m, &tmp_i)); T.child_first = tmp_i;
    CTRY(upload(s, pb->child_count, m, &tmp_i)); T.child_count = tmp_i;
    CTRY(upload(s, yoff.data(), m + 1, &tmp_i)); T.yoff = tmp_i;
    CTRY(upload(s, pb->dyn_idx, n, &tmp_i)); T.dyn_idx = tmp_i;
    CTRY(upload(s, pb->cost_idx, n, &tmp_i)); T.cost_idx = tmp_i;
    CTRY(upload(s, pb->leafcost_idx, nl, &tmp_i)); T.leafcost_idx = tmp_i;
    CTRY(upload(s, L.has_nl_rect ? pb->nl_rect_idx : nullptr, m, &tmp_i)); T.nl_rect_idx = tmp_i;
    {   // packed topology records of the chain nodes (k_dual_chain): child, cost row of the child, y offset, rectangle row
        std::vector<int4> recs(std::max(1, m - s->chain_first));
        for (int i = s->chain_first; i < m; ++i) {
            const int j = s->child_first[i];
            recs[i - s->chain_first] = make_int4(j, s->cost_idx[j], yoff[i], L.has_nl_rect ? pb->nl_rect_idx[i] : 0);
        }
        CTRY(upload(s, recs.data(), recs.size(), &s->chain_recs));
        s->chain_yo0 = yoff[std::min(s->chain_first, m)];
        s->chain_stride = s->chain_first < m ? s->child_first[s->chain_first] - s->chain_first : -1;
        for (int i = s->chain_first; i < m; ++i)
            if (s->child_first[i] - i != s->chain_stride) s->chain_stride = -1;
    }
    CTRY(upload(s, L.has_leaf_rect ? pb->leaf_rect_idx : nullptr, nl, &tmp_i)); T.leaf_rect_idx = tmp_i;
    CTRY(upload(s, pb->cls, m, &tmp_i)); T.cls = tmp_i;
    CTRY(upload(s, pb->cond_prob, n, &tmp_d)); T.cond_prob = tmp_d;
    CTRY(upload(s, pb->risk_alpha, m, &tmp_d)); T.risk_alpha = tmp_d;
    CTRY(upload(s, pb->A, (size_t)pb->num_dyn * nx * nx, &tmp_d)); M.A = tmp_d;
    CTRY(upload(s, pb->B, (size_t)pb->num_dyn * nx * nu, &tmp_d)); M.B = tmp_d;
    {
        const int nxu = nx + nu;
        std::vector<double> cat((size_t)pb->num_dyn * nx * nxu), catT(cat.size());
        for (int t = 0; t < pb->num_dyn; ++t)
            for (int l = 0; l < nx; ++l) {
                for (int k = 0; k < nx; ++k) {
                    const double a_lk = pb->A[((size_t)t * nx + l) * nx + k];
                    cat[((size_t)t * nx + l) * nxu + k] = a_lk;            // row l = [A[l][:], B[l][:]]
                    catT[((size_t)t * nxu + k) * nx + l] = a_lk;           // row k<nx = A[:][k]
                }
                for (int a = 0; a < nu; ++a) {
                    const double b_la = pb->B[((size_t)t * nx + l) * nu + a];
                    cat[((size_t)t * nx + l) * nxu + nx + a] = b_la;
                    catT[((size_t)t * nxu + nx + a) * nx + l] = b_la;      // row nx+a = B[:][a]
                }
            }
        CTRY(upload(s, cat.data(), cat.size(), &tmp_d)); M.ABcat = tmp_d;
        CTRY(upload(s, catT.data(), catT.size(), &tmp_d)); M.ABcatT = tmp_d;
        auto sq = transpose_tab(pb->sqrtQ, pb->num_cost, nx, nx);
        CTRY(upload(s, sq.data(), sq.size(), &tmp_d)); M.sqT = tmp_d;
        auto sr = transpose_tab(pb->sqrtR, pb->num_cost, nu, nu);
        CTRY(upload(s, sr.data(), sr.size(), &tmp_d)); M.srT = tmp_d;
        auto sf = transpose_tab(pb->sqrtQf, pb->num_leafcost, nx, nx);
        CTRY(upload(s, sf.data(), sf.size(), &tmp_d)); M.sqfT = tmp_d;
        M.sq_diag = is_diag(pb->sqrtQ, pb->num_cost, nx);
        M.sr_diag = is_diag(pb->sqrtR, pb->num_cost, nu);
        M.sqf_diag = is_diag(pb->sqrtQf, pb->num_leafcost, nx);
        s->diag_costs = M.sq_diag && M.sr_diag && M.sqf_diag;
        auto diag_of = [](const double *tab, int count, int dim) {
            std::vector<double> d((size_t)count * dim);
            for (int t = 0; t < count; ++t)
                for (int k = 0; k < dim; ++k) d[(size_t)t * dim + k] = tab[((size_t)t * dim + k) * dim + k];
            return d;
        };
        auto dq = diag_of(pb->sqrtQ, pb->num_cost, nx), dr = diag_of(pb->sqrtR, pb->num_cost, nu),
             df = diag_of(pb->sqrtQf, pb->num_leafcost, nx);
        CTRY(upload(s, dq.data(), dq.size(), &tmp_d)); M.sq_d = tmp_d;
        CTRY(upload(s, dr.data(), dr.size(), &tmp_d)); M.sr_d = tmp_d;
        CTRY(upload(s, df.data(), df.size(), &tmp_d)); M.sqf_d = tmp_d;
    }
    CTRY(upload(s, L.has_nl_rect ? pb->nl_lo : nullptr, (size_t)pb->num_nl_rect * (nx + nu), &tmp_d)); M.nl_lo = tmp_d;
    CTRY(upload(s, L.has_nl_rect ? pb->nl_hi : nullptr, (size_t)pb->num_nl_rect * (nx + nu), &tmp_d)); M.nl_hi = tmp_d;
    CTRY(upload(s, L.has_leaf_rect ? pb->leaf_lo : nullptr, (size_t)pb->num_leaf_rect * nx, &tmp_d)); M.leaf_lo = tmp_d;
    CTRY(upload(s, L.has_leaf_rect ? pb->leaf_hi : nullptr, (size_t)pb->num_leaf_rect * nx, &tmp_d)); M.leaf_hi = tmp_d;
    return RB_OK;
}

// rb_create, part: class structure of the offline factorisation, iterates and scratch
int create_classes_and_iterates(rb_solver *s, const rb_problem *pb) {
    const int n = pb->n, m = pb->m, nx = pb->nx, nu = pb->nu, nl = n - m;
    Layout &L = s->P.L;
    Topo &T = s->P.t;
    Tabs &M = s->P.m;
    (void)n; (void)m; (void)nx; (void)nu; (void)nl; (void)L; (void)T; (void)M;
    // ---- class structure for the offline factorisation: representative = first node of the class
    {
        std::vector<int> rep(s->num_cls, -1);
        for (int i = 0; i < m; ++i)
            if (rep[s->cls[i]] < 0) rep[s->cls[i]] = i;
        std::vector<int> level(s->num_cls, 0);
        s->cls_child_ptr.assign(1, 0);
        for (int c = 0; c < s->num_cls; ++c) {
            if (rep[c] < 0) {
                return fail(s, RB_ERR_INVALID, "class without nodes");
            }
            const int i = rep[c];
            for (int j = s->child_first[i]; j < s->child_first[i] + s->child_count[i]; ++j) {
                s->cls_child_dyn.push_back(s->dyn_idx[j]);
                s->cls_child_cls.push_back(j < m ? s->cls[j] : -1);
            }
            s->cls_child_ptr.push_back((int)s->cls_child_dyn.size());
        }
        int max_level = 0;
        for (int c = s->num_cls - 1; c >= 0; --c) {  // children have larger class ids: already final when visited
            int lv = 0;
            for (int k = s->cls_child_ptr[c]; k < s->cls_child_ptr[c + 1]; ++k)
                if (s->cls_child_cls[k] >= 0) lv = std::max(lv, level[s->cls_child_cls[k]] + 1);
            level[c] = lv;
            max_level = std::max(max_level, lv);
        }
        s->level_ptr.assign(max_level + 2, 0);
        for (int c = 0; c < s->num_cls; ++c) ++s->level_ptr[level[c] + 1];
        for (int l = 0; l <= max_level; ++l) s->level_ptr[l + 1] += s->level_ptr[l];
        s->level_list.resize(s->num_cls);
        std::vector<int> cur(s->level_ptr.begin(), s->level_ptr.end() - 1);
        for (int c = 0; c < s->num_cls; ++c) s->level_list[cur[level[c]]++] = c;
        CTRY(upload(s, s->cls_child_ptr.data(), s->cls_child_ptr.size(), &s->d_cls_child_ptr));
        CTRY(upload(s, s->cls_child_dyn.data(), s->cls_child_dyn.size(), &s->d_cls_child_dyn));
        CTRY(upload(s, s->cls_child_cls.data(), s->cls_child_cls.size(), &s->d_cls_child_cls));
        CTRY(upload(s, s->level_list.data(), s->level_list.size(), &s->d_level_list));
    }
    CTRY(dev_zero(s, (size_t)s->num_cls * nx * nx, &s->Ptab));
    CTRY(dev_zero(s, (size_t)s->num_cls * nu * nx, &s->Ktab));
    CTRY(dev_zero(s, (size_t)s->num_cls * (nx + nu) * nu, &s->KRcatT));
    M.K = s->Ktab; M.KRcatT = s->KRcatT;
    // ---- iterates and scratch
    const size_t B = (size_t)L.batch;
    for (int w = 0; w < 2; ++w) {
        CTRY(dev_zero(s, B * L.np_pad, &s->prim[w]));
        CTRY(dev_zero(s, B * L.nd_pad, &s->dual[w]));
    }
    CTRY(dev_zero(s, B * n * nx, &s->q));
    CTRY(dev_zero(s, B * m * nu, &s->r));
    CTRY(dev_zero(s, B * nx, &s->x0));
    CTRY(dev_zero(s, 1, &s->ctrl));
    CTRY(dev_zero(s, B * 6 * 2, &s->slots));   // second half: the other parity of the pipelined sharded loop
    CTRY(dev_zero(s, B * 6, &s->last));
    if (B * 6 <= 1024 && L.nx <= 1024 &&
        cudaHostAlloc((void **)&s->h_last, B * 6 * sizeof(double), cudaHostAllocMapped) == cudaSuccess) {
        if (cudaHostGetDevicePointer((void **)&s->h_last_dev, s->h_last, 0) != cudaSuccess) {
            cudaFreeHost(s->h_last);
            s->h_last = s->h_last_dev = nullptr;
        }
    } else {
        cudaGetLastError();
        s->h_last = s->h_last_dev = nullptr;
    }
    CTRY(dev_zero(s, 1, &s->status));
    return RB_OK;
}

// rb_create, part: the sweep plan (cut stages, chain tiles, tree descriptors, fused launch) and the tensor-core fragment tables
int create_sweep_plan(rb_solver *s, const rb_problem *pb) {
    const int n = pb->n, m = pb->m, nx = pb->nx, nu = pb->nu, nl = n - m;
    Layout &L = s->P.L;
    Topo &T = s->P.t;
    Tabs &M = s->P.m;
    (void)n; (void)m; (void)nx; (void)nu; (void)nl; (void)L; (void)T; (void)M;
    // ---- sweep plan (sweeps.cu): first cut at the first stage with >= 64 nodes; second cut where the tree turns into
    //      chains (below the stopping time of a Markov tree) if there are >= 200 of them (cfg2, 243 chains: 13 222 vs 12 208 it/s with the chain level), else -- if the tree keeps
    //      branching -- at the first stage with >= 2048 nodes and >= 8x the first cut
    {
        SweepPlan &pl = s->plan;
        auto width = [&](int t) { return s->stage_off[t + 1] - s->stage_off[t]; };
        if ((pb->sweep_cut1_min != 0 || pb->sweep_cut2_min != 0) && pb->shard_world > 1) {
            return fail(s, RB_ERR_INVALID, "sweep_cut*_min must be 0 for a sharded problem");
        }
        const int cut1_min = pb->sweep_cut1_min > 0 ? pb->sweep_cut1_min : 64;
        const int cut2_min = pb->sweep_cut2_min > 0 ? pb->sweep_cut2_min : 200;
        int c1 = L.num_stages, c2 = L.num_stages;
        for (int t = 0; t < L.num_stages; ++t)
            if (width(t) >= cut1_min) {
                c1 = t;
                break;
            }
        int c_chain = L.num_stages - 1;   // first stage from which every node has at most one child
        while (c_chain > 0) {
            bool chains = true;
            for (int i = s->stage_off[c_chain - 1]; i < s->stage_off[c_chain] && chains; ++i) chains = s->child_count[i] <= 1;
            if (!chains) break;
            --c_chain;
        }
        if (c_chain > c1 && c_chain < L.num_stages - 1 && L.num_stages - c_chain <= 64 && width(c_chain) >= cut2_min) {
            c2 = c_chain;
            // The top is ONE CTA and every level-0 subtree is one CTA; a stage of either costs what its parents cost.
            // Balance them: the first cut goes where the widest parent stage of the top, width(c - 1), and the widest
            // parent stage of a subtree, width(c2 - 1) / width(c), are closest (cfg3: stage 3 either way; cfg5, 3 modes:
            // stage 3 with 27 subtrees of 1 + 3 + 9 nodes instead of stage 4 with a 27-parent stage in the top).
            if (pb->sweep_cut1_min == 0) {
                auto cost = [&](int c) { return std::max(width(c - 1), width(c2 - 1) / std::max(1, width(c))); };
                for (int c = c1 - 1; c >= 1 && width(c) >= 8; --c)
                    if (cost(c) < cost(c1)) c1 = c;
            }
        } else {
            for (int t = c1 + 1; t < L.num_stages; ++t)
                if (width(t) >= std::max(2048, cut2_min) && width(t) >= 8 * width(c1)) {
                    c2 = t;
                    break;
                }
        }
        pl.t_top = c1;
        pl.num_levels = c1 >= L.num_stages ? 0 : (c2 >= L.num_stages ? 1 : 2);
        int *d_so = nullptr;
        CTRY(upload(s, s->stage_off.data(), s->stage_off.size(), &d_so));
        pl.stage_off = d_so;
        const int cuts[3] = {c1, c2, L.num_stages};
        for (int v = 0; v < pl.num_levels; ++v) {
            SweepLevel &lv = pl.lv[v];
            lv.t_lo = cuts[v];
            lv.depth = (v + 1 < pl.num_levels ? cuts[v + 1] : L.num_stages) - cuts[v];
            lv.num_sub = width(lv.t_lo);
            std::vector<int> lo((size_t)lv.num_sub * lv.depth), hi(lo.size());
            int max_width = 1;
            for (int c = 0; c < lv.num_sub; ++c) {
                int a = s->stage_off[lv.t_lo] + c, b = a + 1;
                for (int d = 0; d < lv.depth; ++d) {
                    lo[(size_t)c * lv.depth + d] = a;
                    hi[(size_t)c * lv.depth + d] = b;
                    max_width = std::max(max_width, b - a);
                    if (d + 1 < lv.depth) {   // children of [a, b) are one contiguous range of the next stage
                        const int na = s->child_first[a], nb = s->child_first[b - 1] + s->child_count[b - 1];
                        a = na;
                        b = nb;
                    }
                }
            }
            // widest stage of any subtree, in nodes or in children (rows of the group-shared buffer)
            int cap = 1;
            for (int c = 0; c < lv.num_sub; ++c)
                for (int d = 0; d < lv.depth; ++d) {
                    const int a = lo[(size_t)c * lv.depth + d], b = hi[(size_t)c * lv.depth + d];
                    cap = std::max(cap, b - a);
                    if (a < m) cap = std::max(cap, s->child_first[b - 1] + s->child_count[b - 1] - s->child_first[a]);
                }
            lv.stage_cap = max_width == 1 ? 0 : cap;
            lv.chain = (max_width == 1 && lv.depth <= 64 && lv.t_lo + lv.depth == L.num_stages) ? 1 : 0;
            lv.warps_per_sub = std::min(16, max_width == 1 ? 1 : cap);
            lv.subs_per_cta = std::max(1, 8 / lv.warps_per_sub);
            int *d_lo = nullptr, *d_hi = nullptr;
            CTRY(upload(s, lo.data(), lo.size(), &d_lo));
            CTRY(upload(s, hi.data(), hi.size(), &d_hi));
            lv.lo = d_lo;
            lv.hi = d_hi;
            lv.tiles = nullptr;
            lv.tile_meta = nullptr;
            lv.num_tiles = 0;
            if (lv.chain && chain_mma_supported(nx, nu)) {
                s->chain_lo[v] = lo;
                CTRY(build_chain_tiles(s, lo, lv.depth, 0, lv.num_sub, &lv.tiles, &lv.num_tiles, &lv.tile_meta));
            }
            if (!lv.chain) CTRY(build_tree_level(s, lo, hi, lv.depth, lv.num_sub, false, &s->tree_lv[v], &s->tree_smem[1 + v]));
        }
        {   // the top of the tree as one subtree: stages [0, t_top)
            std::vector<int> lo(pl.t_top), hi(pl.t_top);
            for (int t = 0; t < pl.t_top; ++t) {
                lo[t] = s->stage_off[t];
                hi[t] = s->stage_off[t + 1];
            }
            CTRY(build_tree_level(s, lo, hi, pl.t_top, 1, true, &s->tree_top, &s->tree_smem[0]));
        }
        size_t tree_need = std::max(s->tree_smem[0], std::max(s->tree_smem[1], s->tree_smem[2]));
        if (s->tree_top.desc && pl.num_levels > 0 && s->tree_lv[0].desc && s->tree_top.resident == s->tree_lv[0].resident) {
            // level 0 with the footprint of a two-way pass (tables of both directions, r kept)
            s->fuse_threads = 32 * std::max(s->tree_top.warps, s->tree_lv[0].warps);
            s->fuse_smem = std::max(tree_smem_bytes(s->tree_top, nx, nu, s->fuse_threads / 32, true),
                                    tree_smem_bytes(s->tree_lv[0], nx, nu, s->fuse_threads / 32, true));
            if (s->fuse_smem <= 200 * 1024) {
                tree_need = std::max(tree_need, s->fuse_smem);
                s->fuse_ok = true;
            }
        }
        CTRYC(tree_kernels_set_smem((int)tree_need));
        if (s->fuse_ok) {
            s->fuse_ok = tree_fused_fits(nx, nu, s->tree_top.resident != 0, s->fuse_threads, s->fuse_smem,
                                         (s->tree_lv[0].num_sub + 1) * L.batch);
            if (s->fuse_ok) CTRY(dev_zero(s, (size_t)2 * L.batch, &s->tree_sync));
            if (s->fuse_ok) CTRY(dev_zero(s, (size_t)2 * L.batch, &s->overlap_sync));
        }
        int top_max = 1;
        for (int t = 0; t < pl.t_top; ++t) top_max = std::max(top_max, std::max(width(t), t + 1 < L.num_stages ? width(t + 1) : 1));
        pl.top_cap = top_max;
        s->top_warps = std::min(32, std::max(1, top_max));
        const size_t per_warp = (size_t)(2 * (nx + nu) + 32) * sizeof(double);
        size_t need = per_warp * s->top_warps + (size_t)pl.top_cap * (nx + nu) * sizeof(double);
        for (int v = 0; v < pl.num_levels; ++v)
            need = std::max(need, per_warp * pl.lv[v].warps_per_sub * pl.lv[v].subs_per_cta +
                                      (size_t)pl.lv[v].subs_per_cta * pl.lv[v].stage_cap * (nx + nu) * sizeof(double));
        if (need > 200 * 1024) {
            return fail(s, RB_ERR_INVALID, "sweep stage buffers do not fit in shared memory");
        }
        s->sweep_smem_max = need;
        CTRYC(sweep_kernels_set_smem((int)need));
        // lane-major MMA fragment tables for the tiled chain levels (filled by rb_offline)
        bool tiled = false;
        for (int v = 0; v < pl.num_levels; ++v) tiled = tiled || pl.lv[v].num_tiles > 0;
        if (tiled) {
            int f_ab, f_abt, f_k, f_kr;
            chain_mma_frag_counts(nx, nu, &f_ab, &f_abt, &f_k, &f_kr);
            double *t_ab = nullptr, *t_abt = nullptr, *t_k = nullptr, *t_kr = nullptr;
            // two images of every table: k-major, then output-block-major (launch_chain_mma_frags)
            CTRY(dev_zero(s, (size_t)2 * s->num_dyn * f_ab * 32, &t_ab));
            CTRY(dev_zero(s, (size_t)2 * s->num_dyn * f_abt * 32, &t_abt));
            CTRY(dev_zero(s, (size_t)2 * s->num_cls * f_k * 32, &t_k));
            CTRY(dev_zero(s, (size_t)2 * s->num_cls * f_kr * 32, &t_kr));
            s->P.m.fragAB = t_ab;
            s->P.m.fragABT = t_abt;
            s->P.m.fragK = t_k;
            s->P.m.fragKR = t_kr;
            s->P.m.fragAB4 = t_ab + (size_t)s->num_dyn * f_ab * 32;
            s->P.m.fragABT4 = t_abt + (size_t)s->num_dyn * f_abt * 32;
            s->P.m.fragK4 = t_k + (size_t)s->num_cls * f_k * 32;
            s->P.m.fragKR4 = t_kr + (size_t)s->num_cls * f_kr * 32;
            size_t mma_need = 0;
            for (int v = 0; v < pl.num_levels; ++v)
                if (pl.lv[v].num_tiles > 0)
                    mma_need = std::max(mma_need, std::max(chain_mma_smem_bytes(nx, nu, pl.lv[v].depth, true),
                                                           chain_mma_smem_bytes(nx, nu, pl.lv[v].depth, false)));
            CTRYC(chain_mma_set_smem((int)mma_need));
        }
    }
    return RB_OK;
}

// rb_create, part: node tiles of the warp-per-node passes (fused.cu)
int create_node_tiles(rb_solver *s, const rb_problem *pb) {
    const int n = pb->n, m = pb->m, nx = pb->nx, nu = pb->nu, nl = n - m;
    Layout &L = s->P.L;
    Topo &T = s->P.t;
    Tabs &M = s->P.m;
    (void)n; (void)m; (void)nx; (void)nu; (void)nl; (void)L; (void)T; (void)M;
    // ---- node tiles: runs of consecutive nodes (<= 32 nodes and <= 64 edges; leaves: <= 64 nodes) -----------------
    {
        std::vector<int2> tiles;
        const int kTileNodes = 32, kTileEdges = 64, kTileLeaves = 64;
        size_t need_p = 0, need_d = 0;
        auto ev = [](long long c) { return (size_t)(c + 2); };   // chunks are widened to 16-byte boundaries
        int i = 0;
        while (i < m) {
            int j = i, edges = 0;
            while (j < m && j - i < kTileNodes && (j == i || edges + s->child_count[j] <= kTileEdges)) edges += s->child_count[j++];
            tiles.push_back(make_int2(i, j));
            const long long nN = j - i, nE = edges, ny = 2 * nE + nN;
            need_p = std::max(need_p, ev(nN * nx) + ev(nN * nu) + 2 * ev(ny) + 2 * ev(nE) + ev(nN) + 3 * ev(nE) +
                                          ev(nE * nx) + ev(nE * nu) + 2 * ev(nE) + ev(nN * (nx + nu)));
            need_d = std::max(need_d, 2 * ev(nN * nx) + 2 * ev(nN * nu) + 3 * ev(ny) + 2 * ev(nE) + 3 * ev(nN) +
                                          ev(nE * nx) + ev(nE * nu) + 2 * ev(nE) + ev(nN * (nx + nu)));
            i = j;
        }
        for (i = m; i < n; i += kTileLeaves) {
            const int j = std::min(n, i + kTileLeaves);
            tiles.push_back(make_int2(i, j));
            const long long nN = j - i;
            need_p = std::max(need_p, 3 * ev(nN * nx));
            need_d = std::max(need_d, 4 * ev(nN * nx) + 4 * ev(nN));
        }
        int2 *d_tiles = nullptr;
        CTRY(upload(s, tiles.data(), tiles.size(), &d_tiles));
        s->tiles.tiles = d_tiles;
        s->tiles.num_tiles = (int)tiles.size();
        s->tiles.rowlen = (std::max(nx, nu) + 1) & ~1;
        s->primal_smem = sizeof(double) * (need_p + (size_t)8 * 4 * s->tiles.rowlen);
        s->dual_smem = sizeof(double) * (need_d + (size_t)8 * kDualRowsHost * s->tiles.rowlen);
        if (s->dual_smem > 200 * 1024 || s->primal_smem > 200 * 1024) {
            return fail(s, RB_ERR_INVALID, "tile does not fit in shared memory");
        }
        CTRYC(tile_kernels_set_smem(s->primal_smem, s->dual_smem));
    }
    s->kernels_per_iter = 0;   // see iter_launches()
    return RB_OK;
}

// rb_create, part: subtree sharding: the rank's node lists, restricted sweep levels and exchange buffers
int create_shard_plan(rb_solver *s, const rb_problem *pb) {
    const int n = pb->n, m = pb->m, nx = pb->nx, nu = pb->nu, nl = n - m;
    Layout &L = s->P.L;
    Topo &T = s->P.t;
    Tabs &M = s->P.m;
    (void)n; (void)m; (void)nx; (void)nu; (void)nl; (void)L; (void)T; (void)M;
    // ---- subtree sharding: rank r owns a contiguous block of the level-0 subtrees, everybody replicates the top
    if (pb->shard_world > 1) {
        const SweepPlan &pl = s->plan;
        const int W = pb->shard_world, R = pb->shard_rank;
        if (R < 0 || R >= W || L.batch != 1 || pl.num_levels < 1 || pl.lv[0].num_sub < W) {
            return fail(s, RB_ERR_INVALID, "subtree sharding needs batch == 1 and at least one cut-stage subtree per rank");
        }
        const int C = pl.lv[0].num_sub;
        std::vector<int> bounds(W + 1);
        for (int r = 0; r <= W; ++r) bounds[r] = (int)((long long)r * C / W);
        ShardPlan &sp = s->shard;
        sp.rank = R; sp.world = W;
        sp.cut_first = s->stage_off[pl.t_top];
        sp.cut_lo = bounds[R]; sp.cut_hi = bounds[R + 1];
        sp.cap = 0;
        for (int r = 0; r < W; ++r) sp.cap = std::max(sp.cap, bounds[r + 1] - bounds[r]);
        int *d_bounds = nullptr;
        CTRY(upload(s, bounds.data(), bounds.size(), &d_bounds));
        sp.cut_bounds = d_bounds;
        // owned nodes: the descendants of the owned cut nodes, one contiguous range per stage
        std::vector<int> owned, top, both;
        for (int i = 0; i < s->stage_off[pl.t_top]; ++i) top.push_back(i);
        std::vector<std::pair<int, int>> stage_range(L.num_stages, {0, 0});
        int a = sp.cut_first + sp.cut_lo, b = sp.cut_first + sp.cut_hi;
        for (int t = pl.t_top; t < L.num_stages; ++t) {
            stage_range[t] = {a, b};
            for (int i = a; i < b; ++i) owned.push_back(i);
            if (t + 1 < L.num_stages) {
                const int na = s->child_first[a], nb = s->child_first[b - 1] + s->child_count[b - 1];
                a = na;
                b = nb;
            }
        }
        both = top;
        both.insert(both.end(), owned.begin(), owned.end());
        CTRY(upload(s, owned.data(), owned.size(), &s->owned_nodes));
        CTRY(upload(s, top.data(), top.size(), &s->top_nodes));
        CTRY(upload(s, both.data(), both.size(), &s->dual_nodes));
        s->n_owned = (int)owned.size();
        s->n_top = (int)top.size();
        // sweep levels restricted to the owned subtrees (a contiguous block of every level)
        for (int v = 0; v < pl.num_levels; ++v) {
            SweepLevel lv = pl.lv[v];
            const int first = s->stage_off[lv.t_lo];
            const int sa = stage_range[lv.t_lo].first - first, sb = stage_range[lv.t_lo].second - first;
            lv.lo += (size_t)sa * lv.depth;
            lv.hi += (size_t)sa * lv.depth;
            lv.num_sub = sb - sa;
            if (lv.num_tiles > 0) {   // tiles of the owned chains only (indices relative to the shifted lo / hi)
                std::vector<int> sub(s->chain_lo[v].begin() + (size_t)sa * lv.depth, s->chain_lo[v].begin() + (size_t)sb * lv.depth);
                CTRY(build_chain_tiles(s, sub, lv.depth, 0, lv.num_sub, &lv.tiles, &lv.num_tiles, &lv.tile_meta));
            }
            s->shard_lv[v] = lv;
            s->shard_tree_lv[v] = s->tree_lv[v];
            if (s->tree_lv[v].desc) {
                s->shard_tree_lv[v].desc += (size_t)sa * s->tree_lv[v].desc_stride;
                s->shard_tree_lv[v].num_sub = sb - sa;
            }
        }
        s->xchg_count = (size_t)sp.cap * (nx + 1) + 6;
        CTRY(dev_zero(s, s->xchg_count, &s->xchg_send));
        CTRY(dev_zero(s, s->xchg_count * W, &s->xchg_recv));
        {   // node lists of the pipelined sharded loop: the chain nodes [chain_first, m) go through k_dual_chain (the rank's
            // columns of every chain stage: OwnMap), everything else the rank owns -- and the replicated top -- through the
            // general lane pass; the kernel projection takes the nonleaf nodes of either set
            std::vector<int> own_nl, top_nl, lane;
            const int cf = s->chain_stride > 0 ? s->chain_first : m;
            for (int i : top) {
                if (i < m) top_nl.push_back(i);
                lane.push_back(i);
            }
            for (int i : owned) {
                if (i < m) own_nl.push_back(i);
                if (i < cf || i >= m) lane.push_back(i);
            }
            CTRY(upload(s, own_nl.data(), own_nl.size(), &s->own_nonleaf));
            CTRY(upload(s, top_nl.data(), top_nl.size(), &s->top_nonleaf));
            CTRY(upload(s, lane.data(), lane.size(), &s->own_lane));
            s->n_own_nonleaf = (int)own_nl.size();
            s->n_top_nonleaf = (int)top_nl.size();
            s->n_own_lane = (int)lane.size();
            if (cf < m) {   // stage of chain_first and the rank's columns there
                int t = 0;
                while (s->stage_off[t + 1] <= cf) ++t;
                const int wdt = s->stage_off[t + 1] - s->stage_off[t];
                if (s->stage_off[t] == cf && wdt == s->chain_stride && stage_range[t].second > stage_range[t].first) {
                    s->own_chain = OwnMap{stage_range[t].second - stage_range[t].first, stage_range[t].first - cf, wdt};
                    s->n_own_chain = (m - cf) / wdt * s->own_chain.w;
                }
            }
        }
        s->sharded = true;
    }
    return RB_OK;
}

#undef CTRY
#undef CTRYC

int need_offline(rb_solver *s) {
    if (!s->have_offline) return fail(s, RB_ERR_STATE, "rb_offline() has not been run");
    return RB_OK;
}

}  // namespace

// ====================================================================================================================
extern "C" {

const char *rb_last_error(const rb_solver *s) { return s ? s->err.c_str() : g_create_error.c_str(); }

int rb_create(const rb_problem *pb, rb_solver **out) {
    if (!pb || !out) return fail(nullptr, RB_ERR_INVALID, "null argument");
    *out = nullptr;
    if (pb->n < 2 || pb->m < 1 || pb->m >= pb->n || pb->nx < 1 || pb->nu < 1 || pb->batch < 1)
        return fail(nullptr, RB_ERR_INVALID, "bad sizes");
    if (pb->nx > kMaxDim || pb->nu > kMaxDim)
        return fail(nullptr, RB_ERR_INVALID, "nx and nu must be <= 64 in this build (kMaxDim)");
    if ((pb->num_nl_rect > 0 && !pb->nl_rect_idx) || (pb->num_leaf_rect > 0 && !pb->leaf_rect_idx))
        return fail(nullptr, RB_ERR_INVALID, "rectangle tables without index arrays");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, RB_ERR_NO_DEVICE,
                    std::string("no CUDA device (raocp_b200 has no CPU fallback): ") + cudaGetErrorString(e));
    if (pb->device < 0 || pb->device >= ndev) return fail(nullptr, RB_ERR_INVALID, "bad device ordinal");

    rb_solver *s = new rb_solver();
    s->device = pb->device;
    auto bail = [&](int rc) {
        g_create_error = s->err;
        rb_destroy(s);
        return rc;
    };
#define TRY(x)                         \
    do {                               \
        int rc_ = (x);                 \
        if (rc_ != RB_OK) return bail(rc_); \
    } while (0)
#define TRYC(call)                                                                                   \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            s->err = std::string(#call) + ": " + cudaGetErrorString(e_);                             \
            return bail(RB_ERR_CUDA);                                                                \
        }                                                                                            \
    } while (0)

    TRYC(cudaSetDevice(s->device));
    TRYC(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    for (auto &q : s->side) TRYC(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
    for (auto &e : s->pev) TRYC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    const int n = pb->n, m = pb->m, nx = pb->nx, nu = pb->nu, nl = n - m;
    s->stage_off.assign(pb->stage_off, pb->stage_off + pb->num_stages + 1);
    s->parent.assign(pb->parent, pb->parent + n);
    s->child_first.assign(pb->child_first, pb->child_first + m);
    s->child_count.assign(pb->child_count, pb->child_count + m);
    s->dyn_idx.assign(pb->dyn_idx, pb->dyn_idx + n);
    s->cost_idx.assign(pb->cost_idx, pb->cost_idx + n);
    s->cls.assign(pb->cls, pb->cls + m);
    s->num_cls = pb->num_cls;
    s->num_dyn = pb->num_dyn;
    s->num_cost = pb->num_cost;
    s->num_leafcost = pb->num_leafcost;
    TRY(create_validate(s, pb));
    TRY(create_layout_and_tables(s, pb));
    TRY(create_classes_and_iterates(s, pb));
    TRY(create_sweep_plan(s, pb));
    TRY(create_node_tiles(s, pb));
    TRY(create_shard_plan(s, pb));
    *out = s;
    return RB_OK;
#undef TRY
#undef TRYC
}

void rb_destroy(rb_solver *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (int i = 0; i < 2; ++i)
        if (s->graph[i]) cudaGraphExecDestroy(s->graph[i]);
    for (void *p : s->allocs) cudaFree(p);
    for (void *p : s->p2p_opened)
        if (p) cudaIpcCloseMemHandle(p);
    if (s->nccl_comm) nccl_comm_destroy(s->nccl_comm);
    if (s->hist) cudaFree(s->hist);
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    if (s->h_last) cudaFreeHost(s->h_last);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    for (auto q : s->side)
        if (q) cudaStreamDestroy(q);
    for (auto e : s->pev)
        if (e) cudaEventDestroy(e);
    delete s;
}

int rb_set_stream(rb_solver *s, void *cuda_stream) {
    if (!s) return RB_ERR_INVALID;
    RB_CUDA(s, cudaStreamSynchronize(s->stream));
    drop_graphs(s);
    if (cuda_stream) {
        if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
        s->stream = (cudaStream_t)cuda_stream;
        s->own_stream = false;
    } else if (!s->own_stream) {
        RB_CUDA(s, cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        s->own_stream = true;
    }
    return RB_OK;
}

int rb_sizes(const rb_solver *s, int64_t *np, int64_t *nd) {
    if (!s) return RB_ERR_INVALID;
    if (np) *np = s->np;
    if (nd) *nd = s->nd;
    return RB_OK;
}

int rb_synchronize(rb_solver *s) {
    if (!s) return RB_ERR_INVALID;
    RB_CUDA(s, cudaStreamSynchronize(s->stream));
    return RB_OK;
}

int rb_launch_count(const rb_solver *s, int64_t *k) {
    if (!s || !k) return RB_ERR_INVALID;
    *k = s->launches;
    return RB_OK;
}

// ---- offline -------------------------------------------------------------------------------------------------------
int rb_offline(rb_solver *s) {
    if (!s) return RB_ERR_INVALID;
    RB_CUDA(s, cudaSetDevice(s->device));
    const size_t smem = offline_smem_bytes(s->P.L.nx, s->P.L.nu);
    RB_CUDA(s, cudaFuncSetAttribute(k_offline_level, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ClassView cv{s->d_cls_child_ptr, s->d_cls_child_dyn, s->d_cls_child_cls, s->d_level_list};
    const int levels = (int)s->level_ptr.size() - 1;
    for (int l = 0; l < levels; ++l) {
        const int begin = s->level_ptr[l], count = s->level_ptr[l + 1] - begin;
        if (count == 0) continue;
        k_offline_level<<<count, 256, smem, s->stream>>>(s->P, cv, begin, count, s->Ptab, s->Ktab, s->KRcatT, s->status);
        RB_LAUNCHED(s, "k_offline_level");
    }
    if (s->P.m.fragK) {
        launch_chain_mma_frags(s->stream, s->P.m, s->P.L.nx, s->P.L.nu, s->num_dyn, s->num_cls, true, true);
        RB_LAUNCHED(s, "k_frag_table");
    }
    int rc = check_status(s);
    if (rc != RB_OK) return rc;
    if (!s->d_tables) {   // the ranges k_prefetch_ranges walks (all read-only after this point)
        const Layout &L = s->P.L;
        const Tabs &M = s->P.m;
        const size_t nx = L.nx, nu = L.nu, nxu = L.nxu, D = s->num_dyn, C = s->num_cls;
        auto add = [&](const double *p, size_t doubles) {
            if (p && doubles) s->h_tables.push_back(PrefetchRange{reinterpret_cast<const char *>(p), (long long)(doubles * sizeof(double))});
        };
        add(M.ABcat, D * nx * nxu);
        add(M.ABcatT, D * nxu * nx);
        add(M.K, C * nu * nx);
        add(M.KRcatT, C * nxu * nu);
        if (M.fragK) {
            int f_ab, f_abt, f_k, f_kr;
            chain_mma_frag_counts(L.nx, L.nu, &f_ab, &f_abt, &f_k, &f_kr);
            add(M.fragAB, 2 * D * f_ab * 32);
            add(M.fragABT, 2 * D * f_abt * 32);
            add(M.fragK, 2 * C * f_k * 32);
            add(M.fragKR, 2 * C * f_kr * 32);
        }
        // worth a launch only when the tables are megabytes (cfg5: 5.7 MB, +4 % cold; cfg3: 0.5 MB, -2 %: measured, gpurun_out r2z)
        long long total = 0;
        for (const auto &t : s->h_tables) total += t.bytes;
        if (total < (2ll << 20) || total > (24ll << 20)) s->h_tables.clear();   // per-node K / R~ (dedup off) is streaming data, not a table
        rc = upload(s, s->h_tables.data(), s->h_tables.size(), &s->d_tables);
        if (rc != RB_OK) return rc;
    }
    s->have_offline = true;
    return RB_OK;
}

int rb_get_offline(rb_solver *s, double *Pm, double *K, double *Rinv) {
    if (!s) return RB_ERR_INVALID;
    int rc = need_offline(s);
    if (rc != RB_OK) return rc;
    const int nx = s->P.L.nx, nu = s->P.L.nu;
    RB_CUDA(s, cudaStreamSynchronize(s->stream));
    if (Pm) RB_CUDA(s, cudaMemcpy(Pm, s->Ptab, (size_t)s->num_cls * nx * nx * sizeof(double), cudaMemcpyDeviceToHost));
    if (K) RB_CUDA(s, cudaMemcpy(K, s->Ktab, (size_t)s->num_cls * nu * nx * sizeof(double), cudaMemcpyDeviceToHost));
    if (Rinv) {
        std::vector<double> t((size_t)s->num_cls * (nx + nu) * nu);
        RB_CUDA(s, cudaMemcpy(t.data(), s->KRcatT, t.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int c = 0; c < s->num_cls; ++c)   // KRcatT[c][nx + b][a] = Rinv[a][b]
            for (int a = 0; a < nu; ++a)
                for (int b = 0; b < nu; ++b)
                    Rinv[((size_t)c * nu + a) * nu + b] = t[((size_t)c * (nx + nu) + nx + b) * nu + a];
    }
    return RB_OK;
}

// ---- iterate access -------------------------------------------------------------------------------------------------
int rb_set_primal(rb_solver *s, int which, const double *compact) {
    if (!s || !compact || which < 0 || which > 1) return RB_ERR_INVALID;
    RB_MATERIALISE(s);
    return copy_segments(s, s->pseg, s->np, s->P.L.np_pad, view_p(s, which), compact, nullptr);
}
int rb_get_primal(rb_solver *s, int which, double *compact) {
    if (!s || !compact || which < 0 || which > 1) return RB_ERR_INVALID;
    return copy_segments(s, s->pseg, s->np, s->P.L.np_pad, view_p(s, which), nullptr, compact);
}
int rb_set_dual(rb_solver *s, int which, const double *compact) {
    if (!s || !compact || which < 0 || which > 1) return RB_ERR_INVALID;
    RB_MATERIALISE(s);
    return copy_segments(s, s->dseg, s->nd, s->P.L.nd_pad, view_d(s, which), compact, nullptr);
}
int rb_get_dual(rb_solver *s, int which, double *compact) {
    if (!s || !compact || which < 0 || which > 1) return RB_ERR_INVALID;
    return copy_segments(s, s->dseg, s->nd, s->P.L.nd_pad, view_d(s, which), nullptr, compact);
}

int rb_set_initial_state(rb_solver *s, const double *x0) {
    if (!s || !x0) return RB_ERR_INVALID;
    const Layout &L = s->P.L;
    RB_CUDA(s, cudaMemcpyAsync(s->x0, x0, (size_t)L.batch * L.nx * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    // old_primal[0] = state (cache.py:81); the current iterate keeps its own x_0 like the reference
    RB_MATERIALISE(s);
    RB_CUDA(s, cudaMemcpy2DAsync(s->prim[s->old_i] + L.px, L.np_pad * sizeof(double), x0, L.nx * sizeof(double),
                                 L.nx * sizeof(double), L.batch, cudaMemcpyHostToDevice, s->stream));
    RB_CUDA(s, cudaStreamSynchronize(s->stream));
    s->have_x0 = true;
    return RB_OK;
}

int rb_update_cache(rb_solver *s) {
    if (!s) return RB_ERR_INVALID;
    if (s->collapsed) return RB_OK;
    // old <- current without moving data: the current buffer becomes the (only meaningful) old one
    std::swap(s->cur_i, s->old_i);
    s->collapsed = true;
    return RB_OK;
}

// ---- linear operator on host vectors --------------------------------------------------------------------------------
static int ensure_staging(rb_solver *s) {
    const Layout &L = s->P.L;
    if (!s->staging_p) {
        int rc = dev_zero(s, (size_t)L.batch * L.np_pad, &s->staging_p);
        if (rc != RB_OK) return rc;
        rc = dev_zero(s, (size_t)L.batch * L.nd_pad, &s->staging_d);
        if (rc != RB_OK) return rc;
    }
    return RB_OK;
}

int rb_apply_L(rb_solver *s, const double *primal_in, double *dual_out) {
    if (!s || !primal_in || !dual_out) return RB_ERR_INVALID;
    int rc = ensure_staging(s);
    if (rc != RB_OK) return rc;
    rc = copy_segments(s, s->pseg, s->np, s->P.L.np_pad, s->staging_p, primal_in, nullptr);
    if (rc != RB_OK) return rc;
    k_l_axpby<<<node_grid(s, s->P.L.n), kThreads, 0, s->stream>>>(s->P, s->staging_p, nullptr, 1.0, 0.0, nullptr,
                                                                 s->staging_d, 0.0, 1.0);
    RB_LAUNCHED(s, "k_l_axpby");
    return copy_segments(s, s->dseg, s->nd, s->P.L.nd_pad, s->staging_d, nullptr, dual_out);
}

int rb_apply_Lt(rb_solver *s, const double *dual_in, double *primal_out) {
    if (!s || !dual_in || !primal_out) return RB_ERR_INVALID;
    int rc = ensure_staging(s);
    if (rc != RB_OK) return rc;
    rc = copy_segments(s, s->dseg, s->nd, s->P.L.nd_pad, s->staging_d, dual_in, nullptr);
    if (rc != RB_OK) return rc;
    k_lt_axpby<<<node_grid(s, s->P.L.n), kThreads, 0, s->stream>>>(s->P, s->staging_d, nullptr, s->staging_p, 0.0, 1.0);
    RB_LAUNCHED(s, "k_lt_axpby");
    return copy_segments(s, s->pseg, s->np, s->P.L.np_pad, s->staging_p, nullptr, primal_out);
}

int rb_lambda_max(rb_solver *s, double *lambda_max) {
    if (!s || !lambda_max) return RB_ERR_INVALID;
    const Layout &L = s->P.L;
    // group the nonleaf nodes by the cost rows of their children
    std::map<std::vector<int>, int> groups;
    std::vector<int> grp_ptr(1, 0), grp_idx;
    for (int i = 0; i < L.m; ++i) {
        std::vector<int> key(s->cost_idx.begin() + s->child_first[i],
                             s->cost_idx.begin() + s->child_first[i] + s->child_count[i]);
        if (groups.emplace(key, (int)groups.size()).second) {
            grp_idx.insert(grp_idx.end(), key.begin(), key.end());
            grp_ptr.push_back((int)grp_idx.size());
        }
    }
    const int ng = (int)groups.size();
    const int dim = std::max(L.nx, L.nu);
    // temporaries in ONE allocation, released on every path
    struct Scratch {
        char *p = nullptr;
        ~Scratch() { if (p) cudaFree(p); }
    } scratch;
    auto up8 = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t b_ptr = up8(grp_ptr.size() * sizeof(int)), b_idx = up8(std::max<size_t>(grp_idx.size(), 1) * sizeof(int)),
                 b_work = up8((size_t)std::max(ng, s->num_leafcost) * dim * dim * sizeof(double)), b_out = 256;
    RB_CUDA(s, cudaMalloc((void **)&scratch.p, b_ptr + b_idx + b_work + b_out));
    int *d_ptr = reinterpret_cast<int *>(scratch.p), *d_idx = reinterpret_cast<int *>(scratch.p + b_ptr);
    double *d_work = reinterpret_cast<double *>(scratch.p + b_ptr + b_idx);
    double *d_out = reinterpret_cast<double *>(scratch.p + b_ptr + b_idx + b_work);
    RB_CUDA(s, cudaMemcpyAsync(d_ptr, grp_ptr.data(), grp_ptr.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    RB_CUDA(s, cudaMemcpyAsync(d_idx, grp_idx.data(), grp_idx.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    RB_CUDA(s, cudaMemsetAsync(d_out, 0, sizeof(double), s->stream));
    for (int kind = 0; kind < 3; ++kind) {
        const int count = kind == 2 ? s->num_leafcost : ng;
        k_gram_eig<<<count, 32, 0, s->stream>>>(s->P, d_ptr, d_idx, kind, count, d_work, d_out, s->status);
        RB_LAUNCHED(s, "k_gram_eig");
    }
    k_risk_block_eig<<<(L.m + 3) / 4, 128, 0, s->stream>>>(s->P, d_out);
    RB_LAUNCHED(s, "k_risk_block_eig");
    RB_CUDA(s, cudaMemcpyAsync(lambda_max, d_out, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    int st = 0;
    RB_CUDA(s, cudaMemcpyAsync(&st, s->status, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    RB_CUDA(s, cudaStreamSynchronize(s->stream));
    if (st & 8) {
        RB_CUDA(s, cudaMemsetAsync(s->status, 0, sizeof(int), s->stream));
        return fail(s, RB_ERR_NUMERIC, "step size: the Jacobi eigenvalue iteration did not converge (non-finite or "
                                       "ill-conditioned cost matrices); lambda_max(L*L) would be under-estimated");
    }
    if (!(*lambda_max == *lambda_max) || *lambda_max <= 0.0)
        return fail(s, RB_ERR_NUMERIC, "step size: lambda_max(L*L) is not a positive finite number");
    return RB_OK;
}

// ---- half steps -------------------------------------------------------------------------------------------------------
int rb_primal_half(rb_solver *s, double alpha) {
    if (!s) return RB_ERR_INVALID;
    RB_MATERIALISE(s);
    k_lt_axpby<<<node_grid(s, s->P.L.n), kThreads, 0, s->stream>>>(s->P, view_d(s, 1), view_p(s, 1), view_p(s, 0), 1.0,
                                                                  -alpha);
    RB_LAUNCHED(s, "k_lt_axpby");
    return RB_OK;
}

int rb_s0_shift(rb_solver *s, double alpha) {
    if (!s) return RB_ERR_INVALID;
    RB_MATERIALISE(s);
    k_s0_shift<<<(s->P.L.batch + 127) / 128, 128, 0, s->stream>>>(s->P, view_p(s, 0), alpha);
    RB_LAUNCHED(s, "k_s0_shift");
    return RB_OK;
}

int rb_project_dynamics(rb_solver *s) {
    if (!s) return RB_ERR_INVALID;
    int rc = need_offline(s);
    if (rc != RB_OK) return rc;
    if (!s->have_x0) return fail(s, RB_ERR_STATE, "initial state not set (cache_initial_state)");
    const Layout &L = s->P.L;
    RB_MATERIALISE(s);
    (void)L;
    int rc2 = launch_sweeps(s, nullptr, view_p(s, 0), s->stream);
    if (rc2 != RB_OK) return rc2;
    s->launches += 1 + 2 * s->plan.num_levels;
    return RB_OK;
}

int rb_project_kernel(rb_solver *s) {
    if (!s) return RB_ERR_INVALID;
    RB_MATERIALISE(s);
    k_kernel_proj<<<node_grid(s, s->P.L.m), kThreads, 0, s->stream>>>(s->P, view_p(s, 0));
    RB_LAUNCHED(s, "k_kernel_proj");
    return RB_OK;
}

int rb_prox_f(rb_solver *s, double alpha) {
    int rc = rb_s0_shift(s, alpha);
    if (rc != RB_OK) return rc;
    rc = rb_project_dynamics(s);
    if (rc != RB_OK) return rc;
    return rb_project_kernel(s);
}

int rb_dual_half(rb_solver *s, double alpha) {
    if (!s) return RB_ERR_INVALID;
    RB_MATERIALISE(s);
    k_l_axpby<<<node_grid(s, s->P.L.n), kThreads, 0, s->stream>>>(s->P, view_p(s, 0), view_p(s, 1), 2.0, -1.0, view_d(s, 1),
                                                                 view_d(s, 0), 1.0, alpha);
    RB_LAUNCHED(s, "k_l_axpby");
    return RB_OK;
}

static int prox_g_mode(rb_solver *s, double alpha, int mode) {
    if (!s) return RB_ERR_INVALID;
    RB_MATERIALISE(s);
    k_prox_g<<<node_grid(s, s->P.L.n), kThreads, 0, s->stream>>>(s->P, view_d(s, 0), alpha, mode, s->status);
    RB_LAUNCHED(s, "k_prox_g");
    return (mode & 12) ? check_status(s) : RB_OK;
}
int rb_prox_g_conj(rb_solver *s, double alpha) { return prox_g_mode(s, alpha, 31); }
int rb_modify_dual(rb_solver *s, double alpha) { return prox_g_mode(s, alpha, 1); }
int rb_add_halves(rb_solver *s) { return prox_g_mode(s, 1.0, 2); }
int rb_project_nonleaf(rb_solver *s) { return prox_g_mode(s, 1.0, 4); }
int rb_project_leaf(rb_solver *s) { return prox_g_mode(s, 1.0, 8); }

int rb_modify_projection(rb_solver *s, double alpha, const double *modified_dual) {
    if (!s || !modified_dual) return RB_ERR_INVALID;
    int rc = ensure_staging(s);
    if (rc != RB_OK) return rc;
    rc = copy_segments(s, s->dseg, s->nd, s->P.L.nd_pad, s->staging_d, modified_dual, nullptr);
    if (rc != RB_OK) return rc;
    const long long count = (long long)s->P.L.batch * s->P.L.nd_pad;
    RB_MATERIALISE(s);
    // d = alpha * (w - d)   (cache.py:392-393)
    k_axpby<<<1184, 256, 0, s->stream>>>(view_d(s, 0), alpha, s->staging_d, -alpha, view_d(s, 0), count);
    RB_LAUNCHED(s, "k_axpby");
    return RB_OK;
}

// ---- residuals, reference formulas with stand-alone kernels ------------------------------------------------------------
int rb_residuals(rb_solver *s, double alpha, double *norms, double *vectors) {
    if (!s || !norms) return RB_ERR_INVALID;
    const Layout &L = s->P.L;
    const long long pc = (long long)L.batch * L.np_pad, dc = (long long)L.batch * L.nd_pad;
    for (int i = 0; i < 6; ++i)
        if (!s->tp[i]) {
            int rc = dev_zero(s, (size_t)pc, &s->tp[i]);
            if (rc != RB_OK) return rc;
        }
    for (int i = 0; i < 4; ++i)
        if (!s->td[i]) {
            int rc = dev_zero(s, (size_t)dc, &s->td[i]);
            if (rc != RB_OK) return rc;
        }
    double *dp = s->tp[0], *lt = s->tp[1], *xi1 = s->tp[2], *pn = s->tp[3], *xi0 = s->tp[4], *delta0 = s->tp[5];
    double *dd = s->td[0], *lp = s->td[1], *xi2 = s->td[2], *delta2 = s->td[3];
    const dim3 g = node_grid(s, L.n);
    cudaStream_t st = s->stream;
    double *p_cur = view_p(s, 0), *p_old = view_p(s, 1), *d_cur = view_d(s, 0), *d_old = view_d(s, 1);
    k_axpby<<<1184, 256, 0, st>>>(dp, 1.0, p_old, -1.0, p_cur, pc);                      // p - p_new
    k_axpby<<<1184, 256, 0, st>>>(dd, 1.0, d_old, -1.0, d_cur, dc);                      // d - d_new
    k_lt_axpby<<<g, kThreads, 0, st>>>(s->P, dd, nullptr, lt, 0.0, 1.0);                 // L*(d - d_new)
    k_div_add<<<1184, 256, 0, st>>>(xi1, dp, alpha, -1.0, lt, pc);                       // xi1
    k_axpby<<<1184, 256, 0, st>>>(pn, 1.0, p_cur, -1.0, p_old, pc);                      // p_new - p = delta1
    k_l_axpby<<<g, kThreads, 0, st>>>(s->P, pn, nullptr, 1.0, 0.0, nullptr, lp, 0.0, 1.0);  // L(p_new - p)
    k_div_add<<<1184, 256, 0, st>>>(xi2, dd, alpha, 1.0, lp, dc);                        // xi2
    k_lt_axpby<<<g, kThreads, 0, st>>>(s->P, xi2, nullptr, lt, 0.0, 1.0);                // L* xi2
    k_axpby<<<1184, 256, 0, st>>>(xi0, 1.0, xi1, 1.0, lt, pc);                           // xi0
    k_axpby<<<1184, 256, 0, st>>>(delta2, 1.0, d_cur, -1.0, d_old, dc);                  // delta2
    k_lt_axpby<<<g, kThreads, 0, st>>>(s->P, delta2, nullptr, lt, 0.0, 1.0);             // L* delta2
    k_axpby<<<1184, 256, 0, st>>>(delta0, 1.0, pn, -1.0, lt, pc);                        // delta0
    s->launches += 12;
    int rc = launch_ok(s, "residual kernels");
    if (rc != RB_OK) return rc;
    RB_CUDA(s, cudaMemsetAsync(s->slots, 0, (size_t)L.batch * 6 * sizeof(double), st));
    const double *vecs[6] = {xi0, xi1, xi2, delta0, pn, delta2};
    for (int i = 0; i < 6; ++i) {
        const bool is_p = (i != 2 && i != 5);
        k_absmax<<<dim3(148, L.batch), 256, 0, st>>>(vecs[i], is_p ? L.np_pad : L.nd_pad, s->slots + i, 6, s->status);
        RB_LAUNCHED(s, "k_absmax");
    }
    RB_CUDA(s, cudaMemcpyAsync(norms, s->slots, (size_t)L.batch * 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB_CUDA(s, cudaMemsetAsync(s->slots, 0, (size_t)L.batch * 6 * sizeof(double), st));
    RB_CUDA(s, cudaStreamSynchronize(st));
    if (vectors) {
        // per instance: [xi0 | xi1 | xi2 | delta0 | delta1 | delta2], compact
        const int64_t per = 4 * s->np + 2 * s->nd;
        int64_t off = 0;
        for (int i = 0; i < 6; ++i) {
            const bool is_p = (i != 2 && i != 5);
            std::vector<rb_solver::Seg> segs = is_p ? s->pseg : s->dseg;
            for (auto &sg : segs) sg.compact += off;
            rc = copy_segments(s, segs, per, is_p ? L.np_pad : L.nd_pad, const_cast<double *>(vecs[i]), nullptr, vectors);
            if (rc != RB_OK) return rc;
            off += is_p ? s->np : s->nd;
        }
    }
    return RB_OK;
}

// ---- fused loop -------------------------------------------------------------------------------------------------------
namespace {

// the kernels of one iteration reading buffer src and writing buffer 1-src.  have_pbar: prim[1-src] already holds
// pbar = p - alpha L* d of this iteration (written by the previous dual pass)
void launch_panel_sweeps(rb_solver *s, cudaStream_t st, double *prim, cudaEvent_t mid = nullptr) {
    const Layout &L = s->P.L;
    const int top = panel_top(s);
    for (int t = L.num_stages - 2; t >= top; --t)
        launch_bp_bwd(st, s->P, s->ctrl, prim, s->pq, s->pr, s->stage_off[t], s->stage_off[t + 1] - s->stage_off[t]);
    if (mid) cudaEventRecord(mid, st);
    if (top > 0) launch_bp_top(st, s->P, s->ctrl, prim, s->pq, s->pr, s->x0, s->plan.stage_off, top);
    for (int t = top; t <= L.num_stages - 2; ++t)
        launch_bp_fwd(st, s->P, s->ctrl, prim, s->pr, s->x0, s->stage_off[t], s->stage_off[t + 1] - s->stage_off[t]);
}

// the batch-innermost path: always pipelined (the dual pass leaves pbar of the next iteration in the old primal buffer)
int enqueue_iteration_panel(rb_solver *s, int src, cudaStream_t st, bool have_pbar) {
    const Layout &L = s->P.L;
    const int dst = 1 - src;
    if (!have_pbar) launch_bp_primal(st, s->P, s->ctrl, s->pprim[src], s->pdual[src], s->pprim[dst]);
    launch_bp_kproj(st, s->P, s->ctrl, s->pprim[dst], s->x0, s->pprim[src]);
    launch_panel_sweeps(s, st, s->pprim[dst]);
    launch_bp_dual(st, s->P, s->ctrl, s->pprim[src], s->pprim[dst], s->pdual[src], s->pdual[dst], s->slots, s->pprim[src], s->pc2);
    launch_check(st, s->P, s->ctrl, s->slots, s->last, s->h_last_dev);
    return launch_ok(s, "panel iteration");
}

// the stopping test rides on the last CTA of the dual passes (batch 1, pipelined single-GPU loop on the lane kernels) once the
// number of those CTAs is known (counted by the first enqueue / capture) and the running loop's control block carries it
bool fused_check_on(const rb_solver *s) {
    return s->fused_check && s->loop_armed && s->arr_expected > 0 && s->P.L.batch == 1 && !s->sharded && !s->panel_live && use_pipe(s);
}

int enqueue_iteration_kernels(rb_solver *s, int src, cudaStream_t st, bool have_pbar) {
    const Layout &L = s->P.L;
    const int dst = 1 - src;
    if (s->panel_live) return enqueue_iteration_panel(s, src, st, have_pbar);
    if (!use_pipe(s)) {
        launch_primal(s, st, src, dst);
        int rc = launch_sweeps(s, s->ctrl, s->prim[dst], st);
        if (rc != RB_OK) return rc;
        launch_dual(s, st, src, dst);
        launch_check(st, s->P, s->ctrl, s->slots, s->last, s->h_last_dev);
        return launch_ok(s, "fused iteration");
    }
    const PipeSplit ps = pipe_split(s);
    const dim3 nb(1, L.batch);
    cudaStream_t s0 = s->side[0], s1 = s->side[1];
    cudaEvent_t *ev = s->pev;
    // stopping test by the last CTA of the dual passes instead of a launch of its own (lane.cu iteration_arrive): the CTAs of this
    // iteration's dual-pass launches are counted here and must add up to what the control block expects
    const bool arrive = have_pbar && fused_check_on(s);
    int ctas = 0;
    auto dual_lane = [&](cudaStream_t q, int first, int count, bool narrow = false) {
        ctas += launch_dual_lane(nb, q, s->P, s->ctrl, s->prim[src], s->prim[dst], s->dual[src], s->dual[dst], s->slots, nullptr,
                                 first, count, s->prim[src], narrow, arrive);
    };
    // side stream 0, under the backward sweeps: the kernel projection in place (or, in the first iteration of a loop, after
    // the stand-alone primal pass on the main stream), then the risk block of the chain nodes' dual pass, which needs y, s only
    const bool risk_split = ps.cf < L.m && s->risk_split;
    if (!have_pbar) launch_primal(s, st, src, dst);
    const bool side0 = have_pbar || risk_split;
    if (side0) {
        RB_CUDA(s, cudaEventRecord(ev[0], st));
        RB_CUDA(s, cudaStreamWaitEvent(s0, ev[0], 0));
        if (s->allow_table_prefetch && s->d_tables && !s->h_tables.empty())
            k_prefetch_ranges<<<32, 256, 0, s0>>>(s->ctrl, s->d_tables, (int)s->h_tables.size());
        if (have_pbar) launch_kproj(L.batch, s0, s->P, s->ctrl, s->prim[dst], s->h_last_dev ? s->x0 : nullptr, s->prim[src]);
        if (risk_split)
            ctas += launch_dual_risk_chain(L.batch, s0, s->P, s->ctrl, s->prim[src], s->prim[dst], s->dual[src], s->dual[dst],
                                           s->slots, ps.cf, L.m - ps.cf, s->chain_stride, s->chain_yo0, s->prim[src],
                                           OwnMap{0, 0, 0}, arrive);
        RB_CUDA(s, cudaEventRecord(ev[1], s0));
    }
    bool early_done = false;
    cudaError_t herr = cudaSuccess;
    auto after_top = [&]() {   // the top of the tree is final: its dual pass runs next to the forward levels
        if (ps.early <= 0) return;
        if ((herr = cudaEventRecord(ev[2], st)) != cudaSuccess) return;
        if ((herr = cudaStreamWaitEvent(s1, ev[2], 0)) != cudaSuccess) return;
        if (side0 && (herr = cudaStreamWaitEvent(s1, ev[1], 0)) != cudaSuccess) return;
        dual_lane(s1, 0, ps.early, true);   // next to the forward chain walker: as few CTAs as possible
        early_done = true;
    };
    auto dual_chain = [&](cudaStream_t q, int first, int count) {
        ctas += launch_dual_chain(L.batch, q, s->P, s->ctrl, s->prim[src], s->prim[dst], s->dual[src], s->dual[dst], s->slots,
                                  s->chain_recs + (first - ps.cf), first, count, s->chain_stride,
                                  s->chain_yo0 + 3 * (first - ps.cf), s->prim[src], risk_split ? 0 : 1, OwnMap{0, 0, 0}, arrive);
    };
    bool piece_done = false;
    auto after_piece = [&]() {   // the first piece of the forward chain walk is done: its nodes' dual pass starts
        if (herr != cudaSuccess) return;
        if ((herr = cudaEventRecord(ev[5], st)) != cudaSuccess) return;
        if ((herr = cudaStreamWaitEvent(s1, ev[5], 0)) != cudaSuccess) return;
        if (side0 && (herr = cudaStreamWaitEvent(s1, ev[1], 0)) != cudaSuccess) return;
        dual_chain(s1, ps.cf, ps.mid - ps.cf);
        piece_done = true;
    };
    int rc = launch_sweeps(s, s->ctrl, s->prim[dst], st, nullptr, nullptr, after_top, ps.split, after_piece);
    if (rc != RB_OK) return rc;
    if (herr != cudaSuccess) return fail(s, RB_ERR_CUDA, std::string("pipelined iteration: ") + cudaGetErrorString(herr));
    if (side0) RB_CUDA(s, cudaStreamWaitEvent(st, ev[1], 0));
    const int early = early_done ? ps.early : 0;
    const int chain_lo = piece_done ? ps.mid : ps.cf;
    if (ps.cf >= L.m) {   // no chain pass: one general pass over the rest
        dual_lane(st, early, L.n - early);
    } else {
        RB_CUDA(s, cudaEventRecord(ev[3], st));
        RB_CUDA(s, cudaStreamWaitEvent(s1, ev[3], 0));
        dual_lane(s1, early, ps.cf - early);
        dual_lane(s1, L.m, L.n - L.m);
        dual_chain(st, chain_lo, L.m - chain_lo);
        early_done = true;   // s1 carries work that st has to wait for
    }
    if (early_done) {
        RB_CUDA(s, cudaEventRecord(ev[4], s1));
        RB_CUDA(s, cudaStreamWaitEvent(st, ev[4], 0));
    }
    if (arrive) {
        if (ctas != s->arr_expected)
            return fail(s, RB_ERR_STATE, "pipelined iteration: the dual passes launched " + std::to_string(ctas) + " CTAs, the control "
                        "block expects " + std::to_string(s->arr_expected));
    } else {
        launch_check(st, s->P, s->ctrl, s->slots, s->last, s->h_last_dev);
        if (have_pbar) s->arr_counted = ctas;   // what an armed loop will expect (rb_loop_begin / build_graphs)
    }
    return launch_ok(s, "pipelined iteration");
}

// the gather step of the sharded loop: q_j and d2_j of the cut nodes and the residual maxima of the previous iteration
// cross NVLink, then the stopping test of the previous iteration runs (identically on every rank)
int shard_exchange(rb_solver *s, int src, cudaStream_t st, double *aux = nullptr, double *sl = nullptr) {
    // aux: the per-cut-node scalar that travels with q_j; default d2_j of the old dual (unpipelined loop), else sbar_j of pbar
    // sl: the residual maxima that travel (and are tested); default sl, the pipelined loop passes the previous parity
    if (!aux) aux = s->dual[src] + s->P.L.d2;
    if (!sl) sl = s->slots;
    if (s->p2p && s->P.L.batch == 1 && s->xchg_fused) {   // push, pull and the stopping test in one launch
        launch_shard_xchg(st, s->P, s->ctrl, s->shard, s->q, aux, sl, s->px, s->last, s->h_last_dev, s->shard_pending);
        s->launches += 1;
        s->shard_pending = false;
        return launch_ok(s, "shard exchange");
    } else if (s->p2p) {
        launch_shard_push(st, s->P, s->ctrl, s->shard, s->q, aux, sl, s->px);
        launch_shard_pull(st, s->P, s->ctrl, s->shard, s->q, aux, sl, s->px);
        s->launches += 3;
    } else {
        k_shard_pack<<<8, 256, 0, st>>>(s->P, s->ctrl, s->shard, s->q, aux, sl, s->xchg_send);
        const int rc = nccl_all_gather_f64(s->xchg_send, s->xchg_recv, s->xchg_count, s->nccl_comm, st);
        if (rc != 0) return fail(s, RB_ERR_CUDA, std::string("ncclAllGather: ") + nccl_error(rc));
        k_shard_unpack<<<8, 256, 0, st>>>(s->P, s->ctrl, s->shard, s->xchg_recv, s->q, aux, sl);
        s->launches += 2;
    }
    if (s->shard_pending) launch_check(st, s->P, s->ctrl, sl, s->last, s->h_last_dev);
    s->shard_pending = false;
    return launch_ok(s, "shard exchange");
}

// one iteration of the subtree-sharded loop (shard.cu): owned primal pass and backward sweep, ONE all-gather, replicated
// top, owned forward sweep, dual pass on owned + top nodes
int enqueue_iteration_sharded(rb_solver *s, int src, cudaStream_t st) {
    if (!s->nccl_comm) return fail(s, RB_ERR_STATE, "rb_shard_init() has not been called");
    if (!use_lane(s)) return fail(s, RB_ERR_INVALID, "subtree sharding currently needs the one-thread-per-node passes "
                                                     "(diagonal cost roots, even nx / nu, <= 8 children per node)");
    const SweepPlan &pl = s->plan;
    const Layout &L = s->P.L;
    const int dst = 1 - src;
    const size_t per_warp = (size_t)(2 * L.nxu + 32) * sizeof(double);
    auto grid = [&](const SweepLevel &lv) { return dim3((lv.num_sub + lv.subs_per_cta - 1) / lv.subs_per_cta, 1); };
    auto threads = [&](const SweepLevel &lv) { return 32 * lv.warps_per_sub * lv.subs_per_cta; };
    auto smem = [&](const SweepLevel &lv) {
        return per_warp * lv.warps_per_sub * lv.subs_per_cta + (size_t)lv.subs_per_cta * lv.stage_cap * L.nxu * sizeof(double);
    };
    launch_primal(s, st, src, dst, s->owned_nodes, s->n_owned);
    for (int v = pl.num_levels - 1; v >= 0; --v)
        if (s->allow_mma && s->shard_lv[v].num_tiles > 0)
            launch_chain_mma_bwd(st, s->P, s->ctrl, s->shard_lv[v], s->prim[dst], s->q, s->r, mma_four_warps(s));
        else if (s->tree_mode > 0 && s->shard_tree_lv[v].desc)
            launch_tree_bwd(dim3(s->shard_tree_lv[v].num_sub, 1), 32 * s->shard_tree_lv[v].warps, s->tree_smem[1 + v], st, s->P,
                            s->ctrl, s->shard_tree_lv[v], s->prim[dst], s->q, s->r);
        else
            launch_sweep_sub_bwd(grid(s->shard_lv[v]), threads(s->shard_lv[v]), smem(s->shard_lv[v]), st, s->P, s->ctrl,
                                 s->shard_lv[v], s->prim[dst], s->q, s->r);
    int rc = shard_exchange(s, src, st);
    if (rc != RB_OK) return rc;
    launch_primal(s, st, src, dst, s->top_nodes, s->n_top);
    if (s->tree_mode > 0 && s->tree_top.desc)
        launch_tree_top(1, 32 * s->tree_top.warps, s->tree_smem[0], st, s->P, s->ctrl, s->tree_top, s->prim[dst], s->q, s->r,
                        s->x0);
    else
        launch_sweep_top(1, 32 * s->top_warps, per_warp * s->top_warps + (size_t)pl.top_cap * L.nxu * sizeof(double), st, s->P,
                         s->ctrl, pl, s->prim[dst], s->q, s->r, s->x0);
    for (int v = 0; v < pl.num_levels; ++v)
        if (s->allow_mma && s->shard_lv[v].num_tiles > 0)
            launch_chain_mma_fwd(st, s->P, s->ctrl, s->shard_lv[v], s->prim[dst], s->r, 0, -1, mma_four_warps(s));
        else if (s->tree_mode > 0 && s->shard_tree_lv[v].desc)
            launch_tree_fwd(dim3(s->shard_tree_lv[v].num_sub, 1), 32 * s->shard_tree_lv[v].warps, s->tree_smem[1 + v], st, s->P,
                            s->ctrl, s->shard_tree_lv[v], s->prim[dst], s->r);
        else
            launch_sweep_sub_fwd(grid(s->shard_lv[v]), threads(s->shard_lv[v]), smem(s->shard_lv[v]), st, s->P, s->ctrl,
                                 s->shard_lv[v], s->prim[dst], s->r);
    launch_dual(s, st, src, dst, s->dual_nodes, s->n_top + s->n_owned);
    s->shard_pending = true;
    return launch_ok(s, "sharded iteration");
}

// The pipelined loop under subtree sharding (needs the peer-memory exchange, so that the whole iteration is kernels and fits a
// CUDA graph): owned kernel projection next to the owned backward sweeps -> push / pull of q_j, sbar_j and the residual maxima
// (+ the stopping test of the previous iteration) -> replicated top (kernel projection, sweep, dual pass) -> owned forward
// sweeps -> owned dual passes, which leave pbar of the next iteration.  The risk block of the chain nodes adds to the residual
// maxima of THIS iteration, so it may only start once the exchange has collected those of the previous one.
// developer aid (RB_SHARD_TIMING=1): CUDA-event stamps on the main stream of the pipelined sharded iteration, plain launches
// instead of graphs, averages printed per rank when the loop ends
struct ShardTiming {
    static constexpr int kRing = 8, kPts = 8;
    cudaEvent_t ev[kRing][kPts] = {};
    int used[kRing] = {};
    const char *label[kPts] = {};
    double sum[kPts] = {};
    long long n = 0, it = 0;
    int slot = 0, pt = 0;
    bool on = false;
    void begin() {
        slot = (int)(it % kRing);
        if (used[slot] > 1) collect(slot);
        pt = 0;
    }
    void stamp(cudaStream_t st, const char *what) {
        if (pt >= kPts) return;
        if (!ev[slot][pt]) cudaEventCreate(&ev[slot][pt]);
        cudaEventRecord(ev[slot][pt], st);
        label[pt] = what;
        ++pt;
    }
    void end() {
        used[slot] = pt;
        ++it;
    }
    void collect(int sl) {
        cudaEventSynchronize(ev[sl][used[sl] - 1]);
        if (it > 24) {   // past the warm-up
            for (int i = 1; i < used[sl]; ++i) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev[sl][i - 1], ev[sl][i]);
                sum[i] += ms * 1e3;
            }
            ++n;
        }
        used[sl] = 0;
    }
    void report(int rank) {
        for (int sl = 0; sl < kRing; ++sl)
            if (used[sl] > 1) collect(sl);
        if (n == 0) return;
        std::string line = "[shard timing, rank " + std::to_string(rank) + ", " + std::to_string(n) + " iterations] us:";
        double tot = 0;
        for (int i = 1; i < kPts && label[i]; ++i) {
            char buf[96];
            snprintf(buf, sizeof buf, " %s %.1f |", label[i], sum[i] / n);
            line += buf;
            tot += sum[i] / n;
        }
        fprintf(stderr, "%s total %.1f\n", line.c_str(), tot);
        for (auto &v : sum) v = 0;
        n = 0;
        it = 0;
    }
};
ShardTiming g_shard_timing;
bool shard_timing_on() {
    static const bool on = getenv("RB_SHARD_TIMING") != nullptr;
    return on;
}

bool shard_pipe(const rb_solver *s) {
    return s->sharded && s->p2p && use_lane(s) && s->allow_pipe && s->tree_mode > 0 && s->tree_top.desc != nullptr;
}
int enqueue_iteration_sharded_pipe(rb_solver *s, int src, cudaStream_t st, bool have_pbar) {
    const SweepPlan &pl = s->plan;
    const Layout &L = s->P.L;
    const int dst = 1 - src;
    const bool w4 = mma_four_warps(s);
    const size_t per_warp = (size_t)(2 * L.nxu + 32) * sizeof(double);
    auto grid = [&](const SweepLevel &lv) { return dim3((lv.num_sub + lv.subs_per_cta - 1) / lv.subs_per_cta, 1); };
    auto threads = [&](const SweepLevel &lv) { return 32 * lv.warps_per_sub * lv.subs_per_cta; };
    auto smem = [&](const SweepLevel &lv) {
        return per_warp * lv.warps_per_sub * lv.subs_per_cta + (size_t)lv.subs_per_cta * lv.stage_cap * L.nxu * sizeof(double);
    };
    cudaStream_t s0 = s->side[0], s1 = s->side[1];
    // residual maxima, double-buffered by the parity of the iteration: this iteration's dual passes fold into `sw` while the
    // exchange sends and tests `stt`, the previous iteration's -- so the risk block of the chain nodes need not wait for the exchange
    double *sw = s->slots + 6 * src, *stt = s->slots + 6 * (1 - src);
    cudaEvent_t *ev = s->pev;
    const int cf = s->n_own_chain > 0 ? s->chain_first : L.m;
    ShardTiming *tm = shard_timing_on() && have_pbar ? &g_shard_timing : nullptr;
    if (tm) {   // not while the iteration is being captured into a graph
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) tm = nullptr;
    }
    if (tm) {
        tm->begin();
        tm->stamp(st, "start");
    }
    // the exchange inside the top-of-the-tree kernel (k_tree_top<.., SHARD>): packets instead of buffers + flags, no launch of its own
    const bool in_top = have_pbar && s->xchg_fused && s->xchg_in_top;
    if (!have_pbar) launch_primal(s, st, src, dst, s->owned_nodes, s->n_owned);   // incl. the kernel projection of the owned nodes
    RB_CUDA(s, cudaEventRecord(ev[0], st));
    RB_CUDA(s, cudaStreamWaitEvent(s0, ev[0], 0));
    if (have_pbar) launch_kproj(1, s0, s->P, s->ctrl, s->prim[dst], nullptr, nullptr, s->own_nonleaf, s->n_own_nonleaf);
    // the risk block of the owned chain nodes (y, s are final after the owned kernel projection) right behind that projection, under
    // the backward walker as in the single-GPU loop: the residual maxima are double-buffered, so it need not wait for the exchange
    const bool risk_early = have_pbar && cf < L.m && s->risk_split;
    if (risk_early)
        launch_dual_risk_chain(1, s0, s->P, s->ctrl, s->prim[src], s->prim[dst], s->dual[src], s->dual[dst], sw, cf, s->n_own_chain,
                               s->chain_stride, s->chain_yo0, s->prim[src], s->own_chain);
    for (int v = pl.num_levels - 1; v >= 0; --v)
        if (s->allow_mma && s->shard_lv[v].num_tiles > 0)
            launch_chain_mma_bwd(st, s->P, s->ctrl, s->shard_lv[v], s->prim[dst], s->q, s->r, w4);
        else if (s->shard_tree_lv[v].desc)
            launch_tree_bwd(dim3(s->shard_tree_lv[v].num_sub, 1), 32 * s->shard_tree_lv[v].warps, s->tree_smem[1 + v], st, s->P,
                            s->ctrl, s->shard_tree_lv[v], s->prim[dst], s->q, s->r);
        else
            launch_sweep_sub_bwd(grid(s->shard_lv[v]), threads(s->shard_lv[v]), smem(s->shard_lv[v]), st, s->P, s->ctrl,
                                 s->shard_lv[v], s->prim[dst], s->q, s->r);
    if (tm) tm->stamp(st, "chain + tree bwd");
    if (!in_top) {
        int rc = shard_exchange(s, src, st, have_pbar ? s->prim[dst] + L.ps : nullptr, stt);
        if (rc != RB_OK) return rc;
        if (tm) tm->stamp(st, "exchange");
    }
    RB_CUDA(s, cudaEventRecord(ev[1], st));          // (not in_top) the maxima of the previous iteration are collected and tested
    RB_CUDA(s, cudaStreamWaitEvent(s0, ev[1], 0));
    if (!risk_early && cf < L.m && s->risk_split)
        launch_dual_risk_chain(1, s0, s->P, s->ctrl, s->prim[src], s->prim[dst], s->dual[src], s->dual[dst], sw, cf,
                               s->n_own_chain, s->chain_stride, s->chain_yo0, s->prim[src], s->own_chain);
    RB_CUDA(s, cudaEventRecord(ev[2], s0));
    // the kernel projection of the top touches y, tau, s only (and x_0 of the OLD iterate): it runs next to the top's sweep, which
    // reads xbar, ubar and writes x, u -- both are needed only by the top's dual pass
    if (in_top) {   // exchange + stopping test + sweep of the top in one launch; sbar_j of the peers' cut nodes lands in the iterate,
                    // so the top's kernel projection follows it (next to the forward sweeps)
        ShardHand sh;
        sh.sp = s->shard;
        sh.px = s->px;
        sh.aux = s->prim[dst] + L.ps;
        sh.slots = stt;
        sh.last = s->last;
        sh.host_last = s->h_last_dev;
        sh.check = s->shard_pending ? 1 : 0;
        launch_tree_top_sharded(32 * s->tree_top.warps, s->tree_smem[0], st, s->P, s->ctrl, s->tree_top, s->prim[dst], s->q, s->r,
                                s->x0, sh);
        s->shard_pending = false;
        if (tm) tm->stamp(st, "exchange + top");
        RB_CUDA(s, cudaEventRecord(ev[3], st));
        RB_CUDA(s, cudaStreamWaitEvent(s1, ev[3], 0));
        launch_kproj(1, s1, s->P, s->ctrl, s->prim[dst], s->h_last_dev ? s->x0 : nullptr, s->prim[src], s->top_nonleaf,
                     s->n_top_nonleaf);
        RB_CUDA(s, cudaEventRecord(ev[6], s1));
    } else {
        if (have_pbar) {
            RB_CUDA(s, cudaStreamWaitEvent(s1, ev[1], 0));
            launch_kproj(1, s1, s->P, s->ctrl, s->prim[dst], s->h_last_dev ? s->x0 : nullptr, s->prim[src], s->top_nonleaf,
                         s->n_top_nonleaf);
            RB_CUDA(s, cudaEventRecord(ev[6], s1));
        } else {
            launch_primal(s, st, src, dst, s->top_nodes, s->n_top);
        }
        launch_tree_top(1, 32 * s->tree_top.warps, s->tree_smem[0], st, s->P, s->ctrl, s->tree_top, s->prim[dst], s->q, s->r, s->x0);
        if (tm) tm->stamp(st, "top");
        RB_CUDA(s, cudaEventRecord(ev[3], st));
        RB_CUDA(s, cudaStreamWaitEvent(s1, ev[3], 0));   // the top is final: its dual pass runs next to the forward sweeps
    }
    launch_dual_lane(dim3(1, 1), s1, s->P, s->ctrl, s->prim[src], s->prim[dst], s->dual[src], s->dual[dst], sw, s->own_lane, 0,
                     s->n_top, s->prim[src], true);
    for (int v = 0; v < pl.num_levels; ++v)
        if (s->allow_mma && s->shard_lv[v].num_tiles > 0)
            launch_chain_mma_fwd(st, s->P, s->ctrl, s->shard_lv[v], s->prim[dst], s->r, 0, -1, w4);
        else if (s->shard_tree_lv[v].desc)
            launch_tree_fwd(dim3(s->shard_tree_lv[v].num_sub, 1), 32 * s->shard_tree_lv[v].warps, s->tree_smem[1 + v], st, s->P,
                            s->ctrl, s->shard_tree_lv[v], s->prim[dst], s->r);
        else
            launch_sweep_sub_fwd(grid(s->shard_lv[v]), threads(s->shard_lv[v]), smem(s->shard_lv[v]), st, s->P, s->ctrl,
                                 s->shard_lv[v], s->prim[dst], s->r);
    if (tm) tm->stamp(st, "tree + chain fwd");
    RB_CUDA(s, cudaStreamWaitEvent(st, ev[2], 0));   // owned kernel projection (+ risk block) done
    if (have_pbar) RB_CUDA(s, cudaStreamWaitEvent(st, ev[6], 0));   // sbar of the cut nodes is projected
    RB_CUDA(s, cudaEventRecord(ev[4], st));
    RB_CUDA(s, cudaStreamWaitEvent(s1, ev[4], 0));
    launch_dual_lane(dim3(1, 1), s1, s->P, s->ctrl, s->prim[src], s->prim[dst], s->dual[src], s->dual[dst], sw,
                     s->own_lane + s->n_top, 0, s->n_own_lane - s->n_top, s->prim[src]);
    if (cf < L.m)
        launch_dual_chain(1, st, s->P, s->ctrl, s->prim[src], s->prim[dst], s->dual[src], s->dual[dst], sw, s->chain_recs, cf,
                          s->n_own_chain, s->chain_stride, s->chain_yo0, s->prim[src], s->risk_split ? 0 : 1, s->own_chain);
    if (tm) tm->stamp(st, "dual chain");
    RB_CUDA(s, cudaEventRecord(ev[5], s1));
    RB_CUDA(s, cudaStreamWaitEvent(st, ev[5], 0));
    if (tm) {
        tm->stamp(st, "join");
        tm->end();
    }
    s->shard_pending = true;
    return launch_ok(s, "pipelined sharded iteration");
}

// capture one iteration per buffer parity into a CUDA graph (the kernel arguments never change afterwards: step size
// and stopping parameters live in the device control block)
int build_graphs(rb_solver *s) {
    // the stopping test by the last CTA of the dual passes needs their CTA count before the iteration is captured: one throw-away
    // capture counts them (enqueue_iteration_kernels, unarmed), the real captures then carry the `arrive` arguments
    const bool want_arrive = s->fused_check && s->P.L.batch == 1 && !s->sharded && !s->panel_live && use_pipe(s);
    if (want_arrive && s->arr_expected == 0 && !s->graph[0] && !s->graph[1]) {
        cudaStream_t cap = nullptr;
        RB_CUDA(s, cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
        RB_CUDA(s, cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
        s->loop_armed = false;
        s->arr_counted = 0;
        const int rc = enqueue_iteration_kernels(s, 0, cap, true);
        cudaGraph_t g = nullptr;
        const cudaError_t e = cudaStreamEndCapture(cap, &g);
        cudaStreamDestroy(cap);
        if (g) cudaGraphDestroy(g);
        if (rc != RB_OK) return rc;
        if (e != cudaSuccess) return fail(s, RB_ERR_CUDA, std::string("graph capture (counting pass): ") + cudaGetErrorString(e));
        s->arr_expected = s->arr_counted;
    }
    s->loop_armed = want_arrive && s->arr_expected > 0;   // (rb_loop_begin sets it again, the same way, for the loop itself)
    for (int src = 0; src < 2; ++src) {
        if (s->graph[src]) continue;
        cudaStream_t cap = nullptr;
        RB_CUDA(s, cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
        RB_CUDA(s, cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
        int rc;
        if (shard_pipe(s)) {   // steady state: pbar is there and the previous iteration is waiting for its stopping test
            s->shard_pending = true;
            rc = enqueue_iteration_sharded_pipe(s, src, cap, true);
        } else {
            rc = enqueue_iteration_kernels(s, src, cap, true);
        }
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamEndCapture(cap, &g);
        cudaStreamDestroy(cap);
        if (rc != RB_OK) return rc;
        if (e != cudaSuccess) return fail(s, RB_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
        e = cudaGraphInstantiate(&s->graph[src], g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return fail(s, RB_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(e));
    }
    return RB_OK;
}

}  // namespace

int rb_loop_begin(rb_solver *s, double alpha, int32_t max_iters, double tol, int32_t hist_capacity) {
    if (!s || max_iters < 0 || hist_capacity < 0) return RB_ERR_INVALID;
    int rc = need_offline(s);
    if (rc != RB_OK) return rc;
    if (!s->have_x0) return fail(s, RB_ERR_STATE, "initial state not set (cache_initial_state)");
    const Layout &L = s->P.L;
    cudaStream_t st = s->stream;
    const bool panel = use_panel(s);
    if (panel != s->panel_live)   // the graphs were captured for the other layout
        drop_graphs(s);
    s->panel_live = panel;
    if (panel && !s->pq) {
        for (int w = 0; w < 2; ++w) {
            rc = dev_zero(s, batch_panel_doubles(L.np_pad, L.batch), &s->pprim[w]);
            if (rc != RB_OK) return rc;
            rc = dev_zero(s, batch_panel_doubles(L.nd_pad, L.batch), &s->pdual[w]);
            if (rc != RB_OK) return rc;
        }
        rc = dev_zero(s, batch_panel_doubles((long long)L.n * L.nx, L.batch), &s->pq);
        if (rc != RB_OK) return rc;
        rc = dev_zero(s, batch_panel_doubles((long long)L.m * L.nu, L.batch), &s->pr);
        if (rc != RB_OK) return rc;
        rc = dev_zero(s, (size_t)L.m * L.nxu, &s->pc2);
        if (rc != RB_OK) return rc;
        launch_bp_c2(st, s->P, s->pc2);
        RB_LAUNCHED(s, "k_bp_c2");
    }
    if (s->use_graphs && (!s->sharded || shard_pipe(s))) {
        const int64_t before = s->launches;
        rc = build_graphs(s);
        s->launches = before;   // captured, not launched
        if (rc != RB_OK) return rc;
    }
    s->shard_pending = false;
    s->mirror_on = false;
    s->pbar_ready = false;
    if (s->hist && s->hist_capacity < hist_capacity) {
        RB_CUDA(s, cudaStreamSynchronize(st));
        cudaFree(s->hist);
        s->hist = nullptr;
    }
    if (hist_capacity > 0 && !s->hist) {
        RB_CUDA(s, cudaMalloc((void **)&s->hist, (size_t)hist_capacity * L.batch * 6 * sizeof(double)));
        s->hist_capacity = hist_capacity;
    }
    if (!s->h_pinned) RB_CUDA(s, cudaMallocHost((void **)&s->h_pinned, 4096));
    Ctrl *hc = reinterpret_cast<Ctrl *>(s->h_pinned);
    std::memset(hc, 0, sizeof(Ctrl));
    hc->max_iters = max_iters;
    hc->tol = tol;
    hc->alpha = alpha;
    hc->hist = hist_capacity > 0 ? s->hist : nullptr;
    hc->hist_capacity = hist_capacity;
    if (s->arr_expected == 0 && s->arr_counted > 0) s->arr_expected = s->arr_counted;   // counted by an earlier loop's plain launches
    s->loop_armed = s->fused_check && s->arr_expected > 0 && L.batch == 1 && !s->sharded && !s->panel_live && use_pipe(s);
    hc->arr_expected = s->loop_armed ? s->arr_expected : 0;
    hc->last = s->last;
    hc->host_last = s->h_last_dev;
    RB_CUDA(s, cudaMemcpyAsync(s->ctrl, hc, sizeof(Ctrl), cudaMemcpyHostToDevice, st));
    RB_CUDA(s, cudaMemsetAsync(s->slots, 0, (size_t)L.batch * 6 * 2 * sizeof(double), st));
    if (s->overlap_sync) RB_CUDA(s, cudaMemsetAsync(s->overlap_sync, 0, (size_t)2 * L.batch * sizeof(int), st));
    // Solver.chock iterates from the OLD iterate (solver.py:29-37); the current one is scratch from here on
    s->collapsed = false;
    s->in_loop = true;
    s->loop_old0 = s->old_i;
    if (s->panel_live) {   // the old iterate into panels (32 instances per panel, the lanes of a warp)
        launch_to_panels(st, s->prim[s->old_i], s->pprim[s->old_i], L.np_pad, L.batch);
        launch_to_panels(st, s->dual[s->old_i], s->pdual[s->old_i], L.nd_pad, L.batch);
        s->launches += 2;
        return launch_ok(s, "panel conversion");
    }
    return RB_OK;
}

int rb_loop_enqueue(rb_solver *s, int32_t count) {
    if (!s || count < 0) return RB_ERR_INVALID;
    if (!s->in_loop) return fail(s, RB_ERR_STATE, "rb_loop_begin() has not been called");
    for (int k = 0; k < count; ++k) {
        const int src = s->old_i;
        if (s->sharded && shard_pipe(s)) {
            if (s->use_graphs && s->pbar_ready && s->shard_pending && !shard_timing_on()) {
                RB_CUDA(s, cudaGraphLaunch(s->graph[src], s->stream));
            } else {
                int rc = enqueue_iteration_sharded_pipe(s, src, s->stream, s->pbar_ready);
                if (rc != RB_OK) return rc;
            }
            s->pbar_ready = true;
            s->shard_pending = true;
            std::swap(s->cur_i, s->old_i);
            continue;
        } else if (s->sharded) {
            int rc = enqueue_iteration_sharded(s, src, s->stream);
            if (rc != RB_OK) return rc;
        } else if (s->use_graphs && (s->pbar_ready || !(use_pipe(s) || s->panel_live))) {
            RB_CUDA(s, cudaGraphLaunch(s->graph[src], s->stream));
        } else {   // plain launches; the first iteration of a pipelined loop has no pbar yet
            int rc = enqueue_iteration_kernels(s, src, s->stream, s->pbar_ready);
            if (rc != RB_OK) return rc;
        }
        s->pbar_ready = (use_pipe(s) || s->panel_live) && !s->sharded;
        // the buffer just written holds the newest iterate: it is the next iteration's "old"
        std::swap(s->cur_i, s->old_i);
    }
    s->launches += (int64_t)count * iter_launches(s);
    return RB_OK;
}

int rb_loop_poll(rb_solver *s, int32_t *iters, int32_t *done, double *last_norms) {
    if (!s) return RB_ERR_INVALID;
    if (!s->in_loop) return fail(s, RB_ERR_STATE, "rb_loop_begin() has not been called");
    const Layout &L = s->P.L;
    if (s->sharded && s->shard_pending) {   // gather and test the residuals of the last enqueued iteration (all ranks)
        int rcx = shard_exchange(s, s->old_i, s->stream, nullptr, shard_pipe(s) ? s->slots + 6 * (1 - s->old_i) : nullptr);
        if (rcx != RB_OK) return rcx;
    }
    Ctrl *hc = reinterpret_cast<Ctrl *>(s->h_pinned);
    double *hn = s->h_pinned + 64;
    if ((size_t)L.batch * 6 > (4096 - 512) / sizeof(double)) hn = nullptr;
    RB_CUDA(s, cudaMemcpyAsync(hc, s->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, s->stream));
    if (last_norms) {
        if (hn) RB_CUDA(s, cudaMemcpyAsync(hn, s->last, (size_t)L.batch * 6 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        else RB_CUDA(s, cudaMemcpyAsync(last_norms, s->last, (size_t)L.batch * 6 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    }
    RB_CUDA(s, cudaStreamSynchronize(s->stream));
    if (last_norms && hn) std::memcpy(last_norms, hn, (size_t)L.batch * 6 * sizeof(double));
    if (iters) *iters = hc->iters;
    if (done) *done = hc->done;
    if (hc->status) {
        RB_CUDA(s, cudaMemsetAsync(&s->ctrl->status, 0, sizeof(int), s->stream));
        if (hc->status & 16) return fail(s, RB_ERR_CUDA, "subtree sharding: the peer-memory exchange timed out (a rank is gone)");
        if (hc->status & 32) return fail(s, RB_ERR_CUDA, "launch overlap: a kernel waited ~0.2 s for the kernel in front of it");
        if (hc->status & 1) return fail(s, RB_ERR_NUMERIC, "Rectangle constraint - 'nan' value cannot be constrained");
        return fail(s, RB_ERR_NUMERIC, "non-finite value in the residuals");
    }
    return RB_OK;
}

int rb_loop_end(rb_solver *s, double *xi_hist, double *delta_hist, int32_t *iters_out, int32_t *status_out) {
    if (!s) return RB_ERR_INVALID;
    if (!s->in_loop) return fail(s, RB_ERR_STATE, "rb_loop_begin() has not been called");
    const Layout &L = s->P.L;
    int32_t iters = 0, done = 0;
    int rc = rb_loop_poll(s, &iters, &done, nullptr);
    s->in_loop = false;
    if (s->sharded && shard_timing_on()) g_shard_timing.report(s->shard.rank);
    if (rc != RB_OK && rc != RB_ERR_NUMERIC) return rc;
    // iterations enqueued after the stopping test fired were no-ops, but the host kept swapping roles while
    // enqueueing: iteration k read buffer (k even ? loop_old0 : 1 - loop_old0) and wrote the other one, so the
    // newest iterate is in the buffer written by iteration iters-1
    const int newest = iters == 0 ? s->loop_old0 : (((iters - 1) % 2 == 0) ? 1 - s->loop_old0 : s->loop_old0);
    s->old_i = newest;
    s->cur_i = 1 - newest;
    s->collapsed = true;
    if (s->panel_live) {   // the newest iterate back into the instance-major buffers the rest of the API works on
        launch_from_panels(s->stream, s->pprim[newest], s->prim[newest], L.np_pad, L.batch);
        launch_from_panels(s->stream, s->pdual[newest], s->dual[newest], L.nd_pad, L.batch);
        s->launches += 2;
        RB_CUDA(s, cudaStreamSynchronize(s->stream));
        int rcp = launch_ok(s, "panel conversion");
        if (rcp != RB_OK) return rcp;
    }
    Ctrl *hc = reinterpret_cast<Ctrl *>(s->h_pinned);
    if ((xi_hist || delta_hist) && hc->hist_capacity > 0) {
        const int rows = std::min(iters, hc->hist_capacity);
        std::vector<double> h((size_t)rows * L.batch * 6);
        RB_CUDA(s, cudaMemcpyAsync(h.data(), s->hist, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        RB_CUDA(s, cudaStreamSynchronize(s->stream));
        for (size_t i = 0; i < (size_t)rows * L.batch; ++i)
            for (int k = 0; k < 3; ++k) {
                if (xi_hist) xi_hist[i * 3 + k] = h[i * 6 + k];
                if (delta_hist) delta_hist[i * 3 + k] = h[i * 6 + 3 + k];
            }
    }
    if (iters_out) *iters_out = iters;
    // reference status (solver.py:166-169): 0 if the final iteration index is < max_iters
    if (status_out) *status_out = (iters - 1 < hc->max_iters) ? 0 : 1;
    return rc;
}

// a loop that failed half way: drain the stream, give the buffers back the roles they had at rb_loop_begin (the iterate of
// that moment is still intact in the "old" buffer of the first iteration only if nothing ran; either way the handle is
// usable again and rb_use_* / rb_loop_begin are accepted), keep the error message of the failing call
static int abort_loop(rb_solver *s, int rc) {
    const std::string msg = s->err;
    cudaStreamSynchronize(s->stream);
    for (auto q : s->side)
        if (q) cudaStreamSynchronize(q);
    cudaGetLastError();
    s->in_loop = false;
    s->pbar_ready = false;
    s->shard_pending = false;
    s->panel_live = false;
    s->old_i = s->loop_old0;
    s->cur_i = 1 - s->loop_old0;
    s->collapsed = true;
    s->err = msg;
    return rc;
}

int rb_iterate(rb_solver *s, double alpha, int32_t max_iters, double tol, int32_t check_every, double *xi_hist,
               double *delta_hist, int32_t hist_capacity, int32_t *iters, int32_t *status) {
    if (!s || max_iters < 0) return RB_ERR_INVALID;
    const bool want_hist = xi_hist || delta_hist;
    const int total = max_iters + 1;
    int rc = rb_loop_begin(s, alpha, max_iters, tol, want_hist ? std::min(total, (int)hist_capacity) : 0);
    if (rc != RB_OK) return rc;
    // the stopping test runs on the device after every iteration; the host only polls every `chunk` iterations and
    // the iterations enqueued past the stopping point are no-ops
    const int chunk = std::max(1, check_every <= 1 ? 64 : (int)check_every);
    int launched = 0;
    while (launched < total) {
        const int todo = std::min(chunk, total - launched);
        rc = rb_loop_enqueue(s, todo);
        if (rc != RB_OK) return abort_loop(s, rc);
        launched += todo;
        int32_t it = 0, done = 0;
        rc = rb_loop_poll(s, &it, &done, nullptr);
        if (rc != RB_OK) return abort_loop(s, rc);
        if (done) break;
    }
    return rb_loop_end(s, xi_hist, delta_hist, iters, status);
}

int rb_iterate_fixed(rb_solver *s, double alpha, int32_t iters, double *norms) {
    if (!s || iters < 1) return RB_ERR_INVALID;
    int rc = rb_loop_begin(s, alpha, iters - 1, -1.0, 0);
    if (rc != RB_OK) return rc;
    rc = rb_loop_enqueue(s, iters);
    if (rc != RB_OK) return abort_loop(s, rc);
    int32_t it = 0, done = 0;
    rc = rb_loop_poll(s, &it, &done, norms);
    if (rc != RB_OK) return abort_loop(s, rc);
    return rb_loop_end(s, nullptr, nullptr, nullptr, nullptr);
}

// one end-to-end step for a host caller: pinned-host x0 -> device, one iteration, six residual norms -> host
int rb_step(rb_solver *s, const double *x0, double *norms) {
    if (!s || !norms) return RB_ERR_INVALID;
    if (!s->in_loop) return fail(s, RB_ERR_STATE, "rb_loop_begin() has not been called");
    const Layout &L = s->P.L;
    // pipelined loop with the mapped mirror: ONE upload (the kernel projection copies x0 into x_0 of the old iterate)
    // and no download (k_check has written the norms to host memory when the stream is idle)
    const bool lean = s->panel_live || (s->h_last_dev && use_pipe(s) && s->pbar_ready);   // the kernel projection copies x0 into x_0
    if (s->h_last_dev && !s->mirror_on) {   // from now on the stopping test mirrors the norms into host memory
        int *one = reinterpret_cast<int *>(s->h_pinned + 256);
        *one = 1;
        RB_CUDA(s, cudaMemcpyAsync(&s->ctrl->mirror, one, sizeof(int), cudaMemcpyHostToDevice, s->stream));
        s->mirror_on = true;
    }
    if (x0) {
        RB_CUDA(s, cudaMemcpyAsync(s->x0, x0, (size_t)L.batch * L.nx * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        if (!lean)
            RB_CUDA(s, cudaMemcpy2DAsync(s->prim[s->old_i] + L.px, L.np_pad * sizeof(double), x0, L.nx * sizeof(double),
                                         L.nx * sizeof(double), L.batch, cudaMemcpyHostToDevice, s->stream));
    }
    int rc = rb_loop_enqueue(s, 1);
    if (rc != RB_OK) return rc;
    if (s->h_last_dev) {
        RB_CUDA(s, cudaStreamSynchronize(s->stream));
        std::memcpy(norms, s->h_last, (size_t)L.batch * 6 * sizeof(double));
    } else {
        RB_CUDA(s, cudaMemcpyAsync(norms, s->last, (size_t)L.batch * 6 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        RB_CUDA(s, cudaStreamSynchronize(s->stream));
    }
    return RB_OK;
}

// one iteration with CUDA events after every launch (plain launches).  ms[0] primal pass; ms[1..] the sweep launches in
// order (backward levels bottom-up, top, forward levels top-down: 1 + 2 * levels entries); then the dual pass (+ stopping
// test).  Unused entries are -1.  Advances the loop by one iteration.
int rb_profile_iteration(rb_solver *s, float *ms) {
    if (!s || !ms) return RB_ERR_INVALID;
    if (!s->in_loop) return fail(s, RB_ERR_STATE, "rb_loop_begin() has not been called");
    const Layout &L = s->P.L;
    cudaStream_t st = s->stream;
    int nsweep = 0;
    cudaEvent_t ev[12];
    for (auto &e : ev) RB_CUDA(s, cudaEventCreate(&e));
    const int src = s->old_i, dst = 1 - src;
    if (s->panel_live) {   // ms[0] kernel projection (+ primal pass), ms[1] backward stages, ms[2] forward stages, ms[3] dual + check
        for (int i = 0; i < 12; ++i) ms[i] = -1.0f;
        RB_CUDA(s, cudaEventRecord(ev[0], st));
        if (!s->pbar_ready) launch_bp_primal(st, s->P, s->ctrl, s->pprim[src], s->pdual[src], s->pprim[dst]);
        launch_bp_kproj(st, s->P, s->ctrl, s->pprim[dst], s->x0, s->pprim[src]);
        RB_CUDA(s, cudaEventRecord(ev[1], st));
        launch_panel_sweeps(s, st, s->pprim[dst], ev[2]);   // ms[1]: backward stage launches; ms[2]: fused top + forward stage launches
        RB_CUDA(s, cudaEventRecord(ev[3], st));
        launch_bp_dual(st, s->P, s->ctrl, s->pprim[src], s->pprim[dst], s->pdual[src], s->pdual[dst], s->slots, s->pprim[src], s->pc2,
                       ev + 5);
        RB_CUDA(s, cudaEventRecord(ev[7], st));
        launch_check(st, s->P, s->ctrl, s->slots, s->last, s->h_last_dev);
        RB_CUDA(s, cudaEventRecord(ev[4], st));
        RB_CUDA(s, cudaStreamSynchronize(st));
        int rcp = launch_ok(s, "profiled panel iteration");
        for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
        cudaEventElapsedTime(&ms[8], ev[3], ev[5]);    // x / u block of the nonleaf nodes
        cudaEventElapsedTime(&ms[9], ev[5], ev[6]);    // risk block
        cudaEventElapsedTime(&ms[10], ev[6], ev[7]);   // leaves
        for (auto &e : ev) cudaEventDestroy(e);
        std::swap(s->cur_i, s->old_i);
        s->pbar_ready = true;
        s->launches += iter_launches(s);
        return rcp;
    }
    const bool pipe = use_pipe(s);
    RB_CUDA(s, cudaEventRecord(ev[0], st));
    if (pipe && s->pbar_ready) launch_kproj(L.batch, st, s->P, s->ctrl, s->prim[dst], s->h_last_dev ? s->x0 : nullptr, s->prim[src]);
    else launch_primal(s, st, src, dst);
    const bool risk_split = pipe && s->risk_split && pipe_split(s).cf < L.m;
    if (risk_split) {
        const PipeSplit ps = pipe_split(s);
        launch_dual_risk_chain(L.batch, st, s->P, s->ctrl, s->prim[src], s->prim[dst], s->dual[src], s->dual[dst], s->slots, ps.cf,
                               L.m - ps.cf, s->chain_stride, s->chain_yo0, s->prim[src]);
    }
    RB_CUDA(s, cudaEventRecord(ev[1], st));
    int rcs = launch_sweeps(s, s->ctrl, s->prim[dst], st, ev + 2, &nsweep);
    if (rcs != RB_OK) return rcs;
    for (int i = 0; i < 12; ++i) ms[i] = -1.0f;
    int nd = 0;   // events ev[8..11] bracket the three kernels of the pipelined dual pass
    if (pipe) {
        const PipeSplit ps = pipe_split(s);
        const dim3 nb(1, L.batch);
        auto dual_lane = [&](int first, int count) {
            launch_dual_lane(nb, st, s->P, s->ctrl, s->prim[src], s->prim[dst], s->dual[src], s->dual[dst], s->slots, nullptr,
                             first, count, s->prim[src]);
        };
        RB_CUDA(s, cudaEventRecord(ev[8], st));
        if (ps.cf >= L.m) {
            dual_lane(0, L.n);
            RB_CUDA(s, cudaEventRecord(ev[9], st));
            nd = 1;
        } else {
            dual_lane(0, ps.cf);
            RB_CUDA(s, cudaEventRecord(ev[9], st));
            launch_dual_chain(L.batch, st, s->P, s->ctrl, s->prim[src], s->prim[dst], s->dual[src], s->dual[dst], s->slots,
                              s->chain_recs, ps.cf, L.m - ps.cf, s->chain_stride, s->chain_yo0, s->prim[src], risk_split ? 0 : 1);
            RB_CUDA(s, cudaEventRecord(ev[10], st));
            dual_lane(L.m, L.n - L.m);
            RB_CUDA(s, cudaEventRecord(ev[11], st));
            nd = 3;
        }
    } else {
        launch_dual(s, st, src, dst);
    }
    launch_check(st, s->P, s->ctrl, s->slots, s->last, s->h_last_dev);
    RB_CUDA(s, cudaEventRecord(ev[2 + nsweep], st));
    RB_CUDA(s, cudaStreamSynchronize(st));
    int rc = launch_ok(s, "profiled iteration");
    for (int i = 0; i < 2 + nsweep; ++i) cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
    for (int i = 0; i < nd; ++i) cudaEventElapsedTime(&ms[8 + i], ev[8 + i], ev[9 + i]);
    for (auto &e : ev) cudaEventDestroy(e);
    std::swap(s->cur_i, s->old_i);
    s->pbar_ready = pipe;
    s->launches += iter_launches(s) - (pipe && pipe_split(s).split > 0 ? 2 : 0);   // profiled: the chain walk in one piece
    return rc;
}

int rb_pipeline_info(const rb_solver *s, int32_t *early_nodes, int32_t *chain_first, int32_t *chain_nodes) {
    if (!s) return RB_ERR_INVALID;
    const bool pipe = use_pipe(s);
    const PipeSplit ps = pipe ? pipe_split(s) : PipeSplit{0, s->P.L.m};
    if (early_nodes) *early_nodes = ps.early;
    if (chain_first) *chain_first = ps.cf;
    if (chain_nodes) *chain_nodes = pipe ? s->P.L.m - ps.cf : 0;
    return RB_OK;
}

int rb_use_pipeline(rb_solver *s, int32_t enable) {
    if (!s) return RB_ERR_INVALID;
    if (s->in_loop) return fail(s, RB_ERR_STATE, "rb_use_pipeline() inside a loop");
    s->allow_pipe = enable != 0;
    s->risk_split = enable != 4;       // 4: pipelined, the risk block inside the chain dual pass (ablation)
    s->pipe_fwd_split = enable == 3;   // 3: additionally the forward chain walk in two pieces (measured slower: kept as an ablation)
    drop_graphs(s);
    return RB_OK;
}

int rb_force_dense_costs(rb_solver *s, int32_t enable) {
    if (!s) return RB_ERR_INVALID;
    // test hook: run the general (dense matrix) cost path even when sqrtQ, sqrtR, sqrtQf are all diagonal
    const Tabs &M = s->P.m;
    s->diag_costs = !enable && M.sq_diag && M.sr_diag && M.sqf_diag;
    drop_graphs(s);
    return RB_OK;
}

int rb_shard_unique_id(char *id128) {
    if (!id128) return RB_ERR_INVALID;
    const char *why = nccl_load();
    if (why) return fail(nullptr, RB_ERR_INVALID, why);
    NcclId id;
    const int rc = nccl_unique_id(&id);
    if (rc != 0) return fail(nullptr, RB_ERR_CUDA, std::string("ncclGetUniqueId: ") + nccl_error(rc));
    std::memcpy(id128, id.internal, 128);
    return RB_OK;
}

int rb_shard_init(rb_solver *s, const char *id128) {
    if (!s || !id128) return RB_ERR_INVALID;
    if (!s->sharded) return fail(s, RB_ERR_STATE, "the problem was not created with shard_world > 1");
    const char *why = nccl_load();
    if (why) return fail(s, RB_ERR_INVALID, why);
    NcclId id;
    std::memcpy(id.internal, id128, 128);
    RB_CUDA(s, cudaSetDevice(s->device));
    const int rc = nccl_comm_init(&s->nccl_comm, s->shard.world, id, s->shard.rank);
    if (rc != 0) return fail(s, RB_ERR_CUDA, std::string("ncclCommInitRank: ") + nccl_error(rc));
    return RB_OK;
}

// peer-memory exchange: every rank exports its receive buffer and flag array (two cudaIpcMemHandle_t, 64 bytes each), the host
// side distributes them (torch.distributed), every rank opens the others'
int rb_shard_p2p_export(rb_solver *s, char *handles128) {
    if (!s || !handles128) return RB_ERR_INVALID;
    if (!s->sharded) return fail(s, RB_ERR_STATE, "the problem was not created with shard_world > 1");
    if (s->shard.world > kMaxPeers) return fail(s, RB_ERR_INVALID, "peer-memory exchange supports at most 8 ranks");
    RB_CUDA(s, cudaSetDevice(s->device));
    const int W = s->shard.world;
    if (!s->p2p_recv) {
        // receive buffers [2][W][xchg_count] doubles, then the packet area [2][W][xchg_count] x 16 bytes (PeerXchg::ll)
        int rc = dev_zero(s, (size_t)2 * W * s->xchg_count * 3, &s->p2p_recv);
        if (rc != RB_OK) return rc;
        rc = dev_zero(s, (size_t)2 * W, &s->p2p_flag);
        if (rc != RB_OK) return rc;
        unsigned long long *seq = nullptr;
        rc = dev_zero(s, 2, &seq);   // [0] sequence number, [1] arrival counter of k_shard_xchg
        if (rc != RB_OK) return rc;
        const unsigned long long one = 1ull;
        RB_CUDA(s, cudaMemcpy(seq, &one, sizeof(one), cudaMemcpyHostToDevice));
        s->px.seq = seq;
    }
    cudaIpcMemHandle_t h[2];
    RB_CUDA(s, cudaIpcGetMemHandle(&h[0], s->p2p_recv));
    RB_CUDA(s, cudaIpcGetMemHandle(&h[1], s->p2p_flag));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    std::memcpy(handles128, h, 128);
    return RB_OK;
}

int rb_shard_p2p_open(rb_solver *s, const char *all_handles) {
    if (!s || !all_handles) return RB_ERR_INVALID;
    if (!s->sharded || !s->p2p_recv) return fail(s, RB_ERR_STATE, "rb_shard_p2p_export() first");
    RB_CUDA(s, cudaSetDevice(s->device));
    const int W = s->shard.world, R = s->shard.rank;
    for (int r = 0; r < W; ++r) {
        if (r == R) {
            s->px.recv[r] = s->p2p_recv;
            s->px.flag[r] = s->p2p_flag;
            s->px.ll[r] = reinterpret_cast<uint4 *>(s->p2p_recv + (size_t)2 * W * s->xchg_count);
            continue;
        }
        cudaIpcMemHandle_t h[2];
        std::memcpy(h, all_handles + (size_t)r * 128, 128);
        void *pr = nullptr, *pf = nullptr;
        RB_CUDA(s, cudaIpcOpenMemHandle(&pr, h[0], cudaIpcMemLazyEnablePeerAccess));
        RB_CUDA(s, cudaIpcOpenMemHandle(&pf, h[1], cudaIpcMemLazyEnablePeerAccess));
        s->p2p_opened[2 * r] = pr;
        s->p2p_opened[2 * r + 1] = pf;
        s->px.recv[r] = static_cast<double *>(pr);
        s->px.flag[r] = static_cast<unsigned long long *>(pf);
        s->px.ll[r] = reinterpret_cast<uint4 *>(s->px.recv[r] + (size_t)2 * W * s->xchg_count);
    }
    s->p2p = true;
    // RAOCP_SHARD_XCHG (ablations): "split": push / pull / stopping test as separate launches; "kernel": one exchange launch
    // (k_shard_xchg) between the level kernels; default: the exchange inside the kernel that sweeps the top of the tree
    const char *mode = getenv("RAOCP_SHARD_XCHG");
    s->xchg_fused = !(mode && std::strcmp(mode, "split") == 0);
    s->xchg_in_top = !(mode && (std::strcmp(mode, "split") == 0 || std::strcmp(mode, "kernel") == 0));
    return RB_OK;
}

int rb_shard_info(const rb_solver *s, int32_t *cut_stage, int32_t *cut_first, int32_t *num_cut, int32_t *chain_stage) {
    if (!s) return RB_ERR_INVALID;
    const SweepPlan &pl = s->plan;
    if (cut_stage) *cut_stage = pl.t_top;
    if (cut_first) *cut_first = pl.t_top < s->P.L.num_stages ? s->stage_off[pl.t_top] : s->P.L.n;
    if (num_cut) *num_cut = pl.num_levels > 0 ? pl.lv[0].num_sub : 0;
    if (chain_stage) *chain_stage = pl.num_levels > 1 ? pl.lv[1].t_lo : -1;
    return RB_OK;
}

int rb_use_lane_kernels(rb_solver *s, int32_t enable) {
    if (!s) return RB_ERR_INVALID;
    s->allow_lane = enable != 0;
    drop_graphs(s);
    return RB_OK;
}

int rb_use_mma_sweeps(rb_solver *s, int32_t enable) {
    if (!s) return RB_ERR_INVALID;
    s->allow_mma = enable != 0;
    s->mma_w4 = enable == 2;   // 2: four warps per tile where instantiated (ablation); other non-zero values: one warp per tile
    s->mma_wide = enable != 3; // 3: wide rows with the one-warp BIG kernels (ablation)
    drop_graphs(s);
    return RB_OK;
}

int rb_use_tree_kernels(rb_solver *s, int32_t enable) {
    if (!s) return RB_ERR_INVALID;
    s->tree_mode = enable < 0 ? 0 : (enable > 2 ? 2 : enable);
    drop_graphs(s);
    return RB_OK;
}

int rb_use_table_prefetch(rb_solver *s, int32_t enable) {
    if (!s) return RB_ERR_INVALID;
    if (s->in_loop) return fail(s, RB_ERR_STATE, "rb_use_table_prefetch() inside a loop");
    s->allow_table_prefetch = enable != 0;
    drop_graphs(s);
    return RB_OK;
}

int rb_use_launch_overlap(rb_solver *s, int32_t enable) {
    if (!s) return RB_ERR_INVALID;
    if (s->in_loop) return fail(s, RB_ERR_STATE, "rb_use_launch_overlap() inside a loop");
    s->allow_overlap = enable != 0;
    drop_graphs(s);
    return RB_OK;
}

int rb_use_batch_panels(rb_solver *s, int32_t enable) {
    if (!s) return RB_ERR_INVALID;
    if (s->in_loop) return fail(s, RB_ERR_STATE, "rb_use_batch_panels() inside a loop");
    s->allow_panel = enable != 0;
    return RB_OK;
}

int rb_use_fused_check(rb_solver *s, int32_t enable) {
    if (!s) return RB_ERR_INVALID;
    if (s->in_loop) return fail(s, RB_ERR_STATE, "rb_use_fused_check() inside a loop");
    s->fused_check = enable != 0;
    drop_graphs(s);
    return RB_OK;
}

int rb_use_graphs(rb_solver *s, int32_t enable) {
    if (!s) return RB_ERR_INVALID;
    s->use_graphs = enable != 0;
    return RB_OK;
}

// ---- stand-alone projections ---------------------------------------------------------------------------------------------
int rb_cone_project(int32_t cone, int32_t dim, const double *in, double *out) {
    if (cone < 0 || cone > 3 || dim < 1 || !in || !out) return fail(nullptr, RB_ERR_INVALID, "bad cone arguments");
    if (dim * sizeof(double) > 48 * 1024) return fail(nullptr, RB_ERR_INVALID, "cone dimension too large");
    double *d_in = nullptr, *d_out = nullptr;
    RB_CUDA(nullptr, cudaMalloc((void **)&d_in, dim * sizeof(double)));
    RB_CUDA(nullptr, cudaMalloc((void **)&d_out, dim * sizeof(double)));
    RB_CUDA(nullptr, cudaMemcpy(d_in, in, dim * sizeof(double), cudaMemcpyHostToDevice));
    k_cone<<<1, 32, dim * sizeof(double)>>>(cone, dim, d_in, d_out);
    RB_CUDA(nullptr, cudaGetLastError());
    RB_CUDA(nullptr, cudaMemcpy(out, d_out, dim * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d_in);
    cudaFree(d_out);
    return RB_OK;
}

int rb_box_project(int32_t dim, const double *in, const double *lo, const double *hi, double *out) {
    if (dim < 1 || !in || !lo || !hi || !out) return fail(nullptr, RB_ERR_INVALID, "bad box arguments");
    double *d = nullptr;
    int *d_status = nullptr;
    RB_CUDA(nullptr, cudaMalloc((void **)&d, 4 * (size_t)dim * sizeof(double)));
    RB_CUDA(nullptr, cudaMalloc((void **)&d_status, sizeof(int)));
    RB_CUDA(nullptr, cudaMemset(d_status, 0, sizeof(int)));
    RB_CUDA(nullptr, cudaMemcpy(d, in, dim * sizeof(double), cudaMemcpyHostToDevice));
    RB_CUDA(nullptr, cudaMemcpy(d + dim, lo, dim * sizeof(double), cudaMemcpyHostToDevice));
    RB_CUDA(nullptr, cudaMemcpy(d + 2 * dim, hi, dim * sizeof(double), cudaMemcpyHostToDevice));
    k_box<<<(dim + 255) / 256, 256>>>(dim, d, d + dim, d + 2 * dim, d + 3 * dim, d_status);
    RB_CUDA(nullptr, cudaGetLastError());
    int st = 0;
    RB_CUDA(nullptr, cudaMemcpy(out, d + 3 * dim, dim * sizeof(double), cudaMemcpyDeviceToHost));
    RB_CUDA(nullptr, cudaMemcpy(&st, d_status, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d);
    cudaFree(d_status);
    if (st) return fail(nullptr, RB_ERR_NUMERIC, "Rectangle constraint - 'nan' value cannot be constrained");
    return RB_OK;
}

}  // extern "C"
