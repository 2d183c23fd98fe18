// sweeps.cu -- projection onto the dynamics set (reference cache.py:259-288) in a handful of launches per iteration
// instead of one per stage.
//
// In breadth-first numbering the descendants of a node form ONE contiguous node range per stage, so a CTA can own a
// whole subtree and walk it stage by stage with block barriers only.  The tree is cut at (up to) two stages:
//   level 0 : subtrees rooted at the first stage with >= 64 nodes
//   level 1 : (optional) subtrees rooted at a much wider stage further down -- below the stopping time of a Markov
//             tree these are chains, one warp walks one chain with no barrier at all
//   top     : the few nodes above level 0, one CTA per problem instance: backward to the root, then forward again.
// Backward: level 1, level 0, top;  forward: top (same launch), level 0, level 1.
//
// Where a subtree branches, the per-stage work is spread over EDGES, not nodes: one warp per child computes that
// child's contribution [A_j' q_j ; B_j' q_j] (backward) or the child's state x_j = A_j x_i + B_j u_i (forward), a second
// short phase per parent adds the contributions up / prepares u_i -- the latency of a stage step is one matrix-vector
// product, not one per child.
//
// The kernels are templates on <NX, NU>: common sizes get fully unrolled matrix-vector loops (node_ops.cuh), any other
// size runs the same code with run-time loops.  The small mode-indexed tables (A, B concatenations) are shared by all
// nodes and stay L1-resident; the class-indexed K, [K R~^-1] stream through L1/L2 (or from HBM when every node is its
// own class).
#include "kernels.cuh"
#include "node_ops.cuh"

namespace rb {

namespace {

struct Group {           // the warps that cooperate on one subtree
    int wl, wps;         // warp index inside the group, warps in the group
    double *scratch;     // warp-private: 2*(nx+nu)+32 doubles
    double *shared;      // group-shared: contributions / parent inputs of one stage, stage_cap*(nx+nu) doubles
};

// one backward stage over nodes [lo, hi) of a subtree; group barrier = __syncthreads (all groups of a CTA walk subtrees
// of equal depth, so they reach every barrier together)
template <int NX, int NU>
__device__ __forceinline__ void stage_bwd(const Layout &L, const Topo &T, const Tabs &M, const double *__restrict__ X,
                                          const double *__restrict__ U, double *__restrict__ Q, double *__restrict__ R,
                                          int lo, int hi, const Group &g, int lane, bool live) {
    const int nx = NX > 0 ? NX : L.nx, nxu = NX > 0 ? NX + NU : L.nxu;
    if (g.wps == 1) {   // a single warp owns the whole stage range (top of a tiny tree)
        if (live)
            for (int node = lo; node < hi; ++node) dyn_bwd_node<NX, NU>(L, T, M, X, U, Q, R, node, lane, g.scratch);
        __syncwarp();
        return;
    }
    const bool leaves = lo >= L.m;
    if (live) {
        if (leaves) {
            for (int node = lo + g.wl; node < hi; node += g.wps) dyn_bwd_node<NX, NU>(L, T, M, X, U, Q, R, node, lane, g.scratch);
        } else {   // phase 1: one warp per child edge
            const int clo = T.child_first[lo], chi = T.child_first[hi - 1] + T.child_count[hi - 1];
            for (int j = clo + g.wl; j < chi; j += g.wps) {
                double *qj = g.scratch;
                for (int k = lane; k < nx; k += 32) qj[k] = Q[j * nx + k];
                __syncwarp();
                bwd_child_contrib<NX, NU>(L, M, T.dyn_idx[j], qj, lane, g.shared + (j - clo) * nxu, false);
                __syncwarp();
            }
        }
    }
    __syncthreads();
    if (live && !leaves) {   // phase 2: one warp per parent adds its children's contributions and finishes the step
        const int clo = T.child_first[lo];
        for (int node = lo + g.wl; node < hi; node += g.wps) {
            double *acc = g.scratch, *rv = g.scratch + nxu;
            const int c0 = T.child_first[node], cc = T.child_count[node];
            for (int k = lane; k < nxu; k += 32) {
                double a = 0.0;
                for (int j = c0; j < c0 + cc; ++j) a += g.shared[(j - clo) * nxu + k];
                acc[k] = a;
            }
            __syncwarp();
            bwd_finish<NX, NU>(L, T, M, X, U, Q, R, node, lane, acc, rv);
        }
    }
    __syncthreads();
}

template <int NX, int NU>
__device__ __forceinline__ void stage_fwd(const Layout &L, const Topo &T, const Tabs &M, double *__restrict__ X,
                                          double *__restrict__ U, const double *__restrict__ R, int lo, int hi,
                                          const Group &g, int lane, bool live) {
    const int nxu = NX > 0 ? NX + NU : L.nxu;
    if (lo >= L.m) return;   // leaves: nothing to do (uniform for the whole CTA: equal depths)
    if (g.wps == 1) {
        if (live)
            for (int node = lo; node < hi; ++node) dyn_fwd_node<NX, NU>(L, T, M, X, U, R, node, lane, g.scratch);
        return;
    }
    if (live)   // phase 1: one warp per parent: u_i and v = [x_i ; u_i] into the group-shared buffer
        for (int node = lo + g.wl; node < hi; node += g.wps)
            fwd_input<NX, NU>(L, T, M, X, U, R, node, lane, g.shared + (node - lo) * nxu, g.scratch);
    __syncthreads();
    if (live) {   // phase 2: one warp per child edge
        const int clo = T.child_first[lo], chi = T.child_first[hi - 1] + T.child_count[hi - 1];
        for (int j = clo + g.wl; j < chi; j += g.wps)
            fwd_child<NX, NU>(L, M, T.dyn_idx[j], g.shared + (T.parent[j] - lo) * nxu, X, j, lane);
    }
    __syncthreads();
}

__device__ __forceinline__ Group make_group(const Layout &L, double *dyn_smem, int warp, int wps, int subs, int stage_cap) {
    const int per_warp = 2 * L.nxu + 32;
    Group g;
    g.wps = wps;
    g.wl = warp % wps;
    g.scratch = dyn_smem + (size_t)warp * per_warp;
    g.shared = dyn_smem + (size_t)wps * subs * per_warp + (size_t)(warp / wps) * stage_cap * L.nxu;
    return g;
}

}  // namespace

template <int NX, int NU>
__global__ void __launch_bounds__(512) k_sweep_sub_bwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                      SweepLevel lv, const double *__restrict__ prim,
                                                      double *__restrict__ q, double *__restrict__ r) {
    if (ctrl && ctrl->done) return;
    extern __shared__ double dyn_smem[];
    const Layout &L = P.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Group g = make_group(L, dyn_smem, warp, lv.warps_per_sub, lv.subs_per_cta, lv.stage_cap);
    const double *Pp = prim + (long long)blockIdx.y * L.np_pad;
    double *Q = q + (long long)blockIdx.y * L.n * L.nx;
    double *R = r + (long long)blockIdx.y * L.m * L.nu;
    const int sub = blockIdx.x * lv.subs_per_cta + warp / lv.warps_per_sub;
    const bool live = sub < lv.num_sub;
    const int *lo = lv.lo + (long long)(live ? sub : 0) * lv.depth, *hi = lv.hi + (long long)(live ? sub : 0) * lv.depth;
    if (lv.chain) {   // chains: one warp walks one chain, no barriers
        if (live) chain_bwd<NX, NU>(L, P.t, P.m, Pp + L.px, Pp + L.pu, Q, R, lo, lv.depth, lane, g.scratch);
        return;
    }
    for (int d = lv.depth - 1; d >= 0; --d)
        stage_bwd<NX, NU>(L, P.t, P.m, Pp + L.px, Pp + L.pu, Q, R, lo[d], hi[d], g, lane, live);
}

template <int NX, int NU>
__global__ void __launch_bounds__(512) k_sweep_sub_fwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                      SweepLevel lv, double *__restrict__ prim,
                                                      const double *__restrict__ r) {
    if (ctrl && ctrl->done) return;
    extern __shared__ double dyn_smem[];
    const Layout &L = P.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Group g = make_group(L, dyn_smem, warp, lv.warps_per_sub, lv.subs_per_cta, lv.stage_cap);
    double *Pp = prim + (long long)blockIdx.y * L.np_pad;
    const double *R = r + (long long)blockIdx.y * L.m * L.nu;
    const int sub = blockIdx.x * lv.subs_per_cta + warp / lv.warps_per_sub;
    const bool live = sub < lv.num_sub;
    const int *lo = lv.lo + (long long)(live ? sub : 0) * lv.depth, *hi = lv.hi + (long long)(live ? sub : 0) * lv.depth;
    if (lv.chain) {
        if (live) chain_fwd<NX, NU>(L, P.t, P.m, Pp + L.px, Pp + L.pu, R, lo, lv.depth, lane, g.scratch);
        return;
    }
    for (int d = 0; d < lv.depth; ++d) stage_fwd<NX, NU>(L, P.t, P.m, Pp + L.px, Pp + L.pu, R, lo[d], hi[d], g, lane, live);
}

template <int NX, int NU>
__global__ void __launch_bounds__(1024) k_sweep_top(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                   SweepPlan plan, double *__restrict__ prim, double *__restrict__ q,
                                                   double *__restrict__ r, const double *__restrict__ x0) {
    if (ctrl && ctrl->done) return;
    extern __shared__ double dyn_smem[];
    const Layout &L = P.L;
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Group g = make_group(L, dyn_smem, warp, warps, 1, plan.top_cap);
    double *Pp = prim + (long long)blockIdx.x * L.np_pad;
    double *Q = q + (long long)blockIdx.x * L.n * L.nx;
    double *R = r + (long long)blockIdx.x * L.m * L.nu;
    for (int t = plan.t_top - 1; t >= 0; --t)   // backward over the stages above the first cut
        stage_bwd<NX, NU>(L, P.t, P.m, Pp + L.px, Pp + L.pu, Q, R, plan.stage_off[t], plan.stage_off[t + 1], g, lane, true);
    if (warps == 1) __syncwarp();
    // x_0 <- initial state (cache.py:282), then forward down to the cut stage
    for (int k = threadIdx.x; k < L.nx; k += blockDim.x) Pp[L.px + k] = x0[blockIdx.x * L.nx + k];
    __syncthreads();
    for (int t = 0; t < plan.t_top; ++t)
        stage_fwd<NX, NU>(L, P.t, P.m, Pp + L.px, Pp + L.pu, R, plan.stage_off[t], plan.stage_off[t + 1], g, lane, true);
}

// ---- host launchers: pick the instantiation for (nx, nu) --------------------------------------------------------------
#define RB_DIMS(X) X(2, 1) X(3, 2) X(4, 2) X(8, 4) X(10, 5) X(20, 10)

cudaError_t sweep_kernels_set_smem(int bytes) {
    cudaError_t e = cudaSuccess;
#define RB_SET(NX, NU)                                                                                                   \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_sub_bwd<NX, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_sub_fwd<NX, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_top<NX, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    RB_DIMS(RB_SET)
    RB_SET(0, 0)
#undef RB_SET
    return e;
}

void launch_sweep_sub_bwd(dim3 grid, int threads, size_t smem, cudaStream_t st, const Params &P, const Ctrl *ctrl,
                          const SweepLevel &lv, const double *prim, double *q, double *r) {
#define RB_GO(NX, NU)                                                                                 \
    if (P.L.nx == NX && P.L.nu == NU) {                                                               \
        k_sweep_sub_bwd<NX, NU><<<grid, threads, smem, st>>>(P, ctrl, lv, prim, q, r);               \
        return;                                                                                       \
    }
    RB_DIMS(RB_GO)
#undef RB_GO
    k_sweep_sub_bwd<0, 0><<<grid, threads, smem, st>>>(P, ctrl, lv, prim, q, r);
}

void launch_sweep_sub_fwd(dim3 grid, int threads, size_t smem, cudaStream_t st, const Params &P, const Ctrl *ctrl,
                          const SweepLevel &lv, double *prim, const double *r) {
#define RB_GO(NX, NU)                                                                                 \
    if (P.L.nx == NX && P.L.nu == NU) {                                                               \
        k_sweep_sub_fwd<NX, NU><<<grid, threads, smem, st>>>(P, ctrl, lv, prim, r);                  \
        return;                                                                                       \
    }
    RB_DIMS(RB_GO)
#undef RB_GO
    k_sweep_sub_fwd<0, 0><<<grid, threads, smem, st>>>(P, ctrl, lv, prim, r);
}

void launch_sweep_top(int grid, int threads, size_t smem, cudaStream_t st, const Params &P, const Ctrl *ctrl,
                      const SweepPlan &plan, double *prim, double *q, double *r, const double *x0) {
#define RB_GO(NX, NU)                                                                                 \
    if (P.L.nx == NX && P.L.nu == NU) {                                                               \
        k_sweep_top<NX, NU><<<grid, threads, smem, st>>>(P, ctrl, plan, prim, q, r, x0);             \
        return;                                                                                       \
    }
    RB_DIMS(RB_GO)
#undef RB_GO
    k_sweep_top<0, 0><<<grid, threads, smem, st>>>(P, ctrl, plan, prim, q, r, x0);
}

}  // namespace rb
