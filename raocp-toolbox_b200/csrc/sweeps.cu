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
// The small mode-indexed tables (A, B concatenations) are shared by all nodes and stay L1-resident; the class-indexed
// K, [K R~^-1] stream through L1/L2 (or from HBM when every node is its own class).
#include "kernels.cuh"
#include "node_ops.cuh"

namespace rb {

__global__ void __launch_bounds__(512) k_sweep_sub_bwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                      SweepLevel lv, const double *__restrict__ prim,
                                                      double *__restrict__ q, double *__restrict__ r) {
    if (ctrl && ctrl->done) return;
    extern __shared__ double dyn_smem[];
    const Layout &L = P.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *scratch = dyn_smem + (size_t)warp * (2 * L.nxu + 32);
    const double *Pp = prim + (long long)blockIdx.y * L.np_pad;
    double *Q = q + (long long)blockIdx.y * L.n * L.nx;
    double *R = r + (long long)blockIdx.y * L.m * L.nu;
    const int sub = blockIdx.x * lv.subs_per_cta + warp / lv.warps_per_sub, wl = warp % lv.warps_per_sub;
    const bool live = sub < lv.num_sub;
    const int *lo = lv.lo + (long long)(live ? sub : 0) * lv.depth, *hi = lv.hi + (long long)(live ? sub : 0) * lv.depth;
    for (int d = lv.depth - 1; d >= 0; --d) {
        if (live)
            for (int node = lo[d] + wl; node < hi[d]; node += lv.warps_per_sub)
                dyn_bwd_node(L, P.t, P.m, Pp + L.px, Pp + L.pu, Q, R, node, lane, scratch);
        if (lv.warps_per_sub > 1) __syncthreads();
    }
}

__global__ void __launch_bounds__(512) k_sweep_sub_fwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                      SweepLevel lv, double *__restrict__ prim,
                                                      const double *__restrict__ r) {
    if (ctrl && ctrl->done) return;
    extern __shared__ double dyn_smem[];
    const Layout &L = P.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *scratch = dyn_smem + (size_t)warp * (2 * L.nxu + 32);
    double *Pp = prim + (long long)blockIdx.y * L.np_pad;
    const double *R = r + (long long)blockIdx.y * L.m * L.nu;
    const int sub = blockIdx.x * lv.subs_per_cta + warp / lv.warps_per_sub, wl = warp % lv.warps_per_sub;
    const bool live = sub < lv.num_sub;
    const int *lo = lv.lo + (long long)(live ? sub : 0) * lv.depth, *hi = lv.hi + (long long)(live ? sub : 0) * lv.depth;
    for (int d = 0; d < lv.depth; ++d) {
        if (live)
            for (int node = lo[d] + wl; node < hi[d]; node += lv.warps_per_sub)
                if (node < L.m) dyn_fwd_node(L, P.t, P.m, Pp + L.px, Pp + L.pu, R, node, lane, scratch);
        if (lv.warps_per_sub > 1) __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) k_sweep_top(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                   SweepPlan plan, double *__restrict__ prim, double *__restrict__ q,
                                                   double *__restrict__ r, const double *__restrict__ x0) {
    if (ctrl && ctrl->done) return;
    extern __shared__ double dyn_smem[];
    const Layout &L = P.L;
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *scratch = dyn_smem + (size_t)warp * (2 * L.nxu + 32);
    double *Pp = prim + (long long)blockIdx.x * L.np_pad;
    double *Q = q + (long long)blockIdx.x * L.n * L.nx;
    double *R = r + (long long)blockIdx.x * L.m * L.nu;
    for (int t = plan.t_top - 1; t >= 0; --t) {   // backward over the stages above the first cut
        for (int node = plan.stage_off[t] + warp; node < plan.stage_off[t + 1]; node += warps)
            dyn_bwd_node(L, P.t, P.m, Pp + L.px, Pp + L.pu, Q, R, node, lane, scratch);
        __syncthreads();
    }
    // x_0 <- initial state (cache.py:282), then forward down to the cut stage
    for (int k = threadIdx.x; k < L.nx; k += blockDim.x) Pp[L.px + k] = x0[blockIdx.x * L.nx + k];
    __syncthreads();
    for (int t = 0; t < plan.t_top; ++t) {
        for (int node = plan.stage_off[t] + warp; node < plan.stage_off[t + 1]; node += warps)
            if (node < L.m) dyn_fwd_node(L, P.t, P.m, Pp + L.px, Pp + L.pu, R, node, lane, scratch);
        __syncthreads();
    }
}

}  // namespace rb
