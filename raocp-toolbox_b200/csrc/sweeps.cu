// sweeps.cu -- projection onto the dynamics set (reference cache.py:259-288) in THREE launches per iteration instead
// of one per stage:
//
//   k_sweep_sub_bwd : one CTA per subtree below the cut stage t_s (the first stage with >= 64 nodes).  In breadth-first
//                     numbering the descendants of a node form one contiguous node range per stage, so the CTA walks
//                     its ranges from the leaves up with a block barrier per stage; below the stopping time of a
//                     Markov tree the ranges are single nodes (chains) and each warp simply walks its chain.
//   k_sweep_top     : one CTA per problem instance for the few nodes above the cut: backward to the root, then
//                     forward again down to the cut stage.
//   k_sweep_sub_fwd : the subtrees again, forward.
//
// The mode-indexed dynamics tables (A, A', B, B') are staged in shared memory when they fit (they are read by every
// node); the class-indexed K, K', R~^-1 stream through L1/L2 (or from HBM when every node is its own class).
#include "kernels.cuh"
#include "node_ops.cuh"

namespace rb {

__device__ __forceinline__ void stage_tables(const Params &P, const SweepPlan &plan, Tabs &tb, double *smem_tabs) {
    tb = P.m;
    if (!plan.tabs_in_smem) return;
    const int nx = P.L.nx, nu = P.L.nu;
    const long long na = (long long)plan.num_dyn * nx * nx, nb = (long long)plan.num_dyn * nx * nu;
    double *sA = smem_tabs, *sAT = sA + na, *sB = sAT + na, *sBT = sB + nb;
    for (long long i = threadIdx.x; i < na; i += blockDim.x) {
        sA[i] = P.m.A[i];
        sAT[i] = P.m.AT[i];
    }
    for (long long i = threadIdx.x; i < nb; i += blockDim.x) {
        sB[i] = P.m.B[i];
        sBT[i] = P.m.BT[i];
    }
    tb.A = sA;
    tb.AT = sAT;
    tb.B = sB;
    tb.BT = sBT;
    __syncthreads();
}

__global__ void k_sweep_sub_bwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl, SweepPlan plan,
                                const double *__restrict__ prim, double *__restrict__ q, double *__restrict__ r) {
    if (ctrl && ctrl->done) return;
    extern __shared__ double dyn_smem[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double(*rows)[kMaxDim] = reinterpret_cast<double(*)[kMaxDim]>(dyn_smem) + warp * 4;
    Tabs tb;
    stage_tables(P, plan, tb, dyn_smem + (size_t)warps * 4 * kMaxDim);
    const double *Pp = prim + (long long)blockIdx.y * P.L.np_pad;
    double *Q = q + (long long)blockIdx.y * P.L.n * P.L.nx;
    double *R = r + (long long)blockIdx.y * P.L.m * P.L.nu;
    const int *lo = plan.sub_lo + (long long)blockIdx.x * plan.depth, *hi = plan.sub_hi + (long long)blockIdx.x * plan.depth;
    for (int d = plan.depth - 1; d >= 0; --d) {
        for (int node = lo[d] + warp; node < hi[d]; node += warps) dyn_bwd_node(P.L, P.t, tb, Pp, Q, R, node, lane, rows);
        if (warps > 1) __syncthreads();
    }
}

__global__ void k_sweep_sub_fwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl, SweepPlan plan,
                                double *__restrict__ prim, const double *__restrict__ r) {
    if (ctrl && ctrl->done) return;
    extern __shared__ double dyn_smem[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double(*rows)[kMaxDim] = reinterpret_cast<double(*)[kMaxDim]>(dyn_smem) + warp * 4;
    Tabs tb;
    stage_tables(P, plan, tb, dyn_smem + (size_t)warps * 4 * kMaxDim);
    double *Pp = prim + (long long)blockIdx.y * P.L.np_pad;
    const double *R = r + (long long)blockIdx.y * P.L.m * P.L.nu;
    const int *lo = plan.sub_lo + (long long)blockIdx.x * plan.depth, *hi = plan.sub_hi + (long long)blockIdx.x * plan.depth;
    for (int d = 0; d < plan.depth - 1; ++d) {   // the last stage of a subtree are leaves
        for (int node = lo[d] + warp; node < hi[d]; node += warps)
            if (node < P.L.m) dyn_fwd_node(P.L, P.t, tb, Pp, R, node, lane, rows);
        if (warps > 1) __syncthreads();
    }
}

__global__ void k_sweep_top(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl, SweepPlan plan,
                            double *__restrict__ prim, double *__restrict__ q, double *__restrict__ r,
                            const double *__restrict__ x0) {
    if (ctrl && ctrl->done) return;
    extern __shared__ double dyn_smem[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double(*rows)[kMaxDim] = reinterpret_cast<double(*)[kMaxDim]>(dyn_smem) + warp * 4;
    Tabs tb;
    stage_tables(P, plan, tb, dyn_smem + (size_t)warps * 4 * kMaxDim);
    double *Pp = prim + (long long)blockIdx.x * P.L.np_pad;
    double *Q = q + (long long)blockIdx.x * P.L.n * P.L.nx;
    double *R = r + (long long)blockIdx.x * P.L.m * P.L.nu;
    // backward over the stages above the cut
    for (int t = plan.t_s - 1; t >= 0; --t) {
        for (int node = plan.stage_off[t] + warp; node < plan.stage_off[t + 1]; node += warps)
            dyn_bwd_node(P.L, P.t, tb, Pp, Q, R, node, lane, rows);
        __syncthreads();
    }
    // x_0 <- initial state (cache.py:282), then forward down to the cut stage
    for (int k = threadIdx.x; k < P.L.nx; k += blockDim.x) Pp[P.L.px + k] = x0[blockIdx.x * P.L.nx + k];
    __syncthreads();
    const int last = plan.t_s < P.L.num_stages - 1 ? plan.t_s : P.L.num_stages - 1;
    for (int t = 0; t < last; ++t) {
        for (int node = plan.stage_off[t] + warp; node < plan.stage_off[t + 1]; node += warps)
            dyn_fwd_node(P.L, P.t, tb, Pp, R, node, lane, rows);
        __syncthreads();
    }
}

}  // namespace rb
