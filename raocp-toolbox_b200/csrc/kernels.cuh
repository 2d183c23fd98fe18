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
__global__ void k_dyn_bwd(const __grid_constant__ Params P, const double *__restrict__ prim, double *__restrict__ q,
                          double *__restrict__ r, int lo, int hi);
__global__ void k_dyn_fwd(const __grid_constant__ Params P, double *__restrict__ prim, const double *__restrict__ r,
                          int lo, int hi);
__global__ void k_set_root(const __grid_constant__ Params P, double *__restrict__ prim, const double *__restrict__ x0);
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
__global__ void k_cone(int cone, int dim, const double *__restrict__ in, double *__restrict__ out);
__global__ void k_box(int dim, const double *__restrict__ in, const double *__restrict__ lo, const double *__restrict__ hi,
                      double *__restrict__ out, int *__restrict__ status);

// ---- fused.cu: Solver.chock's loop body --------------------------------------------------------------------------
// control block shared by the kernels of the fused loop
struct Ctrl {
    int done;        // set by k_check when the stopping test of solver.py:156-161 fires; later launches are no-ops
    int iters;       // iterations executed so far
    int status;      // bit 0: NaN met in a rectangle projection; bit 1: non-finite residual
    int pad;
};
__global__ void k_fused_primal(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                               const double *__restrict__ p_old, const double *__restrict__ d_old,
                               double *__restrict__ p_new, double alpha);
__global__ void k_fused_bwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                            const double *__restrict__ prim, double *__restrict__ q, double *__restrict__ r, int lo, int hi);
__global__ void k_fused_fwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl, double *__restrict__ prim,
                            const double *__restrict__ r, const double *__restrict__ x0, int lo, int hi);
__global__ void k_fused_dual(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl, const double *__restrict__ p_old,
                             const double *__restrict__ p_new, const double *__restrict__ d_old,
                             double *__restrict__ d_new, double alpha, double *__restrict__ slots);
__global__ void k_check(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl, double *__restrict__ slots,
                        double *__restrict__ last, double *__restrict__ hist, int hist_capacity, int max_iters,
                        double tol);

// ---- offline.cu --------------------------------------------------------------------------------------------------
struct ClassView {
    const int *child_ptr;   // [num_cls+1]
    const int *child_dyn;   // dyn table row of every child of the class representative
    const int *child_cls;   // class of that child, -1 = leaf (P = I)
    const int *level_list;  // classes of the level being processed
};
__global__ void k_offline_level(const __grid_constant__ Params P, ClassView cv, int level_begin, int level_count,
                                double *__restrict__ Ptab, double *__restrict__ Ktab, double *__restrict__ KTtab,
                                double *__restrict__ RinvTtab, int *__restrict__ status);
size_t offline_smem_bytes(int nx, int nu);

// lambda_max(L* L) pieces
__global__ void k_gram_eig(const __grid_constant__ Params P, const int *__restrict__ grp_ptr, const int *__restrict__ grp_idx,
                           int kind, int num_groups, double *__restrict__ work, double *__restrict__ out_max);
__global__ void k_risk_block_eig(const __grid_constant__ Params P, double *__restrict__ out_max);

}  // namespace rb
