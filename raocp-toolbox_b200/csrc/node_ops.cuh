// node_ops.cuh -- per-node device bodies shared by the stand-alone kernels (ops.cu) and the fused loop (fused.cu).
#pragma once
#include "common.cuh"

namespace rb {

// Backward DP step of the projection onto the dynamics set at one node (reference cache.py:264-280).
//   leaf:     q_i = -xbar_i
//   nonleaf:  r_i = ubar_i - sum_j B_j' q_j;   q_i = -xbar_i - K_i' r_i + sum_j A_j' q_j
// which equals the reference's q_i = -xbar_i + K_i'(d_i - ubar_i) + sum_j (A_j+B_jK_i)'(P_j B_j d_i + q_j) with
// d_i = R~_i^-1 r_i (DESIGN.md "DP identities").  sm: 4 warp-private rows of kMaxDim doubles.
__device__ __forceinline__ void dyn_bwd_node(const Layout &L, const Topo &T, const Tabs &M,
                                             const double *__restrict__ Pp, double *__restrict__ Q,
                                             double *__restrict__ R, int node, int lane, double (*sm)[kMaxDim]) {
    const int nx = L.nx, nu = L.nu;
    if (node >= L.m) {
        for (int k = lane; k < nx; k += 32) Q[(long long)node * nx + k] = -Pp[L.px + (long long)node * nx + k];
        return;
    }
    double *qj = sm[0], *sig = sm[1], *aq = sm[2], *rv = sm[3];
    for (int k = lane; k < nu; k += 32) sig[k] = 0.0;
    for (int k = lane; k < nx; k += 32) aq[k] = 0.0;
    const int c0 = T.child_first[node], cc = T.child_count[node];
    for (int j = c0; j < c0 + cc; ++j) {
        for (int k = lane; k < nx; k += 32) qj[k] = Q[(long long)j * nx + k];
        __syncwarp();
        const int di = T.dyn_idx[j];
        mv_acc(M.B + (long long)di * nx * nu, qj, nu, nx, sig, 1.0, lane);   // B' q  (B row-major is (B')^T)
        mv_acc(M.A + (long long)di * nx * nx, qj, nx, nx, aq, 1.0, lane);    // A' q
        __syncwarp();
    }
    for (int k = lane; k < nu; k += 32) {
        const double rk = Pp[L.pu + (long long)node * nu + k] - sig[k];
        rv[k] = rk;
        R[(long long)node * nu + k] = rk;
    }
    __syncwarp();
    const int cl = T.cls[node];
    for (int k = lane; k < nx; k += 32) aq[k] -= Pp[L.px + (long long)node * nx + k];
    mv_acc(M.K + (long long)cl * nu * nx, rv, nx, nu, aq, -1.0, lane);       // - K' r  (K row-major is (K')^T)
    for (int k = lane; k < nx; k += 32) Q[(long long)node * nx + k] = aq[k];
    __syncwarp();
}

// Forward DP step at one nonleaf node (reference cache.py:282-288):
//   u_i = K_i x_i + R~_i^-1 r_i;   x_j = A_j x_i + B_j u_i   ( = (A_j+B_jK_i) x_i + B_j d_i of the reference )
__device__ __forceinline__ void dyn_fwd_node(const Layout &L, const Topo &T, const Tabs &M, double *__restrict__ Pp,
                                             const double *__restrict__ R, int node, int lane, double (*sm)[kMaxDim]) {
    const int nx = L.nx, nu = L.nu;
    double *xi = sm[0], *rv = sm[1], *ui = sm[2], *xj = sm[3];
    for (int k = lane; k < nx; k += 32) xi[k] = Pp[L.px + (long long)node * nx + k];
    for (int k = lane; k < nu; k += 32) rv[k] = R[(long long)node * nu + k];
    __syncwarp();
    const int cl = T.cls[node];
    mv_set(M.RinvT + (long long)cl * nu * nu, rv, nu, nu, ui, lane);         // d = R~^-1 r
    __syncwarp();
    mv_acc(M.KT + (long long)cl * nx * nu, xi, nu, nx, ui, 1.0, lane);       // + K x
    __syncwarp();
    for (int k = lane; k < nu; k += 32) Pp[L.pu + (long long)node * nu + k] = ui[k];
    const int c0 = T.child_first[node], cc = T.child_count[node];
    for (int j = c0; j < c0 + cc; ++j) {
        const int di = T.dyn_idx[j];
        mv_set(M.AT + (long long)di * nx * nx, xi, nx, nx, xj, lane);
        __syncwarp();
        mv_acc(M.BT + (long long)di * nu * nx, ui, nx, nu, xj, 1.0, lane);
        __syncwarp();
        for (int k = lane; k < nx; k += 32) Pp[L.px + (long long)j * nx + k] = xj[k];
        __syncwarp();
    }
}

// Projection of (y_i, tau_ch(i), s_ch(i)) onto ker [E' -I -I] (reference cache.py:290-317), in place.  For AVaR
// (risks.py:28-35) M = [a I, -I, 1, -I, -I] and M M' = (a^2+3) I + 1 1', hence proj = v - M'(M M')^-1 M v in closed form.
__device__ __forceinline__ void kernel_projection(const Params &P, double *Pp, int node, int lane) {
    const Layout &L = P.L;
    const int c0 = P.t.child_first[node], cc = P.t.child_count[node];
    const int yo = P.t.yoff[node];
    const double a = P.t.risk_alpha[node];
    const double ylast = Pp[L.py + yo + 2 * cc];
    double rsum = 0.0;
    for (int e = lane; e < cc; e += 32)
        rsum += a * Pp[L.py + yo + e] - Pp[L.py + yo + cc + e] + ylast - Pp[L.ptau + c0 + e] - Pp[L.ps + c0 + e];
    rsum = warp_sum(rsum);
    const double den = a * a + 3.0;
    const double shift = rsum / (den + (double)cc);
    double wsum = 0.0;
    for (int e = lane; e < cc; e += 32) {
        const double ya = Pp[L.py + yo + e], yb = Pp[L.py + yo + cc + e];
        const double tj = Pp[L.ptau + c0 + e], sj = Pp[L.ps + c0 + e];
        const double w = ((a * ya - yb + ylast - tj - sj) - shift) / den;
        Pp[L.py + yo + e] = ya - a * w;
        Pp[L.py + yo + cc + e] = yb + w;
        Pp[L.ptau + c0 + e] = tj + w;
        Pp[L.ps + c0 + e] = sj + w;
        wsum += w;
    }
    wsum = warp_sum(wsum);
    if (lane == 0) Pp[L.py + yo + 2 * cc] = ylast - wsum;
}


}  // namespace rb

namespace rb {
// convenience overloads on the full parameter block
__device__ __forceinline__ void dyn_bwd_node(const Params &P, const double *__restrict__ Pp, double *__restrict__ Q,
                                             double *__restrict__ R, int node, int lane, double (*sm)[kMaxDim]) {
    dyn_bwd_node(P.L, P.t, P.m, Pp, Q, R, node, lane, sm);
}
__device__ __forceinline__ void dyn_fwd_node(const Params &P, double *__restrict__ Pp, const double *__restrict__ R,
                                             int node, int lane, double (*sm)[kMaxDim]) {
    dyn_fwd_node(P.L, P.t, P.m, Pp, R, node, lane, sm);
}
}  // namespace rb
