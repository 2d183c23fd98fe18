// node_ops.cuh -- per-node device bodies of the DP sweeps and of the kernel projection, shared by ops.cu / sweeps.cu.
#pragma once
#include "common.cuh"

namespace rb {

// Backward DP step of the projection onto the dynamics set at one node (reference cache.py:264-280).
//   leaf:     q_i = -xbar_i
//   nonleaf:  r_i = ubar_i - sum_j B_j' q_j;   q_i = -xbar_i - K_i' r_i + sum_j A_j' q_j
// which equals the reference's q_i = -xbar_i + K_i'(d_i - ubar_i) + sum_j (A_j+B_jK_i)'(P_j B_j d_i + q_j) with
// d_i = R~_i^-1 r_i (DESIGN.md "DP identities").  X / U: row bases of xbar / ubar (index node*nx+k, node*nu+k).
// scratch: 2*(nx+nu) warp-private doubles.
__device__ __forceinline__ void dyn_bwd_node(const Layout &L, const Topo &T, const Tabs &M, const double *__restrict__ X,
                                             const double *__restrict__ U, double *__restrict__ Q,
                                             double *__restrict__ R, int node, int lane, double *scratch) {
    const int nx = L.nx, nu = L.nu, nxu = L.nxu;
    if (node >= L.m) {
        for (int k = lane; k < nx; k += 32) Q[node * nx + k] = -X[node * nx + k];
        return;
    }
    double *qj = scratch, *acc = scratch + nxu;   // qj: nx (+ rv: nu behind it), acc: nxu
    double *rv = qj + nx;
    for (int k = lane; k < nxu; k += 32) acc[k] = 0.0;
    const int c0 = T.child_first[node], cc = T.child_count[node];
    for (int j = c0; j < c0 + cc; ++j) {
        for (int k = lane; k < nx; k += 32) qj[k] = Q[j * nx + k];
        __syncwarp();
        const double *C = M.ABcat + (long long)T.dyn_idx[j] * nx * nxu;
        for (int k = lane; k < nxu; k += 32) acc[k] += mv_row(C, qj, nxu, nx, k);   // [A'q ; B'q]
        __syncwarp();
    }
    for (int k = lane; k < nu; k += 32) {
        const double rk = U[node * nu + k] - acc[nx + k];
        rv[k] = rk;
        R[node * nu + k] = rk;
    }
    __syncwarp();
    const double *Kc = M.K + (long long)T.cls[node] * nu * nx;
    for (int k = lane; k < nx; k += 32) Q[node * nx + k] = acc[k] - X[node * nx + k] - mv_row(Kc, rv, nx, nu, k);
    __syncwarp();
}

// out[a] (a < rows) = sum_l MT[l*rows + a] v[l] with the reduction split over G = 32/rows lane groups when the output
// is narrow (rows <= 16), so that narrow products (u = K x + R~^-1 r, nu lanes) still use the whole warp.
// part: 32 warp-private doubles.  Result valid in lanes a < rows (returned), all lanes must call.
__device__ __forceinline__ double mv_split(const double *__restrict__ MT, const double *v, int rows, int cols, int lane,
                                           double *part) {
    if (rows > 16) return lane < rows ? mv_row(MT, v, rows, cols, lane) : 0.0;   // (rows <= 32 here)
    const int G = 32 / rows, g = lane / rows, a = lane - g * rows;
    const int seg = (cols + G - 1) / G;
    double p = 0.0;
    if (g < G) {
        const int l0 = g * seg, l1 = min(cols, l0 + seg);
        double p1 = 0.0;
        int l = l0;
        for (; l + 2 <= l1; l += 2) {
            p = fma(MT[(long long)l * rows + a], v[l], p);
            p1 = fma(MT[(long long)(l + 1) * rows + a], v[l + 1], p1);
        }
        if (l < l1) p = fma(MT[(long long)l * rows + a], v[l], p);
        p += p1;
    }
    part[lane] = p;
    __syncwarp();
    double out = 0.0;
    if (lane < rows)
        for (int gg = 0; gg < G; ++gg) out += part[gg * rows + lane];
    __syncwarp();
    return out;
}

// Forward DP step at one nonleaf node (reference cache.py:282-288):
//   u_i = K_i x_i + R~_i^-1 r_i;   x_j = A_j x_i + B_j u_i   ( = (A_j+B_jK_i) x_i + B_j d_i of the reference )
// X / U are the x / u row bases (x_i is read, u_i and the children's x_j are written).
// scratch: (nx+nu) + 32 warp-private doubles.
__device__ __forceinline__ void dyn_fwd_node(const Layout &L, const Topo &T, const Tabs &M, double *__restrict__ X,
                                             double *__restrict__ U, const double *__restrict__ R, int node, int lane,
                                             double *scratch) {
    const int nx = L.nx, nu = L.nu, nxu = L.nxu;
    double *v = scratch, *part = scratch + nxu;   // v = [x_i ; r_i] then [x_i ; u_i]
    for (int k = lane; k < nx; k += 32) v[k] = X[node * nx + k];
    for (int k = lane; k < nu; k += 32) v[nx + k] = R[node * nu + k];
    __syncwarp();
    const double *KR = M.KRcatT + (long long)T.cls[node] * nxu * nu;
    if (nu <= 32) {
        const double ua = mv_split(KR, v, nu, nxu, lane, part);
        if (lane < nu) {
            v[nx + lane] = ua;            // r is dead after mv_split's internal barrier
            U[node * nu + lane] = ua;
        }
    } else {
        double u0 = mv_row(KR, v, nu, nxu, lane), u1 = lane + 32 < nu ? mv_row(KR, v, nu, nxu, lane + 32) : 0.0;
        __syncwarp();
        v[nx + lane] = u0;
        U[node * nu + lane] = u0;
        if (lane + 32 < nu) {
            v[nx + lane + 32] = u1;
            U[node * nu + lane + 32] = u1;
        }
    }
    __syncwarp();
    const int c0 = T.child_first[node], cc = T.child_count[node];
    for (int j = c0; j < c0 + cc; ++j) {
        const double *C = M.ABcatT + (long long)T.dyn_idx[j] * nxu * nx;
        for (int k = lane; k < nx; k += 32) X[j * nx + k] = mv_row(C, v, nx, nxu, k);
    }
    __syncwarp();
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
