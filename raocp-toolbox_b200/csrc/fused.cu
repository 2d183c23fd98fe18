// fused.cu -- the node-parallel passes of Solver.chock's loop body (reference solver.py:124-161) as tiled, fused
// sm_100a kernels (FP64), plus the device-side stopping test.
//
// One Chambolle-Pock iteration =
//   k_primal_tile  : pbar = p - alpha L* d  (solver.py:27-39)  +  s_0 -= alpha (cache.py:253-257)
//                    + projection of (y, tau, s) onto the risk kernels (cache.py:290-317)
//   sweeps.cu      : projection of (xbar, ubar) onto the dynamics set (cache.py:259-288), three launches
//   k_dual_tile    : dbar = d + alpha L(2 p+ - p) (solver.py:44-58), prox of g* (cache.py:321-393) and all six
//                    residual inf-norms of solver.py:63-95,137-141 in ONE pass over the duals
//   k_check        : stopping test of solver.py:156-161 on the device; once it fires every later launch is a no-op,
//                    so the host can enqueue iterations without synchronising and still stop at the exact iteration.
//
// Tiling: a CTA owns a run of consecutive nodes and, because children are numbered contiguously, a run of
// consecutive edges.  Every input segment of the tile is therefore ONE contiguous chunk of the node-major layout:
// the CTA first streams all chunks into shared memory with all threads (many independent, fully coalesced loads in
// flight -- the passes are HBM-bound), then one warp per node does the small dense algebra out of shared memory and
// writes its output rows straight back to HBM.  Several CTAs are resident per SM, so one tile's loads overlap another
// tile's arithmetic.  Iterates live in two device copies that swap roles every iteration; only the residual norms
// leave the device.
#include "kernels.cuh"
#include "node_ops.cuh"

namespace rb {

// segment base pointers; every pointer is indexed with GLOBAL indices (node * nx + k, yoff[i] + e, edge j - 1, leaf
// index ...) whether it points into HBM or into a re-based shared-memory chunk
struct PrimalView {
    const double *x, *u, *y, *tau, *s;
};
struct DualView {
    const double *d1, *d2, *d3, *d4, *d5, *d6, *d7, *d11, *d12, *d13, *d14;
    const double *d2c;   // d2 of the CHILDREN of the tile's nodes (primal pass only)
};
struct PrimalOut {
    double *x, *u, *y, *tau, *s;
};
struct DualOut {
    double *d1, *d2, *d3, *d4, *d5, *d6, *d7, *d11, *d12, *d13, *d14;
};

__device__ __forceinline__ PrimalView primal_view(const Layout &L, const double *p) {
    return {p + L.px, p + L.pu, p + L.py, p + L.ptau, p + L.ps};
}
__device__ __forceinline__ PrimalOut primal_out(const Layout &L, double *p) {
    return {p + L.px, p + L.pu, p + L.py, p + L.ptau, p + L.ps};
}
__device__ __forceinline__ DualView dual_view(const Layout &L, const double *d) {
    return {d + L.d1, d + L.d2, d + L.d3, d + L.d4, d + L.d5, d + L.d6, d + L.d7, d + L.d11, d + L.d12, d + L.d13,
            d + L.d14, d + L.d2};
}
__device__ __forceinline__ DualOut dual_out(const Layout &L, double *d) {
    return {d + L.d1, d + L.d2, d + L.d3, d + L.d4, d + L.d5, d + L.d6, d + L.d7, d + L.d11, d + L.d12, d + L.d13, d + L.d14};
}

// bump allocator over the CTA's dynamic shared memory; stage() copies [first, first+count) of a segment with all threads
// and returns a pointer re-based so that the segment's GLOBAL indices work
struct Stager {
    double *cursor;
    __device__ __forceinline__ const double *stage(const double *seg, long long first, long long count) {
        if (count <= 0) return seg;
        double *dst = cursor;
        cursor += (count + 1) & ~1LL;
        const double *src = seg + first;
        for (long long i = threadIdx.x; i < count; i += blockDim.x) dst[i] = src[i];
        return dst - first;
    }
};

// cost-matrix application with a diagonal fast path (diagonal weights are the common case in MPC; the tables are
// classified once at rb_create)
__device__ __forceinline__ void cost_mv2(const double *__restrict__ MT, int diag, const double *v, const double *w, int dim,
                                         int k, double &rv, double &rw) {
    if (diag) {
        const double m = MT[(long long)k * dim + k];
        rv = m * v[k];
        rw = m * w[k];
    } else {
        mv_row2(MT, v, w, dim, dim, k, rv, rw);
    }
}
__device__ __forceinline__ double cost_mv(const double *__restrict__ MT, int diag, const double *v, int dim, int k) {
    return diag ? MT[(long long)k * dim + k] * v[k] : mv_row(MT, v, dim, dim, k);
}

// ====================================================================================================================
// primal pass
// ====================================================================================================================
// nonleaf node: xbar_i, ubar_i, ybar_i, and the children's taubar_j, sbar_j (edge quantities are owned by the parent's
// warp) with the kernel projection (cache.py:290-317) fused in.  s_0 / tau_0 are handled by the caller.
// rows: 4 warp-private shared rows of rowlen doubles.
__device__ __forceinline__ void primal_nonleaf_node(const Layout &L, const Topo &T, const Tabs &M, const PrimalView &po,
                                                    const DualView &d, const PrimalOut &pn, double alpha, int node,
                                                    int lane, double *rows, int rowlen) {
    const int nx = L.nx, nu = L.nu;
    double *v3 = rows, *v4 = rows + rowlen, *ax = rows + 2 * rowlen, *au = rows + 3 * rowlen;
    const int c0 = T.child_first[node], cc = T.child_count[node];
    for (int k = lane; k < nx; k += 32) ax[k] = L.has_nl_rect ? d.d7[(long long)node * L.nxu + k] : 0.0;
    for (int k = lane; k < nu; k += 32) au[k] = L.has_nl_rect ? d.d7[(long long)node * L.nxu + nx + k] : 0.0;
    for (int j = c0; j < c0 + cc; ++j) {
        const long long e = j - 1;
        for (int k = lane; k < nx; k += 32) v3[k] = d.d3[e * nx + k];
        for (int k = lane; k < nu; k += 32) v4[k] = d.d4[e * nu + k];
        __syncwarp();
        const int ci = T.cost_idx[j];
        for (int k = lane; k < nx; k += 32) ax[k] += cost_mv(M.sqT + (long long)ci * nx * nx, M.sq_diag, v3, nx, k);
        for (int k = lane; k < nu; k += 32) au[k] += cost_mv(M.srT + (long long)ci * nu * nu, M.sr_diag, v4, nu, k);
        __syncwarp();
    }
    for (int k = lane; k < nx; k += 32) {
        const long long idx = (long long)node * nx + k;
        pn.x[idx] = po.x[idx] - alpha * ax[k];
    }
    for (int k = lane; k < nu; k += 32) {
        const long long idx = (long long)node * nu + k;
        pn.u[idx] = po.u[idx] - alpha * au[k];
    }
    const double d2v = d.d2[node];
    const int yo = T.yoff[node];
    const double a = T.risk_alpha[node];
    // For AVaR M = [a I, -I, 1, -I, -I] and M M' = (a^2+3) I + 1 1', so proj = v - M'(M M')^-1 M v in closed form
    const double ylast_bar = po.y[yo + 2 * cc] - alpha * (d.d1[yo + 2 * cc] - d2v);
    double rsum = 0.0;
    for (int e = lane; e < cc; e += 32) {
        const int j = c0 + e;
        const double ya = po.y[yo + e] - alpha * (d.d1[yo + e] - T.cond_prob[j] * d2v);
        const double yb = po.y[yo + cc + e] - alpha * d.d1[yo + cc + e];
        const double tj = po.tau[j] - alpha * (0.5 * (d.d5[j - 1] + d.d6[j - 1]));
        const double lts = j < L.m ? d.d2c[j] : 0.5 * (d.d12[j - L.m] + d.d13[j - L.m]);
        const double sj = po.s[j] - alpha * lts;
        rsum += a * ya - yb + ylast_bar - tj - sj;
    }
    rsum = warp_sum(rsum);
    const double den = a * a + 3.0;
    const double shift = rsum / (den + (double)cc);
    double wsum = 0.0;
    for (int e = lane; e < cc; e += 32) {   // same arithmetic as above, bit for bit
        const int j = c0 + e;
        const double ya = po.y[yo + e] - alpha * (d.d1[yo + e] - T.cond_prob[j] * d2v);
        const double yb = po.y[yo + cc + e] - alpha * d.d1[yo + cc + e];
        const double tj = po.tau[j] - alpha * (0.5 * (d.d5[j - 1] + d.d6[j - 1]));
        const double lts = j < L.m ? d.d2c[j] : 0.5 * (d.d12[j - L.m] + d.d13[j - L.m]);
        const double sj = po.s[j] - alpha * lts;
        const double w = ((a * ya - yb + ylast_bar - tj - sj) - shift) / den;
        pn.y[yo + e] = ya - a * w;
        pn.y[yo + cc + e] = yb + w;
        pn.tau[j] = tj + w;
        pn.s[j] = sj + w;
        wsum += w;
    }
    wsum = warp_sum(wsum);
    if (lane == 0) pn.y[yo + 2 * cc] = ylast_bar - wsum;
    __syncwarp();
}

__device__ __forceinline__ void primal_leaf_node(const Layout &L, const Topo &T, const Tabs &M, const PrimalView &po,
                                                 const DualView &d, const PrimalOut &pn, double alpha, int node, int lane,
                                                 double *rows) {
    const int nx = L.nx;
    const long long li = node - L.m;
    double *v = rows;
    for (int k = lane; k < nx; k += 32) v[k] = d.d11[li * nx + k];
    __syncwarp();
    const double *sqfT = M.sqfT + (long long)T.leafcost_idx[li] * nx * nx;
    for (int k = lane; k < nx; k += 32) {
        double acc = cost_mv(sqfT, M.sqf_diag, v, nx, k);
        if (L.has_leaf_rect) acc += d.d14[li * nx + k];
        const long long idx = (long long)node * nx + k;
        pn.x[idx] = po.x[idx] - alpha * acc;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(256) k_primal_tile(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                    TilePlan plan, const double *__restrict__ p_old,
                                                    const double *__restrict__ d_old, double *__restrict__ p_new) {
    if (ctrl->done) return;
    const double alpha = ctrl->alpha;
    const Layout &L = P.L;
    extern __shared__ double dsm[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lo = plan.tiles[blockIdx.x].x, hi = plan.tiles[blockIdx.x].y;
    const double *Pg = p_old + (long long)blockIdx.y * L.np_pad;
    const double *Dg = d_old + (long long)blockIdx.y * L.nd_pad;
    const PrimalOut pn = primal_out(L, p_new + (long long)blockIdx.y * L.np_pad);
    const int nx = L.nx, nu = L.nu;
    const int rowlen = plan.rowlen;
    double *rows = dsm + (size_t)warp * 4 * rowlen;
    Stager st{dsm + (size_t)warps * 4 * rowlen};
    PrimalView po = primal_view(L, Pg);
    DualView d = dual_view(L, Dg);
    if (lo < L.m) {
        const long long nN = hi - lo;
        const long long c0 = P.t.child_first[lo], c1 = P.t.child_first[hi - 1] + P.t.child_count[hi - 1], nE = c1 - c0;
        const long long y0 = P.t.yoff[lo], ny = P.t.yoff[hi] - y0;
        po.x = st.stage(po.x, (long long)lo * nx, nN * nx);
        po.u = st.stage(po.u, (long long)lo * nu, nN * nu);
        po.y = st.stage(po.y, y0, ny);
        po.tau = st.stage(po.tau, c0, nE);
        po.s = st.stage(po.s, c0, nE);
        d.d1 = st.stage(d.d1, y0, ny);
        d.d2 = st.stage(d.d2, lo, nN);
        const long long cm = c1 < L.m ? c1 : L.m;             // nonleaf children: [c0, cm); leaf children: [max(c0,m), c1)
        d.d2c = st.stage(d.d2c, c0, cm - c0);
        if (c1 > L.m) {
            const long long l0 = (c0 > L.m ? c0 : L.m) - L.m, l1 = c1 - L.m;
            d.d12 = st.stage(d.d12, l0, l1 - l0);
            d.d13 = st.stage(d.d13, l0, l1 - l0);
        }
        d.d3 = st.stage(d.d3, (c0 - 1) * nx, nE * nx);
        d.d4 = st.stage(d.d4, (c0 - 1) * nu, nE * nu);
        d.d5 = st.stage(d.d5, c0 - 1, nE);
        d.d6 = st.stage(d.d6, c0 - 1, nE);
        if (L.has_nl_rect) d.d7 = st.stage(d.d7, (long long)lo * L.nxu, nN * L.nxu);
        __syncthreads();
        for (int node = lo + warp; node < hi; node += warps) {
            primal_nonleaf_node(L, P.t, P.m, po, d, pn, alpha, node, lane, rows, rowlen);
            if (node == 0 && lane == 0) {
                pn.s[0] = (Pg[L.ps] - alpha * d.d2[0]) - alpha;   // s_0: half step, then prox of alpha * identity
                pn.tau[0] = Pg[L.ptau] - alpha * Pg[L.ptau];      // tau_0 (always 0; same arithmetic as the reference)
            }
        }
    } else {
        const long long nN = hi - lo, l0 = lo - L.m;
        po.x = st.stage(po.x, (long long)lo * nx, nN * nx);
        d.d11 = st.stage(d.d11, l0 * nx, nN * nx);
        if (L.has_leaf_rect) d.d14 = st.stage(d.d14, l0 * nx, nN * nx);
        __syncthreads();
        for (int node = lo + warp; node < hi; node += warps) primal_leaf_node(L, P.t, P.m, po, d, pn, alpha, node, lane, rows);
    }
}

// ====================================================================================================================
// dual pass + residuals
// ====================================================================================================================
// residual bookkeeping: six running maxima per lane (xi0, xi1, xi2, delta0, delta1, delta2)
struct Resid {
    double v[6];
    int nan;
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < 6; ++i) v[i] = 0.0;
        nan = 0;
    }
    __device__ __forceinline__ void put(int slot, double x) {
        nan |= (x != x);
        v[slot] = fmax(v[slot], fabs(x));
    }
    // one dual entry: old value, new value, lpp = [L(p+ - p)] entry; gives dd = d - d+ and the xi2 entry
    __device__ __forceinline__ void dual(double dold, double dnew, double lpp, double alpha, double &dd, double &xi2) {
        dd = dold - dnew;
        xi2 = dd / alpha + lpp;
        put(2, xi2);
        put(5, dnew - dold);
    }
    // one primal entry: old value po, new value pn, g1 = [L*(d - d+)] entry, g2 = [L* xi2] entry
    __device__ __forceinline__ void primal(double po, double pn, double g1, double g2, double alpha) {
        const double xi1 = (po - pn) / alpha - g1;
        put(1, xi1);
        put(0, xi1 + g2);
        const double d1 = pn - po;
        put(4, d1);
        put(3, d1 + g1);   // delta0 = delta1 - L*(d+ - d) = delta1 + L*(d - d+)
    }
};

// Moreau step for one dual entry.  dbar = d_old + alpha * lz (solver.py:55-58); w = dbar / alpha (+ shift)
// (cache.py:329-347).  Returns w.
__device__ __forceinline__ double dual_w(double d_old, double lz, double alpha, double shift) {
    const double dbar = d_old + lz * alpha;
    return dbar / alpha + shift;
}

// warp-private scratch rows of the dual pass (W spans three rows: nx + nu + 2 <= 3 * rowlen)
enum { ZX, DX, ZU, DU, V1, V2, V1U, V2U, G1, G2, G1U, G2U, WROW, kDualRows = WROW + 3 };
static_assert(kDualRows == kDualRowsHost, "host shared-memory sizing out of sync");

__device__ __forceinline__ void dual_nonleaf_node(const Layout &L, const Topo &T, const Tabs &M, const PrimalView &po,
                                                  const PrimalView &pn, const DualView &dold, const DualOut &dn,
                                                  double alpha, int node, int lane, double *rows, int rowlen, Resid &R,
                                                  int &bad) {
    const int nx = L.nx, nu = L.nu;
    double *zx = rows + ZX * rowlen, *dx = rows + DX * rowlen, *zu = rows + ZU * rowlen, *du = rows + DU * rowlen;
    double *v1 = rows + V1 * rowlen, *v2 = rows + V2 * rowlen, *v1u = rows + V1U * rowlen, *v2u = rows + V2U * rowlen;
    double *g1 = rows + G1 * rowlen, *g2 = rows + G2 * rowlen, *g1u = rows + G1U * rowlen, *g2u = rows + G2U * rowlen;
    double *w = rows + WROW * rowlen;
    for (int k = lane; k < nx; k += 32) {
        const double xo = po.x[(long long)node * nx + k], xn = pn.x[(long long)node * nx + k];
        zx[k] = 2 * xn - xo;
        dx[k] = xn - xo;
        g1[k] = 0.0;
        g2[k] = 0.0;
    }
    for (int k = lane; k < nu; k += 32) {
        const double uo = po.u[(long long)node * nu + k], un = pn.u[(long long)node * nu + k];
        zu[k] = 2 * un - uo;
        du[k] = un - uo;
        g1u[k] = 0.0;
        g2u[k] = 0.0;
    }
    __syncwarp();
    const int c0 = T.child_first[node], cc = T.child_count[node];
    const int dim = nx + nu + 2;
    for (int j = c0; j < c0 + cc; ++j) {
        const long long e = j - 1;
        const int ci = T.cost_idx[j];
        const double *sqT = M.sqT + (long long)ci * nx * nx;
        const double *srT = M.srT + (long long)ci * nu * nu;
        // pass A: L z and L (p+ - p) on this edge, then w = (d + alpha L z) / alpha
        for (int k = lane; k < nx; k += 32) {
            double la, lb;
            cost_mv2(sqT, M.sq_diag, zx, dx, nx, k, la, lb);
            w[k] = dual_w(dold.d3[e * nx + k], la, alpha, 0.0);
            v2[k] = lb;
        }
        for (int k = lane; k < nu; k += 32) {
            double la, lb;
            cost_mv2(srT, M.sr_diag, zu, du, nu, k, la, lb);
            w[nx + k] = dual_w(dold.d4[e * nu + k], la, alpha, 0.0);
            v2u[k] = lb;
        }
        const double to = po.tau[j], tn = pn.tau[j];
        if (lane == 0) {
            const double ht = 0.5 * (2 * tn - to);
            w[nx + nu] = dual_w(dold.d5[e], ht, alpha, -0.5);
            w[nx + nu + 1] = dual_w(dold.d6[e], ht, alpha, 0.5);
        }
        __syncwarp();
        const SocResult sr = soc_classify(w, dim, lane);
        for (int k = lane; k < nx; k += 32) {
            const double dnew = alpha * (w[k] - soc_entry(sr, w[k], false));
            dn.d3[e * nx + k] = dnew;
            double dd, xi2;
            R.dual(dold.d3[e * nx + k], dnew, v2[k], alpha, dd, xi2);
            v1[k] = dd;
            v2[k] = xi2;
        }
        for (int k = lane; k < nu; k += 32) {
            const double dnew = alpha * (w[nx + k] - soc_entry(sr, w[nx + k], false));
            dn.d4[e * nu + k] = dnew;
            double dd, xi2;
            R.dual(dold.d4[e * nu + k], dnew, v2u[k], alpha, dd, xi2);
            v1u[k] = dd;
            v2u[k] = xi2;
        }
        if (lane == 0) {
            const double w5 = w[nx + nu], w6 = w[nx + nu + 1];
            const double dn5 = alpha * (w5 - soc_entry(sr, w5, false));
            const double dn6 = alpha * (w6 - soc_entry(sr, w6, true));
            dn.d5[e] = dn5;
            dn.d6[e] = dn6;
            const double hdt = 0.5 * (tn - to);
            double dd5, dd6, xi25, xi26;
            R.dual(dold.d5[e], dn5, hdt, alpha, dd5, xi25);
            R.dual(dold.d6[e], dn6, hdt, alpha, dd6, xi26);
            R.primal(to, tn, 0.5 * (dd5 + dd6), 0.5 * (xi25 + xi26), alpha);
        }
        __syncwarp();
        // pass B: child -> parent sums  g1 += sqrtQ_j dd3_j,  g2 += sqrtQ_j xi2_3j  (same for R / d4)
        for (int k = lane; k < nx; k += 32) {
            double a1, a2;
            cost_mv2(sqT, M.sq_diag, v1, v2, nx, k, a1, a2);
            g1[k] += a1;
            g2[k] += a2;
        }
        for (int k = lane; k < nu; k += 32) {
            double a1, a2;
            cost_mv2(srT, M.sr_diag, v1u, v2u, nu, k, a1, a2);
            g1u[k] += a1;
            g2u[k] += a2;
        }
        __syncwarp();
    }
    // d7: rectangle on [x; u] (cache.py:367-371)
    if (L.has_nl_rect) {
        const long long ri = (long long)T.nl_rect_idx[node] * L.nxu;
        for (int k = lane; k < L.nxu; k += 32) {
            const bool isx = k < nx;
            const double zk = isx ? zx[k] : zu[k - nx];
            const double dk = isx ? dx[k] : du[k - nx];
            const long long idx = (long long)node * L.nxu + k;
            const double dol = dold.d7[idx];
            const double wv = dual_w(dol, zk, alpha, 0.0);
            const double dnew = alpha * (wv - box_clip(wv, M.nl_lo[ri + k], M.nl_hi[ri + k], &bad));
            dn.d7[idx] = dnew;
            double dd, xi2;
            R.dual(dol, dnew, dk, alpha, dd, xi2);
            if (isx) {
                g1[k] += dd;
                g2[k] += xi2;
            } else {
                g1u[k - nx] += dd;
                g2u[k - nx] += xi2;
            }
        }
        __syncwarp();
    }
    for (int k = lane; k < nx; k += 32)
        R.primal(po.x[(long long)node * nx + k], pn.x[(long long)node * nx + k], g1[k], g2[k], alpha);
    for (int k = lane; k < nu; k += 32)
        R.primal(po.u[(long long)node * nu + k], pn.u[(long long)node * nu + k], g1u[k], g2u[k], alpha);
    // d1, d2 (risk blocks) and the y_i, s_i residual rows
    const int yo = T.yoff[node];
    double dot_z = 0.0, dot_d = 0.0;
    for (int e = lane; e < 2 * cc + 1; e += 32) {
        const double yold = po.y[yo + e], ynew = pn.y[yo + e];
        const double b = e < cc ? T.cond_prob[c0 + e] : (e == 2 * cc ? 1.0 : 0.0);
        dot_z = fma(b, 2 * ynew - yold, dot_z);
        dot_d = fma(b, ynew - yold, dot_d);
    }
    dot_z = warp_sum(dot_z);
    dot_d = warp_sum(dot_d);
    const double so = po.s[node], sn = pn.s[node];
    const double do2 = dold.d2[node];
    const double w2 = dual_w(do2, (2 * sn - so) - dot_z, alpha, 0.0);
    const double dn2 = alpha * (w2 - fmax(0.0, w2));
    const double dd2 = do2 - dn2;
    const double xi22 = dd2 / alpha + ((sn - so) - dot_d);
    if (lane == 0) {
        dn.d2[node] = dn2;
        R.put(2, xi22);
        R.put(5, dn2 - do2);
        R.primal(so, sn, dd2, xi22, alpha);   // s_i of a nonleaf node: L* row is d2_i
    }
    for (int e = lane; e < 2 * cc + 1; e += 32) {
        const double yold = po.y[yo + e], ynew = pn.y[yo + e];
        const double b = e < cc ? T.cond_prob[c0 + e] : (e == 2 * cc ? 1.0 : 0.0);
        const double do1 = dold.d1[yo + e];
        const double wv = dual_w(do1, 2 * ynew - yold, alpha, 0.0);
        const double zv = e < 2 * cc ? fmax(0.0, wv) : wv;
        const double dnew = alpha * (wv - zv);
        dn.d1[yo + e] = dnew;
        double dd, xi2;
        R.dual(do1, dnew, ynew - yold, alpha, dd, xi2);
        R.primal(yold, ynew, dd - b * dd2, xi2 - b * xi22, alpha);
    }
    __syncwarp();
}

__device__ __forceinline__ void dual_leaf_node(const Layout &L, const Topo &T, const Tabs &M, const PrimalView &po,
                                               const PrimalView &pn, const DualView &dold, const DualOut &dn, double alpha,
                                               int node, int lane, double *rows, int rowlen, Resid &R, int &bad) {
    const int nx = L.nx;
    double *zx = rows + ZX * rowlen, *dx = rows + DX * rowlen, *v1 = rows + V1 * rowlen, *v2 = rows + V2 * rowlen;
    double *w = rows + WROW * rowlen;
    const long long li = node - L.m;
    for (int k = lane; k < nx; k += 32) {
        const double xo = po.x[(long long)node * nx + k], xn = pn.x[(long long)node * nx + k];
        zx[k] = 2 * xn - xo;
        dx[k] = xn - xo;
    }
    __syncwarp();
    const double *sqfT = M.sqfT + (long long)T.leafcost_idx[li] * nx * nx;
    const int dim = nx + 2;
    for (int k = lane; k < nx; k += 32) {
        double la, lb;
        cost_mv2(sqfT, M.sqf_diag, zx, dx, nx, k, la, lb);
        w[k] = dual_w(dold.d11[li * nx + k], la, alpha, 0.0);
        v2[k] = lb;
    }
    const double so = po.s[node], sn = pn.s[node];
    if (lane == 0) {
        const double hs = 0.5 * (2 * sn - so);
        w[nx] = dual_w(dold.d12[li], hs, alpha, -0.5);
        w[nx + 1] = dual_w(dold.d13[li], hs, alpha, 0.5);
    }
    __syncwarp();
    const SocResult sr = soc_classify(w, dim, lane);
    for (int k = lane; k < nx; k += 32) {
        const double dnew = alpha * (w[k] - soc_entry(sr, w[k], false));
        dn.d11[li * nx + k] = dnew;
        double dd, xi2;
        R.dual(dold.d11[li * nx + k], dnew, v2[k], alpha, dd, xi2);
        v1[k] = dd;
        v2[k] = xi2;
    }
    if (lane == 0) {
        const double w12 = w[nx], w13 = w[nx + 1];
        const double dn12 = alpha * (w12 - soc_entry(sr, w12, false));
        const double dn13 = alpha * (w13 - soc_entry(sr, w13, true));
        dn.d12[li] = dn12;
        dn.d13[li] = dn13;
        const double hds = 0.5 * (sn - so);
        double dd12, dd13, xa, xb;
        R.dual(dold.d12[li], dn12, hds, alpha, dd12, xa);
        R.dual(dold.d13[li], dn13, hds, alpha, dd13, xb);
        R.primal(so, sn, 0.5 * (dd12 + dd13), 0.5 * (xa + xb), alpha);
    }
    __syncwarp();
    for (int k = lane; k < nx; k += 32) {
        double a1, a2;
        cost_mv2(sqfT, M.sqf_diag, v1, v2, nx, k, a1, a2);
        if (L.has_leaf_rect) {
            const long long ri = (long long)T.leaf_rect_idx[li] * nx;
            const long long idx = li * nx + k;
            const double dol = dold.d14[idx];
            const double wv = dual_w(dol, zx[k], alpha, 0.0);
            const double dnew = alpha * (wv - box_clip(wv, M.leaf_lo[ri + k], M.leaf_hi[ri + k], &bad));
            dn.d14[idx] = dnew;
            double dd, xi2;
            R.dual(dol, dnew, dx[k], alpha, dd, xi2);
            a1 += dd;
            a2 += xi2;
        }
        R.primal(po.x[(long long)node * nx + k], pn.x[(long long)node * nx + k], a1, a2, alpha);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(256) k_dual_tile(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl,
                                                  TilePlan plan, const double *__restrict__ p_old,
                                                  const double *__restrict__ p_new, const double *__restrict__ d_old,
                                                  double *__restrict__ d_new, double *__restrict__ slots) {
    if (ctrl->done) return;
    const double alpha = ctrl->alpha;
    const Layout &L = P.L;
    extern __shared__ double dsm[];
    __shared__ double blockmax[8][6];
    __shared__ int blockflags;
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) blockflags = 0;
    const int lo = plan.tiles[blockIdx.x].x, hi = plan.tiles[blockIdx.x].y;
    const double *Pog = p_old + (long long)blockIdx.y * L.np_pad;
    const double *Png = p_new + (long long)blockIdx.y * L.np_pad;
    const double *Dog = d_old + (long long)blockIdx.y * L.nd_pad;
    const DualOut dn = dual_out(L, d_new + (long long)blockIdx.y * L.nd_pad);
    const int nx = L.nx, nu = L.nu;
    const int rowlen = plan.rowlen;
    double *rows = dsm + (size_t)warp * kDualRows * rowlen;
    Stager st{dsm + (size_t)warps * kDualRows * rowlen};
    PrimalView po = primal_view(L, Pog), pn = primal_view(L, Png);
    DualView d = dual_view(L, Dog);
    Resid R;
    R.init();
    int bad = 0;
    if (lo < L.m) {
        const long long nN = hi - lo;
        const long long c0 = P.t.child_first[lo], c1 = P.t.child_first[hi - 1] + P.t.child_count[hi - 1], nE = c1 - c0;
        const long long y0 = P.t.yoff[lo], ny = P.t.yoff[hi] - y0;
        po.x = st.stage(po.x, (long long)lo * nx, nN * nx);
        pn.x = st.stage(pn.x, (long long)lo * nx, nN * nx);
        po.u = st.stage(po.u, (long long)lo * nu, nN * nu);
        pn.u = st.stage(pn.u, (long long)lo * nu, nN * nu);
        po.y = st.stage(po.y, y0, ny);
        pn.y = st.stage(pn.y, y0, ny);
        po.tau = st.stage(po.tau, c0, nE);
        pn.tau = st.stage(pn.tau, c0, nE);
        po.s = st.stage(po.s, lo, nN);
        pn.s = st.stage(pn.s, lo, nN);
        d.d1 = st.stage(d.d1, y0, ny);
        d.d2 = st.stage(d.d2, lo, nN);
        d.d3 = st.stage(d.d3, (c0 - 1) * nx, nE * nx);
        d.d4 = st.stage(d.d4, (c0 - 1) * nu, nE * nu);
        d.d5 = st.stage(d.d5, c0 - 1, nE);
        d.d6 = st.stage(d.d6, c0 - 1, nE);
        if (L.has_nl_rect) d.d7 = st.stage(d.d7, (long long)lo * L.nxu, nN * L.nxu);
        __syncthreads();
        for (int node = lo + warp; node < hi; node += warps)
            dual_nonleaf_node(L, P.t, P.m, po, pn, d, dn, alpha, node, lane, rows, rowlen, R, bad);
    } else {
        const long long nN = hi - lo, l0 = lo - L.m;
        po.x = st.stage(po.x, (long long)lo * nx, nN * nx);
        pn.x = st.stage(pn.x, (long long)lo * nx, nN * nx);
        po.s = st.stage(po.s, lo, nN);
        pn.s = st.stage(pn.s, lo, nN);
        d.d11 = st.stage(d.d11, l0 * nx, nN * nx);
        d.d12 = st.stage(d.d12, l0, nN);
        d.d13 = st.stage(d.d13, l0, nN);
        if (L.has_leaf_rect) d.d14 = st.stage(d.d14, l0 * nx, nN * nx);
        __syncthreads();
        for (int node = lo + warp; node < hi; node += warps)
            dual_leaf_node(L, P.t, P.m, po, pn, d, dn, alpha, node, lane, rows, rowlen, R, bad);
    }
    // block-level reduction of the six maxima, one atomic per slot per block
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const double mval = warp_max(R.v[i]);
        if (lane == 0) blockmax[warp][i] = mval;
    }
    const int anynan = __any_sync(0xffffffffu, R.nan);
    const int anybad = __any_sync(0xffffffffu, bad);
    __syncthreads();
    if (lane == 0 && (anynan || anybad)) atomicOr(&blockflags, (anynan ? 2 : 0) | (anybad ? 1 : 0));
    __syncthreads();
    if (threadIdx.x < 6) {
        double mval = blockmax[0][threadIdx.x];
        for (int wv = 1; wv < warps; ++wv) mval = fmax(mval, blockmax[wv][threadIdx.x]);
        atomic_max_nonneg(slots + (long long)blockIdx.y * 6 + threadIdx.x, mval);
    }
    if (threadIdx.x == 0 && blockflags) atomicOr(&ctrl->status, blockflags);
}

// ----------------------------------------------------------------------------------------------------------------
// stopping test (solver.py:137-161).  slots: [batch][6] maxima of the iteration that just finished; they are copied
// into the history and reset.  The loop stops when the iteration index reaches max_iters or every instance has
// max(xi0, xi1, xi2) <= tol.
__global__ void k_check(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl, double *__restrict__ slots,
                        double *__restrict__ last) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (ctrl->done) return;
    const int it = ctrl->iters;
    double *hist = ctrl->hist;
    const int hist_capacity = ctrl->hist_capacity, max_iters = ctrl->max_iters;
    const double tol = ctrl->tol;
    bool all_ok = true;
    for (int b = 0; b < P.L.batch; ++b) {
        double *s = slots + (long long)b * 6;
        const double err = fmax(s[0], fmax(s[1], s[2]));
        if (!(err <= tol)) all_ok = false;
        if (hist && it < hist_capacity)
            for (int i = 0; i < 6; ++i) hist[((long long)it * P.L.batch + b) * 6 + i] = s[i];
        for (int i = 0; i < 6; ++i) {
            last[b * 6 + i] = s[i];
            s[i] = 0.0;
        }
    }
    ctrl->iters = it + 1;
    if (it >= max_iters || all_ok) ctrl->done = 1;
}

}  // namespace rb
