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
#include "views.cuh"

namespace rb {

// ====================================================================================================================
// primal pass
// ====================================================================================================================
// nonleaf node: [xbar_i; ubar_i] (one concatenated pass, nx+nu lanes), ybar_i and the children's taubar_j, sbar_j (edge
// quantities are owned by the parent's warp) with the kernel projection (cache.py:290-317) fused in.  s_0 / tau_0 are
// handled by the caller.  DIAG: the cost square roots are diagonal (classified once at rb_create) -> no matvec.
// scratch: 2 * (nx+nu) warp-private doubles.
template <bool DIAG>
__device__ __forceinline__ void primal_nonleaf_node(const Layout &L, const Topo &T, const Tabs &M, const PrimalView &po,
                                                    const DualView &d, const PrimalOut &pn, double alpha, int node,
                                                    int lane, double *scratch) {
    const int nx = L.nx, nu = L.nu, nxu = L.nxu;
    const int c0 = T.child_first[node], cc = T.child_count[node];
    double *v = scratch, *acc = scratch + nxu;
    const double *d7 = d.d7 + node * nxu;
    if (DIAG) {
        for (int k = lane; k < nxu; k += 32) {
            const bool isx = k < nx;
            double a = L.has_nl_rect ? d7[k] : 0.0;
            for (int j = c0; j < c0 + cc; ++j) {
                const int ci = T.cost_idx[j];
                a += isx ? M.sq_d[ci * nx + k] * d.d3[(j - 1) * nx + k] : M.sr_d[ci * nu + k - nx] * d.d4[(j - 1) * nu + k - nx];
            }
            if (isx) pn.x[node * nx + k] = po.x[node * nx + k] - alpha * a;
            else pn.u[node * nu + k - nx] = po.u[node * nu + k - nx] - alpha * a;
        }
    } else {
        for (int k = lane; k < nxu; k += 32) acc[k] = L.has_nl_rect ? d7[k] : 0.0;
        for (int j = c0; j < c0 + cc; ++j) {
            for (int k = lane; k < nxu; k += 32) v[k] = k < nx ? d.d3[(j - 1) * nx + k] : d.d4[(j - 1) * nu + k - nx];
            __syncwarp();
            const int ci = T.cost_idx[j];
            for (int k = lane; k < nxu; k += 32)
                acc[k] += k < nx ? mv_row(M.sqT + (long long)ci * nx * nx, v, nx, nx, k)
                                 : mv_row(M.srT + (long long)ci * nu * nu, v + nx, nu, nu, k - nx);
            __syncwarp();
        }
        for (int k = lane; k < nxu; k += 32) {
            if (k < nx) pn.x[node * nx + k] = po.x[node * nx + k] - alpha * acc[k];
            else pn.u[node * nu + k - nx] = po.u[node * nu + k - nx] - alpha * acc[k];
        }
    }
    const double d2v = d.d2[node];
    const int yo = T.yoff[node];
    const double a = T.risk_alpha[node];
    // For AVaR M = [a I, -I, 1, -I, -I] and M M' = (a^2+3) I + 1 1', so proj = v - M'(M M')^-1 M v in closed form
    const double ylast_bar = po.y[yo + 2 * cc] - alpha * (d.d1[yo + 2 * cc] - d2v);
    const double den = a * a + 3.0;
    if (cc <= 32) {   // one child per lane: everything stays in registers
        const int e = lane, j = c0 + e;
        double ya = 0.0, yb = 0.0, tj = 0.0, sj = 0.0, res = 0.0;
        if (e < cc) {
            ya = po.y[yo + e] - alpha * (d.d1[yo + e] - T.cond_prob[j] * d2v);
            yb = po.y[yo + cc + e] - alpha * d.d1[yo + cc + e];
            tj = po.tau[j] - alpha * (0.5 * (d.d5[j - 1] + d.d6[j - 1]));
            const double lts = j < L.m ? d.d2c[j] : 0.5 * (d.d12[j - L.m] + d.d13[j - L.m]);
            sj = po.s[j] - alpha * lts;
            res = a * ya - yb + ylast_bar - tj - sj;
        }
        const double rsum = cc == 1 ? __shfl_sync(0xffffffffu, res, 0) : warp_sum(res);
        const double shift = rsum / (den + (double)cc);
        double w = 0.0;
        if (e < cc) {
            w = (res - shift) / den;
            pn.y[yo + e] = ya - a * w;
            pn.y[yo + cc + e] = yb + w;
            pn.tau[j] = tj + w;
            pn.s[j] = sj + w;
        }
        const double wsum = cc == 1 ? __shfl_sync(0xffffffffu, w, 0) : warp_sum(w);
        if (lane == 0) pn.y[yo + 2 * cc] = ylast_bar - wsum;
    } else {
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
    }
    __syncwarp();
}

template <bool DIAG>
__device__ __forceinline__ void primal_leaf_node(const Layout &L, const Topo &T, const Tabs &M, const PrimalView &po,
                                                 const DualView &d, const PrimalOut &pn, double alpha, int node, int lane,
                                                 double *scratch) {
    const int nx = L.nx;
    const int li = node - L.m;
    const int lc = T.leafcost_idx[li];
    if (!DIAG) {
        for (int k = lane; k < nx; k += 32) scratch[k] = d.d11[li * nx + k];
        __syncwarp();
    }
    for (int k = lane; k < nx; k += 32) {
        double acc = DIAG ? M.sqf_d[lc * nx + k] * d.d11[li * nx + k]
                          : mv_row(M.sqfT + (long long)lc * nx * nx, scratch, nx, nx, k);
        if (L.has_leaf_rect) acc += d.d14[li * nx + k];
        pn.x[node * nx + k] = po.x[node * nx + k] - alpha * acc;
    }
    __syncwarp();
}

template <bool DIAG>
__global__ void __launch_bounds__(256, 3) k_primal_tile(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
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
    double *rows = dsm + (size_t)warp * 4 * rowlen;   // 4 * rowlen >= 2 * (nx + nu)
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
        st.finish();
        for (int node = lo + warp; node < hi; node += warps) {
            primal_nonleaf_node<DIAG>(L, P.t, P.m, po, d, pn, alpha, node, lane, rows);
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
        st.finish();
        for (int node = lo + warp; node < hi; node += warps)
            primal_leaf_node<DIAG>(L, P.t, P.m, po, d, pn, alpha, node, lane, rows);
    }
}

// ====================================================================================================================
// dual pass + residuals
// ====================================================================================================================
// residual bookkeeping: six running maxima per lane (xi0, xi1, xi2, delta0, delta1, delta2).  Non-negative doubles
// order like their bit patterns and NaN patterns sit above +inf, so an integer max keeps a NaN instead of dropping it
// (fmax would): a non-finite iterate reaches the stopping test as a NaN residual.
struct Resid {
    unsigned long long v[6];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < 6; ++i) v[i] = 0ull;
    }
    __device__ __forceinline__ void put(int slot, double x) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(fabs(x));
        v[slot] = b > v[slot] ? b : v[slot];
    }
    // one dual entry: dd = d - d+, lpp = [L(p+ - p)] entry; returns the xi2 entry
    __device__ __forceinline__ double dual(double dd, double lpp, double inv_alpha) {
        const double xi2 = fma(dd, inv_alpha, lpp);
        put(2, xi2);
        put(5, dd);        // |delta2| = |d+ - d|
        return xi2;
    }
    // one primal entry: dp = p+ - p, g1 = [L*(d - d+)] entry, g2 = [L* xi2] entry
    __device__ __forceinline__ void primal(double dp, double g1, double g2, double inv_alpha) {
        const double xi1 = -fma(dp, inv_alpha, g1);   // (p - p+)/alpha - L*(d - d+)
        put(1, xi1);
        put(0, xi1 + g2);
        put(4, dp);
        put(3, dp + g1);   // delta0 = delta1 - L*(d+ - d) = delta1 + L*(d - d+)
    }
};

// Moreau step for one dual entry: dbar = d_old + alpha * lz (solver.py:55-58); w = dbar / alpha (+ shift)
// (cache.py:329-347).  The division is a multiplication by 1/alpha (<= 1 ulp from the reference's quotient).
__device__ __forceinline__ double dual_w(double d_old, double lz, double alpha, double inv_alpha, double shift) {
    return fma(lz, alpha, d_old) * inv_alpha + shift;
}

// warp-private scratch of the dual pass, in units of rowlen doubles (rowlen >= max(nx, nu), even):
//   Z, DL : [z_x; z_u] = 2 p+ - p and [d_x; d_u] = p+ - p of the node          (2 rows each: nx+nu <= 2 rowlen)
//   G1, G2: child -> parent sums [L*(d - d+)] and [L* xi2] on the [x; u] rows    (2 rows each)
//   W, LB : w and L(p+ - p) on one edge, nx+nu+2 entries                        (3 rows each; dense path: V1/V2 alias)
enum { SZ = 0, SDL = 2, SG1 = 4, SG2 = 6, SW = 8, SLB = 11, kDualRows = 15 };
static_assert(kDualRows == kDualRowsHost, "host shared-memory sizing out of sync");

template <bool DIAG>
__device__ __forceinline__ void dual_nonleaf_node(const Layout &L, const Topo &T, const Tabs &M, const PrimalView &po,
                                                  const PrimalView &pn, const DualView &dold, const DualOut &dn,
                                                  double alpha, double inv_alpha, int node, int lane, double *rows,
                                                  int rowlen, Resid &R, int &bad) {
    const int nx = L.nx, nu = L.nu, nxu = L.nxu, E = nxu + 2;
    double *z = rows + SZ * rowlen, *dl = rows + SDL * rowlen, *g1 = rows + SG1 * rowlen, *g2 = rows + SG2 * rowlen;
    double *w = rows + SW * rowlen, *lbv = rows + SLB * rowlen;
    const double *xo = po.x + node * nx, *xn = pn.x + node * nx, *uo = po.u + node * nu - nx, *un = pn.u + node * nu - nx;
    for (int k = lane; k < nxu; k += 32) {
        const double o = k < nx ? xo[k] : uo[k], nw = k < nx ? xn[k] : un[k];
        z[k] = 2 * nw - o;
        dl[k] = nw - o;
        g1[k] = 0.0;
        g2[k] = 0.0;
    }
    __syncwarp();
    const int c0 = T.child_first[node], cc = T.child_count[node];
    for (int j = c0; j < c0 + cc; ++j) {
        const int e0 = j - 1, ci = T.cost_idx[j];
        const double *d3 = dold.d3 + e0 * nx, *d4 = dold.d4 + e0 * nu - nx;
        const double to = po.tau[j], tn = pn.tau[j];
        // pass A over the concatenated edge vector [d3; d4; d5; d6]: w = (d + alpha L z) / alpha (+-1/2), |.|^2 of all
        // but the last entry (the cone's t)
        double ss = 0.0;
        for (int e = lane; e < E; e += 32) {
            double la, lb, dol, shift = 0.0;
            if (e < nx) {
                if (DIAG) {
                    const double mm = M.sq_d[ci * nx + e];
                    la = mm * z[e];
                    lb = mm * dl[e];
                } else {
                    mv_row2(M.sqT + (long long)ci * nx * nx, z, dl, nx, nx, e, la, lb);
                }
                dol = d3[e];
            } else if (e < nxu) {
                if (DIAG) {
                    const double mm = M.sr_d[ci * nu + e - nx];
                    la = mm * z[e];
                    lb = mm * dl[e];
                } else {
                    mv_row2(M.srT + (long long)ci * nu * nu, z + nx, dl + nx, nu, nu, e - nx, la, lb);
                }
                dol = d4[e];
            } else {
                la = 0.5 * (2 * tn - to);
                lb = 0.5 * (tn - to);
                dol = e == nxu ? dold.d5[e0] : dold.d6[e0];
                shift = e == nxu ? -0.5 : 0.5;
            }
            const double wv = dual_w(dol, la, alpha, inv_alpha, shift);
            w[e] = wv;
            lbv[e] = lb;
            if (e < E - 1) ss = fma(wv, wv, ss);
        }
        ss = warp_sum(ss);   // also orders the scratch writes above before the reads below
        __syncwarp();
        // SecondOrderCone.project (cones.py:113-132): same branch order
        const double r = sqrt(ss), t = w[E - 1];
        const int mode = r <= t ? 0 : (r <= -t ? 1 : 2);
        const double t_new = (r + t) / 2;
        const double scale = mode == 2 ? t_new / r : 0.0;   // reference: t_new * (w / r) entrywise
        double tdd = 0.0, txi = 0.0;
        for (int e = lane; e < E; e += 32) {
            const double wv = w[e];
            const double zv = mode == 0 ? wv : (mode == 1 ? 0.0 : (e == E - 1 ? t_new : scale * wv));
            const double dnew = alpha * (wv - zv);
            double dol;
            if (e < nx) {
                dol = d3[e];
                dn.d3[e0 * nx + e] = dnew;
            } else if (e < nxu) {
                dol = d4[e];
                dn.d4[e0 * nu + e - nx] = dnew;
            } else if (e == nxu) {
                dol = dold.d5[e0];
                dn.d5[e0] = dnew;
            } else {
                dol = dold.d6[e0];
                dn.d6[e0] = dnew;
            }
            const double dd = dol - dnew;
            const double xi2 = R.dual(dd, lbv[e], inv_alpha);
            if (e < nxu) {
                if (DIAG) {   // child -> parent sums with a diagonal cost root: entrywise
                    const double mm = e < nx ? M.sq_d[ci * nx + e] : M.sr_d[ci * nu + e - nx];
                    g1[e] += mm * dd;
                    g2[e] += mm * xi2;
                } else {
                    w[e] = dd;      // reuse as V1 / V2 for pass B (each lane rewrites only its own entries)
                    lbv[e] = xi2;
                }
            } else {
                tdd = dd;
                txi = xi2;
            }
        }
        // tau_j residual rows need d5 and d6 together: entries nxu and nxu+1 sit in neighbouring lanes
        {
            const int l5 = nxu & 31, l6 = (nxu + 1) & 31;
            const double dd5 = __shfl_sync(0xffffffffu, tdd, l5), dd6 = __shfl_sync(0xffffffffu, tdd, l6);
            const double x5 = __shfl_sync(0xffffffffu, txi, l5), x6 = __shfl_sync(0xffffffffu, txi, l6);
            if (lane == 0) R.primal(tn - to, 0.5 * (dd5 + dd6), 0.5 * (x5 + x6), inv_alpha);
        }
        if (!DIAG) {
            __syncwarp();
            for (int k = lane; k < nxu; k += 32) {
                double a1, a2;
                if (k < nx) mv_row2(M.sqT + (long long)ci * nx * nx, w, lbv, nx, nx, k, a1, a2);
                else mv_row2(M.srT + (long long)ci * nu * nu, w + nx, lbv + nx, nu, nu, k - nx, a1, a2);
                g1[k] += a1;
                g2[k] += a2;
            }
        }
        __syncwarp();
    }
    // d7: rectangle on [x; u] (cache.py:367-371), then the x / u residual rows
    const long long ri = L.has_nl_rect ? (long long)T.nl_rect_idx[node] * nxu : 0;
    for (int k = lane; k < nxu; k += 32) {
        double a1 = g1[k], a2 = g2[k];
        if (L.has_nl_rect) {
            const double dol = dold.d7[node * nxu + k];
            const double wv = dual_w(dol, z[k], alpha, inv_alpha, 0.0);
            const double dnew = alpha * (wv - box_clip(wv, M.nl_lo[ri + k], M.nl_hi[ri + k], &bad));
            dn.d7[node * nxu + k] = dnew;
            const double dd = dol - dnew;
            a1 += dd;
            a2 += R.dual(dd, dl[k], inv_alpha);
        }
        R.primal(dl[k], a1, a2, inv_alpha);
    }
    // d1, d2 (risk blocks) and the y_i, s_i residual rows
    const int yo = T.yoff[node], ny = 2 * cc + 1;
    double dot_z = 0.0, dot_d = 0.0;
    for (int e = lane; e < ny; e += 32) {
        const double yold = po.y[yo + e], ynew = pn.y[yo + e];
        const double b = e < cc ? T.cond_prob[c0 + e] : (e == 2 * cc ? 1.0 : 0.0);
        dot_z = fma(b, 2 * ynew - yold, dot_z);
        dot_d = fma(b, ynew - yold, dot_d);
    }
    dot_z = warp_sum(dot_z);
    dot_d = warp_sum(dot_d);
    const double so = po.s[node], sn = pn.s[node];
    const double do2 = dold.d2[node];
    const double w2 = dual_w(do2, (2 * sn - so) - dot_z, alpha, inv_alpha, 0.0);
    const double dn2 = alpha * (w2 - fmax(0.0, w2));
    const double dd2 = do2 - dn2;
    const double xi22 = fma(dd2, inv_alpha, (sn - so) - dot_d);
    if (lane == 0) {
        dn.d2[node] = dn2;
        R.put(2, xi22);
        R.put(5, dd2);
        R.primal(sn - so, dd2, xi22, inv_alpha);   // s_i of a nonleaf node: its L* row is d2_i
    }
    for (int e = lane; e < ny; e += 32) {
        const double yold = po.y[yo + e], ynew = pn.y[yo + e];
        const double b = e < cc ? T.cond_prob[c0 + e] : (e == 2 * cc ? 1.0 : 0.0);
        const double do1 = dold.d1[yo + e];
        const double wv = dual_w(do1, 2 * ynew - yold, alpha, inv_alpha, 0.0);
        const double zv = e < 2 * cc ? fmax(0.0, wv) : wv;
        const double dnew = alpha * (wv - zv);
        dn.d1[yo + e] = dnew;
        const double dd = do1 - dnew;
        const double xi2 = R.dual(dd, ynew - yold, inv_alpha);
        R.primal(ynew - yold, dd - b * dd2, xi2 - b * xi22, inv_alpha);
    }
    __syncwarp();
}

template <bool DIAG>
__device__ __forceinline__ void dual_leaf_node(const Layout &L, const Topo &T, const Tabs &M, const PrimalView &po,
                                               const PrimalView &pn, const DualView &dold, const DualOut &dn, double alpha,
                                               double inv_alpha, int node, int lane, double *rows, int rowlen, Resid &R,
                                               int &bad) {
    const int nx = L.nx, E = nx + 2;
    double *z = rows + SZ * rowlen, *dl = rows + SDL * rowlen, *w = rows + SW * rowlen, *lbv = rows + SLB * rowlen;
    double *g1 = rows + SG1 * rowlen, *g2 = rows + SG2 * rowlen;
    const int li = node - L.m, lc = T.leafcost_idx[li];
    const double *xo = po.x + node * nx, *xn = pn.x + node * nx, *d11 = dold.d11 + li * nx;
    for (int k = lane; k < nx; k += 32) {
        z[k] = 2 * xn[k] - xo[k];
        dl[k] = xn[k] - xo[k];
    }
    __syncwarp();
    const double so = po.s[node], sn = pn.s[node];
    double ss = 0.0;
    for (int e = lane; e < E; e += 32) {   // concatenated [d11; d12; d13], SOC with t = d13 (cache.py:375-386)
        double la, lb, dol, shift = 0.0;
        if (e < nx) {
            if (DIAG) {
                const double mm = M.sqf_d[lc * nx + e];
                la = mm * z[e];
                lb = mm * dl[e];
            } else {
                mv_row2(M.sqfT + (long long)lc * nx * nx, z, dl, nx, nx, e, la, lb);
            }
            dol = d11[e];
        } else {
            la = 0.5 * (2 * sn - so);
            lb = 0.5 * (sn - so);
            dol = e == nx ? dold.d12[li] : dold.d13[li];
            shift = e == nx ? -0.5 : 0.5;
        }
        const double wv = dual_w(dol, la, alpha, inv_alpha, shift);
        w[e] = wv;
        lbv[e] = lb;
        if (e < E - 1) ss = fma(wv, wv, ss);
    }
    ss = warp_sum(ss);
    __syncwarp();
    const double r = sqrt(ss), t = w[E - 1];
    const int mode = r <= t ? 0 : (r <= -t ? 1 : 2);
    const double t_new = (r + t) / 2;
    const double scale = mode == 2 ? t_new / r : 0.0;
    double tdd = 0.0, txi = 0.0;
    for (int e = lane; e < E; e += 32) {
        const double wv = w[e];
        const double zv = mode == 0 ? wv : (mode == 1 ? 0.0 : (e == E - 1 ? t_new : scale * wv));
        const double dnew = alpha * (wv - zv);
        double dol;
        if (e < nx) {
            dol = d11[e];
            dn.d11[li * nx + e] = dnew;
        } else if (e == nx) {
            dol = dold.d12[li];
            dn.d12[li] = dnew;
        } else {
            dol = dold.d13[li];
            dn.d13[li] = dnew;
        }
        const double dd = dol - dnew;
        const double xi2 = R.dual(dd, lbv[e], inv_alpha);
        if (e < nx) {
            if (DIAG) {
                const double mm = M.sqf_d[lc * nx + e];
                g1[e] = mm * dd;
                g2[e] = mm * xi2;
            } else {
                w[e] = dd;
                lbv[e] = xi2;
            }
        } else {
            tdd = dd;
            txi = xi2;
        }
    }
    {
        const int l12 = nx & 31, l13 = (nx + 1) & 31;
        const double dd12 = __shfl_sync(0xffffffffu, tdd, l12), dd13 = __shfl_sync(0xffffffffu, tdd, l13);
        const double xa = __shfl_sync(0xffffffffu, txi, l12), xb = __shfl_sync(0xffffffffu, txi, l13);
        if (lane == 0) R.primal(sn - so, 0.5 * (dd12 + dd13), 0.5 * (xa + xb), inv_alpha);
    }
    if (!DIAG) __syncwarp();
    const long long ri = L.has_leaf_rect ? (long long)T.leaf_rect_idx[li] * nx : 0;
    for (int k = lane; k < nx; k += 32) {
        double a1, a2;
        if (DIAG) {
            a1 = g1[k];
            a2 = g2[k];
        } else {
            mv_row2(M.sqfT + (long long)lc * nx * nx, w, lbv, nx, nx, k, a1, a2);
        }
        if (L.has_leaf_rect) {
            const double dol = dold.d14[li * nx + k];
            const double wv = dual_w(dol, z[k], alpha, inv_alpha, 0.0);
            const double dnew = alpha * (wv - box_clip(wv, M.leaf_lo[ri + k], M.leaf_hi[ri + k], &bad));
            dn.d14[li * nx + k] = dnew;
            const double dd = dol - dnew;
            a1 += dd;
            a2 += R.dual(dd, dl[k], inv_alpha);
        }
        R.primal(dl[k], a1, a2, inv_alpha);
    }
    __syncwarp();
}

template <bool DIAG>
__global__ void __launch_bounds__(256, 2) k_dual_tile(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl,
                                                     TilePlan plan, const double *__restrict__ p_old,
                                                     const double *__restrict__ p_new, const double *__restrict__ d_old,
                                                     double *__restrict__ d_new, double *__restrict__ slots) {
    if (ctrl->done) return;
    const double alpha = ctrl->alpha, inv_alpha = 1.0 / alpha;
    const Layout &L = P.L;
    extern __shared__ double dsm[];
    __shared__ unsigned long long blockmax[8][6];
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
        st.finish();
        for (int node = lo + warp; node < hi; node += warps)
            dual_nonleaf_node<DIAG>(L, P.t, P.m, po, pn, d, dn, alpha, inv_alpha, node, lane, rows, rowlen, R, bad);
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
        st.finish();
        for (int node = lo + warp; node < hi; node += warps)
            dual_leaf_node<DIAG>(L, P.t, P.m, po, pn, d, dn, alpha, inv_alpha, node, lane, rows, rowlen, R, bad);
    }
    // block-level reduction of the six maxima (as bit patterns), one atomic per slot per block
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        unsigned long long mval = R.v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, mval, o);
            mval = other > mval ? other : mval;
        }
        if (lane == 0) blockmax[warp][i] = mval;
    }
    const int anybad = __any_sync(0xffffffffu, bad);
    __syncthreads();
    if (lane == 0 && anybad) atomicOr(&blockflags, 1);
    __syncthreads();
    if (threadIdx.x < 6) {
        unsigned long long mval = blockmax[0][threadIdx.x];
        for (int wv = 1; wv < warps; ++wv) mval = blockmax[wv][threadIdx.x] > mval ? blockmax[wv][threadIdx.x] : mval;
        atomicMax(reinterpret_cast<unsigned long long *>(slots + (long long)blockIdx.y * 6 + threadIdx.x), mval);
    }
    if (threadIdx.x == 0 && blockflags) atomicOr(&ctrl->status, blockflags);
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) ctrl->pending = 1;
}

cudaError_t tile_kernels_set_smem(size_t primal_bytes, size_t dual_bytes) {
    cudaError_t e = cudaFuncSetAttribute(k_primal_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)primal_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_primal_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)primal_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_dual_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dual_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_dual_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dual_bytes);
    return e;
}

void launch_primal_tile(bool diag, dim3 grid, size_t smem, cudaStream_t st, const Params &P, const Ctrl *ctrl,
                        const TilePlan &plan, const double *p_old, const double *d_old, double *p_new) {
    if (diag) k_primal_tile<true><<<grid, 256, smem, st>>>(P, ctrl, plan, p_old, d_old, p_new);
    else k_primal_tile<false><<<grid, 256, smem, st>>>(P, ctrl, plan, p_old, d_old, p_new);
}

void launch_dual_tile(bool diag, dim3 grid, size_t smem, cudaStream_t st, const Params &P, Ctrl *ctrl, const TilePlan &plan,
                      const double *p_old, const double *p_new, const double *d_old, double *d_new, double *slots) {
    if (diag) k_dual_tile<true><<<grid, 256, smem, st>>>(P, ctrl, plan, p_old, p_new, d_old, d_new, slots);
    else k_dual_tile<false><<<grid, 256, smem, st>>>(P, ctrl, plan, p_old, p_new, d_old, d_new, slots);
}

// ----------------------------------------------------------------------------------------------------------------
// stopping test (solver.py:137-161).  slots: [batch][6] maxima of the iteration that just finished; they are copied
// into the history and reset.  The loop stops when the iteration index reaches max_iters or every instance has
// max(xi0, xi1, xi2) <= tol.
__global__ void k_check(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl, double *__restrict__ slots,
                        double *__restrict__ last, double *__restrict__ host_last) {
    // host_last (may be null): mapped pinned host memory; the norms land there as well, so that a host caller that
    // synchronises after every iteration (rb_step) needs no device-to-host copy of its own
    // One warp.  The control block and the first instances' slots are requested together (the kernel is a chain of
    // dependent global round trips otherwise, and it sits on the critical path of every iteration); a launch with
    // nothing pending -- the loop tests iteration k at the head of iteration k + 1 and again before the host reads the
    // norms -- is a no-op.
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const int batch = P.L.batch;
    const Ctrl c = *ctrl;
    const int total = batch * 6;
    double mine = lane < total ? slots[lane] : 0.0;
    if (c.done || !c.pending) return;
    const int it = c.iters;
    bool ok = true, nan = false;
    for (int base = 0; base < total; base += 32) {
        const int i = base + lane;
        if (base > 0) mine = i < total ? slots[i] : 0.0;
        if (i < total) {
            if (mine != mine) nan = true;
            if (i % 6 < 3 && !(mine <= c.tol)) ok = false;   // max(xi0, xi1, xi2) <= tol  <=>  each of them is
            if (c.hist && it < c.hist_capacity) c.hist[(long long)it * total + i] = mine;
            last[i] = mine;
            if (host_last && c.mirror) host_last[i] = mine;
            slots[i] = 0.0;
        }
    }
    const bool all_ok = __all_sync(0xffffffffu, ok);
    const bool any_nan = __any_sync(0xffffffffu, nan);
    if (lane == 0) {
        if (any_nan) ctrl->status |= 2;   // a NaN maximum: some iterate entry is not finite
        ctrl->iters = it + 1;
        ctrl->pending = 0;
        if (it >= c.max_iters || all_ok) ctrl->done = 1;
    }
}

// k_check for thousands of instances: the one-warp kernel walks batch x 6 slots serially (330 us at batch 4096 -- a fifth of a
// cfg4 iteration).  One thread per slot; every CTA folds its verdict into the control block and the last one to arrive closes the
// iteration exactly like k_check does.
__global__ void __launch_bounds__(256) k_check_wide(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl,
                                                    double *__restrict__ slots, double *__restrict__ last,
                                                    double *__restrict__ host_last) {
    __shared__ int sh_fail, sh_nan;
    const Ctrl c = *ctrl;              // read before this CTA arrives, i.e. before the closing CTA changes it
    if (c.done || !c.pending) return;
    if (threadIdx.x == 0) sh_fail = sh_nan = 0;
    __syncthreads();
    const int total = P.L.batch * 6;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) {
        const double mine = slots[i];
        if (mine != mine) sh_nan = 1;
        if (i % 6 < 3 && !(mine <= c.tol)) sh_fail = 1;
        if (c.hist && c.iters < c.hist_capacity) c.hist[(long long)c.iters * total + i] = mine;
        last[i] = mine;
        if (host_last && c.mirror) host_last[i] = mine;
        slots[i] = 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (sh_fail) atomicOr(&ctrl->chk_fail, 1);
        if (sh_nan) atomicOr(&ctrl->chk_nan, 1);
        __threadfence();
        if (atomicAdd(&ctrl->chk_arrived, 1) == (int)gridDim.x - 1) {   // the last CTA: every verdict is in
            __threadfence();
            const int fail = atomicExch(&ctrl->chk_fail, 0), nan = atomicExch(&ctrl->chk_nan, 0);
            ctrl->chk_arrived = 0;
            if (nan) ctrl->status |= 2;
            ctrl->iters = c.iters + 1;
            ctrl->pending = 0;
            if (c.iters >= c.max_iters || !fail) ctrl->done = 1;
        }
    }
}

void launch_check(cudaStream_t st, const Params &P, Ctrl *ctrl, double *slots, double *last, double *host_last) {
    const int total = P.L.batch * 6;
    if (total <= 256) k_check<<<1, 32, 0, st>>>(P, ctrl, slots, last, host_last);
    else k_check_wide<<<(total + 255) / 256, 256, 0, st>>>(P, ctrl, slots, last, host_last);
}

}  // namespace rb
