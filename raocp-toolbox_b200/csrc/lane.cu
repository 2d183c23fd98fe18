// lane.cu -- the node-parallel passes of one Chambolle-Pock iteration with ONE THREAD PER TREE NODE (FP64, sm_100a).
//
// The warp-per-node tile kernels of fused.cu spend >85 % of their instructions on index arithmetic, shuffles and
// divergent per-lane branches (profiles/r1_kernel_evolution.md): a node carries only ~130 doubles.  Here every thread
// owns a node and streams its rows sequentially with 16-byte loads: all 32 lanes do identical useful work, there are no
// shuffles and no shared-memory scratch rows, and the instruction count is 5x lower (227 warp instructions per node
// for the dual pass).  Consecutive lanes own consecutive nodes, i.e. rows that are contiguous in the node-major layout;
// a 32-byte sector fetched for a lane is consumed by that lane's next load.  What bounds these kernels now is the
// L1TEX wavefront rate (one wavefront per row touched by a request) and the small number of warps a 6e4-node tree
// offers (13 per SM); a shared-memory staged variant measured no faster (profiles/r1_kernel_evolution.md), so the
// simple form is kept.  Used when the cost square roots are diagonal (rb_create classifies the tables), nx and nu are
// even and no node has more than kLaneMaxChildren children; otherwise the general warp-per-node tile kernels run.
//
//   k_primal_lane : pbar = p - alpha L* d (solver.py:27-39), s_0 -= alpha (cache.py:253-257), kernel projection
//                   (cache.py:290-317)
//   k_dual_lane   : dbar = d + alpha L(2 p+ - p) (solver.py:44-58), prox of g* (cache.py:321-393), all six residual
//                   inf-norms (solver.py:63-95,137-141)
#include "kernels.cuh"

namespace rb {

namespace {

struct ResidLane {
    unsigned long long v[6];   // bit patterns of non-negative doubles order like the doubles; NaN sits above +inf
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < 6; ++i) v[i] = 0ull;
    }
    __device__ __forceinline__ void put(int slot, double x) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(fabs(x));
        v[slot] = b > v[slot] ? b : v[slot];
    }
    __device__ __forceinline__ double dual(double dd, double lpp, double inv_alpha) {   // dd = d - d+
        const double xi2 = fma(dd, inv_alpha, lpp);
        put(2, xi2);
        put(5, dd);
        return xi2;
    }
    __device__ __forceinline__ void primal(double dp, double g1, double g2, double inv_alpha) {   // dp = p+ - p
        const double xi1 = -fma(dp, inv_alpha, g1);
        put(1, xi1);
        put(0, xi1 + g2);
        put(4, dp);
        put(3, dp + g1);
    }
};

__device__ __forceinline__ double dual_w(double d_old, double lz, double alpha, double inv_alpha) {
    return fma(lz, alpha, d_old) * inv_alpha;
}

// Rows are walked two doubles (16 bytes) at a time: nx and nu are even on this path, so every row start is 16-byte
// aligned, a lane consumes a 32-byte sector in two consecutive loads, and a warp request moves 512 useful bytes.
__device__ __forceinline__ double2 ld2(const double *__restrict__ p, int k2) {
    return *reinterpret_cast<const double2 *>(p + 2 * k2);
}
__device__ __forceinline__ void st2(double *__restrict__ p, int k2, double a, double b) {
    *reinterpret_cast<double2 *>(p + 2 * k2) = make_double2(a, b);
}

}  // namespace

// ====================================================================================================================
__global__ void __launch_bounds__(kLaneThreads) k_primal_lane(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                             const double *__restrict__ p_old,
                                                             const double *__restrict__ d_old, double *__restrict__ p_new,
                                                             const int *__restrict__ node_list, int count) {
    if (ctrl->done) return;
    const double alpha = ctrl->alpha;
    const Layout &L = P.L;
    const Topo &T = P.t;
    const Tabs &M = P.m;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= count) return;
    const int node = node_list ? node_list[gid] : gid;
    const double *Po = p_old + (long long)blockIdx.y * L.np_pad;
    const double *D = d_old + (long long)blockIdx.y * L.nd_pad;
    double *Pn = p_new + (long long)blockIdx.y * L.np_pad;
    const int nx = L.nx, nu = L.nu, nxu = L.nxu;
    if (node >= L.m) {   // leaf: xbar = x - alpha (sqrtQf d11 + d14)   (operators.py:89-92)
        const int li = node - L.m;
        const double *sq = M.sqf_d + T.leafcost_idx[li] * nx;
        const double *d11 = D + L.d11 + (long long)li * nx, *d14 = D + L.d14 + (long long)li * nx;
        const double *xo = Po + L.px + (long long)node * nx;
        double *xn = Pn + L.px + (long long)node * nx;
#pragma unroll 2
        for (int k2 = 0; k2 < nx / 2; ++k2) {
            const double2 m2 = ld2(sq, k2), a2 = ld2(d11, k2), o2 = ld2(xo, k2);
            double acc0 = m2.x * a2.x, acc1 = m2.y * a2.y;
            if (L.has_leaf_rect) {
                const double2 b2 = ld2(d14, k2);
                acc0 += b2.x;
                acc1 += b2.y;
            }
            st2(xn, k2, o2.x - alpha * acc0, o2.y - alpha * acc1);
        }
        return;
    }
    const int c0 = T.child_first[node], cc = T.child_count[node];
    {   // [xbar; ubar] = [x; u] - alpha (Gamma' d7 + sum_j sqrt(Q_j, R_j) [d3_j; d4_j])   (operators.py:74-87)
        const double *xo = Po + L.px + (long long)node * nx, *uo = Po + L.pu + (long long)node * nu;
        double *xn = Pn + L.px + (long long)node * nx, *un = Pn + L.pu + (long long)node * nu;
        const double *d7 = D + L.d7 + (long long)node * nxu;
#pragma unroll 2
        for (int k2 = 0; k2 < nx / 2; ++k2) {
            const double2 o2 = ld2(xo, k2);
            double2 acc = L.has_nl_rect ? ld2(d7, k2) : make_double2(0.0, 0.0);
            for (int j = c0; j < c0 + cc; ++j) {
                const double2 m2 = ld2(M.sq_d + T.cost_idx[j] * nx, k2), v2 = ld2(D + L.d3 + (long long)(j - 1) * nx, k2);
                acc.x = fma(m2.x, v2.x, acc.x);
                acc.y = fma(m2.y, v2.y, acc.y);
            }
            st2(xn, k2, o2.x - alpha * acc.x, o2.y - alpha * acc.y);
        }
#pragma unroll 2
        for (int k2 = 0; k2 < nu / 2; ++k2) {
            const double2 o2 = ld2(uo, k2);
            double2 acc = L.has_nl_rect ? ld2(d7 + nx, k2) : make_double2(0.0, 0.0);
            for (int j = c0; j < c0 + cc; ++j) {
                const double2 m2 = ld2(M.sr_d + T.cost_idx[j] * nu, k2), v2 = ld2(D + L.d4 + (long long)(j - 1) * nu, k2);
                acc.x = fma(m2.x, v2.x, acc.x);
                acc.y = fma(m2.y, v2.y, acc.y);
            }
            st2(un, k2, o2.x - alpha * acc.x, o2.y - alpha * acc.y);
        }
    }
    // ybar_i, the children's taubar_j / sbar_j, and the projection onto ker [E' -I -I] (cache.py:290-317).  For AVaR
    // M = [a I, -I, 1, -I, -I], M M' = (a^2+3) I + 1 1', so proj = v - M'(M M')^-1 M v in closed form.
    const double d2v = D[L.d2 + node];
    const int yo = T.yoff[node];
    const double a = T.risk_alpha[node];
    const double *yold = Po + L.py + yo, *d1 = D + L.d1 + yo;
    double *ynew = Pn + L.py + yo;
    const double ylast_bar = yold[2 * cc] - alpha * (d1[2 * cc] - d2v);
    const double den = a * a + 3.0;
    double rsum = 0.0;
    for (int e = 0; e < cc; ++e) {
        const int j = c0 + e;
        const double ya = yold[e] - alpha * (d1[e] - T.cond_prob[j] * d2v);
        const double yb = yold[cc + e] - alpha * d1[cc + e];
        const double tj = Po[L.ptau + j] - alpha * (0.5 * (D[L.d5 + j - 1] + D[L.d6 + j - 1]));
        const double lts = j < L.m ? D[L.d2 + j] : 0.5 * (D[L.d12 + j - L.m] + D[L.d13 + j - L.m]);
        const double sj = Po[L.ps + j] - alpha * lts;
        rsum += a * ya - yb + ylast_bar - tj - sj;
    }
    const double shift = rsum / (den + (double)cc);
    double wsum = 0.0;
    for (int e = 0; e < cc; ++e) {   // the same arithmetic again, bit for bit
        const int j = c0 + e;
        const double ya = yold[e] - alpha * (d1[e] - T.cond_prob[j] * d2v);
        const double yb = yold[cc + e] - alpha * d1[cc + e];
        const double tj = Po[L.ptau + j] - alpha * (0.5 * (D[L.d5 + j - 1] + D[L.d6 + j - 1]));
        const double lts = j < L.m ? D[L.d2 + j] : 0.5 * (D[L.d12 + j - L.m] + D[L.d13 + j - L.m]);
        const double sj = Po[L.ps + j] - alpha * lts;
        const double w = ((a * ya - yb + ylast_bar - tj - sj) - shift) / den;
        ynew[e] = ya - a * w;
        ynew[cc + e] = yb + w;
        Pn[L.ptau + j] = tj + w;
        Pn[L.ps + j] = sj + w;
        wsum += w;
    }
    ynew[2 * cc] = ylast_bar - wsum;
    if (node == 0) {
        Pn[L.ps] = (Po[L.ps] - alpha * d2v) - alpha;   // s_0: half step, then prox of alpha * identity
        Pn[L.ptau] = Po[L.ptau] - alpha * Po[L.ptau];  // tau_0 (always 0; same arithmetic as the reference)
    }
}

// ====================================================================================================================
__global__ void __launch_bounds__(kLaneThreads) k_dual_lane(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl,
                                                           const double *__restrict__ p_old, const double *__restrict__ p_new,
                                                           const double *__restrict__ d_old, double *__restrict__ d_new,
                                                           double *__restrict__ slots, const int *__restrict__ node_list,
                                                           int count) {
    if (ctrl->done) return;
    const double alpha = ctrl->alpha, inv_alpha = 1.0 / alpha;
    const Layout &L = P.L;
    const Topo &T = P.t;
    const Tabs &M = P.m;
    // per-thread SOC results of the children, [slot][thread]: scale (projection = scale * w on all but the last entry)
    // and the projected last entry
    __shared__ double soc_scale[kLaneMaxChildren][kLaneThreads];
    __shared__ double soc_last[kLaneMaxChildren][kLaneThreads];
    __shared__ unsigned long long blockmax[kLaneThreads / 32][6];
    __shared__ int blockflags;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) blockflags = 0;
    const int gid = blockIdx.x * blockDim.x + tid;
    const int node = gid < count ? (node_list ? node_list[gid] : gid) : L.n;   // L.n = nothing to do
    const double *Po = p_old + (long long)blockIdx.y * L.np_pad;
    const double *Pn = p_new + (long long)blockIdx.y * L.np_pad;
    const double *Do = d_old + (long long)blockIdx.y * L.nd_pad;
    double *Dn = d_new + (long long)blockIdx.y * L.nd_pad;
    const int nx = L.nx, nu = L.nu, nxu = L.nxu;
    ResidLane R;
    R.init();
    int bad = 0;

    if (node < L.m) {
        const int c0 = T.child_first[node], cc = T.child_count[node];
        const double *xo = Po + L.px + (long long)node * nx, *xn = Pn + L.px + (long long)node * nx;
        const double *uo = Po + L.pu + (long long)node * nu, *un = Pn + L.pu + (long long)node * nu;
        // ---- phase 1: classify the second-order-cone block of every child edge (cones.py:113-132) ---------------------
        for (int jj = 0; jj < cc; ++jj) {
            const int j = c0 + jj;
            const long long e0 = j - 1;
            const double *sq = M.sq_d + T.cost_idx[j] * nx, *sr = M.sr_d + T.cost_idx[j] * nu;
            const double *d3 = Do + L.d3 + e0 * nx, *d4 = Do + L.d4 + e0 * nu;
            double ss = 0.0, ss1 = 0.0;
#pragma unroll 4
            for (int k2 = 0; k2 < nx / 2; ++k2) {
                const double2 o2 = ld2(xo, k2), n2 = ld2(xn, k2), d2 = ld2(d3, k2), m2 = ld2(sq, k2);
                const double w0 = dual_w(d2.x, m2.x * (2 * n2.x - o2.x), alpha, inv_alpha);
                const double w1 = dual_w(d2.y, m2.y * (2 * n2.y - o2.y), alpha, inv_alpha);
                ss = fma(w0, w0, ss);
                ss1 = fma(w1, w1, ss1);
            }
#pragma unroll 4
            for (int k2 = 0; k2 < nu / 2; ++k2) {
                const double2 o2 = ld2(uo, k2), n2 = ld2(un, k2), d2 = ld2(d4, k2), m2 = ld2(sr, k2);
                const double w0 = dual_w(d2.x, m2.x * (2 * n2.x - o2.x), alpha, inv_alpha);
                const double w1 = dual_w(d2.y, m2.y * (2 * n2.y - o2.y), alpha, inv_alpha);
                ss = fma(w0, w0, ss);
                ss1 = fma(w1, w1, ss1);
            }
            ss += ss1;
            const double to = Po[L.ptau + j], tn = Pn[L.ptau + j];
            const double ht = 0.5 * (2 * tn - to);
            const double w5 = dual_w(Do[L.d5 + e0], ht, alpha, inv_alpha) - 0.5;
            const double w6 = dual_w(Do[L.d6 + e0], ht, alpha, inv_alpha) + 0.5;
            ss = fma(w5, w5, ss);
            const double r = sqrt(ss);
            double scale, last;
            if (r <= w6) {          // inside the cone: projection = w
                scale = 1.0;
                last = w6;
            } else if (r <= -w6) {  // inside the polar cone: projection = 0
                scale = 0.0;
                last = 0.0;
            } else {
                last = (r + w6) / 2;
                scale = last / r;   // reference: last * (w / r) entrywise
            }
            soc_scale[jj][tid] = scale;
            soc_last[jj][tid] = last;
        }
        // ---- phase 2: one pass over the [x; u] rows: d3/d4 of every child, d7, and the x / u residual rows ----------------
        const long long ri = L.has_nl_rect ? (long long)T.nl_rect_idx[node] * nxu : 0;
        const double *d7o = Do + L.d7 + (long long)node * nxu;
        double *d7n = Dn + L.d7 + (long long)node * nxu;
        // one entry of the [x; u] rows: old / new primal value, old d7 value, bounds -> new d7 value, residual rows;
        // d3 / d4 of the children are handled by the caller through `edge`
        auto entry = [&](double o, double nw, double d7old, double lo_b, double hi_b, double g1, double g2, double &d7new) {
            const double z = 2 * nw - o, dlt = nw - o;
            if (L.has_nl_rect) {   // rectangle on [x; u] (cache.py:367-371)
                const double wv = dual_w(d7old, z, alpha, inv_alpha);
                d7new = alpha * (wv - box_clip(wv, lo_b, hi_b, &bad));
                const double dd = d7old - d7new;
                g1 += dd;
                g2 += R.dual(dd, dlt, inv_alpha);
            }
            R.primal(dlt, g1, g2, inv_alpha);
        };
        auto edge = [&](double mm, double z, double dlt, double dol, double scale, double &dnew, double &g1, double &g2) {
            const double wv = dual_w(dol, mm * z, alpha, inv_alpha);
            dnew = alpha * (wv - scale * wv);
            const double dd = dol - dnew;
            const double xi2 = R.dual(dd, mm * dlt, inv_alpha);
            g1 = fma(mm, dd, g1);
            g2 = fma(mm, xi2, g2);
        };
        for (int part = 0; part < 2; ++part) {   // part 0: x rows with d3 / sqrtQ, part 1: u rows with d4 / sqrtR
            const int w = part == 0 ? nx : nu, off7 = part == 0 ? 0 : nx;
            const double *po_row = part == 0 ? xo : uo, *pn_row = part == 0 ? xn : un;
            const double *mtab = part == 0 ? M.sq_d : M.sr_d;
            const long long seg = part == 0 ? L.d3 : L.d4;
#pragma unroll 2
            for (int k2 = 0; k2 < w / 2; ++k2) {
                const double2 o2 = ld2(po_row, k2), n2 = ld2(pn_row, k2);
                double2 d7v = make_double2(0.0, 0.0), lo2 = d7v, hi2 = d7v;
                if (L.has_nl_rect) {
                    d7v = ld2(d7o + off7, k2);
                    lo2 = ld2(M.nl_lo + ri + off7, k2);
                    hi2 = ld2(M.nl_hi + ri + off7, k2);
                }
                const double z0 = 2 * n2.x - o2.x, z1 = 2 * n2.y - o2.y, dl0 = n2.x - o2.x, dl1 = n2.y - o2.y;
                double g10 = 0.0, g20 = 0.0, g11 = 0.0, g21 = 0.0;
                for (int jj = 0; jj < cc; ++jj) {
                    const int j = c0 + jj;
                    const double2 m2 = ld2(mtab + T.cost_idx[j] * w, k2);
                    const double *dseg = Do + seg + (long long)(j - 1) * w;
                    const double2 dol2 = ld2(dseg, k2);
                    const double scale = soc_scale[jj][tid];
                    double dn0, dn1;
                    edge(m2.x, z0, dl0, dol2.x, scale, dn0, g10, g20);
                    edge(m2.y, z1, dl1, dol2.y, scale, dn1, g11, g21);
                    st2(Dn + seg + (long long)(j - 1) * w, k2, dn0, dn1);
                }
                double dn70 = 0.0, dn71 = 0.0;
                entry(o2.x, n2.x, d7v.x, lo2.x, hi2.x, g10, g20, dn70);
                entry(o2.y, n2.y, d7v.y, lo2.y, hi2.y, g11, g21, dn71);
                if (L.has_nl_rect) st2(d7n + off7, k2, dn70, dn71);
            }
        }
        // ---- d5, d6 and the tau_j residual rows ----------------------------------------------------------------------------
        for (int jj = 0; jj < cc; ++jj) {
            const int j = c0 + jj;
            const long long e0 = j - 1;
            const double to = Po[L.ptau + j], tn = Pn[L.ptau + j];
            const double ht = 0.5 * (2 * tn - to), hdt = 0.5 * (tn - to);
            const double do5 = Do[L.d5 + e0], do6 = Do[L.d6 + e0];
            const double w5 = dual_w(do5, ht, alpha, inv_alpha) - 0.5;
            const double w6 = dual_w(do6, ht, alpha, inv_alpha) + 0.5;
            const double dn5 = alpha * (w5 - soc_scale[jj][tid] * w5);
            const double dn6 = alpha * (w6 - soc_last[jj][tid]);
            Dn[L.d5 + e0] = dn5;
            Dn[L.d6 + e0] = dn6;
            const double dd5 = do5 - dn5, dd6 = do6 - dn6;
            const double x5 = R.dual(dd5, hdt, inv_alpha), x6 = R.dual(dd6, hdt, inv_alpha);
            R.primal(tn - to, 0.5 * (dd5 + dd6), 0.5 * (x5 + x6), inv_alpha);
        }
        // ---- d1, d2 (risk blocks) and the y_i, s_i residual rows ----------------------------------------------------------
        const int yo = T.yoff[node], ny = 2 * cc + 1;
        const double *yold = Po + L.py + yo, *ynew = Pn + L.py + yo, *d1o = Do + L.d1 + yo;
        double *d1n = Dn + L.d1 + yo;
        double dot_z = 0.0, dot_d = 0.0;
        for (int e = 0; e < cc; ++e) {
            const double b = T.cond_prob[c0 + e];
            dot_z = fma(b, 2 * ynew[e] - yold[e], dot_z);
            dot_d = fma(b, ynew[e] - yold[e], dot_d);
        }
        dot_z += 2 * ynew[2 * cc] - yold[2 * cc];
        dot_d += ynew[2 * cc] - yold[2 * cc];
        const double so = Po[L.ps + node], sn = Pn[L.ps + node];
        const double do2 = Do[L.d2 + node];
        const double w2 = dual_w(do2, (2 * sn - so) - dot_z, alpha, inv_alpha);
        const double dn2 = alpha * (w2 - fmax(0.0, w2));
        Dn[L.d2 + node] = dn2;
        const double dd2 = do2 - dn2;
        const double xi22 = fma(dd2, inv_alpha, (sn - so) - dot_d);
        R.put(2, xi22);
        R.put(5, dd2);
        R.primal(sn - so, dd2, xi22, inv_alpha);   // s_i of a nonleaf node: its L* row is d2_i
        for (int e = 0; e < ny; ++e) {
            const double b = e < cc ? T.cond_prob[c0 + e] : (e == 2 * cc ? 1.0 : 0.0);
            const double dy = ynew[e] - yold[e];
            const double do1 = d1o[e];
            const double wv = dual_w(do1, 2 * ynew[e] - yold[e], alpha, inv_alpha);
            const double zv = e < 2 * cc ? fmax(0.0, wv) : wv;   // dual of R_+^{2c} x {0} (risks.py:32-33)
            const double dnew = alpha * (wv - zv);
            d1n[e] = dnew;
            const double dd = do1 - dnew;
            const double xi2 = R.dual(dd, dy, inv_alpha);
            R.primal(dy, dd - b * dd2, xi2 - b * xi22, inv_alpha);
        }
    } else if (node < L.n) {
        // ---- leaf: SOC on [d11; d12; d13] (cache.py:375-386), rectangle on d14, x_i and s_i residual rows -----------------
        const int li = node - L.m;
        const double *sq = M.sqf_d + T.leafcost_idx[li] * nx;
        const double *xo = Po + L.px + (long long)node * nx, *xn = Pn + L.px + (long long)node * nx;
        const double *d11o = Do + L.d11 + (long long)li * nx, *d14o = Do + L.d14 + (long long)li * nx;
        double *d11n = Dn + L.d11 + (long long)li * nx, *d14n = Dn + L.d14 + (long long)li * nx;
        const double so = Po[L.ps + node], sn = Pn[L.ps + node];
        const double hs = 0.5 * (2 * sn - so), hds = 0.5 * (sn - so);
        const double do12 = Do[L.d12 + li], do13 = Do[L.d13 + li];
        double ss = 0.0, ss1 = 0.0;
#pragma unroll 4
        for (int k2 = 0; k2 < nx / 2; ++k2) {
            const double2 o2 = ld2(xo, k2), n2 = ld2(xn, k2), d2 = ld2(d11o, k2), m2 = ld2(sq, k2);
            const double w0 = dual_w(d2.x, m2.x * (2 * n2.x - o2.x), alpha, inv_alpha);
            const double w1 = dual_w(d2.y, m2.y * (2 * n2.y - o2.y), alpha, inv_alpha);
            ss = fma(w0, w0, ss);
            ss1 = fma(w1, w1, ss1);
        }
        ss += ss1;
        const double w12 = dual_w(do12, hs, alpha, inv_alpha) - 0.5;
        const double w13 = dual_w(do13, hs, alpha, inv_alpha) + 0.5;
        ss = fma(w12, w12, ss);
        const double r = sqrt(ss);
        double scale, last;
        if (r <= w13) {
            scale = 1.0;
            last = w13;
        } else if (r <= -w13) {
            scale = 0.0;
            last = 0.0;
        } else {
            last = (r + w13) / 2;
            scale = last / r;
        }
        const long long ri = L.has_leaf_rect ? (long long)T.leaf_rect_idx[li] * nx : 0;
        auto leaf_entry = [&](double o, double nw, double mm, double dol, double dol14, double lo_b, double hi_b,
                              double &dnew, double &dn14) {
            const double z = 2 * nw - o, dlt = nw - o;
            const double wv = dual_w(dol, mm * z, alpha, inv_alpha);
            dnew = alpha * (wv - scale * wv);
            const double dd = dol - dnew;
            const double xi2 = R.dual(dd, mm * dlt, inv_alpha);
            double g1 = mm * dd, g2 = mm * xi2;
            if (L.has_leaf_rect) {
                const double wv14 = dual_w(dol14, z, alpha, inv_alpha);
                dn14 = alpha * (wv14 - box_clip(wv14, lo_b, hi_b, &bad));
                const double dd14 = dol14 - dn14;
                g1 += dd14;
                g2 += R.dual(dd14, dlt, inv_alpha);
            }
            R.primal(dlt, g1, g2, inv_alpha);
        };
#pragma unroll 2
        for (int k2 = 0; k2 < nx / 2; ++k2) {
            const double2 o2 = ld2(xo, k2), n2 = ld2(xn, k2), m2 = ld2(sq, k2), dol2 = ld2(d11o, k2);
            double2 d14v = make_double2(0.0, 0.0), lo2 = d14v, hi2 = d14v;
            if (L.has_leaf_rect) {
                d14v = ld2(d14o, k2);
                lo2 = ld2(M.leaf_lo + ri, k2);
                hi2 = ld2(M.leaf_hi + ri, k2);
            }
            double dn0, dn1, q0 = 0.0, q1 = 0.0;
            leaf_entry(o2.x, n2.x, m2.x, dol2.x, d14v.x, lo2.x, hi2.x, dn0, q0);
            leaf_entry(o2.y, n2.y, m2.y, dol2.y, d14v.y, lo2.y, hi2.y, dn1, q1);
            st2(d11n, k2, dn0, dn1);
            if (L.has_leaf_rect) st2(d14n, k2, q0, q1);
        }
        const double dn12 = alpha * (w12 - scale * w12), dn13 = alpha * (w13 - last);
        Dn[L.d12 + li] = dn12;
        Dn[L.d13 + li] = dn13;
        const double dd12 = do12 - dn12, dd13 = do13 - dn13;
        const double xa = R.dual(dd12, hds, inv_alpha), xb = R.dual(dd13, hds, inv_alpha);
        R.primal(sn - so, 0.5 * (dd12 + dd13), 0.5 * (xa + xb), inv_alpha);
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
    if (tid < 6) {
        unsigned long long mval = blockmax[0][tid];
        for (int wv = 1; wv < kLaneThreads / 32; ++wv) mval = blockmax[wv][tid] > mval ? blockmax[wv][tid] : mval;
        atomicMax(reinterpret_cast<unsigned long long *>(slots + (long long)blockIdx.y * 6 + tid), mval);
    }
    if (tid == 0 && blockflags) atomicOr(&ctrl->status, blockflags);
}

}  // namespace rb
