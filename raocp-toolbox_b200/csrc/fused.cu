// fused.cu -- the loop body of Solver.chock (reference solver.py:124-161) as fused sm_100a kernels (FP64).
//
// One Chambolle-Pock iteration =
//   k_fused_primal : pbar = p - alpha L* d  (solver.py:27-39)  +  s_0 -= alpha (cache.py:253-257)
//                    + projection of (y, tau, s) onto the risk kernels (cache.py:290-317)          [node-parallel]
//   k_fused_bwd/fwd: projection of (xbar, ubar) onto the dynamics set, backward / forward DP sweeps
//                    (cache.py:259-288)                                                            [stage by stage]
//   k_fused_dual   : dbar = d + alpha L(2 p+ - p) (solver.py:44-58), prox of g* (cache.py:321-393) and all six
//                    residual inf-norms of solver.py:63-95,137-141 in ONE pass over the duals       [node-parallel]
//   k_check        : stopping test of solver.py:156-161 on the device; once it fires every later launch is a no-op,
//                    so the host can enqueue iterations without synchronising and still stop at the exact iteration.
// Iterates live in two device copies (A, B) that swap roles every iteration; only the residual norms leave the device.
#include "kernels.cuh"
#include "node_ops.cuh"

namespace rb {

// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_fused_primal(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                          const double *__restrict__ p_old,
                                                          const double *__restrict__ d_old, double *__restrict__ p_new,
                                                          double alpha) {
    if (ctrl->done) return;
    const Layout &L = P.L;
    __shared__ double sm[kWarpsPerBlock][4][kMaxDim];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int node = blockIdx.x * kWarpsPerBlock + warp;
    if (node >= L.n) return;
    const double *D = d_old + (long long)blockIdx.y * L.nd_pad;
    const double *Po = p_old + (long long)blockIdx.y * L.np_pad;
    double *Pn = p_new + (long long)blockIdx.y * L.np_pad;
    double *v3 = sm[warp][0], *v4 = sm[warp][1], *ax = sm[warp][2], *au = sm[warp][3];
    const int nx = L.nx, nu = L.nu;

    if (node < L.m) {
        const int c0 = P.t.child_first[node], cc = P.t.child_count[node];
        for (int k = lane; k < nx; k += 32) ax[k] = L.has_nl_rect ? D[L.d7 + (long long)node * L.nxu + k] : 0.0;
        for (int k = lane; k < nu; k += 32) au[k] = L.has_nl_rect ? D[L.d7 + (long long)node * L.nxu + nx + k] : 0.0;
        for (int j = c0; j < c0 + cc; ++j) {
            const long long e = j - 1;
            for (int k = lane; k < nx; k += 32) v3[k] = D[L.d3 + e * nx + k];
            for (int k = lane; k < nu; k += 32) v4[k] = D[L.d4 + e * nu + k];
            __syncwarp();
            const int ci = P.t.cost_idx[j];
            mv_acc(P.m.sqT + (long long)ci * nx * nx, v3, nx, nx, ax, 1.0, lane);
            mv_acc(P.m.srT + (long long)ci * nu * nu, v4, nu, nu, au, 1.0, lane);
            __syncwarp();
        }
        for (int k = lane; k < nx; k += 32) {
            const long long idx = L.px + (long long)node * nx + k;
            Pn[idx] = Po[idx] - alpha * ax[k];
        }
        for (int k = lane; k < nu; k += 32) {
            const long long idx = L.pu + (long long)node * nu + k;
            Pn[idx] = Po[idx] - alpha * au[k];
        }
        // ybar_i, and the children's taubar_j, sbar_j (edge quantities are owned by the parent's warp)
        const double d2v = D[L.d2 + node];
        const int yo = P.t.yoff[node];
        for (int e = lane; e < 2 * cc + 1; e += 32) {
            const double b = e < cc ? P.t.cond_prob[c0 + e] : (e == 2 * cc ? 1.0 : 0.0);
            Pn[L.py + yo + e] = Po[L.py + yo + e] - alpha * (D[L.d1 + yo + e] - b * d2v);
        }
        for (int e = lane; e < cc; e += 32) {
            const int j = c0 + e;
            Pn[L.ptau + j] = Po[L.ptau + j] - alpha * (0.5 * (D[L.d5 + j - 1] + D[L.d6 + j - 1]));
            const double lts = j < L.m ? D[L.d2 + j] : 0.5 * (D[L.d12 + j - L.m] + D[L.d13 + j - L.m]);
            Pn[L.ps + j] = Po[L.ps + j] - alpha * lts;
        }
        if (node == 0 && lane == 0) {
            Pn[L.ps] = (Po[L.ps] - alpha * d2v) - alpha;   // s_0: half step, then prox of alpha * identity
            Pn[L.ptau] = Po[L.ptau] - alpha * Po[L.ptau];  // tau_0 (always 0; same arithmetic as the reference)
        }
        __syncwarp();
        kernel_projection(P, Pn, node, lane);
    } else {
        const long long li = node - L.m;
        for (int k = lane; k < nx; k += 32) {
            v3[k] = D[L.d11 + li * nx + k];
            ax[k] = L.has_leaf_rect ? D[L.d14 + li * nx + k] : 0.0;
        }
        __syncwarp();
        mv_acc(P.m.sqfT + (long long)P.t.leafcost_idx[li] * nx * nx, v3, nx, nx, ax, 1.0, lane);
        __syncwarp();
        for (int k = lane; k < nx; k += 32) {
            const long long idx = L.px + (long long)node * nx + k;
            Pn[idx] = Po[idx] - alpha * ax[k];
        }
    }
}

// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_fused_bwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                       const double *__restrict__ prim, double *__restrict__ q,
                                                       double *__restrict__ r, int lo, int hi) {
    if (ctrl->done) return;
    __shared__ double sm[kWarpsPerBlock][4][kMaxDim];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int node = lo + blockIdx.x * kWarpsPerBlock + warp;
    if (node >= hi) return;
    dyn_bwd_node(P, prim + (long long)blockIdx.y * P.L.np_pad, q + (long long)blockIdx.y * P.L.n * P.L.nx,
                 r + (long long)blockIdx.y * P.L.m * P.L.nu, node, lane, sm[warp]);
}

__global__ void __launch_bounds__(kThreads) k_fused_fwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                       double *__restrict__ prim, const double *__restrict__ r,
                                                       const double *__restrict__ x0, int lo, int hi) {
    if (ctrl->done) return;
    __shared__ double sm[kWarpsPerBlock][4][kMaxDim];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int node = lo + blockIdx.x * kWarpsPerBlock + warp;
    if (node >= hi || node >= P.L.m) return;
    double *Pp = prim + (long long)blockIdx.y * P.L.np_pad;
    if (node == 0) {  // x_0 <- initial state (cache.py:282)
        for (int k = lane; k < P.L.nx; k += 32) Pp[P.L.px + k] = x0[blockIdx.y * P.L.nx + k];
        __syncwarp();
    }
    dyn_fwd_node(P, Pp, r + (long long)blockIdx.y * P.L.m * P.L.nu, node, lane, sm[warp]);
}

// ----------------------------------------------------------------------------------------------------------------
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
    // primal entry: old value po, new value pn, g1 = [L*(d - d+)] entry, g2 = [L* xi2] entry
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

__global__ void __launch_bounds__(kThreads) k_fused_dual(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl,
                                                        const double *__restrict__ p_old, const double *__restrict__ p_new,
                                                        const double *__restrict__ d_old, double *__restrict__ d_new,
                                                        double alpha, double *__restrict__ slots) {
    if (ctrl->done) return;
    const Layout &L = P.L;
    // warp-private rows
    enum { ZX, DX, ZU, DU, V1, V2, V1U, V2U, G1, G2, G1U, G2U, kRows };
    __shared__ double sm[kWarpsPerBlock][kRows][kMaxDim];
    __shared__ double wbuf[kWarpsPerBlock][2 * kMaxDim + 2];
    __shared__ double blockmax[kWarpsPerBlock][6];
    __shared__ int blocknan;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int node = blockIdx.x * kWarpsPerBlock + warp;
    if (threadIdx.x == 0) blocknan = 0;
    const double *Po = p_old + (long long)blockIdx.y * L.np_pad;
    const double *Pn = p_new + (long long)blockIdx.y * L.np_pad;
    const double *Do = d_old + (long long)blockIdx.y * L.nd_pad;
    double *Dn = d_new + (long long)blockIdx.y * L.nd_pad;
    const int nx = L.nx, nu = L.nu;
    double *zx = sm[warp][ZX], *dx = sm[warp][DX], *zu = sm[warp][ZU], *du = sm[warp][DU];
    double *v1 = sm[warp][V1], *v2 = sm[warp][V2], *v1u = sm[warp][V1U], *v2u = sm[warp][V2U];
    double *g1 = sm[warp][G1], *g2 = sm[warp][G2], *g1u = sm[warp][G1U], *g2u = sm[warp][G2U];
    double *w = wbuf[warp];
    Resid R;
    R.init();
    int bad = 0;

    if (node < L.n) {
        for (int k = lane; k < nx; k += 32) {
            const double xo = Po[L.px + (long long)node * nx + k], xn = Pn[L.px + (long long)node * nx + k];
            zx[k] = 2 * xn - xo;
            dx[k] = xn - xo;
            g1[k] = 0.0;
            g2[k] = 0.0;
        }
    }

    if (node < L.m) {
        for (int k = lane; k < nu; k += 32) {
            const double uo = Po[L.pu + (long long)node * nu + k], un = Pn[L.pu + (long long)node * nu + k];
            zu[k] = 2 * un - uo;
            du[k] = un - uo;
            g1u[k] = 0.0;
            g2u[k] = 0.0;
        }
        __syncwarp();
        const int c0 = P.t.child_first[node], cc = P.t.child_count[node];
        const int dim = nx + nu + 2;
        for (int j = c0; j < c0 + cc; ++j) {
            const long long e = j - 1;
            const int ci = P.t.cost_idx[j];
            const double *sqT = P.m.sqT + (long long)ci * nx * nx;
            const double *srT = P.m.srT + (long long)ci * nu * nu;
            // pass A: L z and L (p+ - p) on this edge, then w = (d + alpha L z) / alpha
            for (int k = lane; k < nx; k += 32) {
                double la = 0.0, lb = 0.0;
                for (int l = 0; l < nx; ++l) {
                    const double mkl = __ldg(sqT + (long long)l * nx + k);
                    la = fma(mkl, zx[l], la);
                    lb = fma(mkl, dx[l], lb);
                }
                w[k] = dual_w(Do[L.d3 + e * nx + k], la, alpha, 0.0);
                v2[k] = lb;
            }
            for (int k = lane; k < nu; k += 32) {
                double la = 0.0, lb = 0.0;
                for (int l = 0; l < nu; ++l) {
                    const double mkl = __ldg(srT + (long long)l * nu + k);
                    la = fma(mkl, zu[l], la);
                    lb = fma(mkl, du[l], lb);
                }
                w[nx + k] = dual_w(Do[L.d4 + e * nu + k], la, alpha, 0.0);
                v2u[k] = lb;
            }
            const double to = Po[L.ptau + j], tn = Pn[L.ptau + j];
            if (lane == 0) {
                const double ht = 0.5 * (2 * tn - to);
                w[nx + nu] = dual_w(Do[L.d5 + e], ht, alpha, -0.5);
                w[nx + nu + 1] = dual_w(Do[L.d6 + e], ht, alpha, 0.5);
            }
            __syncwarp();
            const SocResult sr = soc_classify(w, dim, lane);
            for (int k = lane; k < nx; k += 32) {
                const double dn = alpha * (w[k] - soc_entry(sr, w[k], false));
                const double dol = Do[L.d3 + e * nx + k];
                Dn[L.d3 + e * nx + k] = dn;
                const double dd = dol - dn;
                const double xi2 = dd / alpha + v2[k];
                R.put(2, xi2);
                R.put(5, dn - dol);
                v1[k] = dd;
                v2[k] = xi2;
            }
            for (int k = lane; k < nu; k += 32) {
                const double dn = alpha * (w[nx + k] - soc_entry(sr, w[nx + k], false));
                const double dol = Do[L.d4 + e * nu + k];
                Dn[L.d4 + e * nu + k] = dn;
                const double dd = dol - dn;
                const double xi2 = dd / alpha + v2u[k];
                R.put(2, xi2);
                R.put(5, dn - dol);
                v1u[k] = dd;
                v2u[k] = xi2;
            }
            if (lane == 0) {
                const double w5 = w[nx + nu], w6 = w[nx + nu + 1];
                const double dn5 = alpha * (w5 - soc_entry(sr, w5, false));
                const double dn6 = alpha * (w6 - soc_entry(sr, w6, true));
                const double do5 = Do[L.d5 + e], do6 = Do[L.d6 + e];
                Dn[L.d5 + e] = dn5;
                Dn[L.d6 + e] = dn6;
                const double dd5 = do5 - dn5, dd6 = do6 - dn6;
                const double hdt = 0.5 * (tn - to);
                const double xi25 = dd5 / alpha + hdt, xi26 = dd6 / alpha + hdt;
                R.put(2, xi25);
                R.put(2, xi26);
                R.put(5, dn5 - do5);
                R.put(5, dn6 - do6);
                R.primal(to, tn, 0.5 * (dd5 + dd6), 0.5 * (xi25 + xi26), alpha);
            }
            __syncwarp();
            // pass B: child -> parent sums  g1 += sqrtQ_j dd3_j,  g2 += sqrtQ_j xi2_3j  (same for R / d4)
            for (int k = lane; k < nx; k += 32) {
                double a1 = 0.0, a2 = 0.0;
                for (int l = 0; l < nx; ++l) {
                    const double mkl = __ldg(sqT + (long long)l * nx + k);
                    a1 = fma(mkl, v1[l], a1);
                    a2 = fma(mkl, v2[l], a2);
                }
                g1[k] += a1;
                g2[k] += a2;
            }
            for (int k = lane; k < nu; k += 32) {
                double a1 = 0.0, a2 = 0.0;
                for (int l = 0; l < nu; ++l) {
                    const double mkl = __ldg(srT + (long long)l * nu + k);
                    a1 = fma(mkl, v1u[l], a1);
                    a2 = fma(mkl, v2u[l], a2);
                }
                g1u[k] += a1;
                g2u[k] += a2;
            }
            __syncwarp();
        }
        // d7: rectangle on [x; u] (cache.py:367-371)
        if (L.has_nl_rect) {
            const long long ri = (long long)P.t.nl_rect_idx[node] * L.nxu;
            for (int k = lane; k < L.nxu; k += 32) {
                const bool isx = k < nx;
                const double zk = isx ? zx[k] : zu[k - nx];
                const double dk = isx ? dx[k] : du[k - nx];
                const long long idx = L.d7 + (long long)node * L.nxu + k;
                const double dol = Do[idx];
                const double wv = dual_w(dol, zk, alpha, 0.0);
                const double dn = alpha * (wv - box_clip(wv, P.m.nl_lo[ri + k], P.m.nl_hi[ri + k], &bad));
                Dn[idx] = dn;
                const double dd = dol - dn;
                const double xi2 = dd / alpha + dk;
                R.put(2, xi2);
                R.put(5, dn - dol);
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
            R.primal(Po[L.px + (long long)node * nx + k], Pn[L.px + (long long)node * nx + k], g1[k], g2[k], alpha);
        for (int k = lane; k < nu; k += 32)
            R.primal(Po[L.pu + (long long)node * nu + k], Pn[L.pu + (long long)node * nu + k], g1u[k], g2u[k], alpha);
        // d1, d2 (risk blocks) and the y_i, s_i residual rows
        const int yo = P.t.yoff[node];
        double dot_z = 0.0, dot_d = 0.0;
        for (int e = lane; e < 2 * cc + 1; e += 32) {
            const double yold = Po[L.py + yo + e], ynew = Pn[L.py + yo + e];
            const double b = e < cc ? P.t.cond_prob[c0 + e] : (e == 2 * cc ? 1.0 : 0.0);
            dot_z = fma(b, 2 * ynew - yold, dot_z);
            dot_d = fma(b, ynew - yold, dot_d);
        }
        dot_z = warp_sum(dot_z);
        dot_d = warp_sum(dot_d);
        const double so = Po[L.ps + node], sn = Pn[L.ps + node];
        const double do2 = Do[L.d2 + node];
        const double w2 = dual_w(do2, (2 * sn - so) - dot_z, alpha, 0.0);
        const double dn2 = alpha * (w2 - fmax(0.0, w2));
        const double dd2 = do2 - dn2;
        const double xi22 = dd2 / alpha + ((sn - so) - dot_d);
        if (lane == 0) {
            Dn[L.d2 + node] = dn2;
            R.put(2, xi22);
            R.put(5, dn2 - do2);
            R.primal(so, sn, dd2, xi22, alpha);   // s_i of a nonleaf node: L* row is d2_i
        }
        for (int e = lane; e < 2 * cc + 1; e += 32) {
            const double yold = Po[L.py + yo + e], ynew = Pn[L.py + yo + e];
            const double b = e < cc ? P.t.cond_prob[c0 + e] : (e == 2 * cc ? 1.0 : 0.0);
            const double do1 = Do[L.d1 + yo + e];
            const double wv = dual_w(do1, 2 * ynew - yold, alpha, 0.0);
            const double zv = e < 2 * cc ? fmax(0.0, wv) : wv;
            const double dn = alpha * (wv - zv);
            Dn[L.d1 + yo + e] = dn;
            const double dd = do1 - dn;
            const double xi2 = dd / alpha + (ynew - yold);
            R.put(2, xi2);
            R.put(5, dn - do1);
            R.primal(yold, ynew, dd - b * dd2, xi2 - b * xi22, alpha);
        }
    } else if (node < L.n) {
        __syncwarp();
        const long long li = node - L.m;
        const double *sqfT = P.m.sqfT + (long long)P.t.leafcost_idx[li] * nx * nx;
        const int dim = nx + 2;
        for (int k = lane; k < nx; k += 32) {
            double la = 0.0, lb = 0.0;
            for (int l = 0; l < nx; ++l) {
                const double mkl = __ldg(sqfT + (long long)l * nx + k);
                la = fma(mkl, zx[l], la);
                lb = fma(mkl, dx[l], lb);
            }
            w[k] = dual_w(Do[L.d11 + li * nx + k], la, alpha, 0.0);
            v2[k] = lb;
        }
        const double so = Po[L.ps + node], sn = Pn[L.ps + node];
        if (lane == 0) {
            const double hs = 0.5 * (2 * sn - so);
            w[nx] = dual_w(Do[L.d12 + li], hs, alpha, -0.5);
            w[nx + 1] = dual_w(Do[L.d13 + li], hs, alpha, 0.5);
        }
        __syncwarp();
        const SocResult sr = soc_classify(w, dim, lane);
        for (int k = lane; k < nx; k += 32) {
            const double dn = alpha * (w[k] - soc_entry(sr, w[k], false));
            const double dol = Do[L.d11 + li * nx + k];
            Dn[L.d11 + li * nx + k] = dn;
            const double dd = dol - dn;
            const double xi2 = dd / alpha + v2[k];
            R.put(2, xi2);
            R.put(5, dn - dol);
            v1[k] = dd;
            v2[k] = xi2;
        }
        if (lane == 0) {
            const double w12 = w[nx], w13 = w[nx + 1];
            const double dn12 = alpha * (w12 - soc_entry(sr, w12, false));
            const double dn13 = alpha * (w13 - soc_entry(sr, w13, true));
            const double do12 = Do[L.d12 + li], do13 = Do[L.d13 + li];
            Dn[L.d12 + li] = dn12;
            Dn[L.d13 + li] = dn13;
            const double dd12 = do12 - dn12, dd13 = do13 - dn13;
            const double hds = 0.5 * (sn - so);
            const double xa = dd12 / alpha + hds, xb = dd13 / alpha + hds;
            R.put(2, xa);
            R.put(2, xb);
            R.put(5, dn12 - do12);
            R.put(5, dn13 - do13);
            R.primal(so, sn, 0.5 * (dd12 + dd13), 0.5 * (xa + xb), alpha);
        }
        __syncwarp();
        for (int k = lane; k < nx; k += 32) {
            double a1 = 0.0, a2 = 0.0;
            for (int l = 0; l < nx; ++l) {
                const double mkl = __ldg(sqfT + (long long)l * nx + k);
                a1 = fma(mkl, v1[l], a1);
                a2 = fma(mkl, v2[l], a2);
            }
            if (L.has_leaf_rect) {
                const long long ri = (long long)P.t.leaf_rect_idx[li] * nx;
                const long long idx = L.d14 + li * nx + k;
                const double dol = Do[idx];
                const double wv = dual_w(dol, zx[k], alpha, 0.0);
                const double dn = alpha * (wv - box_clip(wv, P.m.leaf_lo[ri + k], P.m.leaf_hi[ri + k], &bad));
                Dn[idx] = dn;
                const double dd = dol - dn;
                const double xi2 = dd / alpha + dx[k];
                R.put(2, xi2);
                R.put(5, dn - dol);
                a1 += dd;
                a2 += xi2;
            }
            R.primal(Po[L.px + (long long)node * nx + k], Pn[L.px + (long long)node * nx + k], a1, a2, alpha);
        }
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
    if (lane == 0 && (anynan || anybad)) atomicOr(&blocknan, (anynan ? 2 : 0) | (anybad ? 1 : 0));
    __syncthreads();
    if (threadIdx.x < 6) {
        double mval = blockmax[0][threadIdx.x];
#pragma unroll
        for (int wv = 1; wv < kWarpsPerBlock; ++wv) mval = fmax(mval, blockmax[wv][threadIdx.x]);
        atomic_max_nonneg(slots + (long long)blockIdx.y * 6 + threadIdx.x, mval);
    }
    if (threadIdx.x == 0 && blocknan) atomicOr(&ctrl->status, blocknan);
}

// ----------------------------------------------------------------------------------------------------------------
// stopping test (solver.py:137-161).  slots: [batch][6] maxima of the iteration that just finished; they are copied
// into the history and reset.  The loop stops when the iteration index reaches max_iters or every instance has
// max(xi0, xi1, xi2) <= tol.
__global__ void k_check(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl, double *__restrict__ slots,
                        double *__restrict__ last, double *__restrict__ hist, int hist_capacity, int max_iters,
                        double tol) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (ctrl->done) return;
    const int it = ctrl->iters;
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
