// node_ops.cuh -- per-node device bodies of the DP sweeps and of the kernel projection, shared by ops.cu / sweeps.cu.
//
// The sweep bodies are templates on <NX, NU>: with compile-time sizes the matrix-vector loops unroll completely (every
// matrix load has an immediate offset and is issued up front, ~3 instructions per multiply-add instead of ~8), with
// NX = 0 the same code runs on run-time sizes.
#pragma once
#include "common.cuh"

namespace rb {

// one output row per lane with compile-time shape: sum_l MT[l*ROWS + k] * v[l]
template <int ROWS, int COLS>
__device__ __forceinline__ double mv_row_t(const double *__restrict__ MT, const double *__restrict__ v, int k) {
    // all matrix loads are issued before the first multiply-add: one L1/L2 round trip per product, not one per term
    double mcol[COLS];
#pragma unroll
    for (int l = 0; l < COLS; ++l) mcol[l] = __ldg(MT + l * ROWS + k);
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int l = 0; l < COLS; ++l) {
        if (l & 1) a1 = fma(mcol[l], v[l], a1);
        else a0 = fma(mcol[l], v[l], a0);
    }
    return a0 + a1;
}

// contribution of child j to its parent's backward step: [A_j' q_j ; B_j' q_j]  (nx+nu entries, entry k -> out[k])
// qj: the child's q staged in a warp-private shared row.  Rows k = lane, lane+32, ...
template <int NX, int NU>
__device__ __forceinline__ void bwd_child_contrib(const Layout &L, const Tabs &M, int dyn, const double *qj, int lane,
                                                  double *out, bool accumulate) {
    if constexpr (NX > 0) {
        constexpr int NXU = NX + NU;
        const double *C = M.ABcat + dyn * (NX * NXU);
        for (int k = lane; k < NXU; k += 32) {
            const double c = mv_row_t<NXU, NX>(C, qj, k);
            out[k] = accumulate ? out[k] + c : c;
        }
    } else {
        const int nx = L.nx, nxu = L.nxu;
        const double *C = M.ABcat + (long long)dyn * nx * nxu;
        for (int k = lane; k < nxu; k += 32) {
            const double c = mv_row(C, qj, nxu, nx, k);
            out[k] = accumulate ? out[k] + c : c;
        }
    }
}

// finish the backward step at nonleaf node i given acc = sum_j [A_j' q_j ; B_j' q_j] (warp-private shared row):
//   r_i = ubar_i - acc[nx:],   q_i = -xbar_i + acc[:nx] - K_i' r_i      (reference cache.py:264-280)
// rv: warp-private shared row (nu).
template <int NX, int NU>
__device__ __forceinline__ void bwd_finish(const Layout &L, const Topo &T, const Tabs &M, const double *__restrict__ X,
                                           const double *__restrict__ U, double *__restrict__ Q, double *__restrict__ R,
                                           int node, int lane, const double *acc, double *rv) {
    const int nx = NX > 0 ? NX : L.nx, nu = NX > 0 ? NU : L.nu;
    for (int k = lane; k < nu; k += 32) {
        const double rk = U[node * nu + k] - acc[nx + k];
        rv[k] = rk;
        R[node * nu + k] = rk;
    }
    __syncwarp();
    const double *Kc = M.K + (long long)T.cls[node] * nu * nx;
    for (int k = lane; k < nx; k += 32) {
        double kr;
        if constexpr (NX > 0) kr = mv_row_t<NX, NU>(Kc, rv, k);
        else kr = mv_row(Kc, rv, nx, nu, k);
        Q[node * nx + k] = acc[k] - X[node * nx + k] - kr;
    }
    __syncwarp();
}

// Backward DP step at one node, all children handled by this warp (chains and small fan-out).
// scratch: 2*(nx+nu) warp-private doubles.
template <int NX, int NU>
__device__ __forceinline__ void dyn_bwd_node(const Layout &L, const Topo &T, const Tabs &M, const double *__restrict__ X,
                                             const double *__restrict__ U, double *__restrict__ Q,
                                             double *__restrict__ R, int node, int lane, double *scratch) {
    const int nx = NX > 0 ? NX : L.nx, nxu = NX > 0 ? NX + NU : L.nxu;
    if (node >= L.m) {   // leaf: q_i = -xbar_i
        for (int k = lane; k < nx; k += 32) Q[node * nx + k] = -X[node * nx + k];
        return;
    }
    double *qj = scratch, *acc = scratch + nxu, *rv = qj + nx;
    const int c0 = T.child_first[node], cc = T.child_count[node];
    for (int j = c0; j < c0 + cc; ++j) {
        for (int k = lane; k < nx; k += 32) qj[k] = Q[j * nx + k];
        __syncwarp();
        bwd_child_contrib<NX, NU>(L, M, T.dyn_idx[j], qj, lane, acc, j > c0);
        __syncwarp();
    }
    bwd_finish<NX, NU>(L, T, M, X, U, Q, R, node, lane, acc, rv);
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
template <int ROWS, int COLS>
__device__ __forceinline__ double mv_split_t(const double *__restrict__ MT, const double *__restrict__ v, int lane,
                                             double *part) {
    if constexpr (ROWS > 16) {
        return lane < ROWS ? mv_row_t<ROWS, COLS>(MT, v, lane) : 0.0;
    } else {
        constexpr int G = 32 / ROWS, SEG = (COLS + G - 1) / G;
        const int g = lane / ROWS, a = lane - g * ROWS;
        double p = 0.0;
        if (g < G) {
            double mcol[SEG];
#pragma unroll
            for (int i = 0; i < SEG; ++i) {
                const int l = g * SEG + i;
                mcol[i] = l < COLS ? __ldg(MT + l * ROWS + a) : 0.0;
            }
#pragma unroll
            for (int i = 0; i < SEG; ++i) {
                const int l = g * SEG + i;
                if (l < COLS) p = fma(mcol[i], v[l], p);
            }
        }
        part[lane] = p;
        __syncwarp();
        double out = 0.0;
        if (lane < ROWS) {
#pragma unroll
            for (int gg = 0; gg < G; ++gg) out += part[gg * ROWS + lane];
        }
        __syncwarp();
        return out;
    }
}

// first half of the forward step at nonleaf node i: u_i = K_i x_i + R~_i^-1 r_i, written to U and left behind x_i in
// v = [x_i ; u_i] (warp-private shared row, nx+nu).  part: 32 warp-private doubles.
template <int NX, int NU>
__device__ __forceinline__ void fwd_input(const Layout &L, const Topo &T, const Tabs &M, const double *__restrict__ X,
                                          double *__restrict__ U, const double *__restrict__ R, int node, int lane,
                                          double *v, double *part) {
    const int nx = NX > 0 ? NX : L.nx, nu = NX > 0 ? NU : L.nu, nxu = nx + nu;
    for (int k = lane; k < nx; k += 32) v[k] = X[node * nx + k];
    for (int k = lane; k < nu; k += 32) v[nx + k] = R[node * nu + k];
    __syncwarp();
    const double *KR = M.KRcatT + (long long)T.cls[node] * nxu * nu;
    if (nu <= 32) {
        double ua;
        if constexpr (NX > 0) ua = mv_split_t<NU, NX + NU>(KR, v, lane, part);
        else ua = mv_split(KR, v, nu, nxu, lane, part);
        if (lane < nu) {
            v[nx + lane] = ua;   // r is dead after mv_split's internal barrier
            U[node * nu + lane] = ua;
        }
    } else {
        const double u0 = mv_row(KR, v, nu, nxu, lane), u1 = lane + 32 < nu ? mv_row(KR, v, nu, nxu, lane + 32) : 0.0;
        __syncwarp();
        v[nx + lane] = u0;
        U[node * nu + lane] = u0;
        if (lane + 32 < nu) {
            v[nx + lane + 32] = u1;
            U[node * nu + lane + 32] = u1;
        }
    }
    __syncwarp();
}

// second half: x_j = A_j x_i + B_j u_i for one child j, from the parent's v = [x_i ; u_i]
template <int NX, int NU>
__device__ __forceinline__ void fwd_child(const Layout &L, const Tabs &M, int dyn, const double *v, double *__restrict__ X,
                                          int j, int lane) {
    if constexpr (NX > 0) {
        const double *C = M.ABcatT + dyn * ((NX + NU) * NX);
        for (int k = lane; k < NX; k += 32) X[j * NX + k] = mv_row_t<NX, NX + NU>(C, v, k);
    } else {
        const int nx = L.nx, nxu = L.nxu;
        const double *C = M.ABcatT + (long long)dyn * nxu * nx;
        for (int k = lane; k < nx; k += 32) X[j * nx + k] = mv_row(C, v, nx, nxu, k);
    }
}

// Forward DP step at one nonleaf node, all children handled by this warp (reference cache.py:282-288):
//   u_i = K_i x_i + R~_i^-1 r_i;   x_j = A_j x_i + B_j u_i   ( = (A_j+B_jK_i) x_i + B_j d_i of the reference )
// scratch: (nx+nu) + 32 warp-private doubles.
template <int NX, int NU>
__device__ __forceinline__ void dyn_fwd_node(const Layout &L, const Topo &T, const Tabs &M, double *__restrict__ X,
                                             double *__restrict__ U, const double *__restrict__ R, int node, int lane,
                                             double *scratch) {
    const int nxu = NX > 0 ? NX + NU : L.nxu;
    double *v = scratch, *part = scratch + nxu;
    fwd_input<NX, NU>(L, T, M, X, U, R, node, lane, v, part);
    const int c0 = T.child_first[node], cc = T.child_count[node];
    for (int j = c0; j < c0 + cc; ++j) fwd_child<NX, NU>(L, M, T.dyn_idx[j], v, X, j, lane);
    __syncwarp();
}

// ---- chain walkers ------------------------------------------------------------------------------------------------------
// Below the stopping time every subtree is a chain; one warp walks it, lane k owns output row k.  Everything a step
// needs besides the previous step's result is fetched ahead of time:
//   * the chain's node ids, dynamics rows and class ids are loaded once (lane d holds depth d, broadcast by shuffle);
//   * with compile-time sizes the lane's column of the dynamics tables ([A | B] for the backward, [A ; B]' for the forward
//     step) lives in REGISTERS for the whole chain (reloaded only if the dynamics row changes along the chain -- never on
//     a Markov tree, where the mode is frozen after the stopping time);
//   * the class-indexed K column / [K R~^-1] segment and the xbar / ubar / r rows of the NEXT step are prefetched into
//     registers while the current step computes;
//   * q / x are carried in shared memory -- q of interior chain nodes never goes to HBM.
// chain[d] = node at depth d (d = 0 is the head).  depth <= 64.
template <int COLS>
__device__ __forceinline__ double dot_col(const double (&m)[COLS], const double *__restrict__ v) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int l = 0; l < COLS; ++l) {
        if ((l & 3) == 0) a0 = fma(m[l], v[l], a0);
        else if ((l & 3) == 1) a1 = fma(m[l], v[l], a1);
        else if ((l & 3) == 2) a2 = fma(m[l], v[l], a2);
        else a3 = fma(m[l], v[l], a3);
    }
    return (a0 + a1) + (a2 + a3);
}
template <int COLS>
__device__ __forceinline__ void load_col(double (&m)[COLS], const double *__restrict__ MT, int rows, int k, bool active) {
#pragma unroll
    for (int l = 0; l < COLS; ++l) m[l] = active ? __ldg(MT + l * rows + k) : 0.0;
}

template <int NX, int NU>
__device__ __forceinline__ void chain_bwd(const Layout &L, const Topo &T, const Tabs &M, const double *__restrict__ X,
                                          const double *__restrict__ U, double *__restrict__ Q, double *__restrict__ R,
                                          const int *__restrict__ chain, int depth, int lane, double *scratch) {
    const int nx = NX > 0 ? NX : L.nx, nu = NX > 0 ? NU : L.nu, nxu = nx + nu;
    double *qj = scratch, *rv = scratch + nx, *acc = scratch + nxu;
    // per-depth metadata: lanes hold depths lane and lane + 32
    int node_a = lane < depth ? chain[lane] : 0, node_b = lane + 32 < depth ? chain[lane + 32] : 0;
    int dyn_a = lane + 1 < depth ? T.dyn_idx[chain[lane + 1]] : 0, dyn_b = lane + 33 < depth ? T.dyn_idx[chain[lane + 33]] : 0;
    int cls_a = (lane < depth && node_a < L.m) ? T.cls[node_a] : 0, cls_b = (lane + 32 < depth && node_b < L.m) ? T.cls[node_b] : 0;
    auto at = [&](int va, int vb, int d) { return d < 32 ? __shfl_sync(0xffffffffu, va, d) : __shfl_sync(0xffffffffu, vb, d - 32); };
    int d = depth - 1;
    int node = at(node_a, node_b, d);
    // rows of the current step (lane k: xbar[k], k < nx; ubar[k], k < nu; second register for widths > 32)
    double xb0 = lane < nx ? X[node * nx + lane] : 0.0, xb1 = lane + 32 < nx ? X[node * nx + lane + 32] : 0.0;
    double ub0 = 0.0, ub1 = 0.0;
    if (node < L.m) {
        ub0 = lane < nu ? U[node * nu + lane] : 0.0;
        ub1 = lane + 32 < nu ? U[node * nu + lane + 32] : 0.0;
    }
    // register-resident table columns (compile-time sizes, nx + nu <= 32): ccol = column `lane` of [A | B] of the current
    // dynamics row, kcol = column `lane` of K of the current step's class (prefetched one step ahead)
    constexpr int CN = NX > 0 ? NX : 1, KN = NX > 0 ? NU : 1;
    constexpr bool kRegs = NX > 0 && NX + NU <= 32;
    double ccol[CN], kcol[KN], kcol_next[KN];
    int dyn_loaded = -1;
    if constexpr (kRegs) {
        if (d > 0 || node < L.m) {
            const int dn = node < L.m ? d : d - 1;   // first nonleaf step
            if (dn >= 0) load_col<KN>(kcol, M.K + (long long)at(cls_a, cls_b, dn) * NU * NX, NX, lane, lane < NX);
        }
    }
    for (; d >= 0; --d) {
        // prefetch the rows (and the K column) of the next step (depth d - 1)
        double nxb0 = 0.0, nxb1 = 0.0, nub0 = 0.0, nub1 = 0.0;
        int next = 0;
        if (d > 0) {
            next = at(node_a, node_b, d - 1);
            nxb0 = lane < nx ? X[next * nx + lane] : 0.0;
            nxb1 = lane + 32 < nx ? X[next * nx + lane + 32] : 0.0;
            nub0 = lane < nu ? U[next * nu + lane] : 0.0;
            nub1 = lane + 32 < nu ? U[next * nu + lane + 32] : 0.0;
            if constexpr (kRegs) {
                if (node < L.m)   // (when the current node is a leaf, kcol already holds the next step's column)
                    load_col<KN>(kcol_next, M.K + (long long)at(cls_a, cls_b, d - 1) * NU * NX, NX, lane, lane < NX);
            }
        }
        if (node >= L.m) {   // leaf: q = -xbar
            if (lane < nx) qj[lane] = -xb0;
            if (lane + 32 < nx) qj[lane + 32] = -xb1;
        } else {
            // acc = [A' q ; B' q] of the single child (the chain's previous node, q in shared memory)
            const int dyn = at(dyn_a, dyn_b, d);
            if constexpr (kRegs) {
                if (dyn != dyn_loaded) {
                    load_col<CN>(ccol, M.ABcat + dyn * (NX * (NX + NU)), NX + NU, lane, lane < NX + NU);
                    dyn_loaded = dyn;
                }
                if (lane < NX + NU) acc[lane] = dot_col<CN>(ccol, qj);
            } else {
                bwd_child_contrib<NX, NU>(L, M, dyn, qj, lane, acc, false);
            }
            __syncwarp();
            // r = ubar - acc[nx:]  (lane a < nu)
            if (lane < nu) {
                const double rk = ub0 - acc[nx + lane];
                rv[lane] = rk;
                R[node * nu + lane] = rk;
            }
            if (lane + 32 < nu) {
                const double rk = ub1 - acc[nx + lane + 32];
                rv[lane + 32] = rk;
                R[node * nu + lane + 32] = rk;
            }
            __syncwarp();
            double q0 = 0.0, q1 = 0.0;
            if constexpr (kRegs) {
                if (lane < NX) q0 = acc[lane] - xb0 - dot_col<KN>(kcol, rv);
            } else {
                const double *Kc = M.K + (long long)at(cls_a, cls_b, d) * nu * nx;
                if (lane < nx) {
                    double kr;
                    if constexpr (NX > 0) kr = mv_row_t<NX, NU>(Kc, rv, lane);
                    else kr = mv_row(Kc, rv, nx, nu, lane);
                    q0 = acc[lane] - xb0 - kr;
                }
                if (lane + 32 < nx) q1 = acc[lane + 32] - xb1 - mv_row(Kc, rv, nx, nu, lane + 32);
            }
            __syncwarp();
            if (lane < nx) qj[lane] = q0;
            if (lane + 32 < nx) qj[lane + 32] = q1;
            if constexpr (kRegs) {
#pragma unroll
                for (int l = 0; l < KN; ++l) kcol[l] = kcol_next[l];
            }
        }
        __syncwarp();
        if (d == 0) {   // only the head's q is needed outside the chain
            if (lane < nx) Q[node * nx + lane] = qj[lane];
            if (lane + 32 < nx) Q[node * nx + lane + 32] = qj[lane + 32];
        }
        node = next;
        xb0 = nxb0; xb1 = nxb1; ub0 = nub0; ub1 = nub1;
    }
}

template <int NX, int NU>
__device__ __forceinline__ void chain_fwd(const Layout &L, const Topo &T, const Tabs &M, double *__restrict__ X,
                                          double *__restrict__ U, const double *__restrict__ R,
                                          const int *__restrict__ chain, int depth, int lane, double *scratch) {
    const int nx = NX > 0 ? NX : L.nx, nu = NX > 0 ? NU : L.nu, nxu = nx + nu;
    double *v = scratch, *part = scratch + nxu;
    int node_a = lane < depth ? chain[lane] : 0, node_b = lane + 32 < depth ? chain[lane + 32] : 0;
    int dyn_a = lane + 1 < depth ? T.dyn_idx[chain[lane + 1]] : 0, dyn_b = lane + 33 < depth ? T.dyn_idx[chain[lane + 33]] : 0;
    int cls_a = (lane < depth && node_a < L.m) ? T.cls[node_a] : 0, cls_b = (lane + 32 < depth && node_b < L.m) ? T.cls[node_b] : 0;
    auto at = [&](int va, int vb, int d) { return d < 32 ? __shfl_sync(0xffffffffu, va, d) : __shfl_sync(0xffffffffu, vb, d - 32); };
    int node = at(node_a, node_b, 0);
    // x of the head was written by the level above; r rows are prefetched one step ahead
    if (lane < nx) v[lane] = X[node * nx + lane];
    if (lane + 32 < nx) v[lane + 32] = X[node * nx + lane + 32];
    double r0 = 0.0, r1 = 0.0;
    if (node < L.m) {
        r0 = lane < nu ? R[node * nu + lane] : 0.0;
        r1 = lane + 32 < nu ? R[node * nu + lane + 32] : 0.0;
    }
    // register-resident tables (compile-time sizes, nu <= 16 so that the split product applies): ct = column `lane` of
    // [A ; B]' of the current dynamics row; kr = this lane's segment of [K R~^-1] of the current class (prefetched)
    constexpr bool kRegs = NX > 0 && NX + NU <= 32 && NU <= 16;
    constexpr int TN = NX > 0 ? NX + NU : 1;
    constexpr int G = kRegs ? 32 / (NU > 0 ? NU : 1) : 1, SEG = kRegs ? (NX + NU + G - 1) / G : 1;
    double ct[TN], kr[SEG], kr_next[SEG];
    int dyn_loaded = -1;
    const int sg = kRegs ? lane / (NU > 0 ? NU : 1) : 0, sa = kRegs ? lane - sg * (NU > 0 ? NU : 1) : 0;
    auto load_kr = [&](double (&dst)[SEG], int cls) {
        const double *KR = M.KRcatT + (long long)cls * (NX + NU) * NU;
#pragma unroll
        for (int i = 0; i < SEG; ++i) {
            const int l = sg * SEG + i;
            dst[i] = (sg < G && l < NX + NU) ? __ldg(KR + l * NU + sa) : 0.0;
        }
    };
    if constexpr (kRegs) {
        if (node < L.m) load_kr(kr, at(cls_a, cls_b, 0));
    }
    for (int d = 0; d + 1 < depth; ++d) {
        if (node >= L.m) break;
        const int child = at(node_a, node_b, d + 1);
        double nr0 = 0.0, nr1 = 0.0;
        if (child < L.m) {
            nr0 = lane < nu ? R[child * nu + lane] : 0.0;
            nr1 = lane + 32 < nu ? R[child * nu + lane + 32] : 0.0;
            if constexpr (kRegs) load_kr(kr_next, at(cls_a, cls_b, d + 1));
        }
        if (lane < nu) v[nx + lane] = r0;
        if (lane + 32 < nu) v[nx + lane + 32] = r1;
        __syncwarp();
        if constexpr (kRegs) {   // u = [K R~^-1] [x ; r], reduction split over G lane groups
            double p = 0.0;
#pragma unroll
            for (int i = 0; i < SEG; ++i) {
                const int l = sg * SEG + i;
                if (l < NX + NU) p = fma(kr[i], v[l], p);
            }
            part[lane] = p;
            __syncwarp();
            double ua = 0.0;
            if (lane < NU) {
#pragma unroll
                for (int gg = 0; gg < G; ++gg) ua += part[gg * NU + lane];
            }
            __syncwarp();
            if (lane < NU) {
                v[NX + lane] = ua;
                U[node * NU + lane] = ua;
            }
        } else {
            const double *KR = M.KRcatT + (long long)at(cls_a, cls_b, d) * nxu * nu;
            if (nu <= 32) {
                double ua;
                if constexpr (NX > 0) ua = mv_split_t<NU, NX + NU>(KR, v, lane, part);
                else ua = mv_split(KR, v, nu, nxu, lane, part);
                if (lane < nu) {
                    v[nx + lane] = ua;
                    U[node * nu + lane] = ua;
                }
            } else {
                const double u0 = mv_row(KR, v, nu, nxu, lane), u1 = lane + 32 < nu ? mv_row(KR, v, nu, nxu, lane + 32) : 0.0;
                __syncwarp();
                v[nx + lane] = u0;
                U[node * nu + lane] = u0;
                if (lane + 32 < nu) {
                    v[nx + lane + 32] = u1;
                    U[node * nu + lane + 32] = u1;
                }
            }
        }
        __syncwarp();
        // x_child = A x + B u, kept in shared memory for the next step and written out
        const int dyn = at(dyn_a, dyn_b, d);
        double x0 = 0.0, x1 = 0.0;
        if constexpr (kRegs) {
            if (dyn != dyn_loaded) {
                load_col<TN>(ct, M.ABcatT + dyn * ((NX + NU) * NX), NX, lane, lane < NX);
                dyn_loaded = dyn;
            }
            if (lane < NX) x0 = dot_col<TN>(ct, v);
        } else if constexpr (NX > 0) {
            const double *C = M.ABcatT + dyn * ((NX + NU) * NX);
            if (lane < NX) x0 = mv_row_t<NX, NX + NU>(C, v, lane);
        } else {
            const double *C = M.ABcatT + (long long)dyn * nxu * nx;
            if (lane < nx) x0 = mv_row(C, v, nx, nxu, lane);
            if (lane + 32 < nx) x1 = mv_row(C, v, nx, nxu, lane + 32);
        }
        __syncwarp();
        if (lane < nx) {
            v[lane] = x0;
            X[child * nx + lane] = x0;
        }
        if (lane + 32 < nx) {
            v[lane + 32] = x1;
            X[child * nx + lane + 32] = x1;
        }
        if constexpr (kRegs) {
#pragma unroll
            for (int i = 0; i < SEG; ++i) kr[i] = kr_next[i];
        }
        node = child;
        r0 = nr0;
        r1 = nr1;
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
