// lane.cu -- the node-parallel passes of one Chambolle-Pock iteration with EIGHT LANES PER TREE NODE (FP64, sm_100a).
//
// History (profiles/r1_kernel_evolution.md): one warp per node spends >85 % of its instructions on index arithmetic,
// shuffles and divergent per-lane branches (a node carries only ~130 doubles); one thread per node cuts the instruction
// count 5x but every warp request then touches 32 different rows (L1TEX wavefront bound, 3.6x L2 over-fetch) and a
// 6e4-node tree offers only 13 warps per SM, each a long serial chain of exposed memory latencies.
// Here a node is owned by an OCTET of lanes, a warp by four consecutive nodes: the eight lanes walk a row two doubles
// (16 bytes) at a time, so an octet request is one contiguous 128-byte line, a warp request four neighbouring rows
// (rows of consecutive nodes are contiguous in the node-major layout) -- fully coalesced, no reliance on L1 -- there
// are 8x more warps to hide latency, the only cross-lane traffic is three shuffles per reduction, and all lanes run
// the same code path.  Used when the cost square roots are diagonal (rb_create classifies the tables), nx and nu are
// even and no node has more than kLaneMaxChildren children; otherwise the general warp-per-node tile kernels run.
//
//   k_primal_lane : pbar = p - alpha L* d (solver.py:27-39), s_0 -= alpha (cache.py:253-257), kernel projection
//                   (cache.py:290-317)
//   k_dual_lane   : dbar = d + alpha L(2 p+ - p) (solver.py:44-58), prox of g* (cache.py:321-393), all six residual
//                   inf-norms (solver.py:63-95,137-141)
#include <algorithm>
#include <cstdlib>

#include "kernels.cuh"

namespace rb {

namespace {

constexpr int kMaxNodesPerCta = kLaneThreads / 4;   // nodes per CTA with the narrowest lane group
constexpr int kChainDualThreads = 128;

struct ResidLane {
    // Running maxima of |x| per residual norm, kept as the (signed) entry that attained them: one DSETP and two selects
    // per update instead of five integer instructions on the bit pattern (sm_100a has no 64-bit or FP64 max).  A
    // comparison with a NaN is false, so NaNs are caught separately: every dual entry feeds some primal entry's xi0
    // through L*, and primal() ORs "xi0 is NaN" into `nan`; bits() then reports a NaN maximum, like the bit-pattern
    // maxima of fused.cu do.
    double v[6];
    int nan;
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < 6; ++i) v[i] = 0.0;
        nan = 0;
    }
    __device__ __forceinline__ void put(int slot, double x) { v[slot] = fabs(x) > fabs(v[slot]) ? x : v[slot]; }
    __device__ __forceinline__ double dual(double dd, double lpp, double inv_alpha) {   // dd = d - d+
        const double xi2 = fma(dd, inv_alpha, lpp);
        put(2, xi2);
        put(5, dd);
        return xi2;
    }
    __device__ __forceinline__ void primal(double dp, double g1, double g2, double inv_alpha) {   // dp = p+ - p
        const double xi1 = -fma(dp, inv_alpha, g1);
        const double xi0 = xi1 + g2;
        put(1, xi1);
        put(0, xi0);
        put(4, dp);
        put(3, dp + g1);
        nan |= xi0 != xi0;
    }
    // bit pattern of the maximum (non-negative doubles order like their bit patterns; a NaN pattern sits above +inf)
    __device__ __forceinline__ unsigned long long bits(int slot) const {
        if (nan) return 0x7ff8000000000000ull;
        return (unsigned long long)__double_as_longlong(v[slot]) & 0x7fffffffffffffffull;
    }
};

// Rectangle._constrain (rectangle.py:50-59) without branches: v if lo <= v <= hi, else lo if v <= lo, else hi if
// v >= hi -- the two selects below give exactly that, also for lo > hi -- else (NaN) v itself with `bad` set
__device__ __forceinline__ double box_clip_select(double v, double lo, double hi, int &bad) {
    double r = v >= hi ? hi : v;
    r = v <= lo ? lo : r;
    bad |= v != v;
    return r;
}

__device__ __forceinline__ double dual_w(double d_old, double lz, double alpha, double inv_alpha) {
    return fma(lz, alpha, d_old) * inv_alpha;
}

// sum over the kOct lanes that share a node (all 32 lanes of the warp take part)
template <int kOct>
__device__ __forceinline__ double oct_sum(double v) {
#pragma unroll
    for (int o = 1; o < kOct; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// maximum over the lane groups of a warp: shuffles are warp-wide, so every group runs the warp's maximum trip count
template <int kOct>
__device__ __forceinline__ int warp_octets_max(int v) {
#pragma unroll
    for (int o = kOct; o < 32; o <<= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// rows are walked two doubles (16 bytes) at a time: nx and nu are even on this path, so every row start is 16-byte
// aligned.  The inputs are read-only for the whole kernel: the non-coherent path lets loads move above stores.
__device__ __forceinline__ double2 ld2(const double *__restrict__ p, int k2) {
    return __ldg(reinterpret_cast<const double2 *>(p + 2 * k2));
}
// L1 prefetch of the line holding *p: the later phases of a pass (tau / d5 / d6 of the children, the risk rows) then hit
// L1 instead of paying a full memory round trip each, one after the other
__device__ __forceinline__ void pf(const double *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void st2(double *__restrict__ p, int k2, double a, double b) {
    *reinterpret_cast<double2 *>(p + 2 * k2) = make_double2(a, b);
}

}  // namespace

// ====================================================================================================================
template <int kOct>
__global__ void __launch_bounds__(kLaneThreads, 3) k_primal_lane(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                             const double *__restrict__ p_old,
                                                             const double *__restrict__ d_old, double *__restrict__ p_new,
                                                             const int *__restrict__ node_list, int count) {
    if (ctrl->done) return;
    const double alpha = ctrl->alpha;
    const Layout &L = P.L;
    const Topo &T = P.t;
    const Tabs &M = P.m;
    const int g = threadIdx.x & (kOct - 1);
    const int slot = (blockIdx.x * blockDim.x + threadIdx.x) / kOct;
    const int node = slot < count ? (node_list ? node_list[slot] : slot) : L.n;   // L.n: idle octet (still shuffles)
    const double *Po = p_old + (long long)blockIdx.y * L.np_pad;
    const double *D = d_old + (long long)blockIdx.y * L.nd_pad;
    double *Pn = p_new + (long long)blockIdx.y * L.np_pad;
    const int nx = L.nx, nu = L.nu, nxu = L.nxu;
    const bool nonleaf = node < L.m;
    if (!nonleaf && node < L.n) {   // leaf: xbar = x - alpha (sqrtQf d11 + d14)   (operators.py:89-92)
        const int li = node - L.m;
        const double *sq = M.sqf_d + T.leafcost_idx[li] * nx;
        const double *d11 = D + L.d11 + (long long)li * nx, *d14 = D + L.d14 + (long long)li * nx;
        const double *xo = Po + L.px + (long long)node * nx;
        double *xn = Pn + L.px + (long long)node * nx;
        for (int k2 = g; k2 < nx / 2; k2 += kOct) {
            const double2 m2 = ld2(sq, k2), a2 = ld2(d11, k2), o2 = ld2(xo, k2);
            double acc0 = m2.x * a2.x, acc1 = m2.y * a2.y;
            if (L.has_leaf_rect) {
                const double2 b2 = ld2(d14, k2);
                acc0 += b2.x;
                acc1 += b2.y;
            }
            st2(xn, k2, o2.x - alpha * acc0, o2.y - alpha * acc1);
        }
    }
    const int c0 = nonleaf ? T.child_first[node] : 0, cc = nonleaf ? T.child_count[node] : 0;
    if (nonleaf) {   // lines of the kernel-projection phase at the end of the pass, fetched while the row phase runs
        const int yo_pf = T.yoff[node];
        if (g == 0) {
            pf(Po + L.py + yo_pf);
            pf(D + L.d1 + yo_pf);
            pf(D + L.d2 + node);
        } else if (g == 1) {
            pf(Po + L.ptau + c0);
            pf(Po + L.ps + c0);
            pf(T.cond_prob + c0);
        } else if (g == 2) {
            pf(D + L.d5 + c0 - 1);
            pf(D + L.d6 + c0 - 1);
        } else if (g == 3) {
            if (c0 < L.m) pf(D + L.d2 + c0);
            else {
                pf(D + L.d12 + c0 - L.m);
                pf(D + L.d13 + c0 - L.m);
            }
        }
    }
    if (nonleaf) {   // [xbar; ubar] = [x; u] - alpha (Gamma' d7 + sum_j sqrt(Q_j, R_j) [d3_j; d4_j])   (operators.py:74-87)
        const double *d7 = D + L.d7 + (long long)node * nxu;
        for (int part = 0; part < 2; ++part) {
            const int w = part == 0 ? nx : nu;
            const double *po_row = part == 0 ? Po + L.px + (long long)node * nx : Po + L.pu + (long long)node * nu;
            double *pn_row = part == 0 ? Pn + L.px + (long long)node * nx : Pn + L.pu + (long long)node * nu;
            const double *mtab = part == 0 ? M.sq_d : M.sr_d;
            const double *dseg = D + (part == 0 ? L.d3 : L.d4);
            for (int k2 = g; k2 < w / 2; k2 += kOct) {
                const double2 o2 = ld2(po_row, k2);
                double2 acc = L.has_nl_rect ? ld2(d7 + (part == 0 ? 0 : nx), k2) : make_double2(0.0, 0.0);
                for (int j = c0; j < c0 + cc; ++j) {
                    const double2 m2 = ld2(mtab + T.cost_idx[j] * w, k2), v2 = ld2(dseg + (long long)(j - 1) * w, k2);
                    acc.x = fma(m2.x, v2.x, acc.x);
                    acc.y = fma(m2.y, v2.y, acc.y);
                }
                st2(pn_row, k2, o2.x - alpha * acc.x, o2.y - alpha * acc.y);
            }
        }
    }
    // ybar_i, the children's taubar_j / sbar_j, and the projection onto ker [E' -I -I] (cache.py:290-317).  For AVaR
    // M = [a I, -I, 1, -I, -I], M M' = (a^2+3) I + 1 1', so proj = v - M'(M M')^-1 M v in closed form.  Children are
    // spread over the lanes of the octet.
    const double d2v = nonleaf ? D[L.d2 + node] : 0.0;
    const int yo = nonleaf ? T.yoff[node] : 0;
    const double a = nonleaf ? T.risk_alpha[node] : 0.0;
    const double *yold = Po + L.py + yo, *d1 = D + L.d1 + yo;
    double *ynew = Pn + L.py + yo;
    const double ylast_bar = nonleaf ? yold[2 * cc] - alpha * (d1[2 * cc] - d2v) : 0.0;
    const double den = a * a + 3.0;
    const int max_rounds = warp_octets_max<kOct>((cc + kOct - 1) / kOct);
    double rsum = 0.0;
    for (int rd = 0; rd < max_rounds; ++rd) {
        const int e = rd * kOct + g;
        double res = 0.0;
        if (e < cc) {
            const int j = c0 + e;
            const double ya = yold[e] - alpha * (d1[e] - T.cond_prob[j] * d2v);
            const double yb = yold[cc + e] - alpha * d1[cc + e];
            const double tj = Po[L.ptau + j] - alpha * (0.5 * (D[L.d5 + j - 1] + D[L.d6 + j - 1]));
            const double lts = j < L.m ? D[L.d2 + j] : 0.5 * (D[L.d12 + j - L.m] + D[L.d13 + j - L.m]);
            const double sj = Po[L.ps + j] - alpha * lts;
            res = a * ya - yb + ylast_bar - tj - sj;
        }
        rsum += oct_sum<kOct>(res);
    }
    const double shift = rsum / (den + (double)cc);
    double wsum = 0.0;
    for (int rd = 0; rd < max_rounds; ++rd) {   // the same arithmetic again, bit for bit
        const int e = rd * kOct + g;
        double w = 0.0;
        if (e < cc) {
            const int j = c0 + e;
            const double ya = yold[e] - alpha * (d1[e] - T.cond_prob[j] * d2v);
            const double yb = yold[cc + e] - alpha * d1[cc + e];
            const double tj = Po[L.ptau + j] - alpha * (0.5 * (D[L.d5 + j - 1] + D[L.d6 + j - 1]));
            const double lts = j < L.m ? D[L.d2 + j] : 0.5 * (D[L.d12 + j - L.m] + D[L.d13 + j - L.m]);
            const double sj = Po[L.ps + j] - alpha * lts;
            w = ((a * ya - yb + ylast_bar - tj - sj) - shift) / den;
            ynew[e] = ya - a * w;
            ynew[cc + e] = yb + w;
            Pn[L.ptau + j] = tj + w;
            Pn[L.ps + j] = sj + w;
        }
        wsum += oct_sum<kOct>(w);
    }
    if (nonleaf && g == 0) {
        ynew[2 * cc] = ylast_bar - wsum;
        if (node == 0) {
            Pn[L.ps] = (Po[L.ps] - alpha * d2v) - alpha;   // s_0: half step, then prox of alpha * identity
            Pn[L.ptau] = Po[L.ptau] - alpha * Po[L.ptau];  // tau_0 (always 0; same arithmetic as the reference)
        }
    }
}

// ====================================================================================================================
// The stopping test without a launch of its own (batch 1): called by ALL threads of a CTA at the end of a dual-pass kernel.  The
// CTA's atomics on `slots` / the status word are ordered before its count; the CTA that completes Ctrl::arr_expected -- the last
// one of the last dual-pass kernel of the iteration, whichever stream that is on -- runs k_check's body (fused.cu; solver.py:137-161)
// with one warp.  Saves the k_check launch and its dependency edge at the end of every iteration -- and costs every CTA of the
// dual passes one atomic round trip of residency at its tail (the count needs its return value, the maxima are fire-and-forget
// reductions): measured slower on cfg3 (9 138 vs 9 525 it/s), so it is an ablation behind rb_use_fused_check(1).
__device__ __noinline__ void iteration_arrive(Ctrl *ctrl, double *slots) {   // (out of line: the callers are register-bound)
    __shared__ int is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(&ctrl->arr_count, 1) == ctrl->arr_expected - 1;
    }
    __syncthreads();
    if (!is_last || threadIdx.x >= 32) return;
    __threadfence();
    const int lane = threadIdx.x;
    const Ctrl c = *ctrl;   // (iters, tol, max_iters, hist, last: not written by any kernel of this iteration)
    const double mine = lane < 6 ? __ldcg(slots + lane) : 0.0;
    const bool nan = mine != mine, bad = lane < 3 && !(mine <= c.tol);
    if (lane < 6) {
        if (c.hist && c.iters < c.hist_capacity) c.hist[(long long)c.iters * 6 + lane] = mine;
        c.last[lane] = mine;
        if (c.host_last && c.mirror) c.host_last[lane] = mine;
        slots[lane] = 0.0;
    }
    const bool all_ok = !__any_sync(0xffffffffu, bad), any_nan = __any_sync(0xffffffffu, nan);
    if (lane == 0) {
        if (any_nan) atomicOr(&ctrl->status, 2);   // a NaN maximum: some iterate entry is not finite
        ctrl->iters = c.iters + 1;
        ctrl->pending = 0;
        ctrl->arr_count = 0;
        if (c.iters >= c.max_iters || all_ok) ctrl->done = 1;
    }
}

template <int kOct, int MINB, int BT = kLaneThreads>
__global__ void __launch_bounds__(BT, MINB) k_dual_lane(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl,
                                                           const double *__restrict__ p_old, const double *__restrict__ p_new,
                                                           const double *__restrict__ d_old, double *__restrict__ d_new,
                                                           double *__restrict__ slots, const int *__restrict__ node_list,
                                                           int first, int count, double *pbar, int arrive) {
    // pbar (may be null): the buffer of p_old; when given, the pass also leaves there pbar = p+ - alpha L* d+, the
    // half step of the NEXT iteration (solver.py:27-39), from the d+ it has in registers.  Every entry of p_old is
    // read and overwritten by the same lane (s_i: all lanes of the group read it, so it is written after the block
    // barrier), and it is never read again afterwards.
    if (ctrl->done) return;
    const double alpha = ctrl->alpha, inv_alpha = 1.0 / alpha;
    const Layout &L = P.L;
    const Topo &T = P.t;
    const Tabs &M = P.m;
    // per-node SOC results of the children, [child][node of the CTA]: scale (projection = scale * w on all but the last
    // entry), the projected last entry, and the child's cost-table row
    __shared__ double soc_scale[kLaneMaxChildren][kMaxNodesPerCta];
    __shared__ double soc_last[kLaneMaxChildren][kMaxNodesPerCta];
    __shared__ int child_cost[kLaneMaxChildren][kMaxNodesPerCta];
    static_assert(BT / kOct <= kMaxNodesPerCta, "per-node shared arrays are sized for kMaxNodesPerCta nodes");
    __shared__ unsigned long long blockmax[BT / 32][6];
    __shared__ int blockflags;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = tid & (kOct - 1), ns = tid / kOct;
    if (tid == 0) blockflags = 0;
    const int slot = (blockIdx.x * blockDim.x + tid) / kOct;
    const int node = slot < count ? (node_list ? node_list[slot] : first + slot) : L.n;   // L.n = idle octet
    const double *Po = p_old + (long long)blockIdx.y * L.np_pad;
    const double *Pn = p_new + (long long)blockIdx.y * L.np_pad;
    const double *Do = d_old + (long long)blockIdx.y * L.nd_pad;
    double *Dn = d_new + (long long)blockIdx.y * L.nd_pad;
    double *Pb = pbar ? pbar + (long long)blockIdx.y * L.np_pad : nullptr;
    double sbar = 0.0;          // s-bar of this node, stored by lane 0 of the group after the block barrier
    bool sbar_set = false;
    const int nx = L.nx, nu = L.nu, nxu = L.nxu;
    ResidLane R;
    R.init();
    int bad = 0;
    const bool nonleaf = node < L.m, leaf = node >= L.m && node < L.n;
    const int c0 = nonleaf ? T.child_first[node] : 0, cc = nonleaf ? T.child_count[node] : 0;
    const int max_cc = warp_octets_max<kOct>(cc);
    const double *xo = Po + L.px + (long long)(node < L.n ? node : 0) * nx, *xn = Pn + L.px + (long long)(node < L.n ? node : 0) * nx;
    const double *uo = Po + L.pu + (long long)(nonleaf ? node : 0) * nu, *un = Pn + L.pu + (long long)(nonleaf ? node : 0) * nu;
    if (nonleaf) {   // one lane of the group per line family; neighbouring nodes share most of these lines
        const int yo_pf = T.yoff[node];
        if (g == 0) {
            pf(Po + L.py + yo_pf);
            pf(Pn + L.py + yo_pf);
            pf(Do + L.d1 + yo_pf);
        } else if (g == 1) {
            pf(Po + L.ps + node);
            pf(Pn + L.ps + node);
            pf(Do + L.d2 + node);
        } else if (g == 2) {
            pf(Po + L.ptau + c0);
            pf(Pn + L.ptau + c0);
            pf(T.cond_prob + c0);
        } else if (g == 3) {
            pf(Do + L.d5 + c0 - 1);
            pf(Do + L.d6 + c0 - 1);
        }
        // every row the two phases walk (the node's x / u / d7 rows, the contiguous d3 / d4 rows of its children), one
        // line per lane and round: the phases below then meet L1 hits (or merge into the miss in flight) instead of
        // one full memory round trip per loop iteration -- what bounds the pass when only a few nodes are in flight
        for (int off = g * 16; off < nx; off += kOct * 16) {
            pf(xo + off);
            pf(xn + off);
        }
        for (int off = g * 16; off < nu; off += kOct * 16) {
            pf(uo + off);
            pf(un + off);
        }
        if (L.has_nl_rect)
            for (int off = g * 16; off < nxu; off += kOct * 16) pf(Do + L.d7 + (long long)node * nxu + off);
        for (int off = g * 16; off < cc * nx; off += kOct * 16) pf(Do + L.d3 + (long long)(c0 - 1) * nx + off);
        for (int off = g * 16; off < cc * nu; off += kOct * 16) pf(Do + L.d4 + (long long)(c0 - 1) * nu + off);
    } else if (leaf) {
        const long long lo_ = (long long)(node - L.m) * nx;
        for (int off = g * 16; off < nx; off += kOct * 16) {
            pf(xo + off);
            pf(xn + off);
            pf(Do + L.d11 + lo_ + off);
            if (L.has_leaf_rect) pf(Do + L.d14 + lo_ + off);
        }
        if (g == 0) {
            pf(Po + L.ps + node);
            pf(Pn + L.ps + node);
            pf(Do + L.d12 + node - L.m);
            pf(Do + L.d13 + node - L.m);
        }
    }

    // ---- phase 1: classify the second-order-cone block of every child edge (cones.py:113-132) -------------------------
    for (int jj = 0; jj < max_cc; ++jj) {
        const bool act = jj < cc;
        const int j = c0 + (act ? jj : 0);
        const long long e0 = j - 1;
        double ss = 0.0, w6 = 0.0;
        if (act) {
            const int ci = T.cost_idx[j];
            const double *sq = M.sq_d + ci * nx, *sr = M.sr_d + ci * nu;
            const double *d3 = Do + L.d3 + e0 * nx, *d4 = Do + L.d4 + e0 * nu;
            for (int k2 = g; k2 < nx / 2; k2 += kOct) {
                const double2 o2 = ld2(xo, k2), n2 = ld2(xn, k2), d2 = ld2(d3, k2), m2 = ld2(sq, k2);
                const double w0 = dual_w(d2.x, m2.x * (2 * n2.x - o2.x), alpha, inv_alpha);
                const double w1 = dual_w(d2.y, m2.y * (2 * n2.y - o2.y), alpha, inv_alpha);
                ss = fma(w0, w0, ss);
                ss = fma(w1, w1, ss);
            }
            for (int k2 = g; k2 < nu / 2; k2 += kOct) {
                const double2 o2 = ld2(uo, k2), n2 = ld2(un, k2), d2 = ld2(d4, k2), m2 = ld2(sr, k2);
                const double w0 = dual_w(d2.x, m2.x * (2 * n2.x - o2.x), alpha, inv_alpha);
                const double w1 = dual_w(d2.y, m2.y * (2 * n2.y - o2.y), alpha, inv_alpha);
                ss = fma(w0, w0, ss);
                ss = fma(w1, w1, ss);
            }
            if (g == 0) {
                const double to = Po[L.ptau + j], tn = Pn[L.ptau + j];
                const double ht = 0.5 * (2 * tn - to);
                const double w5 = dual_w(Do[L.d5 + e0], ht, alpha, inv_alpha) - 0.5;
                w6 = dual_w(Do[L.d6 + e0], ht, alpha, inv_alpha) + 0.5;
                ss = fma(w5, w5, ss);
                child_cost[jj][ns] = ci;
            }
        }
        ss = oct_sum<kOct>(ss);
        if (act && g == 0) {
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
            soc_scale[jj][ns] = scale;
            soc_last[jj][ns] = last;
        }
    }
    __syncwarp();
    if (nonleaf) {
        // ---- phase 2: the [x; u] rows: d3/d4 of every child, d7, and the x / u residual rows ----------------------------
        const long long ri = L.has_nl_rect ? (long long)T.nl_rect_idx[node] * nxu : 0;
        const double *d7o = Do + L.d7 + (long long)node * nxu;
        double *d7n = Dn + L.d7 + (long long)node * nxu;
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
            const double *seg_o = Do + (part == 0 ? L.d3 : L.d4);
            double *seg_n = Dn + (part == 0 ? L.d3 : L.d4);
            for (int k2 = g; k2 < w / 2; k2 += kOct) {
                const double2 o2 = ld2(po_row, k2), n2 = ld2(pn_row, k2);
                double2 d7v = make_double2(0.0, 0.0), lo2 = d7v, hi2 = d7v;
                if (L.has_nl_rect) {
                    d7v = ld2(d7o + off7, k2);
                    lo2 = ld2(M.nl_lo + ri + off7, k2);
                    hi2 = ld2(M.nl_hi + ri + off7, k2);
                }
                const double z0 = 2 * n2.x - o2.x, z1 = 2 * n2.y - o2.y, dl0 = n2.x - o2.x, dl1 = n2.y - o2.y;
                double g10 = 0.0, g20 = 0.0, g11 = 0.0, g21 = 0.0, lt0 = 0.0, lt1 = 0.0;
                for (int jj = 0; jj < cc; ++jj) {
                    const int j = c0 + jj;
                    const double2 m2 = ld2(mtab + child_cost[jj][ns] * w, k2);
                    const double2 dol2 = ld2(seg_o + (long long)(j - 1) * w, k2);
                    const double scale = soc_scale[jj][ns];
                    double dn0, dn1;
                    edge(m2.x, z0, dl0, dol2.x, scale, dn0, g10, g20);
                    edge(m2.y, z1, dl1, dol2.y, scale, dn1, g11, g21);
                    st2(seg_n + (long long)(j - 1) * w, k2, dn0, dn1);
                    lt0 = fma(m2.x, dn0, lt0);
                    lt1 = fma(m2.y, dn1, lt1);
                }
                double dn70 = 0.0, dn71 = 0.0;
                entry(o2.x, n2.x, d7v.x, lo2.x, hi2.x, g10, g20, dn70);
                entry(o2.y, n2.y, d7v.y, lo2.y, hi2.y, g11, g21, dn71);
                if (L.has_nl_rect) st2(d7n + off7, k2, dn70, dn71);
                if (Pb) {   // [xbar; ubar] of the next iteration (operators.py:74-87); dn7 is 0 without rectangles
                    double *pb_row = Pb + (part == 0 ? L.px + (long long)node * nx : L.pu + (long long)node * nu);
                    st2(pb_row, k2, n2.x - alpha * (lt0 + dn70), n2.y - alpha * (lt1 + dn71));
                }
            }
        }
        // ---- d5, d6 and the tau_j residual rows: children spread over the lanes of the octet ----------------------------
        for (int jj = g; jj < cc; jj += kOct) {
            const int j = c0 + jj;
            const long long e0 = j - 1;
            const double to = Po[L.ptau + j], tn = Pn[L.ptau + j];
            const double ht = 0.5 * (2 * tn - to), hdt = 0.5 * (tn - to);
            const double do5 = Do[L.d5 + e0], do6 = Do[L.d6 + e0];
            const double w5 = dual_w(do5, ht, alpha, inv_alpha) - 0.5;
            const double w6 = dual_w(do6, ht, alpha, inv_alpha) + 0.5;
            const double dn5 = alpha * (w5 - soc_scale[jj][ns] * w5);
            const double dn6 = alpha * (w6 - soc_last[jj][ns]);
            Dn[L.d5 + e0] = dn5;
            Dn[L.d6 + e0] = dn6;
            if (Pb) Pb[L.ptau + j] = tn - alpha * (0.5 * (dn5 + dn6));   // taubar_j (operators.py:88)
            const double dd5 = do5 - dn5, dd6 = do6 - dn6;
            const double x5 = R.dual(dd5, hdt, inv_alpha), x6 = R.dual(dd6, hdt, inv_alpha);
            R.primal(tn - to, 0.5 * (dd5 + dd6), 0.5 * (x5 + x6), inv_alpha);
        }
    }
    // ---- d1, d2 (risk blocks) and the y_i, s_i residual rows: the 2c+1 entries spread over the octet ---------------------
    {
        const int yo = nonleaf ? T.yoff[node] : 0, ny = nonleaf ? 2 * cc + 1 : 0;
        const double *yold = Po + L.py + yo, *ynew = Pn + L.py + yo, *d1o = Do + L.d1 + yo;
        double *d1n = Dn + L.d1 + yo;
        const int yrounds = (2 * max_cc + 1 + kOct - 1) / kOct;
        double dot_z = 0.0, dot_d = 0.0;
        for (int rd = 0; rd < yrounds; ++rd) {
            const int e = rd * kOct + g;
            double pz = 0.0, pd = 0.0;
            if (e < ny) {
                const double b = e < cc ? T.cond_prob[c0 + e] : (e == 2 * cc ? 1.0 : 0.0);
                pz = b * (2 * ynew[e] - yold[e]);
                pd = b * (ynew[e] - yold[e]);
            }
            dot_z += oct_sum<kOct>(pz);
            dot_d += oct_sum<kOct>(pd);
        }
        if (nonleaf) {
            const double so = Po[L.ps + node], sn = Pn[L.ps + node];
            const double do2 = Do[L.d2 + node];
            const double w2 = dual_w(do2, (2 * sn - so) - dot_z, alpha, inv_alpha);
            const double dn2 = alpha * (w2 - fmax(0.0, w2));
            const double dd2 = do2 - dn2;
            const double xi22 = fma(dd2, inv_alpha, (sn - so) - dot_d);
            if (g == 0) {
                Dn[L.d2 + node] = dn2;
                R.put(2, xi22);
                R.put(5, dd2);
                R.primal(sn - so, dd2, xi22, inv_alpha);   // s_i of a nonleaf node: its L* row is d2_i
            }
            sbar = sn - alpha * dn2;                 // operators.py:73; the root also takes the prox of alpha * identity
            if (node == 0) sbar -= alpha;            // (cache.py:253-257)
            sbar_set = true;
            for (int e = g; e < ny; e += kOct) {
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
                if (Pb) Pb[L.py + yo + e] = ynew[e] - alpha * (dnew - b * dn2);   // ybar_i (operators.py:72)
            }
            if (Pb && node == 0 && g == 0) Pb[L.ptau] = Pn[L.ptau] - alpha * Pn[L.ptau];   // tau_0 (always 0)
        }
    }
    // ---- leaf: SOC on [d11; d12; d13] (cache.py:375-386), rectangle on d14, x_i and s_i residual rows ---------------------
    if (__any_sync(0xffffffffu, leaf)) {
        const int li = leaf ? node - L.m : 0;
        const double *sq = M.sqf_d + (leaf ? T.leafcost_idx[li] : 0) * nx;
        const double *d11o = Do + L.d11 + (long long)li * nx, *d14o = Do + L.d14 + (long long)li * nx;
        double *d11n = Dn + L.d11 + (long long)li * nx, *d14n = Dn + L.d14 + (long long)li * nx;
        double so = 0.0, sn = 0.0, hs = 0.0, hds = 0.0, do12 = 0.0, do13 = 0.0, w12 = 0.0, w13 = 0.0, ss = 0.0;
        if (leaf) {
            so = Po[L.ps + node];
            sn = Pn[L.ps + node];
            hs = 0.5 * (2 * sn - so);
            hds = 0.5 * (sn - so);
            do12 = Do[L.d12 + li];
            do13 = Do[L.d13 + li];
            for (int k2 = g; k2 < nx / 2; k2 += kOct) {
                const double2 o2 = ld2(xo, k2), n2 = ld2(xn, k2), d2 = ld2(d11o, k2), m2 = ld2(sq, k2);
                const double w0 = dual_w(d2.x, m2.x * (2 * n2.x - o2.x), alpha, inv_alpha);
                const double w1 = dual_w(d2.y, m2.y * (2 * n2.y - o2.y), alpha, inv_alpha);
                ss = fma(w0, w0, ss);
                ss = fma(w1, w1, ss);
            }
            w12 = dual_w(do12, hs, alpha, inv_alpha) - 0.5;
            w13 = dual_w(do13, hs, alpha, inv_alpha) + 0.5;
            if (g == 0) ss = fma(w12, w12, ss);
        }
        ss = oct_sum<kOct>(ss);
        if (leaf) {
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
                                  double &dnew, double &dn14, double &xb) {
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
                xb = nw - alpha * (mm * dnew + dn14);   // xbar of the leaf (operators.py:89-92); dn14 is 0 without rectangles
            };
            for (int k2 = g; k2 < nx / 2; k2 += kOct) {
                const double2 o2 = ld2(xo, k2), n2 = ld2(xn, k2), m2 = ld2(sq, k2), dol2 = ld2(d11o, k2);
                double2 d14v = make_double2(0.0, 0.0), lo2 = d14v, hi2 = d14v;
                if (L.has_leaf_rect) {
                    d14v = ld2(d14o, k2);
                    lo2 = ld2(M.leaf_lo + ri, k2);
                    hi2 = ld2(M.leaf_hi + ri, k2);
                }
                double dn0, dn1, q0 = 0.0, q1 = 0.0, xb0, xb1;
                leaf_entry(o2.x, n2.x, m2.x, dol2.x, d14v.x, lo2.x, hi2.x, dn0, q0, xb0);
                leaf_entry(o2.y, n2.y, m2.y, dol2.y, d14v.y, lo2.y, hi2.y, dn1, q1, xb1);
                st2(d11n, k2, dn0, dn1);
                if (L.has_leaf_rect) st2(d14n, k2, q0, q1);
                if (Pb) st2(Pb + L.px + (long long)node * nx, k2, xb0, xb1);
            }
            {
                const double dn12 = alpha * (w12 - scale * w12), dn13 = alpha * (w13 - last);
                sbar = sn - alpha * (0.5 * (dn12 + dn13));   // operators.py:93
                sbar_set = true;
            }
            if (g == 0) {
                const double dn12 = alpha * (w12 - scale * w12), dn13 = alpha * (w13 - last);
                Dn[L.d12 + li] = dn12;
                Dn[L.d13 + li] = dn13;
                const double dd12 = do12 - dn12, dd13 = do13 - dn13;
                const double xa = R.dual(dd12, hds, inv_alpha), xb = R.dual(dd13, hds, inv_alpha);
                R.primal(sn - so, 0.5 * (dd12 + dd13), 0.5 * (xa + xb), inv_alpha);
            }
        }
    }
    // block-level reduction of the six maxima (as bit patterns), one atomic per slot per block
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        unsigned long long mval = R.bits(i);
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
        for (int wv = 1; wv < BT / 32; ++wv) mval = blockmax[wv][tid] > mval ? blockmax[wv][tid] : mval;
        atomicMax(reinterpret_cast<unsigned long long *>(slots + (long long)blockIdx.y * 6 + tid), mval);
    }
    if (tid == 0 && blockflags) atomicOr(&ctrl->status, blockflags);
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) ctrl->pending = 1;
    if (Pb && sbar_set && g == 0) Pb[L.ps + node] = sbar;
    if (arrive) iteration_arrive(ctrl, slots);
}

// ====================================================================================================================
// Dual pass of the CHAIN part of the tree: nonleaf nodes with exactly one child (every stage past the stopping time,
// 91 % of the nodes of cfg3).  Same arithmetic as k_dual_lane, specialised so that the pass is bound by memory rather
// than by dependent loads and index arithmetic:
//   * templated on <NX, NU, G lanes per node>: [x; u] is walked as ONE row of (NX+NU)/2 pairs, pair k = g + G r in
//     round r (15 pairs over 4 lanes = 4 rounds for cfg3 instead of 3 + 2), all trip counts are compile-time;
//   * every HBM row of the node (x, x+, u, u+, d3/d4 of the edge, d7) is requested at the top of the kernel, before
//     anything is consumed (16 independent 16-byte loads per lane in flight), and is used from registers by both the
//     cone classification and the update -- no second walk over the rows, no shared memory, no block barrier before
//     the final reduction;
//   * one child: the SOC scale is a per-group register, the child->parent sums of L* have one term.
// Writes d+ of the node and of the edge to its child, and pbar = p+ - alpha L* d+ of the next iteration (see k_dual_lane).
// ====================================================================================================================
template <int NX, int NU, int G, int MINB>
__global__ void __launch_bounds__(kChainDualThreads, MINB)
    k_dual_chain(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl, const double *__restrict__ p_old,
                 const double *__restrict__ p_new, const double *__restrict__ d_old, double *__restrict__ d_new,
                 double *__restrict__ slots, const int4 *__restrict__ recs, int first, int count, int stride, int yo0,
                 double *pbar, int with_risk, OwnMap own) {
    // with_risk bit 0 clear: the risk block (d1, d2, ybar, sbar) of these nodes has been done by k_dual_risk_chain; bit 1: the CTAs
    // count themselves into Ctrl::arr_count (iteration_arrive) -- a bit of an existing argument, the kernel is register-bound
    const int arrive = with_risk & 2;
    with_risk &= 1;
    constexpr int HX = NX / 2, K = (NX + NU) / 2, R = (K + G - 1) / G, NXU = NX + NU;
    const Layout &L = P.L;
    const Topo &T = P.t;
    const Tabs &M = P.m;
    __shared__ unsigned long long blockmax[kChainDualThreads / 32][6];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = tid & (G - 1);
    const int slot = (blockIdx.x * blockDim.x + tid) / G;
    const bool active = slot < count;
    const int node = first + own.at(active ? slot : 0);   // idle groups of the last CTA repeat the first node: stores and maxima masked
    // ONE global round trip before the arithmetic starts: the control block, the node's packed topology record (child,
    // cost-table row of the child, -, rectangle row) and the rows themselves are all requested before `done` is tested.
    // In a breadth-first numbering every stage of the chain part has the same width, so the child of node i is
    // i + stride (rb_create checks it; stride <= 0: take it from the record) and y_i starts at yo0 + 3 (i - first).
    const int done = ctrl->done;
    const double alpha = ctrl->alpha;
    const int4 rec = __ldg(recs + (node - first));
    const double *Po = p_old + (long long)blockIdx.y * L.np_pad;
    const double *Pn = p_new + (long long)blockIdx.y * L.np_pad;
    const double *Do = d_old + (long long)blockIdx.y * L.nd_pad;
    double *Dn = d_new + (long long)blockIdx.y * L.nd_pad;
    double *Pb = pbar + (long long)blockIdx.y * L.np_pad;
    const int j = stride > 0 ? node + stride : rec.x;
    const long long e0 = j - 1;
    const int yo = yo0 + 3 * (node - first);
    const bool rect = L.has_nl_rect;

    // ---- all row loads of the node, issued back to back ----------------------------------------------------------------
    double2 o[R], n[R], de[R], mm[R], d7v[R];
    // offsets of pair k of the [x; u] row in the primal buffers / in the d3 | d4 segments
    const long long px_row = L.px + (long long)node * NX, pu_row = L.pu + (long long)node * NU - 2 * HX;
    const long long d3_row = L.d3 + e0 * NX, d4_row = L.d4 + e0 * NU - 2 * HX;
    auto off_p = [&](int kk) { return (kk < HX ? px_row : pu_row) + 2 * kk; };
    auto off_e = [&](int kk) { return (kk < HX ? d3_row : d4_row) + 2 * kk; };
    bool valid[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int k = g + G * r;
        valid[r] = k < K;
        const int kk = valid[r] ? k : 0;
        const bool isx = kk < HX;
        o[r] = __ldg(reinterpret_cast<const double2 *>(Po + off_p(kk)));
        n[r] = __ldg(reinterpret_cast<const double2 *>(Pn + off_p(kk)));
        de[r] = __ldg(reinterpret_cast<const double2 *>(Do + off_e(kk)));
        if (MINB <= 3)
            d7v[r] = rect ? __ldg(reinterpret_cast<const double2 *>(Do + L.d7 + (long long)node * NXU + 2 * kk))
                          : make_double2(0.0, 0.0);
        else if (rect) pf(Do + L.d7 + (long long)node * NXU + 2 * kk);   // 128-register build: d7 is fetched into L1 now and
    }                                                                    // read in the update phase
    // scalars of the node and of the edge (same addresses for the lanes of a group: one request)
    const double to = Po[L.ptau + j], tn = Pn[L.ptau + j], do5 = Do[L.d5 + e0], do6 = Do[L.d6 + e0];
    double so = 0.0, sn = 0.0, do2 = 0.0, yold = 0.0, ynew = 0.0, do1 = 0.0, prob = 0.0;
    if (with_risk) {
        so = Po[L.ps + node], sn = Pn[L.ps + node], do2 = Do[L.d2 + node];
        const int ey = g < 2 ? g : 2;   // y_i = [y_a; y_b; y_last]: lanes 0, 1, 2 take one entry each
        yold = Po[L.py + yo + ey], ynew = Pn[L.py + yo + ey], do1 = Do[L.d1 + yo + ey];
        prob = T.cond_prob[j];
    }
    if (done) return;
    const double inv_alpha = 1.0 / alpha;
    const int ci = rec.y;
    const long long ri = (long long)rec.w * NXU;
#pragma unroll
    for (int r = 0; r < R; ++r) {   // cost-table rows of the edge (a handful of rows per tree: L1 hits)
        const int kk = valid[r] ? g + G * r : 0;
        mm[r] = __ldg(reinterpret_cast<const double2 *>(kk < HX ? M.sq_d + ci * NX + 2 * kk : M.sr_d + ci * NU + 2 * (kk - HX)));
    }

    ResidLane Rs;
    Rs.init();
    int bad = 0;
    // ---- d1, d2 (risk block) and the y_i, s_i residual rows: first, while the rows are still in flight -- it only needs
    //      the scalars, and they are dead (registers free) before the row phases start -----------------------------------------
    if (with_risk) {
        const double b = g == 0 ? prob : (g == 2 ? 1.0 : 0.0);   // b_i = [pi; 0; 1] (risks.py:34-35); lanes >= 3 idle
        const bool own = g < 3;
        const double dot_z = oct_sum<G>(own ? b * (2 * ynew - yold) : 0.0);
        const double dot_d = oct_sum<G>(own ? b * (ynew - yold) : 0.0);
        const double w2 = dual_w(do2, (2 * sn - so) - dot_z, alpha, inv_alpha);
        const double dn2 = alpha * (w2 - fmax(0.0, w2));
        const double dd2 = do2 - dn2;
        const double xi22 = fma(dd2, inv_alpha, (sn - so) - dot_d);
        Rs.put(2, xi22);
        Rs.put(5, dd2);
        Rs.primal(sn - so, dd2, xi22, inv_alpha);
        if (own) {
            const double dy = ynew - yold;
            const double wv = dual_w(do1, 2 * ynew - yold, alpha, inv_alpha);
            const double zv = g < 2 ? fmax(0.0, wv) : wv;   // dual of R_+^{2c} x {0} (risks.py:32-33)
            const double dnew = alpha * (wv - zv);
            const double dd = do1 - dnew;
            const double xi2 = Rs.dual(dd, dy, inv_alpha);
            Rs.primal(dy, dd - b * dd2, xi2 - b * xi22, inv_alpha);
            if (active) {
                Dn[L.d1 + yo + g] = dnew;
                Pb[L.py + yo + g] = ynew - alpha * (dnew - b * dn2);
            }
        }
        __syncwarp();   // every lane of the group has consumed s_i before lane 0 replaces it
        if (g == 0 && active) {
            Dn[L.d2 + node] = dn2;
            Pb[L.ps + node] = sn - alpha * dn2;
        }
    }
    // ---- second-order cone of the edge: [d3; d4; d5; d6] (cones.py:113-132) ---------------------------------------------
    double ss = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const double wx = dual_w(de[r].x, mm[r].x * (2 * n[r].x - o[r].x), alpha, inv_alpha);
        const double wy = dual_w(de[r].y, mm[r].y * (2 * n[r].y - o[r].y), alpha, inv_alpha);
        if (valid[r]) {
            ss = fma(wx, wx, ss);
            ss = fma(wy, wy, ss);
        }
    }
    const double ht = 0.5 * (2 * tn - to), hdt = 0.5 * (tn - to);
    const double w5 = dual_w(do5, ht, alpha, inv_alpha) - 0.5;
    const double w6 = dual_w(do6, ht, alpha, inv_alpha) + 0.5;
    if (g == 0) ss = fma(w5, w5, ss);
    ss = oct_sum<G>(ss);
    double scale, last;
    {
        const double rr = sqrt(ss);
        if (rr <= w6) {          // inside the cone: projection = w
            scale = 1.0;
            last = w6;
        } else if (rr <= -w6) {  // inside the polar cone: projection = 0
            scale = 0.0;
            last = 0.0;
        } else {
            last = (rr + w6) / 2;
            scale = last / rr;
        }
    }
    // ---- the [x; u] row: d3 / d4 of the edge, d7 of the node, residual rows, [xbar; ubar] ---------------------------------
    // `dep` is 0 (scale >= 0) but only known once the cone is classified: the loads of this phase carry it in their
    // index, so that the compiler cannot hoist them above the reduction and keep their registers busy across it
    const int dep = MINB > 3 ? (__double2hiint(scale) >> 31) : 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        double2 lo2 = make_double2(0.0, 0.0), hi2 = lo2;
        if (rect) {
            const int kk = (valid[r] ? g + G * r : 0) + dep;
            lo2 = __ldg(reinterpret_cast<const double2 *>(M.nl_lo + ri + 2 * kk));
            hi2 = __ldg(reinterpret_cast<const double2 *>(M.nl_hi + ri + 2 * kk));
            if (MINB > 3) d7v[r] = __ldg(reinterpret_cast<const double2 *>(Do + L.d7 + (long long)node * NXU + 2 * kk));
        } else if (MINB > 3) {
            d7v[r] = make_double2(0.0, 0.0);
        }
        if (MINB > 3) {   // cost-table row again (L1) instead of 16 registers across the cone phase
            const int kk = (valid[r] ? g + G * r : 0) + dep;
            mm[r] = __ldg(reinterpret_cast<const double2 *>(kk < HX ? M.sq_d + ci * NX + 2 * kk : M.sr_d + ci * NU + 2 * (kk - HX)));
        }
        auto elem = [&](double ov, double nv, double dol, double mv, double d7old, double lo_b, double hi_b, double &dnew,
                        double &d7new, double &pb) {
            const double z = 2 * nv - ov, dlt = nv - ov;
            const double wv = dual_w(dol, mv * z, alpha, inv_alpha);   // as in the cone phase, bit for bit (not kept: registers)
            dnew = alpha * (wv - scale * wv);
            const double dd = dol - dnew;
            const double xi2 = Rs.dual(dd, mv * dlt, inv_alpha);
            double g1 = mv * dd, g2 = mv * xi2;
            d7new = 0.0;
            if (rect) {   // rectangle on [x; u] (cache.py:367-371)
                const double w7 = dual_w(d7old, z, alpha, inv_alpha);
                d7new = alpha * (w7 - box_clip_select(w7, lo_b, hi_b, bad));
                const double dd7 = d7old - d7new;
                g1 += dd7;
                g2 += Rs.dual(dd7, dlt, inv_alpha);
            }
            Rs.primal(dlt, g1, g2, inv_alpha);
            pb = nv - alpha * (mv * dnew + d7new);   // operators.py:74-87 applied to d+
        };
        double dn0, dn1, s0, s1, b0, b1;
        elem(o[r].x, n[r].x, de[r].x, mm[r].x, d7v[r].x, lo2.x, hi2.x, dn0, s0, b0);
        elem(o[r].y, n[r].y, de[r].y, mm[r].y, d7v[r].y, lo2.y, hi2.y, dn1, s1, b1);
        if (valid[r] && active) {
            const int kk = g + G * r;
            *reinterpret_cast<double2 *>(Dn + off_e(kk)) = make_double2(dn0, dn1);
            if (rect) *reinterpret_cast<double2 *>(Dn + L.d7 + (long long)node * NXU + 2 * kk) = make_double2(s0, s1);
            *reinterpret_cast<double2 *>(Pb + off_p(kk)) = make_double2(b0, b1);
        }
    }
    // ---- d5, d6 and the tau_j residual row (every lane computes, lane 0 stores) ---------------------------------------------
    {
        const double dn5 = alpha * (w5 - scale * w5), dn6 = alpha * (w6 - last);
        const double dd5 = do5 - dn5, dd6 = do6 - dn6;
        const double x5 = Rs.dual(dd5, hdt, inv_alpha), x6 = Rs.dual(dd6, hdt, inv_alpha);
        Rs.primal(tn - to, 0.5 * (dd5 + dd6), 0.5 * (x5 + x6), inv_alpha);
        if (g == 0 && active) {
            Dn[L.d5 + e0] = dn5;
            Dn[L.d6 + e0] = dn6;
            Pb[L.ptau + j] = tn - alpha * (0.5 * (dn5 + dn6));
        }
    }
    // The idle groups of the last CTA walked node `first` a second time -- possibly AFTER its own group had replaced p
    // by pbar: whatever they computed must not reach the maxima
    if (!active) {
        Rs.init();
        bad = 0;
    }
    // block-level reduction of the six maxima (as bit patterns), one atomic per slot per block.  Warp level: two
    // 32-bit redux.sync per slot (the high words, then the low words of the lanes that hold the maximal high word)
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const unsigned long long mine = Rs.bits(i);
        const unsigned hi = (unsigned)(mine >> 32), lo = (unsigned)mine;
        const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
        const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
        if (lane == 0) blockmax[warp][i] = ((unsigned long long)mhi << 32) | mlo;
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&ctrl->status, 1);
    __syncthreads();
    if (tid < 6) {
        unsigned long long mval = blockmax[0][tid];
#pragma unroll
        for (int wv = 1; wv < kChainDualThreads / 32; ++wv) mval = blockmax[wv][tid] > mval ? blockmax[wv][tid] : mval;
        atomicMax(reinterpret_cast<unsigned long long *>(slots + (long long)blockIdx.y * 6 + tid), mval);
    }
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) ctrl->pending = 1;
    if (arrive) iteration_arrive(ctrl, slots);
}

// ====================================================================================================================
// The risk block of the chain nodes on its own: d1_i, d2_i (dual of R_+^{2c} x {0} and of R_+; risks.py:32-35,
// cache.py:349-352), the y_i / s_i residual rows, and ybar_i, sbar_i of the next iteration.  It needs y, s only -- final
// once the kernel projection is done -- so it runs right after it on the side stream, under the DP sweeps, and the chain
// dual pass on the critical path is 15 % shorter.  One thread per node (three y entries), a few fat CTAs like
// k_kproj_node (it runs next to the chain walkers).  Same arithmetic, in the same order, as the block in k_dual_chain.
// ====================================================================================================================
__global__ void __launch_bounds__(1024) k_dual_risk_chain(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl,
                                                           const double *__restrict__ p_old, const double *__restrict__ p_new,
                                                           const double *__restrict__ d_old, double *__restrict__ d_new,
                                                           double *__restrict__ slots, int first, int count, int stride,
                                                           int yo0, double *pbar, OwnMap own, int arrive) {
    if (ctrl->done) return;
    const double alpha = ctrl->alpha, inv_alpha = 1.0 / alpha;
    const Layout &L = P.L;
    const Topo &T = P.t;
    __shared__ unsigned long long blockmax[32][6];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double *Po = p_old + (long long)blockIdx.y * L.np_pad;
    const double *Pn = p_new + (long long)blockIdx.y * L.np_pad;
    const double *Do = d_old + (long long)blockIdx.y * L.nd_pad;
    double *Dn = d_new + (long long)blockIdx.y * L.nd_pad;
    double *Pb = pbar + (long long)blockIdx.y * L.np_pad;
    ResidLane Rs;
    Rs.init();
    for (int i0 = blockIdx.x * blockDim.x + tid; i0 < count; i0 += gridDim.x * blockDim.x) {
        const int i = own.at(i0);
        const int node = first + i, yo = yo0 + 3 * i;
        const int j = stride > 0 ? node + stride : T.child_first[node];
        const double prob = T.cond_prob[j];
        double yo_[3], yn_[3], d1_[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            yo_[e] = Po[L.py + yo + e];
            yn_[e] = Pn[L.py + yo + e];
            d1_[e] = Do[L.d1 + yo + e];
        }
        const double so = Po[L.ps + node], sn = Pn[L.ps + node], do2 = Do[L.d2 + node];
        const double b[3] = {prob, 0.0, 1.0};   // b_i = [pi; 0; 1] (risks.py:34-35)
        // b' (2 y+ - y) and b' (y+ - y), summed like the lane-group reduction of k_dual_chain: (e0 + e1) + (e2 + 0)
        const double dot_z = (b[0] * (2 * yn_[0] - yo_[0]) + b[1] * (2 * yn_[1] - yo_[1])) + (b[2] * (2 * yn_[2] - yo_[2]) + 0.0);
        const double dot_d = (b[0] * (yn_[0] - yo_[0]) + b[1] * (yn_[1] - yo_[1])) + (b[2] * (yn_[2] - yo_[2]) + 0.0);
        const double w2 = dual_w(do2, (2 * sn - so) - dot_z, alpha, inv_alpha);
        const double dn2 = alpha * (w2 - fmax(0.0, w2));
        const double dd2 = do2 - dn2;
        const double xi22 = fma(dd2, inv_alpha, (sn - so) - dot_d);
        Rs.put(2, xi22);
        Rs.put(5, dd2);
        Rs.primal(sn - so, dd2, xi22, inv_alpha);
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const double dy = yn_[e] - yo_[e];
            const double wv = dual_w(d1_[e], 2 * yn_[e] - yo_[e], alpha, inv_alpha);
            const double zv = e < 2 ? fmax(0.0, wv) : wv;   // dual of R_+^{2c} x {0} (risks.py:32-33)
            const double dnew = alpha * (wv - zv);
            const double dd = d1_[e] - dnew;
            const double xi2 = Rs.dual(dd, dy, inv_alpha);
            Rs.primal(dy, dd - b[e] * dd2, xi2 - b[e] * xi22, inv_alpha);
            Dn[L.d1 + yo + e] = dnew;
            Pb[L.py + yo + e] = yn_[e] - alpha * (dnew - b[e] * dn2);
        }
        Dn[L.d2 + node] = dn2;
        Pb[L.ps + node] = sn - alpha * dn2;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const unsigned long long mine = Rs.bits(i);
        const unsigned hi = (unsigned)(mine >> 32), lo = (unsigned)mine;
        const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
        const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
        if (lane == 0) blockmax[warp][i] = ((unsigned long long)mhi << 32) | mlo;
    }
    __syncthreads();
    if (tid < 6) {
        unsigned long long mval = blockmax[0][tid];
        for (int wv = 1; wv < (int)(blockDim.x >> 5); ++wv) mval = blockmax[wv][tid] > mval ? blockmax[wv][tid] : mval;
        atomicMax(reinterpret_cast<unsigned long long *>(slots + (long long)blockIdx.y * 6 + tid), mval);
    }
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) ctrl->pending = 1;
    if (arrive) iteration_arrive(ctrl, slots);
}

// ====================================================================================================================
// Projection onto ker [E' -I -I] of every nonleaf node (cache.py:290-317), IN PLACE on a buffer that holds
// (ybar_i, taubar_j, sbar_j) -- the part of prox_f that does not depend on the DP sweeps; it runs next to them on its
// own stream.  One thread per node; the sums follow the association of k_primal_lane's lane-group reductions.
// ====================================================================================================================
__global__ void __launch_bounds__(1024) k_kproj_node(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                      double *__restrict__ prim, const double *__restrict__ x0,
                                                      double *__restrict__ p_old, const int *__restrict__ node_list, int count) {
    // node_list / count: the nonleaf nodes to project (subtree sharding); null = all of them
    if (ctrl->done) return;
    const Layout &L = P.L;
    const Topo &T = P.t;
    double *Pb = prim + (long long)blockIdx.y * L.np_pad;
    if (x0 && blockIdx.x == 0 && threadIdx.x < L.nx)   // x_0 of the old iterate = the initial state (cache.py:79-82)
        p_old[(long long)blockIdx.y * L.np_pad + L.px + threadIdx.x] = x0[(long long)blockIdx.y * L.nx + threadIdx.x];
    const int total = node_list ? count : L.m;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < total; slot += gridDim.x * blockDim.x) {
    const int node = node_list ? node_list[slot] : slot;
    const int c0 = T.child_first[node], cc = T.child_count[node];
    const double a = T.risk_alpha[node], den = a * a + 3.0;
    double *y = Pb + L.py + T.yoff[node], *tau = Pb + L.ptau + c0, *sv = Pb + L.ps + c0;
    const double ylast = y[2 * cc];
    double res[kLaneMaxChildren];
#pragma unroll
    for (int e = 0; e < kLaneMaxChildren; ++e) res[e] = e < cc ? a * y[e] - y[cc + e] + ylast - tau[e] - sv[e] : 0.0;
    const double rsum = ((res[0] + res[1]) + (res[2] + res[3])) + ((res[4] + res[5]) + (res[6] + res[7]));
    const double shift = rsum / (den + (double)cc);
    double wv[kLaneMaxChildren];
#pragma unroll
    for (int e = 0; e < kLaneMaxChildren; ++e) {
        wv[e] = 0.0;
        if (e < cc) {
            wv[e] = (res[e] - shift) / den;
            y[e] = y[e] - a * wv[e];
            y[cc + e] = y[cc + e] + wv[e];
            tau[e] = tau[e] + wv[e];
            sv[e] = sv[e] + wv[e];
        }
    }
    const double wsum = ((wv[0] + wv[1]) + (wv[2] + wv[3])) + ((wv[4] + wv[5]) + (wv[6] + wv[7]));
    y[2 * cc] = ylast - wsum;
    }
}

// ---- host launchers: lanes per node chosen from the row length (pairs of doubles per lane and round) ----------------------
int lane_group_width(int nx) { return nx / 2 <= 12 ? 4 : 8; }

void launch_primal_lane(dim3 nodes_batch, cudaStream_t st, const Params &P, const Ctrl *ctrl, const double *p_old,
                        const double *d_old, double *p_new, const int *node_list, int count) {
    const int G = lane_group_width(P.L.nx);
    const dim3 grid((count * G + kLaneThreads - 1) / kLaneThreads, nodes_batch.y);
    if (G == 4) k_primal_lane<4><<<grid, kLaneThreads, 0, st>>>(P, ctrl, p_old, d_old, p_new, node_list, count);
    else k_primal_lane<8><<<grid, kLaneThreads, 0, st>>>(P, ctrl, p_old, d_old, p_new, node_list, count);
}

int launch_dual_lane(dim3 nodes_batch, cudaStream_t st, const Params &P, Ctrl *ctrl, const double *p_old,
                     const double *p_new, const double *d_old, double *d_new, double *slots, const int *node_list,
                     int first, int count, double *pbar, bool narrow, bool arrive) {
    if (count <= 0) return 0;
    const int arr = arrive ? 1 : 0;
    const int G = lane_group_width(P.L.nx);
    const dim3 grid((count * G + kLaneThreads - 1) / kLaneThreads, nodes_batch.y);
    // A launch over a few thousand nodes (the branching top of the tree, the leaves) is a handful of warps per SM, each
    // a long serial instruction stream: such launches get more lanes per node (16: a row is one round per lane, and
    // there are four times as many warps to spread over the SMs) and the 128-register build (no spills); a launch that
    // fills the GPU runs the 80-register build with three CTAs per SM
    const long long threads4 = (long long)count * 4 * nodes_batch.y;
    static const int force_g = getenv("RB_SMALL_G") ? atoi(getenv("RB_SMALL_G")) : 0;   // ablation knob
    int Gs = threads4 <= 148LL * kLaneThreads / 2 ? 16 : (threads4 <= 148LL * kLaneThreads ? 8 : G);
    static const int early_g = getenv("RB_EARLY_G") ? atoi(getenv("RB_EARLY_G")) : 0;   // ablation knob
    if (narrow && early_g == 85 && G == 4) {   // 8 lanes per node, 512-thread CTAs: the same few SMs, shorter warps
        const dim3 grid_w((count * 8 + 511) / 512, nodes_batch.y);
        k_dual_lane<8, 1, 512><<<grid_w, 512, 0, st>>>(P, ctrl, p_old, p_new, d_old, d_new, slots, node_list, first, count, pbar, arr);
        return (int)(grid_w.x * grid_w.y);
    }
    if (narrow) Gs = early_g == 8 || early_g == 16 ? early_g : G;
    else if (force_g == 4) Gs = G;
    else if (force_g == 8 && Gs == 16) Gs = 8;
    const dim3 grid_s((count * Gs + kLaneThreads - 1) / kLaneThreads, nodes_batch.y);
    const bool roomy = (long long)grid_s.x * grid_s.y <= 2 * 148;
#define RB_GO(G_, B_, GRID_)                                                                                                      \
    {                                                                                                                             \
        k_dual_lane<G_, B_><<<GRID_, kLaneThreads, 0, st>>>(P, ctrl, p_old, p_new, d_old, d_new, slots, node_list, first, count, pbar, arr); \
        return (int)(GRID_.x * GRID_.y);                                                                                          \
    }
    if (roomy && Gs == 16) RB_GO(16, 2, grid_s)
    else if (roomy && Gs == 8) RB_GO(8, 2, grid_s)
    else if (G == 4 && roomy) RB_GO(4, 2, grid)
    else if (G == 4) RB_GO(4, 3, grid)
    else if (roomy) RB_GO(8, 2, grid)
    else RB_GO(8, 3, grid)
#undef RB_GO
}

// sizes with an instantiation of the chain dual pass (lanes per node as lane_group_width picks them)
#define RB_CHAIN_DUAL_DIMS(X) X(4, 2, 4) X(8, 4, 4) X(16, 8, 4) X(20, 10, 4) X(64, 32, 8)

bool dual_chain_supported(int nx, int nu) {
#define RB_HAS(NX, NU, G) \
    if (nx == NX && nu == NU) return true;
    RB_CHAIN_DUAL_DIMS(RB_HAS)
#undef RB_HAS
    return false;
}

int launch_dual_chain(int batch, cudaStream_t st, const Params &P, Ctrl *ctrl, const double *p_old, const double *p_new,
                      const double *d_old, double *d_new, double *slots, const int4 *recs, int first, int count,
                      int stride, int yo0, double *pbar, int with_risk, OwnMap own, bool arrive) {
    if (count <= 0) return 0;
    with_risk = (with_risk ? 1 : 0) | (arrive ? 2 : 0);   // bit 1: the CTAs count themselves in (iteration_arrive)
    // resident CTAs per SM the kernel is compiled for: 3 (168 registers, no spills) or 4 (128 registers, ~40 words
    // spilled to L1); RB_CHAIN_DUAL_MINB overrides for ablation runs
    static const int minb = [] {
        const char *e = getenv("RB_CHAIN_DUAL_MINB");
        return e && atoi(e) == 4 ? 4 : 3;
    }();
    static const bool wide = [] {   // ablation: 8 lanes per node where 4 is the default
        const char *e = getenv("RB_CHAIN_DUAL_G");
        return e && atoi(e) == 8;
    }();
#define RB_GO_G(NX, NU, G)                                                                                           \
    {                                                                                                                \
        const dim3 grid((count * G + kChainDualThreads - 1) / kChainDualThreads, batch);                             \
        if (minb == 4)                                                                                               \
            k_dual_chain<NX, NU, G, 4><<<grid, kChainDualThreads, 0, st>>>(P, ctrl, p_old, p_new, d_old, d_new, slots, recs, first, count, stride, yo0, pbar, with_risk, own); \
        else                                                                                                         \
            k_dual_chain<NX, NU, G, 3><<<grid, kChainDualThreads, 0, st>>>(P, ctrl, p_old, p_new, d_old, d_new, slots, recs, first, count, stride, yo0, pbar, with_risk, own); \
        return (int)(grid.x * grid.y);                                                                               \
    }
#define RB_GO(NX, NU, G)                                                                                             \
    if (P.L.nx == NX && P.L.nu == NU) {                                                                              \
        if (wide && G == 4) RB_GO_G(NX, NU, 8)                                                                       \
        RB_GO_G(NX, NU, G)                                                                                           \
    }
    RB_CHAIN_DUAL_DIMS(RB_GO)
#undef RB_GO
    return 0;
}

int launch_dual_risk_chain(int batch, cudaStream_t st, const Params &P, Ctrl *ctrl, const double *p_old, const double *p_new,
                           const double *d_old, double *d_new, double *slots, int first, int count, int stride, int yo0,
                           double *pbar, OwnMap own, bool arrive) {
    if (count <= 0) return 0;
    const int ctas = std::max(1, std::min((count + 1023) / 1024, std::max(1, 16 / batch)));
    k_dual_risk_chain<<<dim3(ctas, batch), 1024, 0, st>>>(P, ctrl, p_old, p_new, d_old, d_new, slots, first, count, stride, yo0, pbar, own,
                                                          arrive ? 1 : 0);
    return ctas * batch;
}

void launch_kproj(int batch, cudaStream_t st, const Params &P, const Ctrl *ctrl, double *prim, const double *x0, double *p_old,
                  const int *node_list, int count) {
    // The pass runs next to the backward chain walker, which keeps one latency-critical warp per SM sub-partition on
    // ~128 SMs: a few fat CTAs (they land on the SMs the walker leaves idle) disturb it less than a grid spread over all
    // SMs (cfg3, L2 flushed before every iteration: 9086 it/s with 16 CTAs of 1024 threads against 7702 with 230 CTAs of
    // 256).  RB_KPROJ_CTAS: ablation knob.
    static const int cap = getenv("RB_KPROJ_CTAS") ? atoi(getenv("RB_KPROJ_CTAS")) : 16;   // 0: one thread per node, 256-thread CTAs
    const int threads = cap > 0 ? 1024 : 256;
    const int total = node_list ? count : P.L.m;
    if (total <= 0 && !x0) return;
    int ctas = std::max(1, (total + threads - 1) / threads);
    if (cap > 0) ctas = std::min(ctas, std::max(1, cap / batch));
    k_kproj_node<<<dim3(ctas, batch), threads, 0, st>>>(P, ctrl, prim, x0, p_old, node_list, count);
}

}  // namespace rb
