// batch.cu -- the Chambolle-Pock iteration for MANY independent initial states on one small tree (BASELINE.json configs[3]:
// 4096 instances of a 1 023-node tree), in a BATCH-INNERMOST layout.
//
// The instance-major buffers of the other kernels (one contiguous iterate per instance) make a batch of small trees a pile of
// tiny, latency-bound launches: a warp per parent uses 12 of 32 lanes, every CTA stages its own descriptor for 15 nodes
// (profiles/r1_kernel_evolution.md, cfg4: 7.5 % of the HBM roofline).  Here the instances are the LANES:
//
//   panel layout:  element e (padded compact offset, Layout::px ... d14) of instance b  ->  ((b / 32) * stride + e) * 32 + b % 32
//
// so every access of a warp is one coalesced 256-byte row, one thread runs the whole per-node arithmetic of ITS instance in
// registers (no shuffles, no shared memory; the SOC norm, the kernel projection and the residual maxima are per-lane scalars),
// and the operator tables are warp-uniform loads.  The fused loop converts the iterate into panels at rb_loop_begin and back
// at rb_loop_end (two transposes per solve); everything in between -- reference solver.py:124-161 with cache.py:248-393 and
// operators.py:19-94 -- runs on the panels:
//
//   k_bp_primal   (first iteration of a loop)  pbar = p - alpha L* d, s_0 -= alpha, kernel projection   solver.py:27-39, cache.py:253-257,290-317
//   k_bp_kproj    (later iterations)           s_0 -= alpha and the kernel projection in place on the pbar the dual pass left
//   k_bp_bwd / k_bp_fwd (one launch per stage) the DP projection onto the dynamics                        cache.py:259-288
//   k_bp_dual_xu / _risk / _leaf               L, dual half step, prox of g*, six residual maxima, pbar of the next iteration
//                                                                                        solver.py:44-95, cache.py:321-393
//
// One thread of a nonleaf node owns the node's x, u, y, s, d1, d2, d7 and the tau_j, d3_j ... d6_j of its child edges, so
// every L / L* term is thread-local.  Diagonal cost square roots only (the host checks), any number of children.
#include <algorithm>

#include "kernels.cuh"

namespace rb {

namespace {

// resident CTAs per SM the dual kernels are compiled for (register caps); measured on cfg4 x 4096 (gpurun_out r2p): leaf 3 -> 4:
// 318 -> 282 us; risk 3 -> 4: 100 -> 146 us (spills); x / u block 2 -> 3: 439 -> 505 us (spills)
#ifndef RB_BP_LEAF_MINB
#define RB_BP_LEAF_MINB 4
#endif
#ifndef RB_BP_RISK_MINB
#define RB_BP_RISK_MINB 3
#endif
#ifndef RB_BP_XU_MINB
#define RB_BP_XU_MINB 2
#endif
constexpr int kPanel = 32;
constexpr int kBpWarps = 4;   // nodes per CTA and pass of the node loops
constexpr int kBpMaxChildren = 4;   // compile-time bound of the risk kernel (2 c + 1 entries of y_i in registers)

// ---- panel <-> instance-major ----------------------------------------------------------------------------------------------------
// src: [batch][stride] instance-major; dst: panels of 32 instances.  Tile: 32 elements x 32 instances through shared memory,
// both sides coalesced.  Instances >= batch (padding of the last panel) are written as zeros / skipped.
__global__ void __launch_bounds__(256) k_bp_to_panels(const double *__restrict__ src, double *__restrict__ dst, long long stride,
                                                     int batch) {
    __shared__ double tile[32][33];
    const long long e0 = (long long)blockIdx.x * 32;
    const int panel = blockIdx.y, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {   // instance i of the panel, element e0 + tx
        const int b = panel * kPanel + i;
        tile[i][tx] = (b < batch && e0 + tx < stride) ? src[(long long)b * stride + e0 + tx] : 0.0;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8)     // element e0 + k, instance tx
        if (e0 + k < stride) dst[((long long)panel * stride + e0 + k) * kPanel + tx] = tile[tx][k];
}
__global__ void __launch_bounds__(256) k_bp_from_panels(const double *__restrict__ src, double *__restrict__ dst, long long stride,
                                                       int batch) {
    __shared__ double tile[32][33];
    const long long e0 = (long long)blockIdx.x * 32;
    const int panel = blockIdx.y, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8)
        tile[k][tx] = e0 + k < stride ? src[((long long)panel * stride + e0 + k) * kPanel + tx] : 0.0;
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int b = panel * kPanel + i;
        if (b < batch && e0 + tx < stride) dst[(long long)b * stride + e0 + tx] = tile[tx][i];
    }
}

// ---- views -------------------------------------------------------------------------------------------------------------------------
// pointer to element e of this lane's instance
struct PanelPtr {
    double *base;   // panel start + lane
    __device__ __forceinline__ double &operator[](long long e) const { return base[e * kPanel]; }
};
__device__ __forceinline__ PanelPtr panel_view(double *buf, long long stride, int panel, int lane) {
    return PanelPtr{buf + (long long)panel * stride * kPanel + lane};
}
__device__ __forceinline__ PanelPtr panel_view(const double *buf, long long stride, int panel, int lane) {
    return PanelPtr{const_cast<double *>(buf) + (long long)panel * stride * kPanel + lane};
}

// L2 prefetch of `rows` consecutive element rows (256 bytes each) of this warp's panel, starting at element e0: lane l takes
// the l-th 128-byte line.  The dual kernels hold their bytes in flight in registers (255 of them: 8 warps per SM); prefetching the
// NEXT node of the warp's loop into L2 while the current one is computed shortens the latency the registers have to cover, at
// 1 instruction per 32 lines.
__device__ __forceinline__ void prefetch_rows(const double *panel_base, long long e0, int rows, int lane) {
#ifndef RB_BP_NO_PREFETCH
    const char *p = reinterpret_cast<const char *>(panel_base + e0 * kPanel);
    for (int l = lane; l < 2 * rows; l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + (long long)l * 128));
#endif
}

// running maxima of |v| as bit patterns: NaN (0x7ff8...) sorts above +inf and sticks (common.cuh atomic_max_nonneg)
__device__ __forceinline__ void upd(unsigned long long &m, double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(fabs(v));
    m = b > m ? b : m;
}

// the kernel projection of (y_i, tau_children, s_children) onto ker [E' -I -I], E = [a I; -I; 1'] (risks.py:28-31), in closed
// form: M M' = (a^2 + 3) I + 1 1'  (DESIGN.md section 3).  Two passes over the children, everything per lane.
__device__ __forceinline__ void kproj_node(const Topo &T, const Layout &L, const PanelPtr &P, int node) {
    const int c0 = T.child_first[node], cc = T.child_count[node];
    const long long yo = L.py + T.yoff[node];
    const double a = T.risk_alpha[node], A = a * a + 3.0;
    const double ylast = P[yo + 2 * cc];
    double sumz = 0.0;
    for (int j = 0; j < cc; ++j)
        sumz += a * P[yo + j] - P[yo + cc + j] + ylast - P[L.ptau + c0 + j] - P[L.ps + c0 + j];
    const double shift = sumz / (A + cc);
    double sumw = 0.0;
    for (int j = 0; j < cc; ++j) {
        const double ya = P[yo + j], yb = P[yo + cc + j], tj = P[L.ptau + c0 + j], sj = P[L.ps + c0 + j];
        const double w = ((a * ya - yb + ylast - tj - sj) - shift) / A;
        P[yo + j] = ya - a * w;
        P[yo + cc + j] = yb + w;
        P[L.ptau + c0 + j] = tj + w;
        P[L.ps + c0 + j] = sj + w;
        sumw += w;
    }
    P[yo + 2 * cc] = ylast - sumw;
}

// ---- first iteration of a loop: pbar = p - alpha L* d into `out`, then prox_f's node-local parts ---------------------------------------
// pass 1 (every node): xbar, ubar, ybar, sbar of the node, taubar of its child edges; pass 2 (same kernel, after a grid-wide
// dependency is avoided by letting the PARENT's thread recompute nothing: the kernel projection needs sbar of the children,
// which other threads write) -> the projection runs as k_bp_kproj right after this kernel.
template <int NX, int NU>
__global__ void __launch_bounds__(kBpWarps * 32) k_bp_primal(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                            const double *__restrict__ p_old, const double *__restrict__ d_old,
                                                            double *__restrict__ p_out) {
    if (ctrl && ctrl->done) return;
    const Layout &L = P.L;
    const Topo &T = P.t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, panel = blockIdx.y;
    const PanelPtr po = panel_view(p_old, L.np_pad, panel, lane), d = panel_view(d_old, L.nd_pad, panel, lane);
    const PanelPtr pn = panel_view(p_out, L.np_pad, panel, lane);
    const double alpha = ctrl->alpha;
    for (int node = blockIdx.x * kBpWarps + warp; node < L.n; node += gridDim.x * kBpWarps) {
        if (node < L.m) {
            const int c0 = T.child_first[node], cc = T.child_count[node];
            double ax[NX], au[NU];
#pragma unroll
            for (int k = 0; k < NX; ++k) ax[k] = L.has_nl_rect ? d[L.d7 + (long long)node * (NX + NU) + k] : 0.0;
#pragma unroll
            for (int k = 0; k < NU; ++k) au[k] = L.has_nl_rect ? d[L.d7 + (long long)node * (NX + NU) + NX + k] : 0.0;
            for (int j = c0; j < c0 + cc; ++j) {
                const int ci = T.cost_idx[j];
                const long long e = j - 1;
#pragma unroll
                for (int k = 0; k < NX; ++k) ax[k] = fma(__ldg(P.m.sq_d + ci * NX + k), d[L.d3 + e * NX + k], ax[k]);
#pragma unroll
                for (int k = 0; k < NU; ++k) au[k] = fma(__ldg(P.m.sr_d + ci * NU + k), d[L.d4 + e * NU + k], au[k]);
                pn[L.ptau + j] = po[L.ptau + j] - alpha * (0.5 * (d[L.d5 + e] + d[L.d6 + e]));
            }
#pragma unroll
            for (int k = 0; k < NX; ++k) pn[L.px + (long long)node * NX + k] = po[L.px + (long long)node * NX + k] - alpha * ax[k];
#pragma unroll
            for (int k = 0; k < NU; ++k) pn[L.pu + (long long)node * NU + k] = po[L.pu + (long long)node * NU + k] - alpha * au[k];
            const long long yo = T.yoff[node];
            const double d2v = d[L.d2 + node];
            for (int e = 0; e < 2 * cc + 1; ++e) {
                const double b = e < cc ? __ldg(T.cond_prob + c0 + e) : (e == 2 * cc ? 1.0 : 0.0);
                pn[L.py + yo + e] = po[L.py + yo + e] - alpha * (d[L.d1 + yo + e] - b * d2v);
            }
            pn[L.ps + node] = po[L.ps + node] - alpha * d2v;
            if (node == 0) pn[L.ptau] = po[L.ptau];   // tau_0 is not a variable (never written by L*, operators.py:55-94)
        } else {
            const long long li = node - L.m;
            const int ci = T.leafcost_idx[li];
#pragma unroll
            for (int k = 0; k < NX; ++k) {
                double a = __ldg(P.m.sqf_d + ci * NX + k) * d[L.d11 + li * NX + k];
                if (L.has_leaf_rect) a += d[L.d14 + li * NX + k];
                pn[L.px + (long long)node * NX + k] = po[L.px + (long long)node * NX + k] - alpha * a;
            }
            pn[L.ps + node] = po[L.ps + node] - alpha * (0.5 * (d[L.d12 + li] + d[L.d13 + li]));
        }
    }
}

// s_0 -= alpha (cache.py:253-257) and the kernel projection (cache.py:290-317) in place; also x_0 of the OLD iterate <- x0
// (cache_initial_state, cache.py:79-82: rb_step uploads x0 once, instance-major [batch][nx])
__global__ void __launch_bounds__(kBpWarps * 32) k_bp_kproj(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                           double *prim, const double *__restrict__ x0, double *p_old) {
    if (ctrl && ctrl->done) return;
    const Layout &L = P.L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, panel = blockIdx.y;
    const PanelPtr p = panel_view(prim, L.np_pad, panel, lane);
    for (int node = blockIdx.x * kBpWarps + warp; node < L.m; node += gridDim.x * kBpWarps) {
        if (node == 0) {
            p[L.ps] = p[L.ps] - ctrl->alpha;
            const int b = panel * kPanel + lane;
            if (x0 && p_old && b < L.batch) {
                const PanelPtr po = panel_view(p_old, L.np_pad, panel, lane);
                for (int k = 0; k < L.nx; ++k) po[L.px + k] = x0[(long long)b * L.nx + k];
            }
        }
        kproj_node(P.t, L, p, node);
    }
}

// ---- DP sweeps ---------------------------------------------------------------------------------------------------------------------------
// backward at one parent: r = ubar - sum_j B_j' q_j, q = sum_j A_j' q_j - xbar - K' r; a leaf child's q is
// -xbar_j, read in place.  Q: [panel][n * NX], R: [panel][m * NU].  CG: the children's q were written by other warps of THIS
// launch (fused top kernel): read them past the L1.
template <int NX, int NU, bool CG>
__device__ __forceinline__ void bp_bwd_node(const Params &P, const PanelPtr &p, const PanelPtr &Q, const PanelPtr &R, int node) {
    const Layout &L = P.L;
    const Topo &T = P.t;
    const int c0 = T.child_first[node], cc = T.child_count[node];
    double acc[NX + NU];
#pragma unroll
    for (int k = 0; k < NX + NU; ++k) acc[k] = 0.0;
    // the node's own rows and the q of its children TWO AT A TIME, all loads issued before the first use (one memory round trip
    // per pair of children instead of one per child: the wide stages were latency-bound at 2.8 TB/s)
    double xbar[NX], ubar[NU];
#pragma unroll
    for (int k = 0; k < NX; ++k) xbar[k] = p[L.px + (long long)node * NX + k];
#pragma unroll
    for (int a = 0; a < NU; ++a) ubar[a] = p[L.pu + (long long)node * NU + a];
    for (int j0 = c0; j0 < c0 + cc; j0 += 2) {
        const bool two = j0 + 1 < c0 + cc;
        const int jb = two ? j0 + 1 : j0;
        double qa[NX], qb[NX];
#pragma unroll
        for (int l = 0; l < NX; ++l) {
            qa[l] = j0 >= L.m ? -p[L.px + (long long)j0 * NX + l] : (CG ? __ldcg(&Q[(long long)j0 * NX + l]) : Q[(long long)j0 * NX + l]);
            qb[l] = jb >= L.m ? -p[L.px + (long long)jb * NX + l] : (CG ? __ldcg(&Q[(long long)jb * NX + l]) : Q[(long long)jb * NX + l]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h == 1 && !two) break;
            const double *qj = h ? qb : qa;
            const double *C = P.m.ABcat + (long long)T.dyn_idx[h ? jb : j0] * NX * (NX + NU);   // row l = [A[l][:], B[l][:]]
            if constexpr ((NX + NU) % 2 == 0) {   // rows of an even number of doubles: 16-byte warp-uniform loads
#pragma unroll
                for (int l = 0; l < NX; ++l)
#pragma unroll
                    for (int k = 0; k < NX + NU; k += 2) {
                        const double2 c = __ldg(reinterpret_cast<const double2 *>(C + l * (NX + NU) + k));
                        acc[k] = fma(c.x, qj[l], acc[k]);
                        acc[k + 1] = fma(c.y, qj[l], acc[k + 1]);
                    }
            } else {
#pragma unroll
                for (int l = 0; l < NX; ++l)
#pragma unroll
                    for (int k = 0; k < NX + NU; ++k) acc[k] = fma(__ldg(C + l * (NX + NU) + k), qj[l], acc[k]);
            }
        }
    }
    double rv[NU];
#pragma unroll
    for (int a = 0; a < NU; ++a) {
        rv[a] = ubar[a] - acc[NX + a];
        R[(long long)node * NU + a] = rv[a];
    }
    const double *K = P.m.K + (long long)T.cls[node] * NU * NX;   // [nu][nx]
#pragma unroll
    for (int k = 0; k < NX; ++k) {
        double kr = 0.0;
#pragma unroll
        for (int a = 0; a < NU; ++a) kr = fma(__ldg(K + a * NX + k), rv[a], kr);
        Q[(long long)node * NX + k] = acc[k] - xbar[k] - kr;
    }
}

// forward at one parent: u = K x + R~^-1 r, x_j = A_j x + B_j u; x_0 = x0 (instance-major [batch][nx]).  CG: x of the node was
// written by another warp of this launch, over an address this SM may have cached as xbar.
template <int NX, int NU, bool CG>
__device__ __forceinline__ void bp_fwd_node(const Params &P, const PanelPtr &p, const PanelPtr &R, const double *__restrict__ x0, int b,
                                            int node) {
    const Layout &L = P.L;
    const Topo &T = P.t;
    double v[NX + NU];
    if (node == 0) {
#pragma unroll
        for (int k = 0; k < NX; ++k) {
            v[k] = b < L.batch ? x0[(long long)b * NX + k] : 0.0;
            p[L.px + k] = v[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < NX; ++k) v[k] = CG ? __ldcg(&p[L.px + (long long)node * NX + k]) : p[L.px + (long long)node * NX + k];
    }
    double rv[NU];
#pragma unroll
    for (int a = 0; a < NU; ++a) rv[a] = R[(long long)node * NU + a];   // written by this same thread (top kernel) or an earlier launch
    const double *KR = P.m.KRcatT + (long long)T.cls[node] * (NX + NU) * NU;   // row l < nx: K[:][l]; row nx + b: R~^-1[:][b]
#pragma unroll
    for (int a = 0; a < NU; ++a) {
        double u = 0.0;
#pragma unroll
        for (int l = 0; l < NX; ++l) u = fma(__ldg(KR + l * NU + a), v[l], u);
#pragma unroll
        for (int l = 0; l < NU; ++l) u = fma(__ldg(KR + (NX + l) * NU + a), rv[l], u);
        v[NX + a] = u;
        p[L.pu + (long long)node * NU + a] = u;
    }
    const int c0 = T.child_first[node], cc = T.child_count[node];
    for (int j = c0; j < c0 + cc; ++j) {
        const double *CT = P.m.ABcatT + (long long)T.dyn_idx[j] * (NX + NU) * NX;   // row l < nx: A[:][l]; row nx + a: B[:][a]
        if constexpr (NX % 2 == 0) {
            double xj[NX];
#pragma unroll
            for (int k = 0; k < NX; ++k) xj[k] = 0.0;
#pragma unroll
            for (int l = 0; l < NX + NU; ++l)
#pragma unroll
                for (int k = 0; k < NX; k += 2) {
                    const double2 c = __ldg(reinterpret_cast<const double2 *>(CT + l * NX + k));
                    xj[k] = fma(c.x, v[l], xj[k]);
                    xj[k + 1] = fma(c.y, v[l], xj[k + 1]);
                }
#pragma unroll
            for (int k = 0; k < NX; ++k) p[L.px + (long long)j * NX + k] = xj[k];
        } else {
#pragma unroll
            for (int k = 0; k < NX; ++k) {
                double x = 0.0;
#pragma unroll
                for (int l = 0; l < NX + NU; ++l) x = fma(__ldg(CT + l * NX + k), v[l], x);
                p[L.px + (long long)j * NX + k] = x;
            }
        }
    }
}

// one stage per launch (the wide stages)
template <int NX, int NU>
__global__ void __launch_bounds__(kBpWarps * 32) k_bp_bwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                         const double *__restrict__ prim, double *__restrict__ q,
                                                         double *__restrict__ r, int first, int count) {
    if (ctrl && ctrl->done) return;
    const Layout &L = P.L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, panel = blockIdx.y;
    const PanelPtr p = panel_view(prim, L.np_pad, panel, lane);
    const PanelPtr Q = panel_view(q, (long long)L.n * NX, panel, lane), R = panel_view(r, (long long)L.m * NU, panel, lane);
    for (int node = first + blockIdx.x * kBpWarps + warp; node < first + count; node += gridDim.x * kBpWarps)
        bp_bwd_node<NX, NU, false>(P, p, Q, R, node);
}
template <int NX, int NU>
__global__ void __launch_bounds__(kBpWarps * 32) k_bp_fwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                         double *__restrict__ prim, const double *__restrict__ r,
                                                         const double *__restrict__ x0, int first, int count) {
    if (ctrl && ctrl->done) return;
    const Layout &L = P.L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, panel = blockIdx.y;
    const PanelPtr p = panel_view(prim, L.np_pad, panel, lane);
    const PanelPtr R = panel_view(r, (long long)L.m * NU, panel, lane);
    const int b = panel * kPanel + lane;
    for (int node = first + blockIdx.x * kBpWarps + warp; node < first + count; node += gridDim.x * kBpWarps)
        bp_fwd_node<NX, NU, false>(P, p, R, x0, b, node);
}

// The narrow stages [0, t_top) in ONE launch: one CTA per panel, a warp per parent, backward to the root and forward again with a
// block barrier per stage.  As separate launches each of these stages costs 12 - 18 us of pure latency (ncu r2i: a stage of
// <= 32 parents x 128 panels is one dependent stream per warp) -- twelve of them were 190 of the 520 us of cfg4's sweeps.
constexpr int kTopWarps = 32;
template <int NX, int NU>
__global__ void __launch_bounds__(kTopWarps * 32) k_bp_top(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                          double *prim, double *q, double *r, const double *__restrict__ x0,
                                                          const int *__restrict__ stage_off, int t_top) {
    if (ctrl && ctrl->done) return;
    const Layout &L = P.L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, panel = blockIdx.x;
    const PanelPtr p = panel_view(prim, L.np_pad, panel, lane);
    const PanelPtr Q = panel_view(q, (long long)L.n * NX, panel, lane), R = panel_view(r, (long long)L.m * NU, panel, lane);
    const int b = panel * kPanel + lane;
    {   // xbar / ubar of every node this warp will visit, into L2 now: each stage step is otherwise one cold DRAM round trip
        const double *pb_ = prim + (long long)panel * L.np_pad * kPanel;
        for (int node = warp; node < stage_off[t_top]; node += kTopWarps) {
            prefetch_rows(pb_, L.px + (long long)node * NX, NX, lane);
            prefetch_rows(pb_, L.pu + (long long)node * NU, NU, lane);
        }
    }
    for (int t = t_top - 1; t >= 0; --t) {
        for (int node = stage_off[t] + warp; node < stage_off[t + 1]; node += kTopWarps) bp_bwd_node<NX, NU, true>(P, p, Q, R, node);
        __syncthreads();
    }
    for (int t = 0; t < t_top; ++t) {
        for (int node = stage_off[t] + warp; node < stage_off[t + 1]; node += kTopWarps) bp_fwd_node<NX, NU, true>(P, p, R, x0, b, node);
        __syncthreads();
    }
}

// ---- dual pass ---------------------------------------------------------------------------------------------------------------------------
// SecondOrderCone.project (cones.py:113-132, branch order kept) followed by the Moreau step d+ = alpha (w - z): returns the
// factor f and the last entry so that d+_k = f * w_k for the first dim - 1 entries.
struct SocStep {
    double f;       // d+_k = f * w_k,  k < dim - 1
    double dlast;   // d+ of the last entry (the cone's t)
};
__device__ __forceinline__ SocStep soc_step(double ss, double t, double alpha) {
    const double rr = sqrt(ss);
    SocStep o;
    if (rr <= t) {            // inside: z = w
        o.f = 0.0;
        o.dlast = 0.0;
    } else if (rr <= -t) {    // inside the polar: z = 0
        o.f = alpha;
        o.dlast = alpha * t;
    } else {
        const double tn = (rr + t) / 2;
        o.f = alpha * (1.0 - tn / rr);   // w_k - tn * (w_k / r)
        o.dlast = alpha * (t - tn);
    }
    return o;
}

struct Maxima {
    unsigned long long m[6];   // xi0, xi1, xi2, delta0, delta1, delta2
};
// CTA-wide fold of the per-lane maxima (lane = instance) and one atomic per lane and slot
__device__ __forceinline__ void flush_maxima(const Maxima &mx, double *slots, int batch, int panel, int lane, int warp, Ctrl *ctrl,
                                             int bad) {
    __shared__ unsigned long long red[kBpWarps][6][32];
#pragma unroll
    for (int i = 0; i < 6; ++i) red[warp][i][lane] = mx.m[i];
    __syncthreads();
    if (warp == 0) {
        const int b = panel * kPanel + lane;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            unsigned long long v = red[0][i][lane];
#pragma unroll
            for (int w = 1; w < kBpWarps; ++w) v = red[w][i][lane] > v ? red[w][i][lane] : v;
            if (b < batch && v) atomicMax(reinterpret_cast<unsigned long long *>(slots + (long long)b * 6 + i), v);
        }
    }
    if (bad) atomicOr(&ctrl->status, 1);
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) ctrl->pending = 1;
}

// c2[k] of a nonleaf node: the diagonal of L* L on its x / u rows = [rectangle] + sum_j sqrtQ_j[k]^2 (sqrtR_j for the u rows), so
// that  L* xi2 = L*(d - d+) / alpha + c2 (p+ - p)  on these rows needs no accumulator of its own.  Table [m][nx + nu], built once.
__global__ void k_bp_c2(const __grid_constant__ Params P, double *__restrict__ c2) {
    const Layout &L = P.L;
    const int node = blockIdx.x, k = threadIdx.x;
    if (k >= L.nxu) return;
    double v = L.has_nl_rect ? 1.0 : 0.0;
    for (int j = P.t.child_first[node]; j < P.t.child_first[node] + P.t.child_count[node]; ++j) {
        const int ci = P.t.cost_idx[j];
        const double w = k < L.nx ? P.m.sq_d[ci * L.nx + k] : P.m.sr_d[ci * L.nu + k - L.nx];
        v = fma(w, w, v);
    }
    c2[(long long)node * L.nxu + k] = v;
}

// The x / u block of a nonleaf node: d7 (rectangle) and the edges into its children (SOC on [d3_j; d4_j; d5_j; d6_j]), the
// residual rows of x_i, u_i, tau_j, and pbar of those rows.  The children are taken TWO AT A TIME with all their loads issued
// before the first use: a node costs one memory round trip per pair instead of one per child and phase (the first version of
// this kernel, one runtime loop over the children, ran at 2.7 TB/s with 8 warps per SM waiting on ~10 dependent round trips
// per node).
template <int NX, int NU>
__global__ void __launch_bounds__(kBpWarps * 32, RB_BP_XU_MINB) k_bp_dual_xu(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl,
                                                                const double *p_old, const double *__restrict__ p_new,
                                                                const double *__restrict__ d_old, double *__restrict__ d_new,
                                                                double *__restrict__ slots, double *pbar,   // pbar aliases p_old
                                                                const double *__restrict__ c2tab) {
    if (ctrl->done) return;
    const Layout &L = P.L;
    const Topo &T = P.t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, panel = blockIdx.y;
    const PanelPtr po = panel_view(p_old, L.np_pad, panel, lane), pn = panel_view(p_new, L.np_pad, panel, lane);
    const PanelPtr dO = panel_view(d_old, L.nd_pad, panel, lane), dN = panel_view(d_new, L.nd_pad, panel, lane);
    const PanelPtr pb = panel_view(pbar, L.np_pad, panel, lane);
    const double alpha = ctrl->alpha, ia = 1.0 / alpha;
    constexpr int S = NX + NU;
    Maxima mx;
#pragma unroll
    for (int i = 0; i < 6; ++i) mx.m[i] = 0ull;
    int bad = 0;
    const double *po_b = p_old + (long long)panel * L.np_pad * kPanel, *pn_b = p_new + (long long)panel * L.np_pad * kPanel;
    const double *do_b = d_old + (long long)panel * L.nd_pad * kPanel;
    for (int node = blockIdx.x * kBpWarps + warp; node < L.m; node += gridDim.x * kBpWarps) {
        const int c0 = T.child_first[node], cc = T.child_count[node];
        {   // the next node of this warp into L2
            const int nn = node + gridDim.x * kBpWarps;
            if (nn < L.m) {
                prefetch_rows(po_b, L.px + (long long)nn * NX, NX, lane);
                prefetch_rows(pn_b, L.px + (long long)nn * NX, NX, lane);
                prefetch_rows(po_b, L.pu + (long long)nn * NU, NU, lane);
                prefetch_rows(pn_b, L.pu + (long long)nn * NU, NU, lane);
                if (L.has_nl_rect) prefetch_rows(do_b, L.d7 + (long long)nn * S, S, lane);
                const int nc0 = T.child_first[nn], ncc = T.child_count[nn];   // children are consecutive: one run of edges
                prefetch_rows(do_b, L.d3 + (long long)(nc0 - 1) * NX, ncc * NX, lane);
                prefetch_rows(do_b, L.d4 + (long long)(nc0 - 1) * NU, ncc * NU, lane);
            }
        }
        // v = [x; u] old and new; dl = p+ - p, hat = 2 p+ - p
        double vo[S], vn[S];
#pragma unroll
        for (int k = 0; k < S; ++k) {
            const long long o = k < NX ? L.px + (long long)node * NX + k : L.pu + (long long)node * NU + (k - NX);
            vo[k] = po[o];
            vn[k] = pn[o];
        }
        double g1[S], ld[S];   // g1 = L*(d - d+), ld = L* d+ on the node's rows
        if (L.has_nl_rect) {   // d7 = [x; u], Rectangle.project (rectangle.py:29-35)
            const long long o7 = L.d7 + (long long)node * S;
            const double *lo = P.m.nl_lo + (long long)T.nl_rect_idx[node] * S, *hi = P.m.nl_hi + (long long)T.nl_rect_idx[node] * S;
            double d7o[S];
#pragma unroll
            for (int k = 0; k < S; ++k) d7o[k] = dO[o7 + k];
#pragma unroll
            for (int k = 0; k < S; ++k) {
                const double w = __dmul_rn(__fma_rn(alpha, 2.0 * vn[k] - vo[k], d7o[k]), ia);   // one rounding sequence: w - clip(w) is exactly 0 inside the box
                const double dn = alpha * (w - box_clip(w, __ldg(lo + k), __ldg(hi + k), &bad));
                dN[o7 + k] = dn;
                const double dd = d7o[k] - dn;
                upd(mx.m[2], dd * ia + (vn[k] - vo[k]));
                upd(mx.m[5], dd);
                g1[k] = dd;
                ld[k] = dn;
            }
        } else {
#pragma unroll
            for (int k = 0; k < S; ++k) g1[k] = ld[k] = 0.0;
        }
        for (int j0 = c0; j0 < c0 + cc; j0 += 2) {   // the edges into children j0 and j0 + 1 (cache.py:354-365)
            const bool two = j0 + 1 < c0 + cc;
            const int ja = j0, jb = two ? j0 + 1 : j0;   // a lone last child is processed as `a`; `b` shadows it with its stores off
            const long long ea = ja - 1, eb = jb - 1;
            double da[S], db[S];
#pragma unroll
            for (int k = 0; k < S; ++k) {
                da[k] = k < NX ? dO[L.d3 + ea * NX + k] : dO[L.d4 + ea * NU + (k - NX)];
                db[k] = k < NX ? dO[L.d3 + eb * NX + k] : dO[L.d4 + eb * NU + (k - NX)];
            }
            const double toa = po[L.ptau + ja], tna = pn[L.ptau + ja], d5a = dO[L.d5 + ea], d6a = dO[L.d6 + ea];
            const double tob = po[L.ptau + jb], tnb = pn[L.ptau + jb], d5b = dO[L.d5 + eb], d6b = dO[L.d6 + eb];
            const int cia = T.cost_idx[ja], cib = T.cost_idx[jb];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h == 1 && !two) break;
                const long long e = h ? eb : ea;
                const int j = h ? jb : ja, ci = h ? cib : cia;
                double *dd_o = h ? db : da;
                const double to = h ? tob : toa, tn = h ? tnb : tna, d5o = h ? d5b : d5a, d6o = h ? d6b : d6a;
                double ss = 0.0;
#pragma unroll
                for (int k = 0; k < S; ++k) {   // dd_o[k] becomes w[k]; the old value is recovered below as alpha w - alpha c (2 v+ - v)
                    const double c = k < NX ? __ldg(P.m.sq_d + ci * NX + k) : __ldg(P.m.sr_d + ci * NU + (k - NX));
                    const double w = (dd_o[k] + alpha * (c * (2.0 * vn[k] - vo[k]))) * ia;
                    ss = fma(w, w, ss);
                    // keep both: w in a register pair would double the footprint, so the old value is re-read (L1 hit) below
                    dd_o[k] = w;
                }
                const double half = 0.5 * (2.0 * tn - to);
                const double w5 = (d5o + alpha * half) * ia - 0.5, w6 = (d6o + alpha * half) * ia + 0.5;
                ss = fma(w5, w5, ss);
                const SocStep st = soc_step(ss, w6, alpha);
#pragma unroll
                for (int k = 0; k < S; ++k) {
                    const long long o = k < NX ? L.d3 + e * NX + k : L.d4 + e * NU + (k - NX);
                    const double c = k < NX ? __ldg(P.m.sq_d + ci * NX + k) : __ldg(P.m.sr_d + ci * NU + (k - NX));
                    const double dn = st.f * dd_o[k], dold = dO[o];
                    dN[o] = dn;
                    const double dd = dold - dn;
                    upd(mx.m[2], dd * ia + c * (vn[k] - vo[k]));
                    upd(mx.m[5], dd);
                    g1[k] = fma(c, dd, g1[k]);
                    ld[k] = fma(c, dn, ld[k]);
                }
                const double d5n = st.f * w5, d6n = st.dlast;
                dN[L.d5 + e] = d5n;
                dN[L.d6 + e] = d6n;
                const double dd5 = d5o - d5n, dd6 = d6o - d6n, ht = 0.5 * (tn - to);
                const double x25 = dd5 * ia + ht, x26 = dd6 * ia + ht;
                upd(mx.m[2], x25);
                upd(mx.m[2], x26);
                upd(mx.m[5], dd5);
                upd(mx.m[5], dd6);
                const double g1t = 0.5 * (dd5 + dd6), g2t = 0.5 * (x25 + x26);   // tau_j rows: L* = (d5 + d6) / 2
                const double x1t = (to - tn) * ia - g1t;
                upd(mx.m[1], x1t);
                upd(mx.m[0], x1t + g2t);
                upd(mx.m[4], tn - to);
                upd(mx.m[3], (tn - to) + g1t);
                pb[L.ptau + j] = tn - alpha * (0.5 * (d5n + d6n));
            }
        }
        const double *c2 = c2tab + (long long)node * S;
#pragma unroll
        for (int k = 0; k < S; ++k) {
            const double dl = vn[k] - vo[k];
            const double x1 = -dl * ia - g1[k];
            upd(mx.m[1], x1);
            upd(mx.m[0], x1 + (g1[k] * ia + __ldg(c2 + k) * dl));   // L* xi2 = L*(d - d+) / alpha + diag(L* L) (p+ - p)
            upd(mx.m[4], dl);
            upd(mx.m[3], dl + g1[k]);
            const long long o = k < NX ? L.px + (long long)node * NX + k : L.pu + (long long)node * NU + (k - NX);
            pb[o] = vn[k] - alpha * ld[k];
        }
    }
    flush_maxima(mx, slots, L.batch, panel, lane, warp, ctrl, bad);
}

// The risk block of a nonleaf node: d1 (dual of R+^{2c} x {0}: max(0, .) on the first 2c entries, identity on the last,
// risks.py:32-33), d2 (R+), the residual rows of y_i and s_i, and pbar of those rows.  MAXC: compile-time bound on the number of
// children -- all 3 (2 MAXC + 1) + 3 loads of a node are issued before the first use.
template <int MAXC>
__global__ void __launch_bounds__(kBpWarps * 32, RB_BP_RISK_MINB) k_bp_dual_risk(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl,
                                                               const double *p_old, const double *__restrict__ p_new,
                                                               const double *__restrict__ d_old, double *__restrict__ d_new,
                                                               double *__restrict__ slots, double *pbar) {   // pbar aliases p_old
    if (ctrl->done) return;
    const Layout &L = P.L;
    const Topo &T = P.t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, panel = blockIdx.y;
    const PanelPtr po = panel_view(p_old, L.np_pad, panel, lane), pn = panel_view(p_new, L.np_pad, panel, lane);
    const PanelPtr dO = panel_view(d_old, L.nd_pad, panel, lane), dN = panel_view(d_new, L.nd_pad, panel, lane);
    const PanelPtr pb = panel_view(pbar, L.np_pad, panel, lane);
    const double alpha = ctrl->alpha, ia = 1.0 / alpha;
    constexpr int YM = 2 * MAXC + 1;
    Maxima mx;
#pragma unroll
    for (int i = 0; i < 6; ++i) mx.m[i] = 0ull;
    for (int node = blockIdx.x * kBpWarps + warp; node < L.m; node += gridDim.x * kBpWarps) {
        const int c0 = T.child_first[node], cc = T.child_count[node], ny = 2 * cc + 1;
        const long long yo = T.yoff[node];
        double yold[YM], ynew[YM], d1o[YM], b[YM];
#pragma unroll
        for (int e = 0; e < YM; ++e) {
            const bool on = e < ny;
            yold[e] = on ? po[L.py + yo + e] : 0.0;
            ynew[e] = on ? pn[L.py + yo + e] : 0.0;
            d1o[e] = on ? dO[L.d1 + yo + e] : 0.0;
            b[e] = e < cc ? __ldg(T.cond_prob + c0 + e) : (e == ny - 1 ? 1.0 : 0.0);
        }
        const double so = po[L.ps + node], sn = pn[L.ps + node], d2o = dO[L.d2 + node];
        double bh = 0.0, bd = 0.0;   // b' (2 y+ - y),  b' (y+ - y),  b = [pi; 0; 1]
#pragma unroll
        for (int e = 0; e < YM; ++e) {
            bh = fma(b[e], 2.0 * ynew[e] - yold[e], bh);
            bd = fma(b[e], ynew[e] - yold[e], bd);
        }
        const double w2 = (d2o + alpha * ((2.0 * sn - so) - bh)) * ia;
        const double d2n = alpha * (w2 - fmax(w2, 0.0));
        dN[L.d2 + node] = d2n;
        const double dd2 = d2o - d2n, x22 = dd2 * ia + ((sn - so) - bd);
        upd(mx.m[2], x22);
        upd(mx.m[5], dd2);
#pragma unroll
        for (int e = 0; e < YM; ++e) {
            if (e < ny) {
                const double w1 = (d1o[e] + alpha * (2.0 * ynew[e] - yold[e])) * ia;
                const double d1n = e < ny - 1 ? alpha * (w1 - fmax(w1, 0.0)) : 0.0;
                dN[L.d1 + yo + e] = d1n;
                const double dd1 = d1o[e] - d1n, x21 = dd1 * ia + (ynew[e] - yold[e]);
                upd(mx.m[2], x21);
                upd(mx.m[5], dd1);
                const double g1y = dd1 - b[e] * dd2, g2y = x21 - b[e] * x22;
                const double x1 = (yold[e] - ynew[e]) * ia - g1y;
                upd(mx.m[1], x1);
                upd(mx.m[0], x1 + g2y);
                upd(mx.m[4], ynew[e] - yold[e]);
                upd(mx.m[3], (ynew[e] - yold[e]) + g1y);
                pb[L.py + yo + e] = ynew[e] - alpha * (d1n - b[e] * d2n);
            }
        }
        {   // s_i row: L* = d2
            const double x1 = (so - sn) * ia - dd2;
            upd(mx.m[1], x1);
            upd(mx.m[0], x1 + x22);
            upd(mx.m[4], sn - so);
            upd(mx.m[3], (sn - so) + dd2);
            pb[L.ps + node] = sn - alpha * d2n;
        }
    }
    flush_maxima(mx, slots, L.batch, panel, lane, warp, ctrl, 0);
}

template <int NX, int NU>
__global__ void __launch_bounds__(kBpWarps * 32, RB_BP_LEAF_MINB) k_bp_dual_leaf(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl,
                                                               const double *p_old, const double *__restrict__ p_new,
                                                               const double *__restrict__ d_old, double *__restrict__ d_new,
                                                               double *__restrict__ slots, double *pbar) {   // pbar aliases p_old
    if (ctrl->done) return;
    const Layout &L = P.L;
    const Topo &T = P.t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, panel = blockIdx.y;
    const PanelPtr po = panel_view(p_old, L.np_pad, panel, lane), pn = panel_view(p_new, L.np_pad, panel, lane);
    const PanelPtr dO = panel_view(d_old, L.nd_pad, panel, lane), dN = panel_view(d_new, L.nd_pad, panel, lane);
    const PanelPtr pb = panel_view(pbar, L.np_pad, panel, lane);
    const double alpha = ctrl->alpha, ia = 1.0 / alpha;
    Maxima mx;
#pragma unroll
    for (int i = 0; i < 6; ++i) mx.m[i] = 0ull;
    int bad = 0;
    const double *po_b = p_old + (long long)panel * L.np_pad * kPanel, *pn_b = p_new + (long long)panel * L.np_pad * kPanel;
    const double *do_b = d_old + (long long)panel * L.nd_pad * kPanel;
    for (int node = L.m + blockIdx.x * kBpWarps + warp; node < L.n; node += gridDim.x * kBpWarps) {
        const long long li = node - L.m;
        const int ci = T.leafcost_idx[li];
        {   // the next leaf of this warp into L2
            const int nn = node + gridDim.x * kBpWarps;
            if (nn < L.n) {
                prefetch_rows(po_b, L.px + (long long)nn * NX, NX, lane);
                prefetch_rows(pn_b, L.px + (long long)nn * NX, NX, lane);
                prefetch_rows(do_b, L.d11 + (long long)(nn - L.m) * NX, NX, lane);
                if (L.has_leaf_rect) prefetch_rows(do_b, L.d14 + (long long)(nn - L.m) * NX, NX, lane);
            }
        }
        double xo[NX], xn[NX], w11[NX], f[NX];
        double ss = 0.0;
#pragma unroll
        for (int k = 0; k < NX; ++k) {
            xo[k] = po[L.px + (long long)node * NX + k];
            xn[k] = pn[L.px + (long long)node * NX + k];
            f[k] = __ldg(P.m.sqf_d + ci * NX + k);
            w11[k] = (dO[L.d11 + li * NX + k] + alpha * (f[k] * (2.0 * xn[k] - xo[k]))) * ia;
            ss = fma(w11[k], w11[k], ss);
        }
        const double so = po[L.ps + node], sn = pn[L.ps + node];
        const double d12o = dO[L.d12 + li], d13o = dO[L.d13 + li];
        const double half = 0.5 * (2.0 * sn - so);
        const double w12 = (d12o + alpha * half) * ia - 0.5, w13 = (d13o + alpha * half) * ia + 0.5;
        ss = fma(w12, w12, ss);
        const SocStep st = soc_step(ss, w13, alpha);   // SOC on [d11; d12; d13], t = d13 (cache.py:373-385)
        const double *lo = L.has_leaf_rect ? P.m.leaf_lo + (long long)T.leaf_rect_idx[li] * NX : nullptr;
        const double *hi = L.has_leaf_rect ? P.m.leaf_hi + (long long)T.leaf_rect_idx[li] * NX : nullptr;
#pragma unroll
        for (int k = 0; k < NX; ++k) {
            const double d11o = dO[L.d11 + li * NX + k], d11n = st.f * w11[k];
            dN[L.d11 + li * NX + k] = d11n;
            const double dd = d11o - d11n, x2 = dd * ia + f[k] * (xn[k] - xo[k]);
            upd(mx.m[2], x2);
            upd(mx.m[5], dd);
            double g1 = f[k] * dd, g2 = f[k] * x2, ld = f[k] * d11n;
            if (L.has_leaf_rect) {   // d14 = x, Rectangle.project
                const double d14o = dO[L.d14 + li * NX + k];
                const double w = __dmul_rn(__fma_rn(alpha, 2.0 * xn[k] - xo[k], d14o), ia);   // (see the d7 block)
                const double d14n = alpha * (w - box_clip(w, __ldg(lo + k), __ldg(hi + k), &bad));
                dN[L.d14 + li * NX + k] = d14n;
                const double dd4 = d14o - d14n, x24 = dd4 * ia + (xn[k] - xo[k]);
                upd(mx.m[2], x24);
                upd(mx.m[5], dd4);
                g1 += dd4;
                g2 += x24;
                ld += d14n;
            }
            const double x1 = (xo[k] - xn[k]) * ia - g1;
            upd(mx.m[1], x1);
            upd(mx.m[0], x1 + g2);
            upd(mx.m[4], xn[k] - xo[k]);
            upd(mx.m[3], (xn[k] - xo[k]) + g1);
            pb[L.px + (long long)node * NX + k] = xn[k] - alpha * ld;
        }
        const double d12n = st.f * w12, d13n = st.dlast;
        dN[L.d12 + li] = d12n;
        dN[L.d13 + li] = d13n;
        const double dd12 = d12o - d12n, dd13 = d13o - d13n, hs = 0.5 * (sn - so);
        const double x212 = dd12 * ia + hs, x213 = dd13 * ia + hs;
        upd(mx.m[2], x212);
        upd(mx.m[2], x213);
        upd(mx.m[5], dd12);
        upd(mx.m[5], dd13);
        const double g1s = 0.5 * (dd12 + dd13), g2s = 0.5 * (x212 + x213);
        const double x1 = (so - sn) * ia - g1s;
        upd(mx.m[1], x1);
        upd(mx.m[0], x1 + g2s);
        upd(mx.m[4], sn - so);
        upd(mx.m[3], (sn - so) + g1s);
        pb[L.ps + node] = sn - alpha * (0.5 * (d12n + d13n));
    }
    flush_maxima(mx, slots, L.batch, panel, lane, warp, ctrl, bad);
}

}  // namespace

// ---- host side -------------------------------------------------------------------------------------------------------------------------
#define RB_BP_DIMS(X) X(2, 1) X(3, 2) X(4, 2) X(6, 3) X(8, 4) X(10, 5)

bool batch_panel_supported(int nx, int nu) {
#define RB_HAS(NX, NU) \
    if (nx == NX && nu == NU) return true;
    RB_BP_DIMS(RB_HAS)
#undef RB_HAS
    return false;
}

static inline int panels_of(int batch) { return (batch + kPanel - 1) / kPanel; }
size_t batch_panel_doubles(long long stride, int batch) { return (size_t)panels_of(batch) * (size_t)stride * kPanel; }

void launch_to_panels(cudaStream_t st, const double *src, double *dst, long long stride, int batch) {
    k_bp_to_panels<<<dim3((unsigned)((stride + 31) / 32), panels_of(batch)), 256, 0, st>>>(src, dst, stride, batch);
}
void launch_from_panels(cudaStream_t st, const double *src, double *dst, long long stride, int batch) {
    k_bp_from_panels<<<dim3((unsigned)((stride + 31) / 32), panels_of(batch)), 256, 0, st>>>(src, dst, stride, batch);
}

// node loops: enough CTAs per panel that a launch fills the machine a few times over, never more than the nodes need
static dim3 bp_grid(int nodes, int batch) {
    const int panels = panels_of(batch);
    int per_panel = (nodes + kBpWarps - 1) / kBpWarps;
    const int want = std::max(1, (148 * 16 + panels - 1) / panels);
    per_panel = std::max(1, std::min(per_panel, want));
    return dim3(per_panel, panels);
}

void launch_bp_primal(cudaStream_t st, const Params &P, const Ctrl *ctrl, const double *p_old, const double *d_old, double *p_out) {
#define RB_GO(NX, NU)                                                                                                    \
    if (P.L.nx == NX && P.L.nu == NU) {                                                                                  \
        k_bp_primal<NX, NU><<<bp_grid(P.L.n, P.L.batch), kBpWarps * 32, 0, st>>>(P, ctrl, p_old, d_old, p_out);          \
        return;                                                                                                          \
    }
    RB_BP_DIMS(RB_GO)
#undef RB_GO
}
void launch_bp_kproj(cudaStream_t st, const Params &P, const Ctrl *ctrl, double *prim, const double *x0, double *p_old) {
    k_bp_kproj<<<bp_grid(P.L.m, P.L.batch), kBpWarps * 32, 0, st>>>(P, ctrl, prim, x0, p_old);
}
void launch_bp_bwd(cudaStream_t st, const Params &P, const Ctrl *ctrl, const double *prim, double *q, double *r, int first, int count) {
#define RB_GO(NX, NU)                                                                                                    \
    if (P.L.nx == NX && P.L.nu == NU) {                                                                                  \
        k_bp_bwd<NX, NU><<<bp_grid(count, P.L.batch), kBpWarps * 32, 0, st>>>(P, ctrl, prim, q, r, first, count);        \
        return;                                                                                                          \
    }
    RB_BP_DIMS(RB_GO)
#undef RB_GO
}
void launch_bp_fwd(cudaStream_t st, const Params &P, const Ctrl *ctrl, double *prim, const double *r, const double *x0, int first,
                   int count) {
#define RB_GO(NX, NU)                                                                                                    \
    if (P.L.nx == NX && P.L.nu == NU) {                                                                                  \
        k_bp_fwd<NX, NU><<<bp_grid(count, P.L.batch), kBpWarps * 32, 0, st>>>(P, ctrl, prim, r, x0, first, count);       \
        return;                                                                                                          \
    }
    RB_BP_DIMS(RB_GO)
#undef RB_GO
}
void launch_bp_c2(cudaStream_t st, const Params &P, double *c2) { k_bp_c2<<<P.L.m, 128, 0, st>>>(P, c2); }
int batch_panel_max_children() { return kBpMaxChildren; }

void launch_bp_top(cudaStream_t st, const Params &P, const Ctrl *ctrl, double *prim, double *q, double *r, const double *x0,
                   const int *stage_off, int t_top) {
#define RB_GO(NX, NU)                                                                                                    \
    if (P.L.nx == NX && P.L.nu == NU) {                                                                                  \
        k_bp_top<NX, NU><<<panels_of(P.L.batch), kTopWarps * 32, 0, st>>>(P, ctrl, prim, q, r, x0, stage_off, t_top);     \
        return;                                                                                                          \
    }
    RB_BP_DIMS(RB_GO)
#undef RB_GO
}

void launch_bp_dual(cudaStream_t st, const Params &P, Ctrl *ctrl, const double *p_old, const double *p_new, const double *d_old,
                    double *d_new, double *slots, double *pbar, const double *c2, cudaEvent_t *evs) {
#define RB_GO(NX, NU)                                                                                                    \
    if (P.L.nx == NX && P.L.nu == NU) {                                                                                  \
        k_bp_dual_xu<NX, NU><<<bp_grid(P.L.m, P.L.batch), kBpWarps * 32, 0, st>>>(P, ctrl, p_old, p_new, d_old, d_new, slots, pbar, c2); \
        if (evs) cudaEventRecord(evs[0], st);                                                                            \
        k_bp_dual_risk<kBpMaxChildren><<<bp_grid(P.L.m, P.L.batch), kBpWarps * 32, 0, st>>>(P, ctrl, p_old, p_new, d_old, d_new, slots, pbar); \
        if (evs) cudaEventRecord(evs[1], st);                                                                            \
        k_bp_dual_leaf<NX, NU><<<bp_grid(P.L.n - P.L.m, P.L.batch), kBpWarps * 32, 0, st>>>(P, ctrl, p_old, p_new, d_old, d_new, slots, pbar); \
        return;                                                                                                          \
    }
    RB_BP_DIMS(RB_GO)
#undef RB_GO
}

}  // namespace rb
