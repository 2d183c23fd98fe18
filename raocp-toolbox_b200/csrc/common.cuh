// common.cuh -- device-side problem view, layout and warp helpers shared by all raocp_b200 kernels (sm_100a, FP64).
//
// Layout (DESIGN.md "Data layout in HBM"): one padded primal buffer and one padded dual buffer per iterate copy,
// segments in the compact order of include/raocp_b200.h, every segment start rounded up to 16 doubles (128 B),
// instance-major for batch > 1.  Node-major rows: x row i = nx consecutive doubles, so a warp that owns node i reads
// its rows with consecutive lanes (coalesced).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rb {

constexpr int kWarpsPerBlock = 4;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kMaxDim = 64;  // max(nx, nu) supported by the per-warp shared-memory scratch rows

struct Layout {
    int n, m, nleaf, nx, nu, nxu;
    int num_stages, batch;
    int has_nl_rect, has_leaf_rect;
    int ysz;
    // per-instance strides and segment offsets (doubles)
    long long np_pad, nd_pad;
    long long px, pu, py, ptau, ps;
    long long d1, d2, d3, d4, d5, d6, d7, d11, d12, d13, d14;
};

struct Topo {
    const int *parent, *child_first, *child_count, *yoff;
    const int *dyn_idx, *cost_idx, *leafcost_idx, *nl_rect_idx, *leaf_rect_idx, *cls;
    const double *cond_prob, *risk_alpha;
};

// Operator tables.  Every matrix-vector product assigns one OUTPUT row per lane and walks the reduction index, so a
// table is stored with the output index fastest (consecutive lanes read consecutive addresses; conflict-free when the
// table sits in shared memory).  Matrices applied to the same vector are concatenated so that one pass keeps all
// nx+nu lanes busy.
struct Tabs {
    const double *A, *B;       // [num_dyn][nx][nx], [num_dyn][nx][nu] row-major (offline factorisation)
    const double *ABcat;       // [num_dyn][nx][nx+nu]  row l = [A[l][:], B[l][:]]     ->  [A'q ; B'q]
    const double *ABcatT;      // [num_dyn][nx+nu][nx]  row l<nx: A[:][l], row nx+a: B[:][a]   ->  A x + B u
    const double *sqT, *srT;   // transposes of sqrtQ / sqrtR   [num_cost][..]
    const double *sqfT;        // transpose of sqrtQf           [num_leafcost][nx][nx]
    const double *sq_d, *sr_d, *sqf_d;   // their diagonals [..][nx] / [..][nu] (used when the *_diag flag is set)
    const double *nl_lo, *nl_hi, *leaf_lo, *leaf_hi;
    const double *K;           // [num_cls][nu][nx]                                    ->  K' r
    const double *KRcatT;      // [num_cls][nx+nu][nu]  row l<nx: K[:][l], row nx+b: R~^-1[:][b]   ->  K x + R~^-1 r
    int sq_diag, sr_diag, sqf_diag;  // 1 if every matrix of the table is diagonal (fast path)
    // MMA fragment images of ABcat / ABcatT ([num_dyn][F][32]) and K / KRcatT ([num_cls][F][32]) for chain_mma.cu; null
    // when no sweep level is tiled
    const double *fragAB, *fragABT, *fragK, *fragKR;
    // the same fragments ordered output-block-major (f = obi * nk + k): the four-warps-per-tile walkers fetch the block of one
    // warp as one contiguous run
    const double *fragAB4, *fragABT4, *fragK4, *fragKR4;
};

struct Params {
    Layout L;
    Topo t;
    Tabs m;
};

// ---- warp helpers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// One output row per lane: returns sum_l MT[l*rows + k] * v[l].  MT is the TRANSPOSE of the rows x cols matrix being
// applied (so lanes read consecutive addresses for every l; conflict-free when MT is in shared memory), v is a
// warp-private shared row.  Four independent accumulators / loads in flight per lane: the loop is latency-, not
// throughput-bound, and the trip count is a run-time value.
__device__ __forceinline__ double mv_row(const double *__restrict__ MT, const double *v, int rows, int cols, int k) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const double *col = MT + k;
    int l = 0;
    for (; l + 4 <= cols; l += 4) {
        const double m0 = col[(long long)l * rows], m1 = col[(long long)(l + 1) * rows];
        const double m2 = col[(long long)(l + 2) * rows], m3 = col[(long long)(l + 3) * rows];
        a0 = fma(m0, v[l], a0);
        a1 = fma(m1, v[l + 1], a1);
        a2 = fma(m2, v[l + 2], a2);
        a3 = fma(m3, v[l + 3], a3);
    }
    for (; l < cols; ++l) a0 = fma(col[(long long)l * rows], v[l], a0);
    return (a0 + a1) + (a2 + a3);
}
// the same matrix applied to two vectors at once (one pass over the matrix)
__device__ __forceinline__ void mv_row2(const double *__restrict__ MT, const double *v, const double *w, int rows,
                                        int cols, int k, double &rv, double &rw) {
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
    const double *col = MT + k;
    int l = 0;
    for (; l + 2 <= cols; l += 2) {
        const double m0 = col[(long long)l * rows], m1 = col[(long long)(l + 1) * rows];
        a0 = fma(m0, v[l], a0);
        b0 = fma(m0, w[l], b0);
        a1 = fma(m1, v[l + 1], a1);
        b1 = fma(m1, w[l + 1], b1);
    }
    for (; l < cols; ++l) {
        const double m0 = col[(long long)l * rows];
        a0 = fma(m0, v[l], a0);
        b0 = fma(m0, w[l], b0);
    }
    rv = a0 + a1;
    rw = b0 + b1;
}
// out[k] += scale * (M v)[k]   (k = lane, lane+32, ...; out is a warp-private shared row)
__device__ __forceinline__ void mv_acc(const double *__restrict__ MT, const double *v, int rows, int cols,
                                       double *out, double scale, int lane) {
    for (int k = lane; k < rows; k += 32) out[k] += scale * mv_row(MT, v, rows, cols, k);
}
// out[k] = (M v)[k]
__device__ __forceinline__ void mv_set(const double *__restrict__ MT, const double *v, int rows, int cols,
                                       double *out, int lane) {
    for (int k = lane; k < rows; k += 32) out[k] = mv_row(MT, v, rows, cols, k);
}

// ---- projections (device forms of reference cones.py / rectangle.py) -----------------------------------------------
// Rectangle._constrain, rectangle.py:50-59.  NaN cannot be constrained: *bad is set and NaN is returned.
__device__ __forceinline__ double box_clip(double v, double lo, double hi, int *bad) {
    if (lo <= v && v <= hi) return v;
    if (v <= lo) return lo;
    if (v >= hi) return hi;
    *bad = 1;
    return v;
}

// SecondOrderCone.project, cones.py:113-132, on a warp-private shared row w[0..dim-1] whose LAST entry is t.
// Returns the scaling pair (cz, ct): projection = (cz * w[0..dim-2], ct_is_value ? ct : w[dim-1]) encoded as
//   mode 0: projection = w (inside the cone);  mode 1: projection = 0 (inside the polar);  mode 2: boundary, with
//   t_new = (r+t)/2 and first part t_new * (w / r).
struct SocResult {
    int mode;
    double r, t_new;
};
__device__ __forceinline__ SocResult soc_classify(const double *w, int dim, int lane) {
    double ss = 0.0;
    for (int k = lane; k < dim - 1; k += 32) ss = fma(w[k], w[k], ss);
    ss = warp_sum(ss);
    SocResult res;
    res.r = sqrt(ss);
    const double t = w[dim - 1];
    if (res.r <= t) res.mode = 0;
    else if (res.r <= -t) res.mode = 1;
    else res.mode = 2;
    res.t_new = (res.r + t) / 2;
    return res;
}
__device__ __forceinline__ double soc_entry(const SocResult &s, double wk, bool is_last) {
    if (s.mode == 0) return wk;
    if (s.mode == 1) return 0.0;
    return is_last ? s.t_new : s.t_new * (wk / s.r);
}

// non-negative doubles order like their bit patterns, so an unsigned 64-bit atomicMax is an atomic max; NaN (larger
// than +inf as a bit pattern) sticks, which is how a non-finite residual reaches the host.
__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v) {
    atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

__device__ __forceinline__ double absmax_nan(double running, double v) {
    // fmax drops NaN; keep it instead so that a NaN iterate is reported
    const double a = fabs(v);
    return (a != a || running != running) ? __longlong_as_double(0x7ff8000000000000LL) : fmax(running, a);
}

}  // namespace rb
