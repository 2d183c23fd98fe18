// chain_mma.cu -- the chain part of the DP sweeps (reference cache.py:259-288) on the FP64 tensor cores.
//
// Below the stopping time of a Markov scenario tree every subtree is a chain, and chains that went through the same
// modes share every matrix of the recursion (dynamics row and factorisation class per stage).  Eight such chains form a
// TILE, one warp walks a tile, and each step of the recursion is a handful of m8n8k4 FP64 MMAs
//
//      D (8 chains x 8 outputs)  +=  V (8 chains x 4 reduction slots)  *  W (4 slots x 8 outputs)
//
// with the chains as the M dimension.  The state vector of a chain lives in the ACCUMULATOR layout of the MMA for the
// whole walk (lane (g, t) = (lane / 4, lane % 4) holds slots 8v + 2t, 8v + 2t + 1 of chain g for every 8-wide block v);
// because the order of a reduction is free, the k-blocks of the next product are chosen as kappa = (v, j) -> slots
// {8v + 2t + j : t = 0..3}, i.e. exactly what lane t already holds: no shuffle, no shared memory between the steps.  The
// matrix fragments are gathered from the operator tables in that permuted order; the dynamics fragments stay in
// registers for the whole chain, the class-indexed ones (K, [K R~^-1]) and the xbar / ubar / r rows of the next step are
// prefetched while the current step's MMAs issue.
//
// Slots: 0..nx-1 = the state-sized part (q or x), nx..nx+nu-1 = the input-sized part (r or u); nx + nu <= 32.
// Measured on B200 (profiles/microbench/fp64_pipes.cu): DMMA m8n8k4 issues every 16 cycles from a single warp per SM
// sub-partition at the full 37 TFLOP/s, dependent latency 26 cycles -- one warp per sub-partition saturates the pipe, so
// a step costs (#MMA x 16) cycles instead of the ~2000 cycles of the one-warp-per-chain walker.
#include "kernels.cuh"

namespace rb {

namespace {

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1])
                 : "d"(a), "d"(b));
}

template <int NX, int NU>
struct ChainDims {
    static constexpr int S = NX + NU;
    static constexpr int NT = (S + 7) / 8;     // 8-wide slot blocks
    static constexpr int QT = (NX + 7) / 8;    // blocks that contain state slots
    static constexpr int RT0 = NX / 8;         // first block that contains input slots
    static constexpr int RN = NT - RT0;        // blocks that contain input slots
    static constexpr bool VEC = (NX % 2 == 0) && (NU % 2 == 0);   // slot pairs never straddle a segment: 16-byte accesses
};

// state-sized row (length NX) <-> accumulator layout
template <int NX, int NU, bool LDG>
__device__ __forceinline__ void ld_state(const double *__restrict__ row, int t, double (&v)[ChainDims<NX, NU>::QT][2]) {
    using D = ChainDims<NX, NU>;
#pragma unroll
    for (int b = 0; b < D::QT; ++b) {
        const int s0 = 8 * b + 2 * t;
        if constexpr (D::VEC) {
            double2 w = make_double2(0.0, 0.0);
            if (s0 < NX) w = LDG ? __ldg(reinterpret_cast<const double2 *>(row + s0)) : *reinterpret_cast<const double2 *>(row + s0);
            v[b][0] = w.x;
            v[b][1] = w.y;
        } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) v[b][j] = s0 + j < NX ? (LDG ? __ldg(row + s0 + j) : row[s0 + j]) : 0.0;
        }
    }
}
template <int NX, int NU>
__device__ __forceinline__ void st_state(double *__restrict__ row, int t, const double (&v)[ChainDims<NX, NU>::QT][2]) {
    using D = ChainDims<NX, NU>;
#pragma unroll
    for (int b = 0; b < D::QT; ++b) {
        const int s0 = 8 * b + 2 * t;
        if constexpr (D::VEC) {
            if (s0 < NX) *reinterpret_cast<double2 *>(row + s0) = make_double2(v[b][0], v[b][1]);
        } else {
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (s0 + j < NX) row[s0 + j] = v[b][j];
        }
    }
}
// input-sized row (length NU) <-> the slots nx.. of the accumulator layout (block RT0 + i)
template <int NX, int NU>
__device__ __forceinline__ void ld_input(const double *__restrict__ row, int t, double (&v)[ChainDims<NX, NU>::RN][2]) {
    using D = ChainDims<NX, NU>;
#pragma unroll
    for (int i = 0; i < D::RN; ++i) {
        const int a0 = 8 * (D::RT0 + i) + 2 * t - NX;
        if constexpr (D::VEC) {
            double2 w = make_double2(0.0, 0.0);
            if (a0 >= 0 && a0 < NU) w = __ldg(reinterpret_cast<const double2 *>(row + a0));
            v[i][0] = w.x;
            v[i][1] = w.y;
        } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) v[i][j] = (a0 + j >= 0 && a0 + j < NU) ? __ldg(row + a0 + j) : 0.0;
        }
    }
}
template <int NX, int NU>
__device__ __forceinline__ void st_input(double *__restrict__ row, int t, const double (&v)[ChainDims<NX, NU>::RN][2]) {
    using D = ChainDims<NX, NU>;
#pragma unroll
    for (int i = 0; i < D::RN; ++i) {
        const int a0 = 8 * (D::RT0 + i) + 2 * t - NX;
        if constexpr (D::VEC) {
            if (a0 >= 0 && a0 < NU) *reinterpret_cast<double2 *>(row + a0) = make_double2(v[i][0], v[i][1]);
        } else {
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (a0 + j >= 0 && a0 + j < NU) row[a0 + j] = v[i][j];
        }
    }
}

// Fragment of a table for the B operand: lane (tk, n) = (lane % 4, lane / 4) supplies W[slot 8*kb + 2*tk + kj][slot 8*ob + n]
// where W[l][o] = tab[(l - l_off) * ld + (o - o_off)] for l_off <= l < l_off + rows and o_off <= o < o_off + cols, else 0.
__device__ __forceinline__ double frag(const double *__restrict__ tab, int ld, int l_off, int rows, int o_off, int cols, int kb,
                                       int kj, int ob, int lane, double sign) {
    const int l = 8 * kb + 2 * (lane & 3) + kj - l_off, o = 8 * ob + (lane >> 2) - o_off;
    return (l >= 0 && l < rows && o >= 0 && o < cols) ? sign * __ldg(tab + l * ld + o) : 0.0;
}

// per-warp metadata in shared memory: node ids [depth][8], then dynamics row and class of every depth (of chain 0 of
// the tile; the host only groups chains for which they coincide)
struct TileMeta {
    const int *nodes, *dyns, *clss;
    int g, t;
    bool valid;
};
__device__ __forceinline__ TileMeta stage_meta(const Layout &L, const Topo &T, const SweepLevel &lv, int tile, int warp, int lane,
                                               int *smem) {
    int *nodes = smem + warp * (lv.depth * 10), *dyns = nodes + lv.depth * 8, *clss = dyns + lv.depth;
    const int g = lane >> 2, t = lane & 3;
    const int own = lv.tiles[tile * 8 + g];
    const int c = own >= 0 ? own : lv.tiles[tile * 8];   // padding columns shadow chain 0 (loads only, stores are masked)
    for (int d = t; d < lv.depth; d += 4) nodes[d * 8 + g] = lv.lo[(long long)c * lv.depth + d];
    __syncwarp();
    for (int d = lane; d < lv.depth; d += 32) {
        const int n0 = nodes[d * 8];
        dyns[d] = T.dyn_idx[n0];
        clss[d] = n0 < L.m ? T.cls[n0] : -1;
    }
    __syncwarp();
    return TileMeta{nodes, dyns, clss, g, t, own >= 0};
}

// ---- backward:  r = ubar - B'q_child,  q = A'q_child - xbar - K'r    -------------------------------
template <int NX, int NU>
__global__ void __launch_bounds__(128) k_chain_mma_bwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                      SweepLevel lv, const double *__restrict__ prim,
                                                      double *__restrict__ q, double *__restrict__ r) {
    using D = ChainDims<NX, NU>;
    if (ctrl && ctrl->done) return;
    extern __shared__ int meta_smem[];
    const Layout &L = P.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (blockDim.x >> 5) + warp;
    if (tile >= lv.num_tiles) return;
    const TileMeta tm = stage_meta(L, P.t, lv, tile, warp, lane, meta_smem);
    const int t = tm.t, g = tm.g;
    const double *X = prim + (long long)blockIdx.y * L.np_pad + L.px, *U = prim + (long long)blockIdx.y * L.np_pad + L.pu;
    double *Q = q + (long long)blockIdx.y * L.n * NX, *R = r + (long long)blockIdx.y * L.m * NU;

    double w1[2 * D::QT][D::NT];                             // [A | B] of the chain's dynamics row
    double w2[2 * D::RN][D::QT], w2n[2 * D::RN][D::QT];      // -K of the current / next class
    double qs[D::QT][2], xb[D::QT][2], xbn[D::QT][2], ub[D::RN][2], ubn[D::RN][2];
    int dyn_loaded = -1;
    auto load_w2 = [&](double (&w)[2 * D::RN][D::QT], int cls) {
        const double *K = P.m.K + (long long)cls * NU * NX;
#pragma unroll
        for (int i = 0; i < D::RN; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int ob = 0; ob < D::QT; ++ob) w[2 * i + j][ob] = frag(K, NX, NX, NU, 0, NX, D::RT0 + i, j, ob, lane, -1.0);
    };
#pragma unroll
    for (int b = 0; b < D::QT; ++b) qs[b][0] = qs[b][1] = 0.0;
#pragma unroll
    for (int i = 0; i < D::RN; ++i) ub[i][0] = ub[i][1] = ubn[i][0] = ubn[i][1] = 0.0;

    int d = lv.depth - 1;
    int node = tm.nodes[d * 8 + g];
    ld_state<NX, NU, true>(X + (long long)node * NX, t, xb);
    if (tm.clss[d] >= 0) {
        ld_input<NX, NU>(U + (long long)node * NU, t, ub);
        load_w2(w2, tm.clss[d]);
    }
    for (; d >= 0; --d) {
        const int cls = tm.clss[d];
        int next = node;
        if (d > 0) {   // prefetch the rows and the class fragments of the next step
            next = tm.nodes[(d - 1) * 8 + g];
            ld_state<NX, NU, true>(X + (long long)next * NX, t, xbn);
            const int cn = tm.clss[d - 1];
            if (cn >= 0) {
                ld_input<NX, NU>(U + (long long)next * NU, t, ubn);
                load_w2(w2n, cn);
            }
        }
        if (cls < 0) {   // leaf: q = -xbar
#pragma unroll
            for (int b = 0; b < D::QT; ++b) {
                qs[b][0] = -xb[b][0];
                qs[b][1] = -xb[b][1];
            }
        } else {
            const int dyn = d + 1 < lv.depth ? tm.dyns[d + 1] : 0;
            if (dyn != dyn_loaded) {
                const double *C = P.m.ABcat + (long long)dyn * NX * D::S;
#pragma unroll
                for (int kb = 0; kb < D::QT; ++kb)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int ob = 0; ob < D::NT; ++ob) w1[2 * kb + j][ob] = frag(C, D::S, 0, NX, 0, D::S, kb, j, ob, lane, 1.0);
                dyn_loaded = dyn;
            }
            // E = [A'q ; B'q]
            double E[D::NT][2];
#pragma unroll
            for (int ob = 0; ob < D::NT; ++ob) E[ob][0] = E[ob][1] = 0.0;
#pragma unroll
            for (int kb = 0; kb < D::QT; ++kb)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int ob = 0; ob < D::NT; ++ob) dmma(E[ob], qs[kb][j], w1[2 * kb + j][ob]);
            // r = ubar - B'q  on the input slots
            double rr[D::RN][2];
#pragma unroll
            for (int i = 0; i < D::RN; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int a = 8 * (D::RT0 + i) + 2 * t + j - NX;
                    rr[i][j] = (a >= 0 && a < NU) ? ub[i][j] - E[D::RT0 + i][j] : 0.0;
                }
            if (tm.valid) st_input<NX, NU>(R + (long long)node * NU, t, rr);
            // q = A'q - xbar - K'r  on the state slots
#pragma unroll
            for (int b = 0; b < D::QT; ++b)
#pragma unroll
                for (int j = 0; j < 2; ++j) qs[b][j] = (8 * b + 2 * t + j < NX) ? E[b][j] - xb[b][j] : 0.0;
#pragma unroll
            for (int i = 0; i < D::RN; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int ob = 0; ob < D::QT; ++ob) dmma(qs[ob], rr[i][j], w2[2 * i + j][ob]);
        }
        if (d == 0 && tm.valid) st_state<NX, NU>(Q + (long long)node * NX, t, qs);   // only the head's q leaves the chain
        node = next;
#pragma unroll
        for (int b = 0; b < D::QT; ++b) {
            xb[b][0] = xbn[b][0];
            xb[b][1] = xbn[b][1];
        }
#pragma unroll
        for (int i = 0; i < D::RN; ++i) {
            ub[i][0] = ubn[i][0];
            ub[i][1] = ubn[i][1];
        }
#pragma unroll
        for (int k = 0; k < 2 * D::RN; ++k)
#pragma unroll
            for (int ob = 0; ob < D::QT; ++ob) w2[k][ob] = w2n[k][ob];
    }
}

// ---- forward:  u = K x + R~^-1 r,  x_child = A x + B u -------------------------------------------------------------------
template <int NX, int NU>
__global__ void __launch_bounds__(128) k_chain_mma_fwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                      SweepLevel lv, double *__restrict__ prim, const double *__restrict__ r) {
    using D = ChainDims<NX, NU>;
    if (ctrl && ctrl->done) return;
    extern __shared__ int meta_smem[];
    const Layout &L = P.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (blockDim.x >> 5) + warp;
    if (tile >= lv.num_tiles) return;
    const TileMeta tm = stage_meta(L, P.t, lv, tile, warp, lane, meta_smem);
    const int t = tm.t, g = tm.g;
    double *X = prim + (long long)blockIdx.y * L.np_pad + L.px, *U = prim + (long long)blockIdx.y * L.np_pad + L.pu;
    const double *R = r + (long long)blockIdx.y * L.m * NU;

    double w4[2 * D::NT][D::QT];                             // [A ; B]' of the chain's dynamics row
    double w3[2 * D::NT][D::RN], w3n[2 * D::NT][D::RN];      // [K R~^-1] of the current / next class
    double xs[D::QT][2], rr[D::RN][2], rrn[D::RN][2];
    int dyn_loaded = -1;
    auto load_w3 = [&](double (&w)[2 * D::NT][D::RN], int cls) {
        const double *KR = P.m.KRcatT + (long long)cls * D::S * NU;
#pragma unroll
        for (int kb = 0; kb < D::NT; ++kb)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int i = 0; i < D::RN; ++i) w[2 * kb + j][i] = frag(KR, NU, 0, D::S, NX, NU, kb, j, D::RT0 + i, lane, 1.0);
    };
#pragma unroll
    for (int i = 0; i < D::RN; ++i) rr[i][0] = rr[i][1] = rrn[i][0] = rrn[i][1] = 0.0;
    int node = tm.nodes[g];
    ld_state<NX, NU, false>(X + (long long)node * NX, t, xs);   // x of the head was written by the level above
    if (tm.clss[0] >= 0) {
        ld_input<NX, NU>(R + (long long)node * NU, t, rr);
        load_w3(w3, tm.clss[0]);
    }
    for (int d = 0; d + 1 < lv.depth; ++d) {
        if (tm.clss[d] < 0) break;
        const int child = tm.nodes[(d + 1) * 8 + g], cn = tm.clss[d + 1];
        if (cn >= 0) {
            ld_input<NX, NU>(R + (long long)child * NU, t, rrn);
            load_w3(w3n, cn);
        }
        // u = [K R~^-1] [x ; r]  (lands on the input slots)
        double ua[D::RN][2];
#pragma unroll
        for (int i = 0; i < D::RN; ++i) ua[i][0] = ua[i][1] = 0.0;
#pragma unroll
        for (int kb = 0; kb < D::NT; ++kb)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const bool is_x = 8 * kb + 2 * t + j < NX;
                double v = 0.0;
                if (kb < D::QT) v = xs[kb < D::QT ? kb : 0][j];
                if (kb >= D::RT0 && !is_x) v = rr[kb >= D::RT0 ? kb - D::RT0 : 0][j];
#pragma unroll
                for (int i = 0; i < D::RN; ++i) dmma(ua[i], v, w3[2 * kb + j][i]);
            }
        if (tm.valid) st_input<NX, NU>(U + (long long)node * NU, t, ua);
        // x_child = [A B] [x ; u]
        const int dyn = tm.dyns[d + 1];
        if (dyn != dyn_loaded) {
            const double *C = P.m.ABcatT + (long long)dyn * D::S * NX;
#pragma unroll
            for (int kb = 0; kb < D::NT; ++kb)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int ob = 0; ob < D::QT; ++ob) w4[2 * kb + j][ob] = frag(C, NX, 0, D::S, 0, NX, kb, j, ob, lane, 1.0);
            dyn_loaded = dyn;
        }
        double xn[D::QT][2];
#pragma unroll
        for (int b = 0; b < D::QT; ++b) xn[b][0] = xn[b][1] = 0.0;
#pragma unroll
        for (int kb = 0; kb < D::NT; ++kb)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const bool is_x = 8 * kb + 2 * t + j < NX;
                double v = 0.0;
                if (kb < D::QT) v = xs[kb < D::QT ? kb : 0][j];
                if (kb >= D::RT0 && !is_x) v = ua[kb >= D::RT0 ? kb - D::RT0 : 0][j];
#pragma unroll
                for (int ob = 0; ob < D::QT; ++ob) dmma(xn[ob], v, w4[2 * kb + j][ob]);
            }
        if (tm.valid) st_state<NX, NU>(X + (long long)child * NX, t, xn);
        node = child;
#pragma unroll
        for (int b = 0; b < D::QT; ++b) {
            xs[b][0] = xn[b][0];
            xs[b][1] = xn[b][1];
        }
#pragma unroll
        for (int i = 0; i < D::RN; ++i) {
            rr[i][0] = rrn[i][0];
            rr[i][1] = rrn[i][1];
        }
#pragma unroll
        for (int k = 0; k < 2 * D::NT; ++k)
#pragma unroll
            for (int i = 0; i < D::RN; ++i) w3[k][i] = w3n[k][i];
    }
}

}  // namespace

// ---- host side --------------------------------------------------------------------------------------------------------------
#define RB_MMA_DIMS(X) X(2, 1) X(3, 2) X(4, 2) X(8, 4) X(10, 5) X(20, 10)

bool chain_mma_supported(int nx, int nu) {
#define RB_HAS(NX, NU) \
    if (nx == NX && nu == NU) return true;
    RB_MMA_DIMS(RB_HAS)
#undef RB_HAS
    return false;
}

static dim3 mma_grid(const SweepLevel &lv, int batch) { return dim3((lv.num_tiles + 3) / 4, batch); }
static size_t mma_smem(const SweepLevel &lv) { return (size_t)4 * lv.depth * 10 * sizeof(int); }

void launch_chain_mma_bwd(cudaStream_t st, const Params &P, const Ctrl *ctrl, const SweepLevel &lv, const double *prim,
                          double *q, double *r) {
#define RB_GO(NX, NU)                                                                                             \
    if (P.L.nx == NX && P.L.nu == NU) {                                                                           \
        k_chain_mma_bwd<NX, NU><<<mma_grid(lv, P.L.batch), 128, mma_smem(lv), st>>>(P, ctrl, lv, prim, q, r);     \
        return;                                                                                                   \
    }
    RB_MMA_DIMS(RB_GO)
#undef RB_GO
}

void launch_chain_mma_fwd(cudaStream_t st, const Params &P, const Ctrl *ctrl, const SweepLevel &lv, double *prim,
                          const double *r) {
#define RB_GO(NX, NU)                                                                                             \
    if (P.L.nx == NX && P.L.nu == NU) {                                                                           \
        k_chain_mma_fwd<NX, NU><<<mma_grid(lv, P.L.batch), 128, mma_smem(lv), st>>>(P, ctrl, lv, prim, r);        \
        return;                                                                                                   \
    }
    RB_MMA_DIMS(RB_GO)
#undef RB_GO
}

}  // namespace rb
