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
// prefetched while the current step's MMAs issue.  All matrices come from lane-major fragment tables (k_frag_table).
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
                                       int kj, int ob, int lane) {
    const int l = 8 * kb + 2 * (lane & 3) + kj - l_off, o = 8 * ob + (lane >> 2) - o_off;
    return (l >= 0 && l < rows && o >= 0 && o < cols) ? tab[l * ld + o] : 0.0;
}

// The fragments are gathered ONCE per table (after the offline factorisation) into lane-major tables, so that a walker
// fetches its F fragment words of table entry e as one contiguous, unpredicated run: out[(e * 32 + lane) * F + f], with
// f = (2 * kbi + kj) * nob + obi for k-blocks kb0 + kbi and output blocks ob0 + obi.
__global__ void k_frag_table(const double *__restrict__ tab, long long entry_stride, int ld, int l_off, int rows, int o_off,
                             int cols, int kb0, int ob0, int nob, double *__restrict__ out) {
    const int e = blockIdx.x, f = blockIdx.y, F = gridDim.y, lane = threadIdx.x;
    const int obi = f % nob, kj = (f / nob) & 1, kbi = f / (2 * nob);
    out[((long long)e * 32 + lane) * F + f] =
        frag(tab + e * entry_stride, ld, l_off, rows, o_off, cols, kb0 + kbi, kj, ob0 + obi, lane);
}

template <int F>
__device__ __forceinline__ void ld_frags(double (&w)[F], const double *__restrict__ table, int entry, int lane) {
    const double2 *b = reinterpret_cast<const double2 *>(table + ((long long)entry * 32 + lane) * F);
#pragma unroll
    for (int f = 0; f < F; f += 2) {
        const double2 v = __ldg(b + f / 2);
        w[f] = v.x;
        w[f + 1] = v.y;
    }
}

// per-warp metadata in shared memory, copied from the host-built image of the tile (SweepLevel::tile_meta): node ids
// [depth][8] (padding columns shadow chain 0: loads only, stores are masked), the dynamics row and the class of every depth
// (the host only groups chains for which they coincide), 8 validity flags.  One global round trip, overlapped with the
// read of the loop's "done" flag; returns false if the loop has stopped.
struct TileMeta {
    const int *nodes, *dyns, *clss;
    int g, t;
    bool valid;
};
__device__ __forceinline__ int tile_meta_ints(int depth) { return depth * 10 + 8; }
__device__ __forceinline__ bool stage_meta(const SweepLevel &lv, const Ctrl *ctrl, int tile, int warp, int lane, int *smem,
                                           TileMeta &tm) {
    const int n = tile_meta_ints(lv.depth);
    int *dst = smem + warp * n;
    const int *src = lv.tile_meta + (long long)tile * n;
    int mine[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (lane + 32 * i < n) mine[i] = __ldg(src + lane + 32 * i);
    const int done = ctrl ? *reinterpret_cast<const volatile int *>(&ctrl->done) : 0;
    for (int i = lane + 128; i < n; i += 32) dst[i] = __ldg(src + i);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (lane + 32 * i < n) dst[lane + 32 * i] = mine[i];
    __syncwarp();
    tm.nodes = dst;
    tm.dyns = dst + lv.depth * 8;
    tm.clss = tm.dyns + lv.depth;
    tm.g = lane >> 2;
    tm.t = lane & 3;
    tm.valid = dst[lv.depth * 10 + tm.g] != 0;
    return done == 0;
}

// ---- asynchronous row / fragment staging -----------------------------------------------------------------------------------
// A step's inputs (xbar, ubar or r rows of the eight chains and the fragments of the step's class) are copied with
// cp.async into a per-warp ring of kStages shared-memory stages two steps ahead.  Every lane copies exactly the words it
// will read itself, so no warp barrier is needed -- only cp.async.wait_group -- and, unlike register prefetching, the
// copies hold no register scoreboard: the tensor-core instructions of the current step never wait for a copy of a later one.
constexpr int kStages = 3;
// words per lane of one stage, padded to an odd number of 16-byte units: conflict-free 128-bit shared-memory reads
__host__ __device__ constexpr int ring_lane_words(int words) { return (words / 2) % 2 == 0 ? words + 2 : words; }

template <int BYTES>
__device__ __forceinline__ void cp_async(double *smem_dst, const double *gsrc) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// state-sized row: the lane's slots 8b + 2t + {0,1} (b < QT) -> dst[2b + {0,1}]; slots >= NX are never written (zeroed once)
template <int NX, int NU>
__device__ __forceinline__ void cp_state(double *dst, const double *__restrict__ row, int t) {
    using D = ChainDims<NX, NU>;
#pragma unroll
    for (int b = 0; b < D::QT; ++b) {
        const int s0 = 8 * b + 2 * t;
        if constexpr (D::VEC) {
            if (s0 < NX) cp_async<16>(dst + 2 * b, row + s0);
        } else {
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (s0 + j < NX) cp_async<8>(dst + 2 * b + j, row + s0 + j);
        }
    }
}
// input-sized row: the lane's slots of block RT0 + i -> dst[2i + {0,1}]
template <int NX, int NU>
__device__ __forceinline__ void cp_input(double *dst, const double *__restrict__ row, int t) {
    using D = ChainDims<NX, NU>;
#pragma unroll
    for (int i = 0; i < D::RN; ++i) {
        const int a0 = 8 * (D::RT0 + i) + 2 * t - NX;
        if constexpr (D::VEC) {
            if (a0 >= 0 && a0 < NU) cp_async<16>(dst + 2 * i, row + a0);
        } else {
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (a0 + j >= 0 && a0 + j < NU) cp_async<8>(dst + 2 * i + j, row + a0 + j);
        }
    }
}
// the lane's F fragment words of one table entry (F is even: 16-byte copies)
template <int F>
__device__ __forceinline__ void cp_frags(double *dst, const double *__restrict__ table, int entry, int lane) {
    const double *src = table + ((long long)entry * 32 + lane) * F;
#pragma unroll
    for (int f = 0; f < F; f += 2) cp_async<16>(dst + f, src + f);
}
template <int N>
__device__ __forceinline__ void lds_vec(double (&v)[N], const double *src) {
    static_assert(N % 2 == 0, "pairs");
#pragma unroll
    for (int k = 0; k < N; k += 2) {
        const double2 w = *reinterpret_cast<const double2 *>(src + k);
        v[k] = w.x;
        v[k + 1] = w.y;
    }
}

// BIG: acc[o] += a(r) * W[r][o] over the ROWS rows of a lane's fragment block in shared memory (LEN words per row), the
// row r + 1 loaded while the MMAs of row r issue (ptxas otherwise puts every LDS.128 right in front of the two DMMAs that
// use it).  Measured neutral on cfg5 (45 / 50 us either way): ncu has the lone warp of an SM at ~27 cycles per DMMA with a
// third of its samples on the NOP ptxas puts between two DMMAs -- the walker is bound by the DMMA rate of a single warp.
template <int ROWS, int LEN, typename AF>
__device__ __forceinline__ void mma_rows(const double *w, AF a_of_row, double (&acc)[LEN][2]) {
    double wr[2][LEN];
    lds_vec(wr[0], w);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        if (r + 1 < ROWS) lds_vec(wr[(r + 1) & 1], w + (r + 1) * LEN);
        asm volatile("" ::: "memory");
        const double a = a_of_row(r);
#pragma unroll
        for (int o = 0; o < LEN; ++o) dmma(acc[o], a, wr[r & 1][o]);
    }
}

// ---- backward:  r = ubar - B'q_child,  q = A'q_child - xbar - K'r    -------------------------------
// BIG (nx + nu > 32): the fragments do not fit into registers.  One warp per CTA; the [A | B] fragments of the tile's
// dynamics row are copied once into shared memory (lane-major, every lane reads only its own words: no barrier), the class
// fragments are read straight out of the ring stage -- one 16-byte shared-memory load per two MMAs.
template <int NX, int NU, bool BIG>
__global__ void __launch_bounds__(BIG ? 32 : 128) k_chain_mma_bwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                      SweepLevel lv, const double *__restrict__ prim,
                                                      double *__restrict__ q, double *__restrict__ r) {
    using D = ChainDims<NX, NU>;
    extern __shared__ __align__(16) double mma_smem[];
    const Layout &L = P.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    const int tile = blockIdx.x * warps + warp;
    if (tile >= lv.num_tiles) return;
    constexpr int F1 = 2 * D::QT * D::NT, F2 = 2 * D::RN * D::QT;
    constexpr int kLaneWords = ring_lane_words(2 * D::QT + 2 * D::RN + F2);   // per lane and stage: xbar | ubar | K fragments
    constexpr int kAbWords = BIG ? ring_lane_words(F1) : 0;                   // per lane: the [A | B] fragments (BIG)
    static_assert(!BIG || (D::NT % 2 == 0 && D::QT % 2 == 0 && D::RN % 2 == 0), "BIG reads fragment pairs");
    double *ring = mma_smem + ((size_t)warp * kStages * 32 + lane) * kLaneWords;   // stage s: ring + s * 32 * kLaneWords
    TileMeta tm;
    if (!stage_meta(lv, ctrl, tile, warp, lane,
                    reinterpret_cast<int *>(mma_smem + (size_t)warps * kStages * 32 * kLaneWords + (size_t)warps * 32 * kAbWords), tm))
        return;
    const double *w1s = mma_smem + (size_t)warps * kStages * 32 * kLaneWords + ((size_t)warp * 32 + lane) * kAbWords;
    if constexpr (BIG) {   // the oldest commit group: complete before the first step's wait returns
        if (lv.depth > 1) cp_frags<F1>(const_cast<double *>(w1s), P.m.fragAB, tm.dyns[1], lane);
        cp_async_commit();
    }
    const int t = tm.t, g = tm.g;
    const double *X = prim + (long long)blockIdx.y * L.np_pad + L.px, *U = prim + (long long)blockIdx.y * L.np_pad + L.pu;
    double *Q = q + (long long)blockIdx.y * L.n * NX, *R = r + (long long)blockIdx.y * L.m * NU;

#pragma unroll
    for (int s = 0; s < kStages; ++s)
#pragma unroll
        for (int k = 0; k < 2 * D::QT + 2 * D::RN; ++k) ring[s * 32 * kLaneWords + k] = 0.0;   // padding slots stay zero
    auto prefetch = [&](int d) {   // one commit group per step, possibly empty
        if (d >= 0) {
            double *dst = ring + (d % kStages) * 32 * kLaneWords;
            const int node = tm.nodes[d * 8 + g], cls = tm.clss[d];
            cp_state<NX, NU>(dst, X + (long long)node * NX, t);
            if (cls >= 0) {
                cp_input<NX, NU>(dst + 2 * D::QT, U + (long long)node * NU, t);
                cp_frags<F2>(dst + 2 * D::QT + 2 * D::RN, P.m.fragK, cls, lane);
            }
        }
        cp_async_commit();
    };
    int d = lv.depth - 1;
    prefetch(d);
    prefetch(d - 1);
    double w1[BIG ? 2 : F1];   // [A | B] of the chain's dynamics row (one row per tile: build_chain_tiles), fragment (2 kb + j) * NT + ob
    if constexpr (!BIG) {
        if (lv.depth > 1) ld_frags<F1>(w1, P.m.fragAB, tm.dyns[1], lane);
    }
    double qs[D::QT][2];
#pragma unroll
    for (int b = 0; b < D::QT; ++b) qs[b][0] = qs[b][1] = 0.0;

    for (; d >= 0; --d) {
        prefetch(d - 2);
        cp_async_wait<2>();   // the copies of step d have landed
        const double *src = ring + (d % kStages) * 32 * kLaneWords;
        const int node = tm.nodes[d * 8 + g];
        double xb[2 * D::QT];
        lds_vec(xb, src);
        if (tm.clss[d] < 0) {   // leaf: q = -xbar
#pragma unroll
            for (int b = 0; b < D::QT; ++b) {
                qs[b][0] = -xb[2 * b];
                qs[b][1] = -xb[2 * b + 1];
            }
        } else {
            double ub[2 * D::RN], w2[BIG ? 2 : F2];
            lds_vec(ub, src + 2 * D::QT);
            const double *w2s = src + 2 * D::QT + 2 * D::RN;
            if constexpr (!BIG) lds_vec(w2, w2s);
            // E = [A'q ; B'q]
            double E[D::NT][2];
#pragma unroll
            for (int ob = 0; ob < D::NT; ++ob) E[ob][0] = E[ob][1] = 0.0;
            if constexpr (BIG) {
                mma_rows<2 * D::QT, D::NT>(w1s, [&](int r) { return qs[r >> 1][r & 1]; }, E);
            } else {
#pragma unroll
                for (int kb = 0; kb < D::QT; ++kb)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int ob = 0; ob < D::NT; ++ob) dmma(E[ob], qs[kb][j], w1[(2 * kb + j) * D::NT + ob]);
            }
            // r = ubar - B'q  on the input slots (nr = -r feeds the second product)
            double rr[D::RN][2], nr[D::RN][2];
#pragma unroll
            for (int i = 0; i < D::RN; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int a = 8 * (D::RT0 + i) + 2 * t + j - NX;
                    rr[i][j] = (a >= 0 && a < NU) ? ub[2 * i + j] - E[D::RT0 + i][j] : 0.0;
                    nr[i][j] = -rr[i][j];
                }
            if (tm.valid) st_input<NX, NU>(R + (long long)node * NU, t, rr);
            // q = A'q - xbar - K'r  on the state slots
#pragma unroll
            for (int b = 0; b < D::QT; ++b)
#pragma unroll
                for (int j = 0; j < 2; ++j) qs[b][j] = (8 * b + 2 * t + j < NX) ? E[b][j] - xb[2 * b + j] : 0.0;
            if constexpr (BIG) {
                mma_rows<2 * D::RN, D::QT>(w2s, [&](int r) { return nr[r >> 1][r & 1]; }, qs);
            } else {
#pragma unroll
                for (int i = 0; i < D::RN; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int ob = 0; ob < D::QT; ++ob) dmma(qs[ob], nr[i][j], w2[(2 * i + j) * D::QT + ob]);
            }
        }
        if (d == 0 && tm.valid) st_state<NX, NU>(Q + (long long)node * NX, t, qs);   // only the head's q leaves the chain
    }
}

// ---- forward:  u = K x + R~^-1 r,  x_child = A x + B u -------------------------------------------------------------------
template <int NX, int NU, bool BIG>
__global__ void __launch_bounds__(BIG ? 32 : 128) k_chain_mma_fwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                      SweepLevel lv, double *__restrict__ prim, const double *__restrict__ r,
                                                      int d_begin, int d_end) {
    // steps d_begin <= d < d_end of the walk (d = depth below the head of the chain): a launch starts from the x of depth
    // d_begin, which the level above (d_begin = 0) or the previous launch has written, so the walk can be cut into
    // pieces whose nodes' dual pass runs while the next piece walks on
    using D = ChainDims<NX, NU>;
    extern __shared__ __align__(16) double mma_smem[];
    const Layout &L = P.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    const int tile = blockIdx.x * warps + warp;
    if (tile >= lv.num_tiles) return;
    constexpr int F4 = 2 * D::NT * D::QT, F3 = 2 * D::NT * D::RN;
    constexpr int kLaneWords = ring_lane_words(2 * D::RN + F3);                // per lane and stage: r | [K R~^-1] fragments
    constexpr int kAbWords = BIG ? ring_lane_words(F4) : 0;                    // per lane: the [A ; B]' fragments (BIG)
    static_assert(!BIG || (D::NT % 2 == 0 && D::QT % 2 == 0 && D::RN % 2 == 0), "BIG reads fragment pairs");
    double *ring = mma_smem + ((size_t)warp * kStages * 32 + lane) * kLaneWords;
    TileMeta tm;
    if (!stage_meta(lv, ctrl, tile, warp, lane,
                    reinterpret_cast<int *>(mma_smem + (size_t)warps * kStages * 32 * kLaneWords + (size_t)warps * 32 * kAbWords), tm))
        return;
    const double *w4s = mma_smem + (size_t)warps * kStages * 32 * kLaneWords + ((size_t)warp * 32 + lane) * kAbWords;
    if constexpr (BIG) {
        if (lv.depth > 1) cp_frags<F4>(const_cast<double *>(w4s), P.m.fragABT, tm.dyns[1], lane);
        cp_async_commit();
    }
    const int t = tm.t, g = tm.g;
    double *X = prim + (long long)blockIdx.y * L.np_pad + L.px, *U = prim + (long long)blockIdx.y * L.np_pad + L.pu;
    const double *R = r + (long long)blockIdx.y * L.m * NU;

#pragma unroll
    for (int s = 0; s < kStages; ++s)
#pragma unroll
        for (int k = 0; k < 2 * D::RN; ++k) ring[s * 32 * kLaneWords + k] = 0.0;
    auto prefetch = [&](int d) {
        if (d < lv.depth) {
            const int cls = tm.clss[d];
            if (cls >= 0) {
                double *dst = ring + (d % kStages) * 32 * kLaneWords;
                cp_input<NX, NU>(dst, R + (long long)tm.nodes[d * 8 + g] * NU, t);
                cp_frags<F3>(dst + 2 * D::RN, P.m.fragKR, cls, lane);
            }
        }
        cp_async_commit();
    };
    prefetch(d_begin);
    prefetch(d_begin + 1);
    double w4[BIG ? 2 : F4];   // [A ; B]' of the chain's dynamics row, fragment (2 kb + j) * QT + ob
    if constexpr (!BIG) {
        if (lv.depth > 1) ld_frags<F4>(w4, P.m.fragABT, tm.dyns[1], lane);
    }
    double xs[D::QT][2];
    ld_state<NX, NU, false>(X + (long long)tm.nodes[d_begin * 8 + g] * NX, t, xs);   // written by the level above / the previous piece

    for (int d = d_begin; d < d_end && d + 1 < lv.depth && tm.clss[d] >= 0; ++d) {
        prefetch(d + 2);
        cp_async_wait<2>();
        const double *src = ring + (d % kStages) * 32 * kLaneWords;
        const int node = tm.nodes[d * 8 + g], child = tm.nodes[(d + 1) * 8 + g];
        double rr[2 * D::RN], w3[BIG ? 2 : F3];
        lds_vec(rr, src);
        const double *w3s = src + 2 * D::RN;
        if constexpr (!BIG) lds_vec(w3, w3s);
        // u = [K R~^-1] [x ; r]  (lands on the input slots)
        double ua[D::RN][2];
#pragma unroll
        for (int i = 0; i < D::RN; ++i) ua[i][0] = ua[i][1] = 0.0;
        if constexpr (BIG) {
            mma_rows<2 * D::NT, D::RN>(w3s, [&](int r) {
                const int kb = r >> 1, j = r & 1;
                const bool is_x = 8 * kb + 2 * t + j < NX;
                double v = 0.0;
                if (kb < D::QT) v = xs[kb < D::QT ? kb : 0][j];
                if (kb >= D::RT0 && !is_x) v = rr[2 * (kb >= D::RT0 ? kb - D::RT0 : 0) + j];
                return v;
            }, ua);
        } else {
#pragma unroll
            for (int kb = 0; kb < D::NT; ++kb)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const bool is_x = 8 * kb + 2 * t + j < NX;
                    double v = 0.0;
                    if (kb < D::QT) v = xs[kb < D::QT ? kb : 0][j];
                    if (kb >= D::RT0 && !is_x) v = rr[2 * (kb >= D::RT0 ? kb - D::RT0 : 0) + j];
#pragma unroll
                    for (int i = 0; i < D::RN; ++i) dmma(ua[i], v, w3[(2 * kb + j) * D::RN + i]);
                }
        }
        if (tm.valid) st_input<NX, NU>(U + (long long)node * NU, t, ua);
        // x_child = [A B] [x ; u]
        double xn[D::QT][2];
#pragma unroll
        for (int b = 0; b < D::QT; ++b) xn[b][0] = xn[b][1] = 0.0;
        if constexpr (BIG) {
            mma_rows<2 * D::NT, D::QT>(w4s, [&](int r) {
                const int kb = r >> 1, j = r & 1;
                const bool is_x = 8 * kb + 2 * t + j < NX;
                double v = 0.0;
                if (kb < D::QT) v = xs[kb < D::QT ? kb : 0][j];
                if (kb >= D::RT0 && !is_x) v = ua[kb >= D::RT0 ? kb - D::RT0 : 0][j];
                return v;
            }, xn);
        } else {
#pragma unroll
            for (int kb = 0; kb < D::NT; ++kb)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const bool is_x = 8 * kb + 2 * t + j < NX;
                    double v = 0.0;
                    if (kb < D::QT) v = xs[kb < D::QT ? kb : 0][j];
                    if (kb >= D::RT0 && !is_x) v = ua[kb >= D::RT0 ? kb - D::RT0 : 0][j];
#pragma unroll
                    for (int ob = 0; ob < D::QT; ++ob) dmma(xn[ob], v, w4[(2 * kb + j) * D::QT + ob]);
                }
        }
        if (tm.valid) st_state<NX, NU>(X + (long long)child * NX, t, xn);
#pragma unroll
        for (int b = 0; b < D::QT; ++b) {
            xs[b][0] = xn[b][0];
            xs[b][1] = xn[b][1];
        }
    }
}

}  // namespace

// ---- host side --------------------------------------------------------------------------------------------------------------
// (nx, nu, BIG): BIG = fragments in shared memory, one warp per CTA (nx + nu > 32)
#define RB_MMA_DIMS(X) X(2, 1, false) X(3, 2, false) X(4, 2, false) X(8, 4, false) X(10, 5, false) X(20, 10, false) X(64, 32, true)

bool chain_mma_supported(int nx, int nu) {
#define RB_HAS(NX, NU, BIG) \
    if (nx == NX && nu == NU) return true;
    RB_MMA_DIMS(RB_HAS)
#undef RB_HAS
    return false;
}
static bool chain_mma_big(int nx, int nu) {
#define RB_HAS(NX, NU, BIG) \
    if (nx == NX && nu == NU) return BIG;
    RB_MMA_DIMS(RB_HAS)
#undef RB_HAS
    return false;
}

// fragment tables of chain_mma (Tabs::fragAB, fragABT: per dynamics row; fragK, fragKR: per factorisation class)
void chain_mma_frag_counts(int nx, int nu, int *f_ab, int *f_abt, int *f_k, int *f_kr) {
    const int NT = (nx + nu + 7) / 8, QT = (nx + 7) / 8, RT0 = nx / 8, RN = NT - RT0;
    *f_ab = 2 * QT * NT;
    *f_abt = 2 * NT * QT;
    *f_k = 2 * RN * QT;
    *f_kr = 2 * NT * RN;
}
void launch_chain_mma_frags(cudaStream_t st, const Tabs &M, int nx, int nu, int num_dyn, int num_cls, bool dynamics,
                            bool classes) {
    const int S = nx + nu, NT = (S + 7) / 8, QT = (nx + 7) / 8, RT0 = nx / 8, RN = NT - RT0;
    if (dynamics) {
        // [A'q ; B'q]: rows = state slots of q, columns = all slots
        k_frag_table<<<dim3(num_dyn, 2 * QT * NT), 32, 0, st>>>(M.ABcat, (long long)nx * S, S, 0, nx, 0, S, 0, 0, NT,
                                                                const_cast<double *>(M.fragAB));
        // A x + B u: rows = all slots of [x ; u], columns = state slots
        k_frag_table<<<dim3(num_dyn, 2 * NT * QT), 32, 0, st>>>(M.ABcatT, (long long)S * nx, nx, 0, S, 0, nx, 0, 0, QT,
                                                                const_cast<double *>(M.fragABT));
    }
    if (classes) {
        // K'r: rows = input slots (r), columns = state slots
        k_frag_table<<<dim3(num_cls, 2 * RN * QT), 32, 0, st>>>(M.K, (long long)nu * nx, nx, nx, nu, 0, nx, RT0, 0, QT,
                                                                const_cast<double *>(M.fragK));
        // K x + R~^-1 r: rows = all slots of [x ; r], columns = input slots
        k_frag_table<<<dim3(num_cls, 2 * NT * RN), 32, 0, st>>>(M.KRcatT, (long long)S * nu, nu, 0, S, nx, nu, 0, RT0, RN,
                                                                const_cast<double *>(M.fragKR));
    }
}

static dim3 mma_grid(const SweepLevel &lv, int batch, bool big) {
    return big ? dim3(lv.num_tiles, batch) : dim3((lv.num_tiles + 3) / 4, batch);
}
// dynamic shared memory of a CTA (4 warps, or 1 if BIG): the cp.async rings (+ the dynamics fragments if BIG) + the tile metadata
size_t chain_mma_smem_bytes(int nx, int nu, int depth, bool backward) {
    const int NT = (nx + nu + 7) / 8, QT = (nx + 7) / 8, RT0 = nx / 8, RN = NT - RT0;
    const int lane_words = ring_lane_words(backward ? 2 * QT + 2 * RN + 2 * RN * QT : 2 * RN + 2 * NT * RN);
    const bool big = chain_mma_big(nx, nu);
    const int warps = big ? 1 : 4, ab_words = big ? ring_lane_words(2 * QT * NT) : 0;
    return (size_t)warps * 32 * (kStages * lane_words + ab_words) * sizeof(double) + (size_t)warps * (depth * 10 + 8) * sizeof(int);
}
cudaError_t chain_mma_set_smem(int bytes) {
    cudaError_t e = cudaSuccess;
#define RB_SET(NX, NU, BIG)                                                                                                       \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_chain_mma_bwd<NX, NU, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_chain_mma_fwd<NX, NU, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    RB_MMA_DIMS(RB_SET)
#undef RB_SET
    return e;
}

void launch_chain_mma_bwd(cudaStream_t st, const Params &P, const Ctrl *ctrl, const SweepLevel &lv, const double *prim,
                          double *q, double *r) {
#define RB_GO(NX, NU, BIG)                                                                                             \
    if (P.L.nx == NX && P.L.nu == NU) {                                                                           \
        k_chain_mma_bwd<NX, NU, BIG><<<mma_grid(lv, P.L.batch, BIG), BIG ? 32 : 128, chain_mma_smem_bytes(NX, NU, lv.depth, true), st>>>(P, ctrl, lv, prim, q, r);     \
        return;                                                                                                   \
    }
    RB_MMA_DIMS(RB_GO)
#undef RB_GO
}

void launch_chain_mma_fwd(cudaStream_t st, const Params &P, const Ctrl *ctrl, const SweepLevel &lv, double *prim,
                          const double *r, int d_begin, int d_end) {
    if (d_end < 0) d_end = lv.depth;
#define RB_GO(NX, NU, BIG)                                                                                             \
    if (P.L.nx == NX && P.L.nu == NU) {                                                                           \
        k_chain_mma_fwd<NX, NU, BIG><<<mma_grid(lv, P.L.batch, BIG), BIG ? 32 : 128, chain_mma_smem_bytes(NX, NU, lv.depth, false), st>>>(P, ctrl, lv, prim, r, d_begin, d_end);        \
        return;                                                                                                   \
    }
    RB_MMA_DIMS(RB_GO)
#undef RB_GO
}

}  // namespace rb
