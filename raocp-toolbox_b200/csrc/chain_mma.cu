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
#include <algorithm>
#ifdef RB_TRACE
#include <cstdio>
#endif

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
template <int NX, int NU>
__device__ __forceinline__ void ld_state_cg(const double *row, int t, double (&v)[ChainDims<NX, NU>::QT][2]) {
    using D = ChainDims<NX, NU>;
#pragma unroll
    for (int b = 0; b < D::QT; ++b)
#pragma unroll
        for (int j = 0; j < 2; ++j) v[b][j] = 8 * b + 2 * t + j < NX ? __ldcg(row + 8 * b + 2 * t + j) : 0.0;
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

// The fragments are gathered ONCE per table (after the offline factorisation) into PAIR-MAJOR tables: fragment word f of lane
// `lane` of table entry e sits at out[((e * F/2 + f/2) * 32 + lane) * 2 + (f & 1)], with f = (2 * kbi + kj) * nob + obi for
// k-blocks kb0 + kbi and output blocks ob0 + obi.  A warp's 16-byte copy of one pair is then ONE contiguous 512-byte run (four
// 128-byte lines) -- with the lane-major layout of round 1 (out[(e * 32 + lane) * F + f]) every lane of every copy touched a
// line of its own, 32 L1 tag look-ups per instruction, which is what the walkers were actually bound by (ncu r2c: shared-memory
// loads at ~400 cycles behind a queue of uncoalesced LDGSTS).
// ob_major: the word goes to position g = obi * (F / nob) + (2 * kbi + kj) instead of f (tables of the four-warp walkers).
__global__ void k_frag_table(const double *__restrict__ tab, long long entry_stride, int ld, int l_off, int rows, int o_off,
                             int cols, int kb0, int ob0, int nob, double *__restrict__ out, int ob_major) {
    const int e = blockIdx.x, f = blockIdx.y, F = gridDim.y, lane = threadIdx.x;
    const int obi = f % nob, kj = (f / nob) & 1, kbi = f / (2 * nob);
    const int g = ob_major ? obi * (F / nob) + 2 * kbi + kj : f;
    out[(((long long)e * (F / 2) + g / 2) * 32 + lane) * 2 + (g & 1)] =
        frag(tab + e * entry_stride, ld, l_off, rows, o_off, cols, kb0 + kbi, kj, ob0 + obi, lane);
}

// address of fragment word f of `lane` in a pair-major table of F words per lane and entry
template <int F>
__device__ __forceinline__ const double *frag_addr(const double *__restrict__ table, int entry, int lane, int f) {
    return table + (((long long)entry * (F / 2) + f / 2) * 32 + lane) * 2 + (f & 1);
}
template <int F>
__device__ __forceinline__ void ld_frags(double (&w)[F], const double *__restrict__ table, int entry, int lane) {
    static_assert(F % 2 == 0, "pairs");
#pragma unroll
    for (int f = 0; f < F; f += 2) {
        const double2 v = __ldg(reinterpret_cast<const double2 *>(frag_addr<F>(table, entry, lane, f)));
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
    static_assert(F % 2 == 0, "pairs");
#pragma unroll
    for (int f = 0; f < F; f += 2) cp_async<16>(dst + f, frag_addr<F>(table, entry, lane, f));
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

// ---- launch overlap with the fused tree kernel (programmatic dependent launch) ---------------------------------------------------
// The backward walker, the fused tree kernel and the forward walker are a chain of three launches whose heads (launch latency,
// descriptor / metadata / table staging: 3 - 5 us each) do not depend on the kernel before.  With programmatic dependent launch
// the next kernel starts as soon as every CTA of this one is RUNNING (griddepcontrol.launch_dependents at the top), does its
// staging next to it and then waits for the data itself: walk_count[instance] counts the tiles whose head q has been stored
// (the subtree CTAs of the tree kernel wait for all of them), tree_done[instance] the subtree CTAs whose forward pass has
// written the x of the chain heads (the forward walker waits for all of them).  The tree kernel's top CTA resets walk_count,
// the next backward walk resets tree_done.  Null pointers = plain stream-ordered launches (no protocol).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// bounded wait (one lane): ~0.2 s, then status bit 32 instead of a hung GPU
__device__ __forceinline__ void wait_count(const int *counter, int target, const Ctrl *ctrl) {
    const long long t0 = clock64();
    while (ld_acquire_gpu(counter) < target) {
        __nanosleep(64);
        if (clock64() - t0 > 400000000LL) {
            atomicOr(const_cast<int *>(&ctrl->status), 32);
            break;
        }
    }
}

// ---- backward:  r = ubar - B'q_child,  q = A'q_child - xbar - K'r    -------------------------------
// BIG (nx + nu > 32): the fragments do not fit into registers.  One warp per CTA; the [A | B] fragments of the tile's
// dynamics row are copied once into shared memory (lane-major, every lane reads only its own words: no barrier), the class
// fragments are read straight out of the ring stage -- one 16-byte shared-memory load per two MMAs.
template <int NX, int NU, bool BIG>
__global__ void __launch_bounds__(BIG ? 32 : 128) k_chain_mma_bwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                      SweepLevel lv, const double *__restrict__ prim,
                                                      double *__restrict__ q, double *__restrict__ r, int *walk_count,
                                                      int *tree_done) {
    using D = ChainDims<NX, NU>;
    extern __shared__ __align__(16) double mma_smem[];
    pdl_launch_dependents();
    if (tree_done && blockIdx.x == 0 && threadIdx.x == 0) tree_done[blockIdx.y] = 0;   // nobody waits on it any more (see above)
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
    if (walk_count) {   // publish: every lane's part of the head's q, then one count per tile
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(walk_count + blockIdx.y, 1);
    }
}

// ---- forward:  u = K x + R~^-1 r,  x_child = A x + B u -------------------------------------------------------------------
template <int NX, int NU, bool BIG>
__global__ void __launch_bounds__(BIG ? 32 : 128) k_chain_mma_fwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                      SweepLevel lv, double *__restrict__ prim, const double *__restrict__ r,
                                                      int d_begin, int d_end, const int *tree_done, int tree_ctas) {
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
    double w4[BIG ? 2 : F4];   // [A ; B]' of the chain's dynamics row, fragment (2 kb + j) * QT + ob
    if constexpr (!BIG) {
        if (lv.depth > 1) ld_frags<F4>(w4, P.m.fragABT, tm.dyns[1], lane);
    }
    // launched next to the tree kernel (and, transitively, possibly next to the BACKWARD walker, whose r rows the prefetch below
    // reads): its subtree CTAs count up once the x of the chain heads is written -- and they only get there after every tile of
    // the backward walk has been counted
    if (tree_done) {
        if (lane == 0) wait_count(tree_done + blockIdx.y, tree_ctas, ctrl);
        __syncwarp();
    }
    prefetch(d_begin);
    prefetch(d_begin + 1);
    double xs[D::QT][2];
    if (tree_done) {
        ld_state_cg<NX, NU>(X + (long long)tm.nodes[d_begin * 8 + g] * NX, t, xs);   // past the L1: written while this CTA ran
    } else {
        ld_state<NX, NU, false>(X + (long long)tm.nodes[d_begin * 8 + g] * NX, t, xs);   // written by the level above / the previous piece
    }

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


// ---- four warps per tile (nx + nu in (24, 32]: four 8-wide slot blocks) ------------------------------------------------------
// ncu of the one-warp-per-tile walkers on cfg3 (profiles/r2a_crit_*): 250 instructions per step of which 36 are DMMAs, a single
// warp per SM sub-partition at IPC 0.12 -- the step is one long dependent stream (address arithmetic, copies, the DMMA chains one
// after the other), 2 300 cycles of which the tensor pipe is busy 580.  Here a tile is walked by FOUR warps, warp w owning the
// output block w of every product: its share of a step is one accumulator chain per product (6 + 4 dependent DMMAs backward,
// 8 + 8 forward), one 16-byte row copy and 4-8 fragment words.  The state vector (and r / u) is exchanged through shared
// memory in the accumulator layout -- every lane writes the pair it computed and reads the pairs of all blocks as the A
// operands of the next product: two 128-thread named barriers per step, no shuffles.  The four warps of a tile sit on the
// four sub-partitions of the SM and every sub-partition interleaves the warps of the CTA's four tiles.
// Every reduction is split in two independent halves (see below): results differ from the one-warp kernels in the last bits.
constexpr int kWpt = 4;
constexpr int kStages4 = 4;   // ring stages of the four-warp walkers: copies are issued three steps ahead, two may be in flight
__device__ __forceinline__ void tile_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kWpt * 32) : "memory"); }

template <int NX, int NU>
struct Wide4 {
    using D = ChainDims<NX, NU>;
    static_assert(D::NT == kWpt && D::VEC && D::RT0 == 2 && D::RN == 2, "four slot blocks, the last two with input slots");
    static constexpr int F1 = 2 * D::QT * D::NT, F2 = 2 * D::RN * D::QT, F3 = 2 * D::NT * D::RN, F4 = 2 * D::NT * D::QT;
    static constexpr int LW_B = ring_lane_words(4 + 2 * D::RN);               // xbar pair | ubar pair | K fragments of the block
    static constexpr int LW_F = ring_lane_words(2 * D::RN + 2 * D::NT);       // r (all input blocks) | [K R~^-1] fragments
    static constexpr int XW = (D::RN + D::QT) * 64;                           // exchange doubles per tile: r / u blocks, state blocks
    // backward: one ring per block and tile; forward: only the blocks with input slots consume copied inputs
    __host__ __device__ static constexpr size_t ring_doubles(bool backward) {
        return backward ? (size_t)4 * 4 * kStages4 * 32 * LW_B : (size_t)4 * D::RN * kStages4 * 32 * LW_F;
    }
    __host__ __device__ static constexpr size_t smem_bytes(int depth, bool backward) {
        return ring_doubles(backward) * sizeof(double) + (size_t)4 * XW * sizeof(double) + (size_t)4 * (depth * 10 + 8) * sizeof(int);
    }
};

#ifdef RB_TRACE
#define RB_TK(i) do { if (d == trace_d) tk[i] = clock64(); } while (0)
#else
#define RB_TK(i) do {} while (0)
#endif

// Warp w of tile slot s owns block (w + s) % 4: every sub-partition (warp id % 4) then hosts all four blocks once, and the
// DMMA load of a step (6 + 4 | 6 + 4 | 6 + 4 | 6 backward, 8 | 8 | 8 + 8 | 8 forward) is spread evenly over the four
// tensor pipes -- with block = w all warps of one block share a pipe and every dependent DMMA queues behind three others
// (in-kernel clocks: 70 cycles per dependent DMMA instead of 26).
template <int NX, int NU>
__global__ void __launch_bounds__(4 * kWpt * 32) k_chain_mma_bwd_w4(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                                SweepLevel lv, const double *__restrict__ prim,
                                                                double *__restrict__ q, double *__restrict__ r) {
    using D = ChainDims<NX, NU>;
    using W = Wide4<NX, NU>;
    extern __shared__ __align__(16) double mma_smem[];
    const Layout &L = P.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = warp >> 2, ob = (warp + slot) & 3;   // tile of the CTA, output block of this warp
    const int tile = blockIdx.x * 4 + slot;
    if (tile >= lv.num_tiles) return;            // all four warps of the tile leave together
#ifdef RB_TRACE
    const long long t_entry = clock64();
    long long t_meta = 0, t_loop = 0;
#endif
    constexpr int LW = W::LW_B;
    double *ring = mma_smem + ((size_t)(slot * 4 + ob) * kStages4 * 32 + lane) * LW;
    double2 *xr = reinterpret_cast<double2 *>(mma_smem + W::ring_doubles(true) + (size_t)slot * W::XW);   // [RN][32]
    double2 *xq = xr + D::RN * 32;                                                                                    // [QT][32]
    TileMeta tm;   // the four warps write the same image: benign
    const bool live = stage_meta(lv, ctrl, tile, slot, lane,
                                 reinterpret_cast<int *>(mma_smem + W::ring_doubles(true) + (size_t)4 * W::XW), tm);
    tile_bar(1 + slot);
    if (!live) return;
#ifdef RB_TRACE
    t_meta = clock64();
#endif
    const int t = tm.t, g = tm.g;
    const int s0 = 8 * ob + 2 * t, a0 = s0 - NX;
    const bool has_state = s0 < NX, has_input = a0 >= 0 && a0 < NU;
    const bool ob_state = 8 * ob < NX, ob_input = 8 * ob + 8 > NX;   // warp-uniform: the block holds state / input slots
    const double *Xn = prim + (long long)blockIdx.y * L.np_pad + L.px, *Un = prim + (long long)blockIdx.y * L.np_pad + L.pu;
    double *Q = q + (long long)blockIdx.y * L.n * NX, *R = r + (long long)blockIdx.y * L.m * NU;
#pragma unroll
    for (int s = 0; s < kStages4; ++s)
#pragma unroll
        for (int k = 0; k < 4; ++k) ring[s * 32 * LW + k] = 0.0;   // pairs that are never copied stay zero (own ring: before the barrier below)
    // The warp of block 3 (input slots only) has nothing to do while the others run their K'r chains: it is the PRODUCER of
    // the step inputs of all four blocks (xbar / ubar pairs, K fragments), copying into the four rings two steps ahead and
    // waiting for its copies before the step's second barrier.  Same lane -> same words in every ring.
    const bool producer = ob == 3;
    double *rings = mma_smem + ((size_t)(slot * 4) * kStages4 * 32 + lane) * LW;   // ring of block b: rings + b * kStages4 * 32 * LW
    auto prefetch = [&](int d, int stage) {   // producer only; one commit group per step, possibly empty
        if (d >= 0) {
            const int node = tm.nodes[d * 8 + g], cls = tm.clss[d];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                double *dst = rings + ((size_t)b * kStages4 + stage) * 32 * LW;
                const int sb = 8 * b + 2 * t, ab = sb - NX;
                if (sb < NX) cp_async<16>(dst, Xn + (long long)node * NX + sb);
                if (cls >= 0) {
                    if (ab >= 0 && ab < NU) cp_async<16>(dst + 2, Un + (long long)node * NU + ab);
                    if (8 * b < NX) {
#pragma unroll
                        for (int k = 0; k < 2 * D::RN; k += 2)
                            cp_async<16>(dst + 4 + k, frag_addr<W::F2>(P.m.fragK4, cls, lane, b * 2 * D::RN + k));
                    }
                }
            }
        }
        cp_async_commit();
    };
    int d = lv.depth - 1;
    int st_use = 0, st_fill = 3;   // ring stage read by this step / filled for the step three steps on
    tile_bar(1 + slot);            // the rings are zeroed
    if (producer) {
        prefetch(d, 0);
        prefetch(d - 1, 1);
        prefetch(d - 2, 2);
    }
    double w1[2 * D::QT];   // column block `ob` of [A | B] of the tile's dynamics row
#pragma unroll
    for (int k = 0; k < 2 * D::QT; k += 2) {
        const double2 v = lv.depth > 1 ? __ldg(reinterpret_cast<const double2 *>(frag_addr<W::F1>(P.m.fragAB4, tm.dyns[1], lane, ob * 2 * D::QT + k)))
                                       : make_double2(0.0, 0.0);
        w1[k] = v.x;
        w1[k + 1] = v.y;
    }
    double qs[D::QT][2];
#pragma unroll
    for (int b = 0; b < D::QT; ++b) qs[b][0] = qs[b][1] = 0.0;
#ifdef RB_TRACE
    long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int trace_d = 6;
#endif
    if (producer) cp_async_wait<2>();   // the inputs of the first step have landed
    tile_bar(1 + slot);
#ifdef RB_TRACE
    t_loop = clock64();
#endif
    for (; d >= 0; --d) {
        RB_TK(0);
        const double *src = ring + st_use * 32 * LW;
        const int node = tm.nodes[d * 8 + g];
        const double2 xb = *reinterpret_cast<const double2 *>(src);
        double qp[2];
        RB_TK(1);
        if (tm.clss[d] < 0) {   // leaf: q = -xbar
            qp[0] = -xb.x;
            qp[1] = -xb.y;
        } else {
            const double2 ub = *reinterpret_cast<const double2 *>(src + 2);
            // block `ob` of [A'q ; B'q], the reduction in two independent halves (the warps of a sub-partition share one
            // tensor pipe: a dependent DMMA comes back after ~64 cycles, two chains per warp keep the pipe full)
            double E[2] = {0.0, 0.0}, E2[2] = {0.0, 0.0};
#pragma unroll
            for (int k = 0; k < D::QT; ++k) {
                dmma(E, qs[k >> 1][k & 1], w1[k]);
                dmma(E2, qs[(k + D::QT) >> 1][(k + D::QT) & 1], w1[k + D::QT]);
            }
            E[0] += E2[0];
            E[1] += E2[1];
            if (ob_input) {   // r = ubar - B'q on the input slots; -r feeds K'r
                const double r0 = has_input ? ub.x - E[0] : 0.0, r1 = has_input ? ub.y - E[1] : 0.0;
                if (has_input && tm.valid) *reinterpret_cast<double2 *>(R + (long long)node * NU + a0) = make_double2(r0, r1);
                xr[(ob - D::RT0) * 32 + lane] = make_double2(-r0, -r1);
            }
            qp[0] = has_state ? E[0] - xb.x : 0.0;
            qp[1] = has_state ? E[1] - xb.y : 0.0;
            RB_TK(2);
            tile_bar(1 + slot);
            RB_TK(3);
            if (ob_state) {   // q = A'q - xbar - K'r on the state slots
                double w2[2 * D::RN];
                lds_vec(w2, src + 4);
                double q2[2] = {0.0, 0.0};
                static_assert(D::RN == 2, "two input blocks: one chain each");
                const double2 nr0 = xr[lane], nr1 = xr[32 + lane];
                dmma(qp, nr0.x, w2[0]);
                dmma(q2, nr1.x, w2[2]);
                dmma(qp, nr0.y, w2[1]);
                dmma(q2, nr1.y, w2[3]);
                qp[0] += q2[0];
                qp[1] += q2[1];
            }
        }
        if (ob_state) xq[ob * 32 + lane] = make_double2(qp[0], qp[1]);
        RB_TK(4);
        if (producer) {   // off the critical path: the other warps are in their K'r chains
            prefetch(d - 3, st_fill);
            cp_async_wait<2>();   // the inputs of step d - 1 have landed: visible to the consumers after the barrier
        }
        RB_TK(5);
        tile_bar(1 + slot);
        RB_TK(6);
#pragma unroll
        for (int b = 0; b < D::QT; ++b) {
            const double2 v = xq[b * 32 + lane];
            qs[b][0] = v.x;
            qs[b][1] = v.y;
        }
        if (d == 0 && has_state && tm.valid)   // only the head's q leaves the chain
            *reinterpret_cast<double2 *>(Q + (long long)node * NX + s0) = make_double2(qp[0], qp[1]);
        st_use = st_use == kStages4 - 1 ? 0 : st_use + 1;
        st_fill = st_fill == kStages4 - 1 ? 0 : st_fill + 1;
#ifdef RB_TRACE
        if (d == trace_d) tk[7] = clock64() + (long long)(qs[0][0] == 12345.678);
#endif
    }
#ifdef RB_TRACE
    if (blockIdx.x == 5 && slot == 1 && lane == 0 && ctrl && ctrl->iters == 30)
        printf("bwd_w4 ob %d: wait %lld E-phase %lld bar1 %lld Kr-phase %lld prefetch %lld bar2 %lld readback %lld cycles; meta %lld "
               "prologue %lld loop %lld\n", ob, tk[1] - tk[0], tk[2] - tk[1], tk[3] - tk[2], tk[4] - tk[3], tk[5] - tk[4],
               tk[6] - tk[5], tk[7] - tk[6], t_meta - t_entry, t_loop - t_meta, clock64() - t_loop);
#endif
}

// Forward: u is computed by the warps of the blocks with input slots (2, 3), x_child by those of the blocks with state slots
// (0, 1, 2).  The warps of blocks 0 and 1 are idle while u is computed: they are the PRODUCERS of the step inputs of blocks 2
// and 3 (r rows, [K R~^-1] fragments), copying with cp.async into their partner's ring two steps ahead and waiting for their
// own copies before the step's first barrier -- the consumers then never issue or wait for a copy.
template <int NX, int NU>
__global__ void __launch_bounds__(4 * kWpt * 32) k_chain_mma_fwd_w4(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                                SweepLevel lv, double *__restrict__ prim,
                                                                const double *__restrict__ r, int d_begin, int d_end) {
    using D = ChainDims<NX, NU>;
    using W = Wide4<NX, NU>;
    extern __shared__ __align__(16) double mma_smem[];
    const Layout &L = P.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = warp >> 2, ob = (warp + slot) & 3;
    const int tile = blockIdx.x * 4 + slot;
    if (tile >= lv.num_tiles) return;
    constexpr int LW = W::LW_F;
    const bool ob_state = 8 * ob < NX, ob_input = 8 * ob + 8 > NX;
    const int cons = ob_input ? ob : ob + 2;   // the consumer block whose ring this warp reads (consumer) or fills (producer)
    double *ring = mma_smem + ((size_t)(slot * D::RN + cons - D::RT0) * kStages4 * 32 + lane) * LW;
    double2 *xu = reinterpret_cast<double2 *>(mma_smem + W::ring_doubles(false) + (size_t)slot * W::XW);   // [RN][32]
    double2 *xq = xu + D::RN * 32;                                                                                    // [QT][32]
    TileMeta tm;
    const bool live = stage_meta(lv, ctrl, tile, slot, lane,
                                 reinterpret_cast<int *>(mma_smem + W::ring_doubles(false) + (size_t)4 * W::XW), tm);
    tile_bar(1 + slot);
    if (!live) return;
    const int t = tm.t, g = tm.g;
    const int s0 = 8 * ob + 2 * t, a0 = s0 - NX;
    const bool has_state = s0 < NX, has_input = a0 >= 0 && a0 < NU;
    double *X = prim + (long long)blockIdx.y * L.np_pad + L.px, *U = prim + (long long)blockIdx.y * L.np_pad + L.pu;
    const double *R = r + (long long)blockIdx.y * L.m * NU;
    const bool producer = !ob_input;
    auto prefetch = [&](int d, int stage) {   // producers only; one commit group per step, possibly empty
        if (d < lv.depth) {
            const int cls = tm.clss[d];
            if (cls >= 0) {
                double *dst = ring + stage * 32 * LW;
                cp_input<NX, NU>(dst, R + (long long)tm.nodes[d * 8 + g] * NU, t);
#pragma unroll
                for (int k = 0; k < 2 * D::NT; k += 2)
                    cp_async<16>(dst + 2 * D::RN + k, frag_addr<W::F3>(P.m.fragKR4, cls, lane, (cons - D::RT0) * 2 * D::NT + k));
            }
        }
        cp_async_commit();
    };
    if (producer) {
#pragma unroll
        for (int s = 0; s < kStages4; ++s)
#pragma unroll
            for (int k = 0; k < 2 * D::RN; ++k) ring[s * 32 * LW + k] = 0.0;
        prefetch(d_begin, 0);
        prefetch(d_begin + 1, 1);
        prefetch(d_begin + 2, 2);
    }
    double w4[2 * D::NT];   // column block `ob` of [A ; B]' of the tile's dynamics row
#pragma unroll
    for (int k = 0; k < 2 * D::NT; k += 2) {
        const double2 v = (ob_state && lv.depth > 1)
                              ? __ldg(reinterpret_cast<const double2 *>(frag_addr<W::F4>(P.m.fragABT4, tm.dyns[1], lane, ob * 2 * D::NT + k)))
                              : make_double2(0.0, 0.0);
        w4[k] = v.x;
        w4[k + 1] = v.y;
    }
    double xs[D::QT][2];
    ld_state<NX, NU, false>(X + (long long)tm.nodes[d_begin * 8 + g] * NX, t, xs);   // written by the level above / the previous piece
    if (producer) cp_async_wait<2>();   // the inputs of the first step are in the partner's ring
    tile_bar(1 + slot);
    int st_use = 0, st_fill = 3;
#ifdef RB_TRACE
    long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int trace_d = d_begin + 6;
#endif
    for (int d = d_begin; d < d_end && d + 1 < lv.depth && tm.clss[d] >= 0; ++d) {
        RB_TK(0);
        const double *src = ring + st_use * 32 * LW;
        const int node = tm.nodes[d * 8 + g], child = tm.nodes[(d + 1) * 8 + g];
        if (ob_input) {   // block `ob` of u = [K R~^-1] [x ; r]
            double rr[2 * D::RN], w3[2 * D::NT];
            lds_vec(rr, src);
            lds_vec(w3, src + 2 * D::RN);
            double ua[2] = {0.0, 0.0}, ub2[2] = {0.0, 0.0};   // two independent halves of the reduction (see the backward kernel)
#pragma unroll
            for (int kb = 0; kb < D::NT; ++kb)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const bool is_x = 8 * kb + 2 * t + j < NX;
                    double v = 0.0;
                    if (kb < D::QT) v = xs[kb < D::QT ? kb : 0][j];
                    if (kb >= D::RT0 && !is_x) v = rr[2 * (kb >= D::RT0 ? kb - D::RT0 : 0) + j];
                    if (kb & 1) dmma(ub2, v, w3[2 * kb + j]);
                    else dmma(ua, v, w3[2 * kb + j]);
                }
            ua[0] += ub2[0];
            ua[1] += ub2[1];
            if (has_input && tm.valid) *reinterpret_cast<double2 *>(U + (long long)node * NU + a0) = make_double2(ua[0], ua[1]);
            xu[(ob - D::RT0) * 32 + lane] = make_double2(ua[0], ua[1]);
        } else {   // producer: inputs of step d + 3 on their way, those of step d + 1 landed before the barrier
            prefetch(d + 3, st_fill);
            cp_async_wait<2>();
        }
        RB_TK(1);
        tile_bar(1 + slot);
        RB_TK(2);
        if (ob_state) {   // block `ob` of x_child = [A B] [x ; u]
            double2 uu[D::RN];
#pragma unroll
            for (int i = 0; i < D::RN; ++i) uu[i] = xu[i * 32 + lane];
            double xn[2] = {0.0, 0.0}, xn2[2] = {0.0, 0.0};
#pragma unroll
            for (int kb = 0; kb < D::NT; ++kb)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const bool is_x = 8 * kb + 2 * t + j < NX;
                    double v = 0.0;
                    if (kb < D::QT) v = xs[kb < D::QT ? kb : 0][j];
                    if (kb >= D::RT0 && !is_x) v = j == 0 ? uu[kb >= D::RT0 ? kb - D::RT0 : 0].x : uu[kb >= D::RT0 ? kb - D::RT0 : 0].y;
                    if (kb & 1) dmma(xn2, v, w4[2 * kb + j]);
                    else dmma(xn, v, w4[2 * kb + j]);
                }
            xn[0] += xn2[0];
            xn[1] += xn2[1];
            if (has_state && tm.valid) *reinterpret_cast<double2 *>(X + (long long)child * NX + s0) = make_double2(xn[0], xn[1]);
            xq[ob * 32 + lane] = make_double2(xn[0], xn[1]);
        }
        RB_TK(3);
        tile_bar(1 + slot);
        RB_TK(4);
#pragma unroll
        for (int b = 0; b < D::QT; ++b) {
            const double2 v = xq[b * 32 + lane];
            xs[b][0] = v.x;
            xs[b][1] = v.y;
        }
        st_use = st_use == kStages4 - 1 ? 0 : st_use + 1;
        st_fill = st_fill == kStages4 - 1 ? 0 : st_fill + 1;
#ifdef RB_TRACE
        if (d == trace_d) tk[5] = clock64() + (long long)(xs[0][0] == 12345.678);
#endif
    }
#ifdef RB_TRACE
    if (blockIdx.x == 5 && slot == 1 && lane == 0 && ctrl && ctrl->iters == 30)
        printf("fwd_w4 ob %d: u-phase / prefetch %lld bar1 %lld x-phase %lld bar2 %lld readback %lld cycles\n", ob, tk[1] - tk[0],
               tk[2] - tk[1], tk[3] - tk[2], tk[4] - tk[3], tk[5] - tk[4]);
#endif
}


// ---- wide rows (nx, nu multiples of 32): four warps per tile, ONE tile per CTA ------------------------------------------------------
// cfg5 (nx = 64, nu = 32) has 92 tiles: with one warp per tile (BIG kernels above) 92 SMs each run ONE tensor pipe out of four,
// 256 / 288 DMMAs per step back to back (~27 cycles each: 3.5 us per step).  Here the twelve 8-wide slot blocks of a step's
// products are dealt round robin to four warps -- warp w owns state blocks w, w + 4 and input block w: every warp has the same
// share (64 / 72 DMMAs) and sits on its own sub-partition, so a tile uses all four tensor pipes of its SM.  Unlike the 20 x 10
// case (four tiles per SM already fill the pipes; k_chain_mma_*_w4 is an ablation there) this is where splitting a tile pays.
// The [A | B] / [A ; B]' fragments of a warp's blocks fit in registers (48 words per lane), the class fragments and rows come
// through a per-warp cp.async ring two steps ahead, q / r / u / x are exchanged through shared memory in the accumulator layout
// with two block barriers per step.
template <int NX, int NU>
struct WideRows {
    using D = ChainDims<NX, NU>;
    static_assert(NX % 32 == 0 && NU % 32 == 0, "state and input blocks are dealt four at a time");
    static constexpr int SB = D::QT / 4, IB = D::RN / 4;                      // state / input blocks per warp
    static constexpr int K1 = 2 * D::QT, K2 = 2 * D::RN, K34 = 2 * D::NT;     // reduction words of A'q | K'r | K x + R r, A x + B u
    static constexpr int F1 = K1 * D::NT, F2 = K2 * D::QT, F3 = K34 * D::RN, F4 = K34 * D::QT;
    static constexpr int LW_B = ring_lane_words(2 * SB + 2 * IB + SB * K2);   // xbar pairs | ubar pairs | K fragments
    static constexpr int LW_F = ring_lane_words(2 * D::RN + IB * K34);        // r (all input blocks) | [K R~^-1] fragments
    static constexpr int XW = (D::RN + D::QT) * 64;                           // exchange doubles: r / u blocks, state blocks
    __host__ __device__ static constexpr size_t smem_bytes(int depth, bool backward) {
        return (size_t)4 * kStages * 32 * (backward ? LW_B : LW_F) * sizeof(double) + (size_t)XW * sizeof(double) +
               (size_t)(depth * 10 + 8) * sizeof(int);
    }
};

template <int NX, int NU>
__global__ void __launch_bounds__(128) k_chain_mma_bwd_wide(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                            SweepLevel lv, const double *__restrict__ prim,
                                                            double *__restrict__ q, double *__restrict__ r) {
    using D = ChainDims<NX, NU>;
    using W = WideRows<NX, NU>;
    constexpr int SB = W::SB, IB = W::IB, LW = W::LW_B;
    extern __shared__ __align__(16) double mma_smem[];
    const Layout &L = P.L;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    if (tile >= lv.num_tiles) return;
    double *ring = mma_smem + ((size_t)w * kStages * 32 + lane) * LW;
    double2 *xr = reinterpret_cast<double2 *>(mma_smem + (size_t)4 * kStages * 32 * LW);   // [RN][32]
    double2 *xq = xr + D::RN * 32;                                                        // [QT][32]
    TileMeta tm;
    const bool live = stage_meta(lv, ctrl, tile, 0, lane, reinterpret_cast<int *>(mma_smem + (size_t)4 * kStages * 32 * LW + W::XW), tm);
    __syncthreads();
    if (!live) return;
    const int t = tm.t, g = tm.g;
    const double *X = prim + (long long)blockIdx.y * L.np_pad + L.px, *U = prim + (long long)blockIdx.y * L.np_pad + L.pu;
    double *Q = q + (long long)blockIdx.y * L.n * NX, *R = r + (long long)blockIdx.y * L.m * NU;
    auto prefetch = [&](int d) {   // one commit group per step, possibly empty
        if (d >= 0) {
            double *dst = ring + (d % kStages) * 32 * LW;
            const int node = tm.nodes[d * 8 + g], cls = tm.clss[d];
#pragma unroll
            for (int i = 0; i < SB; ++i) cp_async<16>(dst + 2 * i, X + (long long)node * NX + 8 * (w + 4 * i) + 2 * t);
            if (cls >= 0) {
#pragma unroll
                for (int i = 0; i < IB; ++i) cp_async<16>(dst + 2 * SB + 2 * i, U + (long long)node * NU + 8 * (w + 4 * i) + 2 * t);
#pragma unroll
                for (int i = 0; i < SB; ++i)
#pragma unroll
                    for (int k = 0; k < W::K2; k += 2)
                        cp_async<16>(dst + 2 * SB + 2 * IB + i * W::K2 + k, frag_addr<W::F2>(P.m.fragK4, cls, lane, (w + 4 * i) * W::K2 + k));
            }
        }
        cp_async_commit();
    };
    int d = lv.depth - 1;
    prefetch(d);
    prefetch(d - 1);
    // columns of [A | B] for this warp's blocks: state blocks w + 4 i (i < SB), then input blocks RT0 + w + 4 i (i < IB)
    double w1[SB + IB][W::K1];
#pragma unroll
    for (int b = 0; b < SB + IB; ++b) {
        const int ob = b < SB ? w + 4 * b : D::RT0 + w + 4 * (b - SB);
#pragma unroll
        for (int k = 0; k < W::K1; k += 2) {
            const double2 v = lv.depth > 1 ? __ldg(reinterpret_cast<const double2 *>(frag_addr<W::F1>(P.m.fragAB4, tm.dyns[1], lane, ob * W::K1 + k)))
                                           : make_double2(0.0, 0.0);
            w1[b][k] = v.x;
            w1[b][k + 1] = v.y;
        }
    }
    double qs[D::QT][2];
#pragma unroll
    for (int b = 0; b < D::QT; ++b) qs[b][0] = qs[b][1] = 0.0;

    for (; d >= 0; --d) {
        prefetch(d - 2);
        cp_async_wait<2>();   // the copies of step d have landed
        const double *src = ring + (d % kStages) * 32 * LW;
        const int node = tm.nodes[d * 8 + g];
        double qp[SB][2];
        if (tm.clss[d] < 0) {   // leaf: q = -xbar
#pragma unroll
            for (int i = 0; i < SB; ++i) {
                qp[i][0] = -src[2 * i];
                qp[i][1] = -src[2 * i + 1];
            }
        } else {
            double E[SB + IB][2];
#pragma unroll
            for (int b = 0; b < SB + IB; ++b) E[b][0] = E[b][1] = 0.0;
#pragma unroll
            for (int k = 0; k < W::K1; ++k)   // SB + IB independent accumulator chains
#pragma unroll
                for (int b = 0; b < SB + IB; ++b) dmma(E[b], qs[k >> 1][k & 1], w1[b][k]);
#pragma unroll
            for (int i = 0; i < IB; ++i) {   // r = ubar - B'q; -r feeds K'r
                const double r0 = src[2 * SB + 2 * i] - E[SB + i][0], r1 = src[2 * SB + 2 * i + 1] - E[SB + i][1];
                if (tm.valid) *reinterpret_cast<double2 *>(R + (long long)node * NU + 8 * (w + 4 * i) + 2 * t) = make_double2(r0, r1);
                xr[(w + 4 * i) * 32 + lane] = make_double2(-r0, -r1);
            }
#pragma unroll
            for (int i = 0; i < SB; ++i) {
                qp[i][0] = E[i][0] - src[2 * i];
                qp[i][1] = E[i][1] - src[2 * i + 1];
            }
            __syncthreads();
            double nr[W::K2];
#pragma unroll
            for (int i = 0; i < D::RN; ++i) {
                const double2 v = xr[i * 32 + lane];
                nr[2 * i] = v.x;
                nr[2 * i + 1] = v.y;
            }
            const double *w2 = src + 2 * SB + 2 * IB;
#pragma unroll
            for (int k = 0; k < W::K2; ++k)   // q = A'q - xbar - K'r: SB independent chains
#pragma unroll
                for (int i = 0; i < SB; ++i) dmma(qp[i], nr[k], w2[i * W::K2 + k]);
        }
#pragma unroll
        for (int i = 0; i < SB; ++i) xq[(w + 4 * i) * 32 + lane] = make_double2(qp[i][0], qp[i][1]);
        __syncthreads();
#pragma unroll
        for (int b = 0; b < D::QT; ++b) {
            const double2 v = xq[b * 32 + lane];
            qs[b][0] = v.x;
            qs[b][1] = v.y;
        }
        if (d == 0 && tm.valid) {   // only the head's q leaves the chain
#pragma unroll
            for (int i = 0; i < SB; ++i)
                *reinterpret_cast<double2 *>(Q + (long long)node * NX + 8 * (w + 4 * i) + 2 * t) = make_double2(qp[i][0], qp[i][1]);
        }
        __syncthreads();   // xq / xr are rewritten in the next step
    }
}

template <int NX, int NU>
__global__ void __launch_bounds__(128) k_chain_mma_fwd_wide(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl,
                                                            SweepLevel lv, double *__restrict__ prim, const double *__restrict__ r,
                                                            int d_begin, int d_end) {
    using D = ChainDims<NX, NU>;
    using W = WideRows<NX, NU>;
    constexpr int SB = W::SB, IB = W::IB, LW = W::LW_F;
    extern __shared__ __align__(16) double mma_smem[];
    const Layout &L = P.L;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    if (tile >= lv.num_tiles) return;
    double *ring = mma_smem + ((size_t)w * kStages * 32 + lane) * LW;
    double2 *xu = reinterpret_cast<double2 *>(mma_smem + (size_t)4 * kStages * 32 * LW);   // [RN][32]
    double2 *xq = xu + D::RN * 32;                                                        // [QT][32]
    TileMeta tm;
    const bool live = stage_meta(lv, ctrl, tile, 0, lane, reinterpret_cast<int *>(mma_smem + (size_t)4 * kStages * 32 * LW + W::XW), tm);
    __syncthreads();
    if (!live) return;
    const int t = tm.t, g = tm.g;
    double *X = prim + (long long)blockIdx.y * L.np_pad + L.px, *U = prim + (long long)blockIdx.y * L.np_pad + L.pu;
    const double *R = r + (long long)blockIdx.y * L.m * NU;
    auto prefetch = [&](int d) {
        if (d < lv.depth) {
            const int cls = tm.clss[d];
            if (cls >= 0) {
                double *dst = ring + (d % kStages) * 32 * LW;
                const double *row = R + (long long)tm.nodes[d * 8 + g] * NU;
#pragma unroll
                for (int i = 0; i < D::RN; ++i) cp_async<16>(dst + 2 * i, row + 8 * i + 2 * t);
#pragma unroll
                for (int i = 0; i < IB; ++i)
#pragma unroll
                    for (int k = 0; k < W::K34; k += 2)
                        cp_async<16>(dst + 2 * D::RN + i * W::K34 + k, frag_addr<W::F3>(P.m.fragKR4, cls, lane, (w + 4 * i) * W::K34 + k));
            }
        }
        cp_async_commit();
    };
    prefetch(d_begin);
    prefetch(d_begin + 1);
    double w4[SB][W::K34];   // columns of [A ; B]' for this warp's state blocks
#pragma unroll
    for (int i = 0; i < SB; ++i)
#pragma unroll
        for (int k = 0; k < W::K34; k += 2) {
            const double2 v = lv.depth > 1 ? __ldg(reinterpret_cast<const double2 *>(frag_addr<W::F4>(P.m.fragABT4, tm.dyns[1], lane, (w + 4 * i) * W::K34 + k)))
                                           : make_double2(0.0, 0.0);
            w4[i][k] = v.x;
            w4[i][k + 1] = v.y;
        }
    double xs[D::QT][2];
    ld_state<NX, NU, false>(X + (long long)tm.nodes[d_begin * 8 + g] * NX, t, xs);   // written by the level above / the previous piece

    for (int d = d_begin; d < d_end && d + 1 < lv.depth && tm.clss[d] >= 0; ++d) {
        prefetch(d + 2);
        cp_async_wait<2>();
        const double *src = ring + (d % kStages) * 32 * LW;
        const int node = tm.nodes[d * 8 + g], child = tm.nodes[(d + 1) * 8 + g];
        // u = [K R~^-1] [x ; r] on this warp's input blocks; the reduction in two independent halves per block
        double ua[IB][2], ub[IB][2];
#pragma unroll
        for (int i = 0; i < IB; ++i) ua[i][0] = ua[i][1] = ub[i][0] = ub[i][1] = 0.0;
        const double *w3 = src + 2 * D::RN;
#pragma unroll
        for (int kb = 0; kb < D::NT; ++kb)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const double v = kb < D::QT ? xs[kb < D::QT ? kb : 0][j] : src[2 * (kb >= D::QT ? kb - D::QT : 0) + j];
#pragma unroll
                for (int i = 0; i < IB; ++i) {
                    if (kb & 1) dmma(ub[i], v, w3[i * W::K34 + 2 * kb + j]);
                    else dmma(ua[i], v, w3[i * W::K34 + 2 * kb + j]);
                }
            }
#pragma unroll
        for (int i = 0; i < IB; ++i) {
            ua[i][0] += ub[i][0];
            ua[i][1] += ub[i][1];
            if (tm.valid) *reinterpret_cast<double2 *>(U + (long long)node * NU + 8 * (w + 4 * i) + 2 * t) = make_double2(ua[i][0], ua[i][1]);
            xu[(w + 4 * i) * 32 + lane] = make_double2(ua[i][0], ua[i][1]);
        }
        __syncthreads();
        double uu[W::K2];
#pragma unroll
        for (int i = 0; i < D::RN; ++i) {
            const double2 v = xu[i * 32 + lane];
            uu[2 * i] = v.x;
            uu[2 * i + 1] = v.y;
        }
        // x_child = [A B] [x ; u] on this warp's state blocks
        double xn[SB][2];
#pragma unroll
        for (int i = 0; i < SB; ++i) xn[i][0] = xn[i][1] = 0.0;
#pragma unroll
        for (int kb = 0; kb < D::NT; ++kb)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const double v = kb < D::QT ? xs[kb < D::QT ? kb : 0][j] : uu[2 * (kb >= D::QT ? kb - D::QT : 0) + j];
#pragma unroll
                for (int i = 0; i < SB; ++i) dmma(xn[i], v, w4[i][2 * kb + j]);
            }
#pragma unroll
        for (int i = 0; i < SB; ++i) {
            if (tm.valid) *reinterpret_cast<double2 *>(X + (long long)child * NX + 8 * (w + 4 * i) + 2 * t) = make_double2(xn[i][0], xn[i][1]);
            xq[(w + 4 * i) * 32 + lane] = make_double2(xn[i][0], xn[i][1]);
        }
        __syncthreads();
#pragma unroll
        for (int b = 0; b < D::QT; ++b) {
            const double2 v = xq[b * 32 + lane];
            xs[b][0] = v.x;
            xs[b][1] = v.y;
        }
        __syncthreads();   // xq / xu are rewritten in the next step
    }
}

}  // namespace

// ---- host side --------------------------------------------------------------------------------------------------------------
// (nx, nu, BIG): BIG = fragments in shared memory, one warp per CTA (nx + nu > 32)
#define RB_MMA_DIMS(X) X(2, 1, false) X(3, 2, false) X(4, 2, false) X(8, 4, false) X(10, 5, false) X(20, 10, false) X(64, 32, true)

// (nx, nu) walked by four warps per tile (k_chain_mma_*_w4): four slot blocks
#define RB_MMA_W4_DIMS(X) X(20, 10)
bool chain_mma_w4(int nx, int nu) {
#define RB_HAS(NX, NU) \
    if (nx == NX && nu == NU) return true;
    RB_MMA_W4_DIMS(RB_HAS)
#undef RB_HAS
    return false;
}

// (nx, nu) with wide rows walked by four warps per tile, one tile per CTA (k_chain_mma_*_wide): the default for these sizes
#define RB_MMA_WIDE_DIMS(X) X(64, 32)
bool chain_mma_wide(int nx, int nu) {
#define RB_HAS(NX, NU) \
    if (nx == NX && nu == NU) return true;
    RB_MMA_WIDE_DIMS(RB_HAS)
#undef RB_HAS
    return false;
}

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
    for (int obm = 0; obm < 2; ++obm) {   // k-major tables (one warp per tile, BIG), then output-block-major (four warps per tile)
        if (dynamics) {
            // [A'q ; B'q]: rows = state slots of q, columns = all slots
            k_frag_table<<<dim3(num_dyn, 2 * QT * NT), 32, 0, st>>>(M.ABcat, (long long)nx * S, S, 0, nx, 0, S, 0, 0, NT,
                                                                    const_cast<double *>(obm ? M.fragAB4 : M.fragAB), obm);
            // A x + B u: rows = all slots of [x ; u], columns = state slots
            k_frag_table<<<dim3(num_dyn, 2 * NT * QT), 32, 0, st>>>(M.ABcatT, (long long)S * nx, nx, 0, S, 0, nx, 0, 0, QT,
                                                                    const_cast<double *>(obm ? M.fragABT4 : M.fragABT), obm);
        }
        if (classes) {
            // K'r: rows = input slots (r), columns = state slots
            k_frag_table<<<dim3(num_cls, 2 * RN * QT), 32, 0, st>>>(M.K, (long long)nu * nx, nx, nx, nu, 0, nx, RT0, 0, QT,
                                                                    const_cast<double *>(obm ? M.fragK4 : M.fragK), obm);
            // K x + R~^-1 r: rows = all slots of [x ; r], columns = input slots
            k_frag_table<<<dim3(num_cls, 2 * NT * RN), 32, 0, st>>>(M.KRcatT, (long long)S * nu, nu, 0, S, nx, nu, 0, RT0, RN,
                                                                    const_cast<double *>(obm ? M.fragKR4 : M.fragKR), obm);
        }
    }
}

static dim3 mma_grid(const SweepLevel &lv, int batch, bool big) {
    return big ? dim3(lv.num_tiles, batch) : dim3((lv.num_tiles + 3) / 4, batch);
}
// dynamic shared memory of a CTA (4 warps, or 1 if BIG): the cp.async rings (+ the dynamics fragments if BIG) + the tile metadata
size_t chain_mma_smem_bytes(int nx, int nu, int depth, bool backward, bool w4) {
#define RB_WIDE(NX, NU) \
    if (w4 && nx == NX && nu == NU) return WideRows<NX, NU>::smem_bytes(depth, backward);
    RB_MMA_WIDE_DIMS(RB_WIDE)
#undef RB_WIDE
#define RB_W4(NX, NU) \
    if (w4 && nx == NX && nu == NU) return Wide4<NX, NU>::smem_bytes(depth, backward);
    RB_MMA_W4_DIMS(RB_W4)
#undef RB_W4
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
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_chain_mma_fwd<NX, NU, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); \
    /* largest shared-memory carve-out: a CTA of the fused tree kernel (~158 KB) can then share the SM (launch overlap) */             \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_chain_mma_bwd<NX, NU, BIG>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_chain_mma_fwd<NX, NU, BIG>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    RB_MMA_DIMS(RB_SET)
#undef RB_SET
#define RB_SET(NX, NU)                                                                                                         \
    if (e == cudaSuccess) {                                                                                                    \
        const int need = (int)std::max(Wide4<NX, NU>::smem_bytes(64, true), Wide4<NX, NU>::smem_bytes(64, false));             \
        e = cudaFuncSetAttribute(k_chain_mma_bwd_w4<NX, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);               \
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_chain_mma_fwd_w4<NX, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, need); \
    }
    RB_MMA_W4_DIMS(RB_SET)
#undef RB_SET
#define RB_SET(NX, NU)                                                                                                         \
    if (e == cudaSuccess) {                                                                                                    \
        const int need = (int)std::max(WideRows<NX, NU>::smem_bytes(64, true), WideRows<NX, NU>::smem_bytes(64, false));       \
        e = cudaFuncSetAttribute(k_chain_mma_bwd_wide<NX, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);             \
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_chain_mma_fwd_wide<NX, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, need); \
    }
    RB_MMA_WIDE_DIMS(RB_SET)
#undef RB_SET
    return e;
}

void launch_chain_mma_bwd(cudaStream_t st, const Params &P, const Ctrl *ctrl, const SweepLevel &lv, const double *prim,
                          double *q, double *r, bool w4, int *walk_count, int *tree_done) {
#define RB_GOW(NX, NU)                                                                                                         \
    if (w4 && P.L.nx == NX && P.L.nu == NU) {                                                                                  \
        k_chain_mma_bwd_wide<NX, NU><<<dim3(lv.num_tiles, P.L.batch), 128, WideRows<NX, NU>::smem_bytes(lv.depth, true), st>>>(P, ctrl, lv, prim, q, r); \
        return;                                                                                                                \
    }
    RB_MMA_WIDE_DIMS(RB_GOW)
#undef RB_GOW
#define RB_GO4(NX, NU)                                                                                                         \
    if (w4 && P.L.nx == NX && P.L.nu == NU) {                                                        \
        k_chain_mma_bwd_w4<NX, NU><<<dim3((lv.num_tiles + 3) / 4, P.L.batch), 4 * kWpt * 32,                                   \
                                     Wide4<NX, NU>::smem_bytes(lv.depth, true), st>>>(P, ctrl, lv, prim, q, r);               \
        return;                                                                                                                \
    }
    RB_MMA_W4_DIMS(RB_GO4)
#undef RB_GO4
#define RB_GO(NX, NU, BIG)                                                                                             \
    if (P.L.nx == NX && P.L.nu == NU) {                                                                           \
        k_chain_mma_bwd<NX, NU, BIG><<<mma_grid(lv, P.L.batch, BIG), BIG ? 32 : 128, chain_mma_smem_bytes(NX, NU, lv.depth, true, false), st>>>(P, ctrl, lv, prim, q, r, walk_count, tree_done);     \
        return;                                                                                                   \
    }
    RB_MMA_DIMS(RB_GO)
#undef RB_GO
}

// pdl: launch with programmatic stream serialization (the kernel may start while the previous one in the stream still runs; it
// waits for tree_done itself)
template <typename K, typename... A>
static void launch_pdl(K kernel, dim3 grid, int threads, size_t smem, cudaStream_t st, A... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, args...);
}

void launch_chain_mma_fwd(cudaStream_t st, const Params &P, const Ctrl *ctrl, const SweepLevel &lv, double *prim,
                          const double *r, int d_begin, int d_end, bool w4, const int *tree_done, int tree_ctas) {
    if (d_end < 0) d_end = lv.depth;
#define RB_GOW(NX, NU)                                                                                                         \
    if (w4 && P.L.nx == NX && P.L.nu == NU) {                                                                                  \
        k_chain_mma_fwd_wide<NX, NU><<<dim3(lv.num_tiles, P.L.batch), 128, WideRows<NX, NU>::smem_bytes(lv.depth, false), st>>>(P, ctrl, lv, prim, r, d_begin, d_end); \
        return;                                                                                                                \
    }
    RB_MMA_WIDE_DIMS(RB_GOW)
#undef RB_GOW
#define RB_GO4(NX, NU)                                                                                                         \
    if (w4 && P.L.nx == NX && P.L.nu == NU) {                                                        \
        k_chain_mma_fwd_w4<NX, NU><<<dim3((lv.num_tiles + 3) / 4, P.L.batch), 4 * kWpt * 32,                                   \
                                     Wide4<NX, NU>::smem_bytes(lv.depth, false), st>>>(P, ctrl, lv, prim, r, d_begin, d_end);  \
        return;                                                                                                                \
    }
    RB_MMA_W4_DIMS(RB_GO4)
#undef RB_GO4
#define RB_GO(NX, NU, BIG)                                                                                             \
    if (P.L.nx == NX && P.L.nu == NU) {                                                                           \
        if (tree_done)                                                                                            \
            launch_pdl(k_chain_mma_fwd<NX, NU, BIG>, mma_grid(lv, P.L.batch, BIG), BIG ? 32 : 128,                \
                       chain_mma_smem_bytes(NX, NU, lv.depth, false, false), st, P, ctrl, lv, prim, r, d_begin, d_end, tree_done, tree_ctas); \
        else                                                                                                      \
            k_chain_mma_fwd<NX, NU, BIG><<<mma_grid(lv, P.L.batch, BIG), BIG ? 32 : 128, chain_mma_smem_bytes(NX, NU, lv.depth, false, false), st>>>(P, ctrl, lv, prim, r, d_begin, d_end, nullptr, 0);        \
        return;                                                                                                   \
    }
    RB_MMA_DIMS(RB_GO)
#undef RB_GO
}

}  // namespace rb
