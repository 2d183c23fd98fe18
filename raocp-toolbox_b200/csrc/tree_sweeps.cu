// tree_sweeps.cu -- the BRANCHING part of the DP sweeps (reference cache.py:259-288): the top of the tree and the subtree
// level(s) above the chains, with everything a CTA touches resident in shared memory.
//
// A CTA owns one subtree (breadth-first numbering: one contiguous node range per stage).  What made the first version
// of these kernels slow was not arithmetic but a chain of ~8 dependent global-memory round trips per stage (topology,
// q of the children, operator tables, xbar / ubar, ...) at ~800 cycles each.  Here the host packs the topology of every
// subtree into one contiguous DESCRIPTOR; the CTA reads it (round trip 1), then issues cp.async copies of all rows it
// will need -- xbar, ubar (or r), the q of the children below the subtree, the dynamics tables and the K (or [K R~^-1])
// of its nodes -- (round trip 2), and walks the stages out of shared memory with two block barriers per stage.  q of
// interior nodes never goes to HBM; r, u, x are written as they are produced and never read back by the same kernel.
//
// Per stage one warp per parent does the whole step with the parent's children in flight together (independent
// accumulator chains), so a stage costs one matrix-vector latency and ONE block barrier.  The top kernel runs backward
// to the root and forward again inside one launch, r staying in shared memory in between.
//
// Subtrees that do not fit into shared memory (very wide stages or nx, nu beyond ~48) keep using sweeps.cu.
#include "kernels.cuh"
#ifdef RB_TRACE
#include <cstdio>
#endif

namespace rb {

namespace {

template <int BYTES>
__device__ __forceinline__ void cpa(void *smem_dst, const void *gsrc) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(gsrc), "n"(BYTES) : "memory");
}
// block-wide asynchronous copy of `count` doubles; vec: both sides are 16-byte aligned (decided once per kernel from the
// parities of nx / nu, not per call)
__device__ __forceinline__ void stage_rows(double *dst, const double *__restrict__ src, int count, bool vec) {
    if (vec) {
        for (int i = 2 * threadIdx.x; i + 1 < count; i += 2 * blockDim.x) cpa<16>(dst + i, src + i);
        if ((count & 1) && threadIdx.x == 0) cpa<8>(dst + count - 1, src + count - 1);
    } else {
        for (int i = threadIdx.x; i < count; i += blockDim.x) cpa<8>(dst + i, src + i);
    }
}
// one warp per node: the node's class-indexed matrix (len doubles, len even => 16-byte copies) into its own slot
__device__ __forceinline__ void stage_node_tables(double *dst, const double *__restrict__ table, const int *cls, int ns, int len) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    for (int i = warp; i < ns; i += warps) {
        const int c = cls[i];
        if (c < 0) continue;
        const double *src = table + (long long)c * len;
        double *d = dst + (long long)i * len;
        if ((len & 1) == 0) {
            for (int k = 2 * lane; k < len; k += 64) cpa<16>(d + k, src + k);
        } else {
            for (int k = lane; k < len; k += 32) cpa<8>(d + k, src + k);
        }
    }
}
__device__ __forceinline__ void stage_wait() {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
}

// out_k = sum_l MT[l * stride + k] * v[l];  COLS > 0: compile-time length (fully unrolled, loads first)
template <int COLS>
__device__ __forceinline__ double dot_col(const double *MT, int stride, const double *v, int k, int cols) {
    if constexpr (COLS > 0) {
        constexpr int CH = COLS > 32 ? 32 : COLS;   // at most 32 loads in flight per lane (wide rows: nx + nu > 32)
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int l0 = 0; l0 < COLS; l0 += CH) {
            double m[CH];
#pragma unroll
            for (int l = 0; l < CH; ++l)
                if (l0 + l < COLS) m[l] = MT[(l0 + l) * stride + k];
            asm volatile("" ::: "memory");   // loads first (see schedule_fence)
#pragma unroll
            for (int l = 0; l < CH; ++l)
                if (l0 + l < COLS) {
                    if ((l & 3) == 0) a0 = fma(m[l], v[l0 + l], a0);
                    else if ((l & 3) == 1) a1 = fma(m[l], v[l0 + l], a1);
                    else if ((l & 3) == 2) a2 = fma(m[l], v[l0 + l], a2);
                    else a3 = fma(m[l], v[l0 + l], a3);
                }
        }
        return (a0 + a1) + (a2 + a3);
    } else {
        double a0 = 0.0, a1 = 0.0;
        int l = 0;
        for (; l + 1 < cols; l += 2) {
            a0 = fma(MT[l * stride + k], v[l], a0);
            a1 = fma(MT[(l + 1) * stride + k], v[l + 1], a1);
        }
        if (l < cols) a0 = fma(MT[l * stride + k], v[l], a0);
        return a0 + a1;
    }
}

// shared-memory view of one subtree
struct Sub {
    const int *hdr;        // descriptor header: [0] nodes, [1] external children, [2] first external child (global id)
    const int *lo, *w, *off;            // per depth: first global node id, width, first local index
    const int *dyn, *cls, *cfirst, *ccount, *par;   // per local node
    const int *xdyn, *xpar;             // per external child: dynamics row, local index of the parent
    int depth, ns, ne, ext_first;
};
__device__ __forceinline__ Sub sub_view(const int *desc, const TreeLevel &lv) {
    Sub s;
    s.hdr = desc;
    s.depth = lv.depth;
    s.ns = desc[0];
    s.ne = desc[1];
    s.ext_first = desc[2];
    s.lo = desc + 4;
    s.w = s.lo + lv.depth;
    s.off = s.w + lv.depth;
    s.dyn = s.off + lv.depth;
    s.cls = s.dyn + lv.max_nodes;
    s.cfirst = s.cls + lv.max_nodes;
    s.ccount = s.cfirst + lv.max_nodes;
    s.par = s.ccount + lv.max_nodes;
    s.xdyn = s.par + lv.max_nodes;
    s.xpar = s.xdyn + lv.max_ext;
    return s;
}

// carve the dynamic shared memory (all sizes in doubles; the host computes the same total in tree_smem_bytes)
struct Carve {
    double *p;
    __device__ __forceinline__ double *take(long long count) {
        double *r = p;
        p += (count + 1) & ~1LL;
        return r;
    }
};

// chunk length for the batched loads below: the largest of 5, 4, 3, 2 that divides COLS
template <int COLS>
struct Chunk {
    static constexpr int value = COLS % 5 == 0 ? 5 : COLS % 4 == 0 ? 4 : COLS % 3 == 0 ? 3 : COLS % 2 == 0 ? 2 : 1;
};
// keeps the shared-memory loads of a chunk together and ahead of the multiply-adds that use them (ptxas otherwise
// schedules each LDS right before its DFMA and every DFMA eats a full shared-memory latency)
__device__ __forceinline__ void schedule_fence() { asm volatile("" ::: "memory"); }

// sum over a node's children c of  sum_l M_c[l * stride + k] * v_c[l]   with four children in flight (independent
// accumulator chains: the latency of a parent's step is that of ONE matrix-vector product, whatever the branching factor).
// tab: table of matrices (len doubles each) indexed by cdyn[c]; rows: the children's vectors (row_len apart).
template <int COLS>
__device__ __forceinline__ double child_sum(const double *tab, long long len, int stride, const int *cdyn, const double *rows,
                                            int row_len, int c0, int cc, int k, int cols) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int c = c0;
    for (; c + 3 < c0 + cc; c += 4) {
        const double *m0 = tab + cdyn[c] * len + k, *m1 = tab + cdyn[c + 1] * len + k, *m2 = tab + cdyn[c + 2] * len + k,
                     *m3 = tab + cdyn[c + 3] * len + k;
        const double *v0 = rows + (long long)c * row_len, *v1 = v0 + row_len, *v2 = v1 + row_len, *v3 = v2 + row_len;
        if constexpr (COLS > 0) {
            constexpr int CH = Chunk<COLS>::value;
#pragma unroll
            for (int l0 = 0; l0 < COLS; l0 += CH) {
                double m[4][CH], v[4][CH];
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    m[0][i] = m0[(l0 + i) * stride];
                    m[1][i] = m1[(l0 + i) * stride];
                    m[2][i] = m2[(l0 + i) * stride];
                    m[3][i] = m3[(l0 + i) * stride];
                    v[0][i] = v0[l0 + i];
                    v[1][i] = v1[l0 + i];
                    v[2][i] = v2[l0 + i];
                    v[3][i] = v3[l0 + i];
                }
                schedule_fence();
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    a0 = fma(m[0][i], v[0][i], a0);
                    a1 = fma(m[1][i], v[1][i], a1);
                    a2 = fma(m[2][i], v[2][i], a2);
                    a3 = fma(m[3][i], v[3][i], a3);
                }
            }
        } else {
            for (int l = 0; l < cols; ++l) {
                a0 = fma(m0[l * stride], v0[l], a0);
                a1 = fma(m1[l * stride], v1[l], a1);
                a2 = fma(m2[l * stride], v2[l], a2);
                a3 = fma(m3[l * stride], v3[l], a3);
            }
        }
    }
    for (; c < c0 + cc; ++c) a0 += dot_col<COLS>(tab + cdyn[c] * len, stride, rows + (long long)c * row_len, k, cols);
    return (a0 + a1) + (a2 + a3);
}

// Wide rows (nx + nu > 32, operator tables in global memory): a stage whose parents all have the same children pattern
// (same count, same dynamics rows -- every stage of a Markov tree) is BLOCKED over its parents: a warp owns one
// (child position, 32 output entries) pair, loads a table word once (16 loads in flight per lane) and uses it for up to PB
// parents, whose vectors it reads from shared memory (broadcast).  A 9-parent stage then moves the three tables from L2
// once instead of nine times -- a CTA's L2 bandwidth (~64 B/clk) is what bounded these stages.
// acc[p] = sum_l M[l * stride] * v[p * vstride + l], p < PB (rows of parents >= np shadow the last one; results unused)
template <int ROWS, int PB>
__device__ __forceinline__ void blocked_dots(const double *M, int stride, const double *v, int vstride, int np, double (&acc)[9]) {
    const double *vp[PB];
#pragma unroll
    for (int p = 0; p < PB; ++p) {
        vp[p] = v + (long long)(p < np ? p : np - 1) * vstride;
        acc[p] = 0.0;
    }
#pragma unroll
    for (int l0 = 0; l0 < ROWS; l0 += 16) {
        double m[16];
#pragma unroll
        for (int l = 0; l < 16; ++l)
            if (l0 + l < ROWS) m[l] = M[(long long)(l0 + l) * stride];
        schedule_fence();
#pragma unroll
        for (int l = 0; l < 16; ++l)
            if (l0 + l < ROWS) {
#pragma unroll
                for (int p = 0; p < PB; ++p) acc[p] = fma(m[l], vp[p][l0 + l], acc[p]);
            }
    }
}
template <int ROWS>
__device__ __forceinline__ void blocked_dots_np(const double *M, int stride, const double *v, int vstride, int np, double (&acc)[9]) {
    if (np <= 1) blocked_dots<ROWS, 1>(M, stride, v, vstride, np, acc);
    else if (np <= 3) blocked_dots<ROWS, 3>(M, stride, v, vstride, np, acc);
    else blocked_dots<ROWS, 9>(M, stride, v, vstride, np, acc);
}
// is the stage uniform (see above)?  children of consecutive parents are consecutive (breadth-first numbering)
__device__ __forceinline__ bool stage_uniform(const Sub &s, int off, int w, const int *cdyn, int cc) {
    const int cf0 = s.cfirst[off];
    bool ok = cc >= 1;
    for (int p = 0; p < w && ok; ++p) {
        ok = s.ccount[off + p] == cc && s.cfirst[off + p] == cf0 + p * cc;
        for (int j = 0; j < cc && ok; ++j) ok = cdyn[cf0 + p * cc + j] == cdyn[cf0 + j];
    }
    return ok;
}

// ---- backward over the subtree --------------------------------------------------------------------------------------------
// One warp per parent, one block barrier per stage.  xb, ub: staged rows (local node order); qa holds the q of the
// external children on entry; rbuf (optional) keeps r for a forward pass in the same kernel.  The root stage's q is
// written to global memory if Qglob.
// RES: the dynamics table (Ctab) and the per-node K (Ktab, indexed by local node) are in shared memory; otherwise both are
// the global tables (Ktab indexed by class).
template <int NX, int NU, bool RES>
__device__ __forceinline__ void tree_backward(const Layout &L, const Sub &s, const TreeLevel &lv, const double *xb,
                                              const double *ub, double *qa, double *qb, double *scratch, double *rbuf,
                                              const double *Ctab, const double *Ktab, double *__restrict__ Qglob,
                                              double *__restrict__ Rglob) {
    const int nx = NX > 0 ? NX : L.nx, nu = NX > 0 ? NU : L.nu, nxu = nx + nu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    double *qchild = qa, *qcur = qb;
    double *rv = scratch + (long long)warp * 2 * nxu, *acc = rv + nxu;
#ifdef RB_TRACE
    __shared__ long long st_b[8];
#endif
    for (int d = s.depth - 1; d >= 0; --d) {
#ifdef RB_TRACE
        if (threadIdx.x == 0 && d < 8) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(st_b[d]));
#endif
        const int w = s.w[d], off = s.off[d], lo = s.lo[d];
        if (s.cls[off] < 0) {   // leaves: q = -xbar
            for (int i = threadIdx.x; i < w * nx; i += blockDim.x) qcur[i] = -xb[off * nx + i];
        } else {
            const bool bottom = d == s.depth - 1;
            const int *cdyn = bottom ? s.xdyn : s.dyn + s.off[d + 1];
            // A stage with a handful of parents leaves most warps idle while each busy warp is one long serial stream
            // (four children x nx multiply-adds per lane).  With four children per parent and room for four warps per
            // parent, every child goes to a warp of its own and the parent's first warp adds the four partial sums -- in
            // the order child_sum adds them, so the result does not change by a bit.  In-kernel timestamps (RB_TRACE):
            // 4-parent stage 1.5 -> 1.28 us, 1-parent stage 1.1 -> 0.8 us.  (Blocking a 16-parent stage by child index, so
            // that a table word is read once for four parents, was measured too: 2.88 -> 2.56 us for that stage but
            // slower small stages and a slower kernel overall; not kept.)
            bool split = false;
            if constexpr (NX > 0 && NX + NU <= 32) {
                split = w * 4 <= warps;
                for (int p = 0; p < w && split; ++p) split = s.ccount[off + p] == 4;
            }
            bool blocked = false;
            if constexpr (NX > 0 && NX + NU > 32)
                blocked = s.ccount[off] * w <= 2 * warps && stage_uniform(s, off, w, cdyn, s.ccount[off]);
            if (blocked) {
                if constexpr (NX > 0 && NX + NU > 32) {
                    constexpr int PASSES = (NX + NU + 31) / 32;
                    const int cc = s.ccount[off], cf0 = s.cfirst[off], nblk = (w + 8) / 9, nt = cc * PASSES * nblk;
                    // partial[j][p][k] = entry k of [A'q ; B'q] of child j of parent p  (scratch: cc * w * nxu doubles)
                    for (int task = warp; task < nt; task += warps) {
                        const int j = task % cc, t = (task / cc) % PASSES, p0 = 9 * (task / (cc * PASSES));
                        const int k = lane + 32 * t, np = min(9, w - p0);
                        if (k < nxu) {
                            double a[9];
                            blocked_dots_np<NX>(Ctab + (long long)cdyn[cf0 + j] * nx * nxu + k, nxu,
                                                qchild + (long long)(cf0 + p0 * cc + j) * nx, cc * nx, np, a);
#pragma unroll
                            for (int p = 0; p < 9; ++p)
                                if (p < np) scratch[((long long)j * w + p0 + p) * nxu + k] = a[p];
                        }
                    }
                    __syncthreads();
                    for (int p = warp; p < w; p += warps) {
                        const int i = off + p;
                        const double *K = Ktab + (long long)(RES ? i : s.cls[i]) * nu * nx;
                        double a[PASSES];
#pragma unroll
                        for (int t = 0; t < PASSES; ++t) {
                            const int k = lane + 32 * t;
                            a[t] = 0.0;
                            if (k < nxu)
                                for (int j = 0; j < cc; ++j) a[t] += scratch[((long long)j * w + p) * nxu + k];
                        }
                        __syncwarp();   // the partial sums of parent p are consumed: r goes where partial[0][p][nx..) was
                        double *rvp = scratch + (long long)p * nxu + nx;
#pragma unroll
                        for (int t = 0; t < PASSES; ++t) {
                            const int k = lane + 32 * t;
                            if (k >= nx && k < nxu) {
                                const double rk = ub[i * nu + k - nx] - a[t];
                                rvp[k - nx] = rk;
                                if (rbuf) rbuf[i * nu + k - nx] = rk;
                                Rglob[(long long)(lo + p) * nu + k - nx] = rk;
                            }
                        }
                        __syncwarp();
#pragma unroll
                        for (int t = 0; t < PASSES; ++t) {
                            const int k = lane + 32 * t;
                            if (k < nx) qcur[p * nx + k] = a[t] - xb[i * nx + k] - dot_col<NU>(K, nx, rvp, k, nu);
                        }
                    }
                }
            } else if (split) {
                if constexpr (NX > 0 && NX + NU <= 32) {
                    double *part = scratch + (long long)warp * 2 * nxu;
                    if (warp < w * 4 && lane < nxu) {
                        const int c = s.cfirst[off + (warp >> 2)] + (warp & 3);
                        const double *m = Ctab + (long long)cdyn[c] * nx * nxu + lane, *v = qchild + (long long)c * nx;
                        double a = 0.0;
                        constexpr int CH = Chunk<NX>::value;
#pragma unroll
                        for (int l0 = 0; l0 < NX; l0 += CH) {
                            double mm[CH], vv[CH];
#pragma unroll
                            for (int t = 0; t < CH; ++t) {
                                mm[t] = m[(l0 + t) * nxu];
                                vv[t] = v[l0 + t];
                            }
                            schedule_fence();
#pragma unroll
                            for (int t = 0; t < CH; ++t) a = fma(mm[t], vv[t], a);
                        }
                        part[lane] = a;
                    }
                    __syncthreads();
                    if (warp < w * 4 && (warp & 3) == 0) {
                        const int p = warp >> 2, i = off + p;
                        const double *K = Ktab + (long long)(RES ? i : s.cls[i]) * nu * nx;
                        double a = 0.0;
                        if (lane < nxu) a = (part[lane] + part[2 * nxu + lane]) + (part[4 * nxu + lane] + part[6 * nxu + lane]);
                        __syncwarp();   // the partial sums of this warp are consumed: rv (same scratch row) may be written
                        if (lane >= nx && lane < nxu) {
                            const double rk = ub[i * nu + lane - nx] - a;
                            rv[lane - nx] = rk;
                            if (rbuf) rbuf[i * nu + lane - nx] = rk;
                            Rglob[(long long)(lo + p) * nu + lane - nx] = rk;
                        }
                        __syncwarp();
                        if (lane < nx) qcur[p * nx + lane] = a - xb[i * nx + lane] - dot_col<NU>(K, nx, rv, lane, nu);
                    }
                }
            } else
            for (int p = warp; p < w; p += warps) {
                const int i = off + p, c0 = s.cfirst[i], cc = s.ccount[i];
                const double *K = Ktab + (long long)(RES ? i : s.cls[i]) * nu * nx;
                if (NX > 0 && NX + NU <= 32) {   // lane k: entry k of [A'q ; B'q] summed over the children
                    double a = 0.0;
                    if (lane < nxu) a = child_sum<NX>(Ctab, (long long)nx * nxu, nxu, cdyn, qchild, nx, c0, cc, lane, nx);
                    if (lane >= nx && lane < nxu) {
                        const double rk = ub[i * nu + lane - nx] - a;
                        rv[lane - nx] = rk;
                        if (rbuf) rbuf[i * nu + lane - nx] = rk;
                        Rglob[(long long)(lo + p) * nu + lane - nx] = rk;
                    }
                    __syncwarp();
                    if (lane < nx) qcur[p * nx + lane] = a - xb[i * nx + lane] - dot_col<NU>(K, nx, rv, lane, nu);
                } else {
                    for (int k = lane; k < nxu; k += 32)
                        acc[k] = child_sum<NX>(Ctab, (long long)nx * nxu, nxu, cdyn, qchild, nx, c0, cc, k, nx);
                    __syncwarp();
                    for (int a = lane; a < nu; a += 32) {
                        const double rk = ub[i * nu + a] - acc[nx + a];
                        rv[a] = rk;
                        if (rbuf) rbuf[i * nu + a] = rk;
                        Rglob[(long long)(lo + p) * nu + a] = rk;
                    }
                    __syncwarp();
                    for (int k = lane; k < nx; k += 32)
                        qcur[p * nx + k] = acc[k] - xb[i * nx + k] - dot_col<NU>(K, nx, rv, k, nu);
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (d == 0 && Qglob)
            for (int i = threadIdx.x; i < w * nx; i += blockDim.x) Qglob[(long long)lo * nx + i] = qcur[i];
        double *tmp = qchild;
        qchild = qcur;
        qcur = tmp;
    }
#ifdef RB_TRACE
    if (threadIdx.x == 0 && gridDim.x > 1 && blockIdx.x == gridDim.x - 1) {
        long long t_;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
        printf("top bwd stages (deep->root): %lld %lld %lld ns\n", st_b[1] - st_b[2], st_b[0] - st_b[1], t_ - st_b[0]);
    }
#endif
}

// ---- forward over the subtree ----------------------------------------------------------------------------------------------
// One warp per parent: u = K x + R~^-1 r, then x of each child = A x + B u (all children in flight); one block barrier per
// stage.  xa holds x of the root stage on entry; rbuf: r rows (local node order).  Writes u of all nonleaf nodes and x of
// all nodes below the root stage (including the external children) to global memory.
template <int NX, int NU, bool RES>
__device__ __forceinline__ void tree_forward(const Layout &L, const Sub &s, const TreeLevel &lv, double *xa, double *xbuf,
                                             double *scratch, const double *rbuf, const double *CTtab, const double *KRtab,
                                             double *__restrict__ Xglob, double *__restrict__ Uglob) {
    const int nx = NX > 0 ? NX : L.nx, nu = NX > 0 ? NU : L.nu, nxu = nx + nu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    double *xcur = xa, *xnext = xbuf;
    double *v = scratch + (long long)warp * 2 * nxu;
    for (int d = 0; d < s.depth; ++d) {
        const int w = s.w[d], off = s.off[d], lo = s.lo[d];
        if (s.cls[off] < 0) break;   // leaves
        const bool bottom = d == s.depth - 1;
        const int *cdyn = bottom ? s.xdyn : s.dyn + s.off[d + 1];
        const long long cglob = bottom ? s.ext_first : s.lo[d + 1];
        bool split = false;   // as in tree_backward: one warp per child when the stage has only a few parents
        if constexpr (NX > 0 && NX + NU <= 32) {
            split = w * 4 <= warps;
            for (int p = 0; p < w && split; ++p) split = s.ccount[off + p] == 4;
        }
        bool blocked = false;   // as in tree_backward: wide rows, uniform stage
        if constexpr (NX > 0 && NX + NU > 32) blocked = w <= 2 * warps && stage_uniform(s, off, w, cdyn, s.ccount[off]);
        if (blocked) {
            if constexpr (NX > 0 && NX + NU > 32) {
                for (int p = warp; p < w; p += warps) {   // u = K x + R~^-1 r; [x ; u] of parent p into scratch row p
                    const int i = off + p;
                    double *vp = scratch + (long long)p * nxu;
                    for (int k = lane; k < nx; k += 32) vp[k] = xcur[p * nx + k];
                    for (int a = lane; a < nu; a += 32) vp[nx + a] = rbuf[i * nu + a];
                    __syncwarp();
                    const double *KR = KRtab + (long long)(RES ? i : s.cls[i]) * nxu * nu;
                    double ua[(NU + 31) / 32];
#pragma unroll
                    for (int t = 0; t < (NU + 31) / 32; ++t) {
                        const int a = lane + 32 * t;
                        ua[t] = a < nu ? dot_col<NX + NU>(KR, nu, vp, a, nxu) : 0.0;
                    }
                    __syncwarp();
#pragma unroll
                    for (int t = 0; t < (NU + 31) / 32; ++t) {
                        const int a = lane + 32 * t;
                        if (a < nu) {
                            vp[nx + a] = ua[t];
                            Uglob[(long long)(lo + p) * nu + a] = ua[t];
                        }
                    }
                }
                __syncthreads();
                constexpr int PASSES = (NX + 31) / 32;
                const int cc = s.ccount[off], cf0 = s.cfirst[off], nblk = (w + 8) / 9, nt = cc * PASSES * nblk;
                for (int task = warp; task < nt; task += warps) {   // x of child j of up to 9 parents = [A B]_j [x ; u]
                    const int j = task % cc, t = (task / cc) % PASSES, p0 = 9 * (task / (cc * PASSES));
                    const int k = lane + 32 * t, np = min(9, w - p0);
                    if (k < nx) {
                        double a[9];
                        blocked_dots_np<NX + NU>(CTtab + (long long)cdyn[cf0 + j] * nxu * nx + k, nx,
                                                 scratch + (long long)p0 * nxu, nxu, np, a);
#pragma unroll
                        for (int p = 0; p < 9; ++p)
                            if (p < np) {
                                const int c = cf0 + (p0 + p) * cc + j;
                                if (!bottom) xnext[c * nx + k] = a[p];
                                Xglob[(cglob + c) * nx + k] = a[p];
                            }
                    }
                }
            }
        } else if (split) {
            if constexpr (NX > 0 && NX + NU <= 32) {
                if (warp < w * 4 && (warp & 3) == 0) {   // the parent's first warp: u = K x + R~^-1 r, [x; u] into its scratch row
                    const int p = warp >> 2, i = off + p;
                    if (lane < nx) v[lane] = xcur[p * nx + lane];
                    if (lane < nu) v[nx + lane] = rbuf[i * nu + lane];
                    __syncwarp();
                    const double *KR = KRtab + (long long)(RES ? i : s.cls[i]) * nxu * nu;
                    double ua = 0.0;
                    if (lane < nu) ua = dot_col<NX + NU>(KR, nu, v, lane, nxu);
                    __syncwarp();
                    if (lane < nu) {
                        v[nx + lane] = ua;
                        Uglob[(long long)(lo + p) * nu + lane] = ua;
                    }
                }
                __syncthreads();
                if (warp < w * 4 && lane < nx) {   // one child per warp: x_child = [A B] [x; u]
                    const int p = warp >> 2, c = s.cfirst[off + p] + (warp & 3);
                    const double *vp = scratch + (long long)(warp & ~3) * 2 * nxu;
                    const double *m = CTtab + (long long)cdyn[c] * nxu * nx + lane;
                    double a = 0.0;
                    constexpr int CH = Chunk<NX + NU>::value;
#pragma unroll
                    for (int l0 = 0; l0 < NX + NU; l0 += CH) {
                        double mm[CH], vl[CH];
#pragma unroll
                        for (int t = 0; t < CH; ++t) {
                            mm[t] = m[(l0 + t) * NX];
                            vl[t] = vp[l0 + t];
                        }
                        schedule_fence();
#pragma unroll
                        for (int t = 0; t < CH; ++t) a = fma(mm[t], vl[t], a);
                    }
                    if (!bottom) xnext[c * nx + lane] = a;
                    Xglob[(cglob + c) * nx + lane] = a;
                }
            }
        } else
        for (int p = warp; p < w; p += warps) {
            const int i = off + p, c0 = s.cfirst[i], cc = s.ccount[i];
            for (int k = lane; k < nx; k += 32) v[k] = xcur[p * nx + k];
            for (int a = lane; a < nu; a += 32) v[nx + a] = rbuf[i * nu + a];
            __syncwarp();
            const double *KR = KRtab + (long long)(RES ? i : s.cls[i]) * nxu * nu;
            double ua[2];
            int cnt = 0;
            for (int a = lane; a < nu; a += 32) ua[cnt++] = dot_col<(NX > 0 ? NX + NU : 0)>(KR, nu, v, a, nxu);
            __syncwarp();
            cnt = 0;
            for (int a = lane; a < nu; a += 32) {
                v[nx + a] = ua[cnt];
                Uglob[(long long)(lo + p) * nu + a] = ua[cnt++];
            }
            __syncwarp();
            // the children: four at a time, lane k = entry k of x_child
            int c = c0;
            for (; c + 3 < c0 + cc; c += 4) {
                for (int k = lane; k < nx; k += 32) {
                    const double *m0 = CTtab + (long long)cdyn[c] * nxu * nx + k, *m1 = CTtab + (long long)cdyn[c + 1] * nxu * nx + k,
                                 *m2 = CTtab + (long long)cdyn[c + 2] * nxu * nx + k, *m3 = CTtab + (long long)cdyn[c + 3] * nxu * nx + k;
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                    if constexpr (NX > 0) {
                        constexpr int CH = Chunk<NX + NU>::value;
#pragma unroll
                        for (int l0 = 0; l0 < NX + NU; l0 += CH) {
                            double m[4][CH], vl[CH];
#pragma unroll
                            for (int i = 0; i < CH; ++i) {
                                m[0][i] = m0[(l0 + i) * NX];
                                m[1][i] = m1[(l0 + i) * NX];
                                m[2][i] = m2[(l0 + i) * NX];
                                m[3][i] = m3[(l0 + i) * NX];
                                vl[i] = v[l0 + i];
                            }
                            schedule_fence();
#pragma unroll
                            for (int i = 0; i < CH; ++i) {
                                a0 = fma(m[0][i], vl[i], a0);
                                a1 = fma(m[1][i], vl[i], a1);
                                a2 = fma(m[2][i], vl[i], a2);
                                a3 = fma(m[3][i], vl[i], a3);
                            }
                        }
                    } else {
                        for (int l = 0; l < nxu; ++l) {
                            const double vl = v[l];
                            a0 = fma(m0[l * nx], vl, a0);
                            a1 = fma(m1[l * nx], vl, a1);
                            a2 = fma(m2[l * nx], vl, a2);
                            a3 = fma(m3[l * nx], vl, a3);
                        }
                    }
                    if (!bottom) {
                        xnext[(c + 0) * nx + k] = a0;
                        xnext[(c + 1) * nx + k] = a1;
                        xnext[(c + 2) * nx + k] = a2;
                        xnext[(c + 3) * nx + k] = a3;
                    }
                    Xglob[(cglob + c + 0) * nx + k] = a0;
                    Xglob[(cglob + c + 1) * nx + k] = a1;
                    Xglob[(cglob + c + 2) * nx + k] = a2;
                    Xglob[(cglob + c + 3) * nx + k] = a3;
                }
            }
            for (; c < c0 + cc; ++c)
                for (int k = lane; k < nx; k += 32) {
                    const double xk = dot_col<(NX > 0 ? NX + NU : 0)>(CTtab + (long long)cdyn[c] * nxu * nx, nx, v, k, nxu);
                    if (!bottom) xnext[c * nx + k] = xk;
                    Xglob[(cglob + c) * nx + k] = xk;
                }
            __syncwarp();
        }
        __syncthreads();
        double *tmp = xcur;
        xcur = xnext;
        xnext = tmp;
    }
}

// stage the subtree descriptor, then everything the sweep needs
// The loads of the descriptor, of the loop's "done" flag and (fused kernel) of the hand-off flag are independent: they are
// all issued before the first of them is waited for (one global round trip instead of three at the head of the kernel).
// Returns null if the loop has already stopped (uniform over the launch).
__device__ __forceinline__ const int *stage_desc(const TreeLevel &lv, int sub, Carve &cv, const Ctrl *ctrl,
                                                 const int *flag = nullptr, int *flag0 = nullptr) {
    int *desc = reinterpret_cast<int *>(cv.take((lv.desc_stride + 1) / 2));
    const int *src = lv.desc + (long long)sub * lv.desc_stride;
    int mine[2] = {0, 0};
    const int i0 = threadIdx.x, i1 = threadIdx.x + blockDim.x;
    if (i0 < lv.desc_stride) mine[0] = __ldg(src + i0);
    if (i1 < lv.desc_stride) mine[1] = __ldg(src + i1);
    const int done = ctrl ? *reinterpret_cast<const volatile int *>(&ctrl->done) : 0;
    if (flag && threadIdx.x == 0) *flag0 = *reinterpret_cast<const volatile int *>(flag);
    for (int i = threadIdx.x + 2 * blockDim.x; i < lv.desc_stride; i += blockDim.x) desc[i] = __ldg(src + i);
    if (i0 < lv.desc_stride) desc[i0] = mine[0];
    if (i1 < lv.desc_stride) desc[i1] = mine[1];
    if (done) return nullptr;
    __syncthreads();
    return desc;
}

}  // namespace

template <int NX, int NU, bool RES>
__global__ void __launch_bounds__(512) k_tree_bwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl, TreeLevel lv,
                                                 const double *__restrict__ prim, double *__restrict__ q,
                                                 double *__restrict__ r) {
    extern __shared__ __align__(16) double tree_smem[];
    const Layout &L = P.L;
    const int nx = NX > 0 ? NX : L.nx, nu = NX > 0 ? NU : L.nu, nxu = nx + nu;
    const bool vx = (nx & 1) == 0, vu = (nu & 1) == 0;
    Carve cv{tree_smem};
    const int *desc = stage_desc(lv, blockIdx.x, cv, ctrl);
    if (!desc) return;
    const Sub s = sub_view(desc, lv);
    const double *X = prim + (long long)blockIdx.y * L.np_pad + L.px, *U = prim + (long long)blockIdx.y * L.np_pad + L.pu;
    double *Q = q + (long long)blockIdx.y * L.n * nx, *R = r + (long long)blockIdx.y * L.m * nu;
    double *xb = cv.take((long long)lv.max_nodes * nx), *ub = cv.take((long long)lv.max_nodes * nu);
    double *qa = cv.take((long long)lv.max_row * nx), *qb = cv.take((long long)lv.max_row * nx);
    double *scratch = cv.take((long long)(blockDim.x >> 5) * 2 * nxu);
    for (int d = 0; d < s.depth; ++d) {
        stage_rows(xb + s.off[d] * nx, X + (long long)s.lo[d] * nx, s.w[d] * nx, vx);
        if (s.cls[s.off[d]] >= 0) stage_rows(ub + s.off[d] * nu, U + (long long)s.lo[d] * nu, s.w[d] * nu, vu);
    }
    stage_rows(qa, Q + (long long)s.ext_first * nx, s.ne * nx, vx);
    if constexpr (RES) {
        double *ctab = cv.take((long long)lv.num_dyn * nx * nxu), *knode = cv.take((long long)lv.max_nodes * nu * nx);
        stage_rows(ctab, P.m.ABcat, lv.num_dyn * nx * nxu, ((nx * nxu) & 1) == 0);
        stage_node_tables(knode, P.m.K, s.cls, s.ns, nu * nx);
        stage_wait();
        tree_backward<NX, NU, true>(L, s, lv, xb, ub, qa, qb, scratch, nullptr, ctab, knode, Q, R);
    } else {
        stage_wait();
        tree_backward<NX, NU, false>(L, s, lv, xb, ub, qa, qb, scratch, nullptr, P.m.ABcat, P.m.K, Q, R);
    }
}

template <int NX, int NU, bool RES>
__global__ void __launch_bounds__(512) k_tree_fwd(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl, TreeLevel lv,
                                                 double *__restrict__ prim, const double *__restrict__ r) {
    extern __shared__ __align__(16) double tree_smem[];
    const Layout &L = P.L;
    const int nx = NX > 0 ? NX : L.nx, nu = NX > 0 ? NU : L.nu, nxu = nx + nu;
    const bool vx = (nx & 1) == 0, vu = (nu & 1) == 0;
    Carve cv{tree_smem};
    const int *desc = stage_desc(lv, blockIdx.x, cv, ctrl);
    if (!desc) return;
    const Sub s = sub_view(desc, lv);
    double *X = prim + (long long)blockIdx.y * L.np_pad + L.px, *U = prim + (long long)blockIdx.y * L.np_pad + L.pu;
    const double *R = r + (long long)blockIdx.y * L.m * nu;
    double *rbuf = cv.take((long long)lv.max_nodes * nu);
    double *xa = cv.take((long long)lv.max_row * nx), *xbuf = cv.take((long long)lv.max_row * nx);
    double *scratch = cv.take((long long)(blockDim.x >> 5) * 2 * nxu);
    for (int d = 0; d < s.depth; ++d)
        if (s.cls[s.off[d]] >= 0) stage_rows(rbuf + s.off[d] * nu, R + (long long)s.lo[d] * nu, s.w[d] * nu, vu);
    stage_rows(xa, X + (long long)s.lo[0] * nx, s.w[0] * nx, vx);   // x of the root stage comes from the level above
    if constexpr (RES) {
        double *cttab = cv.take((long long)lv.num_dyn * nxu * nx), *krnode = cv.take((long long)lv.max_nodes * nxu * nu);
        stage_rows(cttab, P.m.ABcatT, lv.num_dyn * nxu * nx, ((nx * nxu) & 1) == 0);
        stage_node_tables(krnode, P.m.KRcatT, s.cls, s.ns, nxu * nu);
        stage_wait();
        tree_forward<NX, NU, true>(L, s, lv, xa, xbuf, scratch, rbuf, cttab, krnode, X, U);
    } else {
        stage_wait();
        tree_forward<NX, NU, false>(L, s, lv, xa, xbuf, scratch, rbuf, P.m.ABcatT, P.m.KRcatT, X, U);
    }
}

// ---- the subtree-sharding exchange as 16-byte packets (kernels.cuh PeerXchg::ll) ---------------------------------------------------
__device__ __forceinline__ void ll_store(uint4 *p, double v, unsigned flag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)b), "r"(flag), "r"((unsigned)(b >> 32)),
                 "r"(flag)
                 : "memory");
}
// spins until both halves of the packet carry `flag`; false after ~2 s (a peer is gone)
__device__ __forceinline__ bool ll_load(const uint4 *p, unsigned flag, double *out) {
    unsigned a, fa, b, fb;
    const long long t0 = clock64();
    for (;;) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb) : "l"(p) : "memory");
        if (fa == flag && fb == flag) break;
        if (clock64() - t0 > 4000000000LL) return false;
    }
    *out = __longlong_as_double((long long)(((unsigned long long)b << 32) | a));
    return true;
}

// top of the tree: one CTA per problem instance, backward to the root, x_0 <- initial state (cache.py:282), forward again.
// SHARD (subtree sharding, batch 1): the q_j of the cut nodes come from all ranks.  The kernel first stores this rank's rows, aux
// scalars and residual maxima into every peer's packet area, then -- with its own tables and rows in flight -- polls the peers'
// packets into the shared-memory row buffer, folds the maxima and runs the stopping test of the previous iteration (k_check's
// body, fused.cu); if the loop has stopped it returns before the sweep.
template <int NX, int NU, bool RES, bool SHARD>
__global__ void __launch_bounds__(512) k_tree_top(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl, TreeLevel lv,
                                                 double *__restrict__ prim, double *__restrict__ q, double *__restrict__ r,
                                                 const double *__restrict__ x0, const __grid_constant__ ShardHand sh) {
    extern __shared__ __align__(16) double tree_smem[];
    __shared__ double sh_max[SHARD ? kMaxPeers : 1][6];
    __shared__ int sh_timeout;
    if (SHARD && threadIdx.x == 0) sh_timeout = 0;   // (the barrier of stage_desc follows)
    const Layout &L = P.L;
    const int nx = NX > 0 ? NX : L.nx, nu = NX > 0 ? NU : L.nu, nxu = nx + nu;
    const bool vx = (nx & 1) == 0, vu = (nu & 1) == 0;
    Carve cv{tree_smem};
    const int *desc = stage_desc(lv, 0, cv, ctrl);
    if (!desc) return;
    const Sub s = sub_view(desc, lv);
    double *X = prim + (long long)blockIdx.x * L.np_pad + L.px, *U = prim + (long long)blockIdx.x * L.np_pad + L.pu;
    double *Q = q + (long long)blockIdx.x * L.n * nx, *R = r + (long long)blockIdx.x * L.m * nu;
    double *xb = cv.take((long long)lv.max_nodes * nx), *ub = cv.take((long long)lv.max_nodes * nu);
    double *qa = cv.take((long long)lv.max_row * nx), *qb = cv.take((long long)lv.max_row * nx);
    double *scratch = cv.take((long long)(blockDim.x >> 5) * 2 * nxu);
    double *rbuf = cv.take((long long)lv.max_nodes * nu);
    for (int d = 0; d < s.depth; ++d) {
        stage_rows(xb + s.off[d] * nx, X + (long long)s.lo[d] * nx, s.w[d] * nx, vx);
        if (s.cls[s.off[d]] >= 0) stage_rows(ub + s.off[d] * nu, U + (long long)s.lo[d] * nu, s.w[d] * nu, vu);
    }
    unsigned long long seq = 0ull;
    if constexpr (SHARD) {   // this rank's packets leave first
        const ShardPlan &sp = sh.sp;
        seq = *sh.px.seq;
        const int parity = (int)(seq & 1ull), w = nx + 1, count = sp.cut_hi - sp.cut_lo, len = count * w + 6;
        const long long stride = (long long)sp.cap * w + 6;
        for (int i = threadIdx.x; i < (sp.world - 1) * len; i += blockDim.x) {
            int peer = i / len;
            const int e = i - peer * len;
            if (peer >= sp.rank) ++peer;
            double v;
            long long slot;
            if (e < count * w) {
                const int c = e / w, k = e - c * w, node = sp.cut_first + sp.cut_lo + c;
                v = k < nx ? Q[(long long)node * nx + k] : sh.aux[node];
                slot = e;
            } else {
                v = sh.slots[e - count * w];
                slot = (long long)sp.cap * w + (e - count * w);
            }
            ll_store(sh.px.ll[peer] + ((long long)parity * sp.world + sp.rank) * stride + slot, v, (unsigned)seq);
        }
        // own rows of the cut stage
        for (int i = threadIdx.x; i < count * nx; i += blockDim.x)
            qa[(long long)sp.cut_lo * nx + i] = Q[(long long)(sp.cut_first + sp.cut_lo) * nx + i];
    } else {
        stage_rows(qa, Q + (long long)s.ext_first * nx, s.ne * nx, vx);
    }
    const double *ct = P.m.ABcat, *ctt = P.m.ABcatT, *kt = P.m.K, *krt = P.m.KRcatT;
    if constexpr (RES) {
        double *ctab = cv.take((long long)lv.num_dyn * nx * nxu), *cttab = cv.take((long long)lv.num_dyn * nxu * nx);
        double *knode = cv.take((long long)lv.max_nodes * nu * nx), *krnode = cv.take((long long)lv.max_nodes * nxu * nu);
        stage_rows(ctab, P.m.ABcat, lv.num_dyn * nx * nxu, ((nx * nxu) & 1) == 0);
        stage_rows(cttab, P.m.ABcatT, lv.num_dyn * nxu * nx, ((nx * nxu) & 1) == 0);
        stage_node_tables(knode, P.m.K, s.cls, s.ns, nu * nx);
        stage_node_tables(krnode, P.m.KRcatT, s.cls, s.ns, nxu * nu);
        ct = ctab;
        ctt = cttab;
        kt = knode;
        krt = krnode;
    }
    if constexpr (SHARD) {   // the peers' packets: rows into qa, aux to the iterate (the top's kernel projection reads it), maxima
        const ShardPlan &sp = sh.sp;
        Ctrl *cw = const_cast<Ctrl *>(ctrl);
        const int parity = (int)(seq & 1ull), w = nx + 1;
        const long long stride = (long long)sp.cap * w + 6;
        const uint4 *mine = sh.px.ll[sp.rank] + (long long)parity * sp.world * stride;
        const int per = sp.cap * w + 6;   // packets polled per source (rows beyond the source's count are skipped)
        for (int i = threadIdx.x; i < (sp.world - 1) * per; i += blockDim.x) {
            int src = i / per;
            const int e = i - src * per;
            if (src >= sp.rank) ++src;
            const int lo = sp.cut_bounds[src], cnt = sp.cut_bounds[src + 1] - lo;
            const bool is_max = e >= sp.cap * w;
            if (!is_max && e >= cnt * w) continue;
            double v;
            if (!ll_load(mine + src * stride + e, (unsigned)seq, &v)) {
                sh_timeout = 1;
                v = 0.0;
            }
            if (is_max) {
                sh_max[src][e - sp.cap * w] = v;
            } else {
                const int c = e / w, k = e - c * w;
                if (k < nx) qa[(long long)(lo + c) * nx + k] = v;
                else sh.aux[sp.cut_first + lo + c] = v;
            }
        }
        stage_wait();   // (block barrier inside)
        if (threadIdx.x < 32) {   // global residual maxima (bit patterns of non-negative doubles), then k_check's body
            const int lane = threadIdx.x;
            const Ctrl c = *ctrl;
            unsigned long long best = lane < 6 ? (unsigned long long)__double_as_longlong(sh.slots[lane]) : 0ull;
            if (lane < 6)
                for (int rk = 0; rk < sp.world; ++rk) {
                    if (rk == sp.rank) continue;
                    const unsigned long long v = (unsigned long long)__double_as_longlong(sh_max[rk][lane]);
                    best = v > best ? v : best;
                }
            const double val = __longlong_as_double((long long)best);
            const bool testing = sh.check && c.pending;
            const bool nan = lane < 6 && val != val, bad = lane < 3 && !(val <= c.tol);
            if (lane < 6) {
                if (testing) {
                    if (c.hist && c.iters < c.hist_capacity) c.hist[(long long)c.iters * 6 + lane] = val;
                    sh.last[lane] = val;
                    if (sh.host_last && c.mirror) sh.host_last[lane] = val;
                    sh.slots[lane] = 0.0;
                } else {
                    sh.slots[lane] = val;
                }
            }
            const bool all_ok = !__any_sync(0xffffffffu, bad), any_nan = __any_sync(0xffffffffu, nan);
            if (lane == 0) {
                if (sh_timeout) atomicOr(&cw->status, 16);
                if (testing) {
                    if (any_nan) cw->status |= 2;
                    cw->iters = c.iters + 1;
                    cw->pending = 0;
                    if (c.iters >= c.max_iters || all_ok) cw->done = 1;
                }
                *sh.px.seq = seq + 1ull;
                sh_timeout = (testing && (c.iters >= c.max_iters || all_ok)) ? 2 : 0;   // 2: the loop has stopped
            }
        }
        __syncthreads();
        if (sh_timeout == 2) return;
    } else {
        stage_wait();
    }
    tree_backward<NX, NU, RES>(L, s, lv, xb, ub, qa, qb, scratch, rbuf, ct, kt, Q, R);
    // forward: the q ping-pong buffers of the backward pass become the x ping-pong
    __syncthreads();
    for (int k = threadIdx.x; k < nx; k += blockDim.x) {
        const double v = x0[(long long)blockIdx.x * nx + k];
        qa[k] = v;
        X[k] = v;
    }
    __syncthreads();
    tree_forward<NX, NU, RES>(L, s, lv, qa, qb, scratch, rbuf, ctt, krt, X, U);
}

// ---- level 0 backward + top + level 0 forward in ONE launch ---------------------------------------------------------------------
// CTAs 0..num_sub-1 own the level-0 subtrees, CTA num_sub the top of the tree (all co-resident: cooperative launch).  A
// subtree CTA stages everything for both directions once, walks backward, publishes its root's q and bumps a counter; the
// top CTA (which staged its own data meanwhile) waits for the counter, runs backward and forward, publishes x of the cut
// stage and flips a flag; the subtree CTAs then walk forward with r still in shared memory.  Two launches, two
// descriptor / table stagings and the r round trip through HBM disappear.  sync: [2 * instance] = counter,
// [2 * instance + 1] = flag generation (both self-resetting across launches).
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// block-wide copy of rows that ANOTHER CTA of this launch wrote (L2 path, never the non-coherent L1)
__device__ __forceinline__ void load_peer_rows(double *dst, const double *src, int count) {
    for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = __ldcg(src + i);
}

template <int NX, int NU, bool RES>
__global__ void __launch_bounds__(512) k_tree_fused(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl, TreeLevel lvs,
                                                   TreeLevel lvt, double *__restrict__ prim, double *__restrict__ q,
                                                   double *__restrict__ r, const double *__restrict__ x0, int *__restrict__ sync,
                                                   int *walk_count, int walk_tiles, int *tree_done) {
    // walk_count / walk_tiles / tree_done (null = off): the launch-overlap protocol with the chain walkers (chain_mma.cu).  The
    // kernel is then launched while the backward walker still runs: everything except the q of the chain heads is staged, the
    // subtree CTAs wait for walk_count == walk_tiles before they read it, and count themselves into tree_done once the x of the
    // chain heads is written (the forward walker, launched the same way behind this kernel, waits for that).
    extern __shared__ __align__(16) double tree_smem[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#ifdef RB_TRACE
    long long tr[10];
    int ntr = 0;
#define RB_STAMP() do { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); tr[ntr++] = t_; } while (0)
#else
#define RB_STAMP() do {} while (0)
#endif
    RB_STAMP();
    const Layout &L = P.L;
    const int nx = NX > 0 ? NX : L.nx, nu = NX > 0 ? NU : L.nu, nxu = nx + nu;
    const bool vx = (nx & 1) == 0, vu = (nu & 1) == 0;
    const bool is_top = blockIdx.x == (unsigned)lvs.num_sub;
    const TreeLevel &lv = is_top ? lvt : lvs;
    int *counter = sync + 2 * blockIdx.y, *flag = counter + 1;
    int flag0 = 0;   // the flag's value before this launch's top CTA can have flipped it (thread 0 of the subtree CTAs)
    Carve cv{tree_smem};
    const int *desc = stage_desc(lv, is_top ? 0 : blockIdx.x, cv, ctrl, is_top ? nullptr : flag, &flag0);
    if (!desc) return;
    const Sub s = sub_view(desc, lv);
    double *X = prim + (long long)blockIdx.y * L.np_pad + L.px, *U = prim + (long long)blockIdx.y * L.np_pad + L.pu;
    double *Q = q + (long long)blockIdx.y * L.n * nx, *R = r + (long long)blockIdx.y * L.m * nu;
    double *xb = cv.take((long long)lv.max_nodes * nx), *ub = cv.take((long long)lv.max_nodes * nu);
    double *qa = cv.take((long long)lv.max_row * nx), *qb = cv.take((long long)lv.max_row * nx);
    double *scratch = cv.take((long long)(blockDim.x >> 5) * 2 * nxu);
    double *rbuf = cv.take((long long)lv.max_nodes * nu);
    for (int d = 0; d < s.depth; ++d) {
        stage_rows(xb + s.off[d] * nx, X + (long long)s.lo[d] * nx, s.w[d] * nx, vx);
        if (s.cls[s.off[d]] >= 0) stage_rows(ub + s.off[d] * nu, U + (long long)s.lo[d] * nu, s.w[d] * nu, vu);
    }
    if (!is_top && !walk_count) stage_rows(qa, Q + (long long)s.ext_first * nx, s.ne * nx, vx);   // written by the previous launch
    const double *ct = P.m.ABcat, *ctt = P.m.ABcatT, *kt = P.m.K, *krt = P.m.KRcatT;
    if constexpr (RES) {
        double *ctab = cv.take((long long)lv.num_dyn * nx * nxu), *cttab = cv.take((long long)lv.num_dyn * nxu * nx);
        double *knode = cv.take((long long)lv.max_nodes * nu * nx), *krnode = cv.take((long long)lv.max_nodes * nxu * nu);
        stage_rows(ctab, P.m.ABcat, lv.num_dyn * nx * nxu, ((nx * nxu) & 1) == 0);
        stage_rows(cttab, P.m.ABcatT, lv.num_dyn * nxu * nx, ((nx * nxu) & 1) == 0);
        stage_node_tables(knode, P.m.K, s.cls, s.ns, nu * nx);
        stage_node_tables(krnode, P.m.KRcatT, s.cls, s.ns, nxu * nu);
        ct = ctab;
        ctt = cttab;
        kt = knode;
        krt = krnode;
    }
    RB_STAMP();
    stage_wait();
    RB_STAMP();
    if (!is_top) {
        if (walk_count) {   // the backward walker may still be running: wait for all of its tiles, then fetch the heads' q past the L1
            if (threadIdx.x == 0) {
                const long long t0 = clock64();
                while (ld_acquire(walk_count + blockIdx.y) < walk_tiles) {
                    __nanosleep(64);
                    if (clock64() - t0 > 400000000LL) {   // ~0.2 s: report (status bit 32) instead of hanging the GPU
                        atomicOr(const_cast<int *>(&ctrl->status), 32);
                        break;
                    }
                }
            }
            __syncthreads();
            load_peer_rows(qa, Q + (long long)s.ext_first * nx, s.ne * nx);
            __syncthreads();
        }
        tree_backward<NX, NU, RES>(L, s, lv, xb, ub, qa, qb, scratch, rbuf, ct, kt, Q, R);
        __threadfence();   // every thread publishes its part of the root's q before the CTA signals
        __syncthreads();
        RB_STAMP();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(counter, 1);
            while (ld_acquire(flag) == flag0) {}
        }
        __syncthreads();
        RB_STAMP();
        load_peer_rows(qa, X + (long long)s.lo[0] * nx, s.w[0] * nx);   // x of the root, from the top CTA
        __syncthreads();
        tree_forward<NX, NU, RES>(L, s, lv, qa, qb, scratch, rbuf, ctt, krt, X, U);
        if (tree_done) {   // the x of the chain heads below this subtree is written: the forward walker may read it
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) atomicAdd(tree_done + blockIdx.y, 1);
        }
        RB_STAMP();
#ifdef RB_TRACE
        if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 63) && ctrl && ctrl->iters == 40)
            printf("sub %d: desc+issue %lld wait %lld bwd %lld spin %lld fwd %lld ns\n", blockIdx.x, tr[1] - tr[0], tr[2] - tr[1],
                   tr[3] - tr[2], tr[4] - tr[3], tr[5] - tr[4]);
#endif
    } else {
        if (threadIdx.x == 0) {
            while (ld_acquire(counter) < lvs.num_sub) {}
            if (walk_count) walk_count[blockIdx.y] = 0;   // every subtree CTA is past its wait on the walker's tiles
        }
        __syncthreads();
        RB_STAMP();
        load_peer_rows(qa, Q + (long long)s.ext_first * nx, s.ne * nx);
        __syncthreads();
        RB_STAMP();
        tree_backward<NX, NU, RES>(L, s, lv, xb, ub, qa, qb, scratch, rbuf, ct, kt, Q, R);
        __syncthreads();
        RB_STAMP();
        for (int k = threadIdx.x; k < nx; k += blockDim.x) {
            const double v = x0[(long long)blockIdx.y * nx + k];
            qa[k] = v;
            X[k] = v;
        }
        __syncthreads();
        tree_forward<NX, NU, RES>(L, s, lv, qa, qb, scratch, rbuf, ctt, krt, X, U);
        __threadfence();   // every thread publishes its rows of the cut stage's x before the flag flips
        __syncthreads();
        RB_STAMP();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicExch(counter, 0);
            st_release(flag, ld_acquire(flag) + 1);
        }
#ifdef RB_TRACE
        if (threadIdx.x == 0 && ctrl && ctrl->iters == 40)
            printf("top: desc+issue %lld wait %lld spin %lld loadq %lld bwd %lld fwd %lld ns\n", tr[1] - tr[0], tr[2] - tr[1],
                   tr[3] - tr[2], tr[4] - tr[3], tr[5] - tr[4], tr[6] - tr[5]);
#endif
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------------------
#define RB_TREE_DIMS(X) X(2, 1) X(3, 2) X(4, 2) X(8, 4) X(10, 5) X(20, 10) X(64, 32)

// bytes of dynamic shared memory of the three kernels for a level (the maximum: one attribute for all)
size_t tree_smem_bytes(const TreeLevel &lv, int nx, int nu, int warps, bool top) {
    auto ev = [](long long c) { return (size_t)((c + 1) & ~1LL); };
    const long long nxu = nx + nu;
    size_t tot = ev((lv.desc_stride + 1) / 2);
    const size_t scr = ev((long long)warps * 2 * nxu);
    size_t bwd = ev((long long)lv.max_nodes * nx) + ev((long long)lv.max_nodes * nu) + 2 * ev((long long)lv.max_row * nx) + scr;
    size_t fwd = ev((long long)lv.max_nodes * nu) + 2 * ev((long long)lv.max_row * nx) + scr;
    const size_t tab = lv.resident ? ev((long long)lv.num_dyn * nx * nxu) : 0;
    const size_t kb = lv.resident ? ev((long long)lv.max_nodes * nu * nx) : 0, kf = lv.resident ? ev((long long)lv.max_nodes * nxu * nu) : 0;
    if (top) return (tot + bwd + ev((long long)lv.max_nodes * nu) + 2 * tab + kb + kf) * sizeof(double);
    return (tot + std::max(bwd + tab + kb, fwd + tab + kf)) * sizeof(double);
}

cudaError_t tree_kernels_set_smem(int bytes) {
    cudaError_t e = cudaSuccess;
#define RB_SET1(NX, NU, RES)                                                                                                  \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tree_bwd<NX, NU, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tree_fwd<NX, NU, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tree_top<NX, NU, RES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tree_top<NX, NU, RES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tree_fused<NX, NU, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tree_fused<NX, NU, RES>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
#define RB_SET(NX, NU) RB_SET1(NX, NU, true) RB_SET1(NX, NU, false)
    RB_TREE_DIMS(RB_SET)
    RB_SET(0, 0)
#undef RB_SET1
#undef RB_SET
    return e;
}

void launch_tree_bwd(dim3 grid, int threads, size_t smem, cudaStream_t st, const Params &P, const Ctrl *ctrl, const TreeLevel &lv,
                     const double *prim, double *q, double *r) {
#define RB_GO(NX, NU)                                                                    \
    if (P.L.nx == NX && P.L.nu == NU) {                                                  \
        if (lv.resident) k_tree_bwd<NX, NU, true><<<grid, threads, smem, st>>>(P, ctrl, lv, prim, q, r);   \
        else k_tree_bwd<NX, NU, false><<<grid, threads, smem, st>>>(P, ctrl, lv, prim, q, r);        \
        return;                                                                          \
    }
    RB_TREE_DIMS(RB_GO)
#undef RB_GO
    if (lv.resident) k_tree_bwd<0, 0, true><<<grid, threads, smem, st>>>(P, ctrl, lv, prim, q, r);
    else k_tree_bwd<0, 0, false><<<grid, threads, smem, st>>>(P, ctrl, lv, prim, q, r);
}

void launch_tree_fwd(dim3 grid, int threads, size_t smem, cudaStream_t st, const Params &P, const Ctrl *ctrl, const TreeLevel &lv,
                     double *prim, const double *r) {
#define RB_GO(NX, NU)                                                                    \
    if (P.L.nx == NX && P.L.nu == NU) {                                                  \
        if (lv.resident) k_tree_fwd<NX, NU, true><<<grid, threads, smem, st>>>(P, ctrl, lv, prim, r);   \
        else k_tree_fwd<NX, NU, false><<<grid, threads, smem, st>>>(P, ctrl, lv, prim, r);           \
        return;                                                                          \
    }
    RB_TREE_DIMS(RB_GO)
#undef RB_GO
    if (lv.resident) k_tree_fwd<0, 0, true><<<grid, threads, smem, st>>>(P, ctrl, lv, prim, r);
    else k_tree_fwd<0, 0, false><<<grid, threads, smem, st>>>(P, ctrl, lv, prim, r);
}

void launch_tree_top(int grid, int threads, size_t smem, cudaStream_t st, const Params &P, const Ctrl *ctrl, const TreeLevel &lv,
                     double *prim, double *q, double *r, const double *x0) {
#define RB_GO(NX, NU)                                                                    \
    if (P.L.nx == NX && P.L.nu == NU) {                                                  \
        if (lv.resident) k_tree_top<NX, NU, true, false><<<grid, threads, smem, st>>>(P, ctrl, lv, prim, q, r, x0, ShardHand{});   \
        else k_tree_top<NX, NU, false, false><<<grid, threads, smem, st>>>(P, ctrl, lv, prim, q, r, x0, ShardHand{});    \
        return;                                                                          \
    }
    RB_TREE_DIMS(RB_GO)
#undef RB_GO
    if (lv.resident) k_tree_top<0, 0, true, false><<<grid, threads, smem, st>>>(P, ctrl, lv, prim, q, r, x0, ShardHand{});
    else k_tree_top<0, 0, false, false><<<grid, threads, smem, st>>>(P, ctrl, lv, prim, q, r, x0, ShardHand{});
}

void launch_tree_top_sharded(int threads, size_t smem, cudaStream_t st, const Params &P, Ctrl *ctrl, const TreeLevel &lv, double *prim,
                             double *q, double *r, const double *x0, const ShardHand &sh) {
#define RB_GO(NX, NU)                                                                    \
    if (P.L.nx == NX && P.L.nu == NU) {                                                  \
        if (lv.resident) k_tree_top<NX, NU, true, true><<<1, threads, smem, st>>>(P, ctrl, lv, prim, q, r, x0, sh);   \
        else k_tree_top<NX, NU, false, true><<<1, threads, smem, st>>>(P, ctrl, lv, prim, q, r, x0, sh);    \
        return;                                                                          \
    }
    RB_TREE_DIMS(RB_GO)
#undef RB_GO
    if (lv.resident) k_tree_top<0, 0, true, true><<<1, threads, smem, st>>>(P, ctrl, lv, prim, q, r, x0, sh);
    else k_tree_top<0, 0, false, true><<<1, threads, smem, st>>>(P, ctrl, lv, prim, q, r, x0, sh);
}

// cooperative launch: all (num_sub + 1) x batch CTAs must be co-resident (the top CTA spins on the others)
// walk_count != null: launched with programmatic stream serialization behind the backward chain walker instead (all CTAs of a
// programmatic dependent become resident together once the walker's CTAs are running: the co-residency the hand-off needs)
template <typename K>
static cudaError_t launch_coop(K kernel, dim3 grid, int threads, size_t smem, cudaStream_t st, const Params &P, const Ctrl *ctrl,
                               const TreeLevel &lvs, const TreeLevel &lvt, double *prim, double *q, double *r, const double *x0,
                               int *sync, int *walk_count, int walk_tiles, int *tree_done) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (walk_count) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
    } else {
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
    }
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, P, ctrl, lvs, lvt, prim, q, r, x0, sync, walk_count, walk_tiles, tree_done);
}

cudaError_t launch_tree_fused(int batch, int threads, size_t smem, cudaStream_t st, const Params &P, const Ctrl *ctrl,
                              const TreeLevel &lvs, const TreeLevel &lvt, double *prim, double *q, double *r, const double *x0,
                              int *sync, int *walk_count, int walk_tiles, int *tree_done) {
    const dim3 grid(lvs.num_sub + 1, batch);
#define RB_GO(NX, NU)                                                                                                       \
    if (P.L.nx == NX && P.L.nu == NU)                                                                                        \
        return lvs.resident ? launch_coop(k_tree_fused<NX, NU, true>, grid, threads, smem, st, P, ctrl, lvs, lvt, prim, q, r, x0, sync, walk_count, walk_tiles, tree_done) \
                            : launch_coop(k_tree_fused<NX, NU, false>, grid, threads, smem, st, P, ctrl, lvs, lvt, prim, q, r, x0, sync, walk_count, walk_tiles, tree_done);
    RB_TREE_DIMS(RB_GO)
#undef RB_GO
    return lvs.resident ? launch_coop(k_tree_fused<0, 0, true>, grid, threads, smem, st, P, ctrl, lvs, lvt, prim, q, r, x0, sync, walk_count, walk_tiles, tree_done)
                        : launch_coop(k_tree_fused<0, 0, false>, grid, threads, smem, st, P, ctrl, lvs, lvt, prim, q, r, x0, sync, walk_count, walk_tiles, tree_done);
}

// can the fused launch run?  (co-residency of all CTAs at this block size and shared-memory footprint)
bool tree_fused_fits(int nx, int nu, bool resident, int threads, size_t smem, int ctas) {
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return false;
    cudaError_t e = cudaErrorInvalidValue;
#define RB_OCC(NX, NU)                                                                                                   \
    if (e != cudaSuccess && nx == NX && nu == NU)                                                                        \
        e = resident ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tree_fused<NX, NU, true>, threads, smem) \
                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tree_fused<NX, NU, false>, threads, smem);
    RB_TREE_DIMS(RB_OCC)
#undef RB_OCC
    if (e != cudaSuccess)
        e = resident ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tree_fused<0, 0, true>, threads, smem)
                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tree_fused<0, 0, false>, threads, smem);
    return e == cudaSuccess && (long long)per_sm * sms >= ctas;
}

}  // namespace rb
