// views.cuh -- segment views of the iterates and the shared-memory stager used by the tiled passes (fused.cu, lane.cu).
#pragma once
#include "common.cuh"

namespace rb {

// segment base pointers; every pointer is indexed with GLOBAL indices (node * nx + k, yoff[i] + e, edge j - 1, leaf
// index ...) whether it points into HBM or into a re-based shared-memory chunk
struct PrimalView {
    const double *x, *u, *y, *tau, *s;
};
struct DualView {
    const double *d1, *d2, *d3, *d4, *d5, *d6, *d7, *d11, *d12, *d13, *d14;
    const double *d2c;   // d2 of the CHILDREN of the tile's nodes (primal pass only)
};
struct PrimalOut {
    double *x, *u, *y, *tau, *s;
};
struct DualOut {
    double *d1, *d2, *d3, *d4, *d5, *d6, *d7, *d11, *d12, *d13, *d14;
};

__device__ __forceinline__ PrimalView primal_view(const Layout &L, const double *p) {
    return {p + L.px, p + L.pu, p + L.py, p + L.ptau, p + L.ps};
}
__device__ __forceinline__ PrimalOut primal_out(const Layout &L, double *p) {
    return {p + L.px, p + L.pu, p + L.py, p + L.ptau, p + L.ps};
}
__device__ __forceinline__ DualView dual_view(const Layout &L, const double *d) {
    return {d + L.d1, d + L.d2, d + L.d3, d + L.d4, d + L.d5, d + L.d6, d + L.d7, d + L.d11, d + L.d12, d + L.d13,
            d + L.d14, d + L.d2};
}
__device__ __forceinline__ DualOut dual_out(const Layout &L, double *d) {
    return {d + L.d1, d + L.d2, d + L.d3, d + L.d4, d + L.d5, d + L.d6, d + L.d7, d + L.d11, d + L.d12, d + L.d13, d + L.d14};
}

// Asynchronous staging: cp.async (LDGSTS) 16-byte copies global -> shared, no register round trip, so every thread has
// all of its copies of all chunks in flight before the single wait.
__device__ __forceinline__ void cp_async16(double *smem_dst, const double *gsrc) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// bump allocator over the CTA's dynamic shared memory; stage() enqueues the copy of [first, first+count) of a segment
// (widened to 16-byte boundaries: segment bases are 128-byte aligned and padded) and returns a pointer re-based so that
// the segment's GLOBAL indices work.  finish() waits for all copies of the CTA.
struct Stager {
    double *cursor;
    __device__ __forceinline__ const double *stage(const double *seg, long long first, long long count) {
        if (count <= 0) return seg;
        const long long a = first & ~1LL, end = (first + count + 1) & ~1LL;
        double *dst = cursor;
        cursor += end - a;
        const double *src = seg + a;
        for (long long i = 2 * (long long)threadIdx.x; i < end - a; i += 2 * (long long)blockDim.x) cp_async16(dst + i, src + i);
        return dst - a;
    }
    __device__ __forceinline__ void finish() {
        cp_async_wait_all();
        __syncthreads();
    }
};

// write [first, first+count) of a staged (re-based) chunk back to its global segment with all threads, coalesced
__device__ __forceinline__ void unstage(double *__restrict__ gseg, const double *staged, long long first, long long count) {
    for (long long i = threadIdx.x; i < count; i += blockDim.x) gseg[first + i] = staged[first + i];
}

}  // namespace rb
