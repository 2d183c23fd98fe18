// ops.cu -- the reference's operator / prox methods one by one, as stand-alone sm_100a kernels (FP64).
//
// These back the step-by-step API (Operator.ell / ell_transpose, the Cache.* methods and the four Solver half steps);
// Solver.chock's loop uses the fused kernels of fused.cu instead.  One warp owns one tree node: node-major rows are
// read/written with consecutive lanes, child->parent sums are accumulated by the parent's warp over its (contiguous)
// children, per-node vectors are staged in warp-private shared rows.
#include "kernels.cuh"
#include "node_ops.cuh"

namespace rb {

// ----------------------------------------------------------------------------------------------------------------
// out_p = beta * base_p + gamma * L*(d)        (reference operators.py:55-94; primal half step solver.py:27-39)
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_lt_axpby(const __grid_constant__ Params P, const double *__restrict__ dual,
                                                      const double *__restrict__ base, double *__restrict__ out,
                                                      double beta, double gamma) {
    const Layout &L = P.L;
    __shared__ double sm[kWarpsPerBlock][4][kMaxDim];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int node = blockIdx.x * kWarpsPerBlock + warp;
    if (node >= L.n) return;
    const double *D = dual + (long long)blockIdx.y * L.nd_pad;
    const double *Bp = base ? base + (long long)blockIdx.y * L.np_pad : nullptr;
    double *O = out + (long long)blockIdx.y * L.np_pad;
    double *v3 = sm[warp][0], *v4 = sm[warp][1], *ax = sm[warp][2], *au = sm[warp][3];
    const int nx = L.nx, nu = L.nu;
    auto blend = [&](long long idx, double val) { O[idx] = (Bp ? beta * Bp[idx] : 0.0) + gamma * val; };

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
            if (lane == 0) blend(L.ptau + j, 0.5 * (D[L.d5 + e] + D[L.d6 + e]));
        }
        for (int k = lane; k < nx; k += 32) blend(L.px + (long long)node * nx + k, ax[k]);
        for (int k = lane; k < nu; k += 32) blend(L.pu + (long long)node * nu + k, au[k]);
        const double d2v = D[L.d2 + node];
        const int yo = P.t.yoff[node];
        for (int e = lane; e < 2 * cc + 1; e += 32) {
            const double b = e < cc ? P.t.cond_prob[c0 + e] : (e == 2 * cc ? 1.0 : 0.0);
            blend(L.py + yo + e, D[L.d1 + yo + e] - b * d2v);
        }
        if (lane == 0) blend(L.ps + node, d2v);
    } else {
        const long long li = node - L.m;
        for (int k = lane; k < nx; k += 32) {
            v3[k] = D[L.d11 + li * nx + k];
            ax[k] = L.has_leaf_rect ? D[L.d14 + li * nx + k] : 0.0;
        }
        __syncwarp();
        // reference order: sqrtQf @ d11 first, then + Gamma' d14 (operators.py:89-92); addition commutes exactly
        mv_acc(P.m.sqfT + (long long)P.t.leafcost_idx[li] * nx * nx, v3, nx, nx, ax, 1.0, lane);
        __syncwarp();
        for (int k = lane; k < nx; k += 32) blend(L.px + (long long)node * nx + k, ax[k]);
        if (lane == 0) blend(L.ps + node, 0.5 * (D[L.d12 + li] + D[L.d13 + li]));
    }
    // tau_0 is never written by the reference's ell_transpose: it keeps the template value
    if (node == 0 && lane == 0) O[L.ptau] = Bp ? Bp[L.ptau] * (beta + gamma) : 0.0;
}

// ----------------------------------------------------------------------------------------------------------------
// out_d = beta * base_d + gamma * L(c1 * p1 + c2 * p2)     (operators.py:19-53; dual half step solver.py:44-58)
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_l_axpby(const __grid_constant__ Params P, const double *__restrict__ p1,
                                                     const double *__restrict__ p2, double c1, double c2,
                                                     const double *__restrict__ base, double *__restrict__ out,
                                                     double beta, double gamma) {
    const Layout &L = P.L;
    __shared__ double sm[kWarpsPerBlock][3][kMaxDim];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int node = blockIdx.x * kWarpsPerBlock + warp;
    if (node >= L.n) return;
    const double *P1 = p1 + (long long)blockIdx.y * L.np_pad;
    const double *P2 = p2 ? p2 + (long long)blockIdx.y * L.np_pad : nullptr;
    const double *Bd = base ? base + (long long)blockIdx.y * L.nd_pad : nullptr;
    double *O = out + (long long)blockIdx.y * L.nd_pad;
    double *zx = sm[warp][0], *zu = sm[warp][1], *res = sm[warp][2];
    const int nx = L.nx, nu = L.nu;
    auto z = [&](long long idx) { return c1 * P1[idx] + (P2 ? c2 * P2[idx] : 0.0); };
    auto blend = [&](long long idx, double val) { O[idx] = (Bd ? beta * Bd[idx] : 0.0) + gamma * val; };

    for (int k = lane; k < nx; k += 32) zx[k] = z(L.px + (long long)node * nx + k);
    if (node < L.m) {
        for (int k = lane; k < nu; k += 32) zu[k] = z(L.pu + (long long)node * nu + k);
        __syncwarp();
        if (L.has_nl_rect) {
            for (int k = lane; k < nx; k += 32) blend(L.d7 + (long long)node * L.nxu + k, zx[k]);
            for (int k = lane; k < nu; k += 32) blend(L.d7 + (long long)node * L.nxu + nx + k, zu[k]);
        }
        const int c0 = P.t.child_first[node], cc = P.t.child_count[node];
        for (int j = c0; j < c0 + cc; ++j) {
            const long long e = j - 1;
            const int ci = P.t.cost_idx[j];
            mv_set(P.m.sqT + (long long)ci * nx * nx, zx, nx, nx, res, lane);
            for (int k = lane; k < nx; k += 32) blend(L.d3 + e * nx + k, res[k]);
            __syncwarp();
            mv_set(P.m.srT + (long long)ci * nu * nu, zu, nu, nu, res, lane);
            for (int k = lane; k < nu; k += 32) blend(L.d4 + e * nu + k, res[k]);
            __syncwarp();
            if (lane == 0) {
                const double half_tau = 0.5 * z(L.ptau + j);
                blend(L.d5 + e, half_tau);
                blend(L.d6 + e, half_tau);
            }
        }
        const int yo = P.t.yoff[node];
        double dot = 0.0;
        for (int e = lane; e < 2 * cc + 1; e += 32) {
            const double zy = z(L.py + yo + e);
            blend(L.d1 + yo + e, zy);
            const double b = e < cc ? P.t.cond_prob[c0 + e] : (e == 2 * cc ? 1.0 : 0.0);
            dot = fma(b, zy, dot);
        }
        dot = warp_sum(dot);
        if (lane == 0) blend(L.d2 + node, z(L.ps + node) - dot);
    } else {
        __syncwarp();
        const long long li = node - L.m;
        mv_set(P.m.sqfT + (long long)P.t.leafcost_idx[li] * nx * nx, zx, nx, nx, res, lane);
        for (int k = lane; k < nx; k += 32) {
            blend(L.d11 + li * nx + k, res[k]);
            if (L.has_leaf_rect) blend(L.d14 + li * nx + k, zx[k]);
        }
        if (lane == 0) {
            const double half_s = 0.5 * z(L.ps + node);
            blend(L.d12 + li, half_s);
            blend(L.d13 + li, half_s);
        }
    }
}

// the projection onto the dynamics set (cache.py:259-288) lives in sweeps.cu; the kernel projection follows
__global__ void __launch_bounds__(kThreads) k_kernel_proj(const __grid_constant__ Params P, double *__restrict__ prim) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int node = blockIdx.x * kWarpsPerBlock + warp;
    if (node >= P.L.m) return;
    kernel_projection(P, prim + (long long)blockIdx.y * P.L.np_pad, node, lane);
}

// s_0 -= alpha (cache.py:253-257)
__global__ void k_s0_shift(const __grid_constant__ Params P, double *__restrict__ prim, double alpha) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < P.L.batch) prim[(long long)b * P.L.np_pad + P.L.ps] -= alpha;
}

// ----------------------------------------------------------------------------------------------------------------
// prox of g* pieces (cache.py:321-393).  mode bits: 1 = divide by alpha (modify_dual), 2 = add halves,
// 4 = project nonleaf blocks, 8 = project leaf blocks, 16 = Moreau step d = alpha (w - proj) with w = the value
// before projection.  All bits = proximal_of_g_conjugate.  Bits 4/8 alone overwrite d with the projection.
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_prox_g(const __grid_constant__ Params P, double *__restrict__ dual,
                                                    double alpha, int mode, int *__restrict__ status) {
    const Layout &L = P.L;
    __shared__ double sm[kWarpsPerBlock][2 * kMaxDim + 2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int node = blockIdx.x * kWarpsPerBlock + warp;
    if (node >= L.n) return;
    double *D = dual + (long long)blockIdx.y * L.nd_pad;
    double *w = sm[warp];
    const int nx = L.nx, nu = L.nu;
    const bool scale = mode & 1, halves = mode & 2, pnl = mode & 4, plf = mode & 8, moreau = mode & 16;
    int bad = 0;
    auto pre = [&](double v, double shift) { return (scale ? v / alpha : v) + (halves ? shift : 0.0); };
    auto post = [&](double wv, double zv, bool projected) {
        if (!projected) return wv;               // block not projected in this call: keep (possibly scaled) value
        return moreau ? alpha * (wv - zv) : zv;
    };

    if (node > 0) {  // edge into this node: SOC on [d3; d4; d5; d6], t = d6  (cache.py:354-365)
        const long long e = node - 1;
        const int dim = nx + nu + 2;
        for (int k = lane; k < nx; k += 32) w[k] = pre(D[L.d3 + e * nx + k], 0.0);
        for (int k = lane; k < nu; k += 32) w[nx + k] = pre(D[L.d4 + e * nu + k], 0.0);
        if (lane == 0) {
            w[nx + nu] = pre(D[L.d5 + e], -0.5);
            w[nx + nu + 1] = pre(D[L.d6 + e], 0.5);
        }
        __syncwarp();
        SocResult sr;
        if (pnl) sr = soc_classify(w, dim, lane);
        for (int k = lane; k < nx; k += 32) D[L.d3 + e * nx + k] = post(w[k], pnl ? soc_entry(sr, w[k], false) : 0.0, pnl);
        for (int k = lane; k < nu; k += 32)
            D[L.d4 + e * nu + k] = post(w[nx + k], pnl ? soc_entry(sr, w[nx + k], false) : 0.0, pnl);
        if (lane == 0) {
            D[L.d5 + e] = post(w[nx + nu], pnl ? soc_entry(sr, w[nx + nu], false) : 0.0, pnl);
            D[L.d6 + e] = post(w[nx + nu + 1], pnl ? soc_entry(sr, w[nx + nu + 1], true) : 0.0, pnl);
        }
        __syncwarp();
    }
    if (node < L.m) {
        const int cc = P.t.child_count[node];
        const int yo = P.t.yoff[node];
        // dual of R_+^{2c} x {0}: max(0,.) on the first 2c entries, identity on the last (risks.py:32-33)
        for (int e = lane; e < 2 * cc + 1; e += 32) {
            const double wv = pre(D[L.d1 + yo + e], 0.0);
            const double zv = e < 2 * cc ? fmax(0.0, wv) : wv;
            D[L.d1 + yo + e] = post(wv, zv, pnl);
        }
        if (lane == 0) {
            const double wv = pre(D[L.d2 + node], 0.0);
            D[L.d2 + node] = post(wv, fmax(0.0, wv), pnl);
        }
        if (L.has_nl_rect) {
            const long long ri = (long long)P.t.nl_rect_idx[node] * L.nxu;
            for (int k = lane; k < L.nxu; k += 32) {
                const double wv = pre(D[L.d7 + (long long)node * L.nxu + k], 0.0);
                const double zv = pnl ? box_clip(wv, P.m.nl_lo[ri + k], P.m.nl_hi[ri + k], &bad) : 0.0;
                D[L.d7 + (long long)node * L.nxu + k] = post(wv, zv, pnl);
            }
        }
    } else {  // leaf: SOC on [d11; d12; d13] (cache.py:375-386), box on d14
        const long long li = node - L.m;
        const int dim = nx + 2;
        for (int k = lane; k < nx; k += 32) w[k] = pre(D[L.d11 + li * nx + k], 0.0);
        if (lane == 0) {
            w[nx] = pre(D[L.d12 + li], -0.5);
            w[nx + 1] = pre(D[L.d13 + li], 0.5);
        }
        __syncwarp();
        SocResult sr;
        if (plf) sr = soc_classify(w, dim, lane);
        for (int k = lane; k < nx; k += 32) D[L.d11 + li * nx + k] = post(w[k], plf ? soc_entry(sr, w[k], false) : 0.0, plf);
        if (lane == 0) {
            D[L.d12 + li] = post(w[nx], plf ? soc_entry(sr, w[nx], false) : 0.0, plf);
            D[L.d13 + li] = post(w[nx + 1], plf ? soc_entry(sr, w[nx + 1], true) : 0.0, plf);
        }
        if (L.has_leaf_rect) {
            const long long ri = (long long)P.t.leaf_rect_idx[li] * nx;
            for (int k = lane; k < nx; k += 32) {
                const double wv = pre(D[L.d14 + li * nx + k], 0.0);
                const double zv = plf ? box_clip(wv, P.m.leaf_lo[ri + k], P.m.leaf_hi[ri + k], &bad) : 0.0;
                D[L.d14 + li * nx + k] = post(wv, zv, plf);
            }
        }
    }
    if (bad) atomicOr(status, 1);
}

// ----------------------------------------------------------------------------------------------------------------
// element-wise helpers on whole padded buffers
// ----------------------------------------------------------------------------------------------------------------
// out = a * x + b * y  (y may be null)
__global__ void k_axpby(double *__restrict__ out, double a, const double *__restrict__ x, double b,
                        const double *__restrict__ y, long long count) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        out[i] = a * x[i] + (y ? b * y[i] : 0.0);
}
// out = x / a + b * y      (residual formulas divide by alpha like the reference, solver.py:69,76)
__global__ void k_div_add(double *__restrict__ out, const double *__restrict__ x, double a, double b,
                          const double *__restrict__ y, long long count) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        out[i] = x[i] / a + b * y[i];
}
// per-instance inf-norm: slot[inst] = max |x| over the instance's padded buffer (padding is zero)
__global__ void k_absmax(const double *__restrict__ x, long long stride, double *__restrict__ slot, int slot_stride,
                         int *__restrict__ status) {
    const double *X = x + (long long)blockIdx.y * stride;
    double best = 0.0;
    int nan = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < stride; i += (long long)gridDim.x * blockDim.x) {
        const double a = fabs(X[i]);
        nan |= (a != a);
        best = fmax(best, a);
    }
    best = warp_max(best);
    nan = __any_sync(0xffffffffu, nan);
    if ((threadIdx.x & 31) == 0) {
        atomic_max_nonneg(slot + (long long)blockIdx.y * slot_stride, best);
        if (nan) atomicOr(status, 2);
    }
}

// stand-alone cone / box projections of one host vector (cones.py, rectangle.py): a single warp
__global__ void k_cone(int cone, int dim, const double *__restrict__ in, double *__restrict__ out) {
    extern __shared__ double w[];
    const int lane = threadIdx.x;
    if (cone == 3) {
        for (int k = lane; k < dim; k += 32) w[k] = in[k];
        __syncwarp();
        const SocResult sr = soc_classify(w, dim, lane);
        for (int k = lane; k < dim; k += 32) out[k] = soc_entry(sr, w[k], k == dim - 1);
        return;
    }
    for (int k = lane; k < dim; k += 32) {
        const double v = in[k];
        out[k] = cone == 0 ? v : (cone == 1 ? 0.0 : fmax(0.0, v));
    }
}
__global__ void k_box(int dim, const double *__restrict__ in, const double *__restrict__ lo,
                      const double *__restrict__ hi, double *__restrict__ out, int *__restrict__ status) {
    int bad = 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < dim; k += gridDim.x * blockDim.x)
        out[k] = box_clip(in[k], lo[k], hi[k], &bad);
    if (bad) atomicOr(status, 1);
}

// The operator tables of a problem (per-mode dynamics, per-class K and R~^-1, the tensor-core fragment images: 0.2 - 6 MB) are read by
// every sweep kernel through chains of dependent loads; whenever they have dropped out of L2 (the bench flushes it before every
// iteration; in a real solve the 100+ MB of iterates of a large tree do it) each of those loads is a DRAM round trip on the critical
// path.  One small launch at the head of the iteration, on the side stream, pulls them back in.
__global__ void k_prefetch_ranges(const Ctrl *__restrict__ ctrl, const PrefetchRange *__restrict__ ranges, int count) {
    if (ctrl && ctrl->done) return;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (long long)gridDim.x * blockDim.x;
    for (int r = 0; r < count; ++r) {
        const char *p = ranges[r].ptr;
        const long long lines = (ranges[r].bytes + 127) / 128;
        for (long long i = tid; i < lines; i += nthreads) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + i * 128));
    }
}

}  // namespace rb
