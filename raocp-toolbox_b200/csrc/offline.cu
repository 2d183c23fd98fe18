// offline.cu -- Cache._offline on the device (reference cache.py:200-242) and lambda_max(L* L).
//
// offline_projection_dynamics (cache.py:207-233), one CTA per factorisation class, classes of one tree level per
// launch (a class = nodes sharing the same (dynamics, child-class) list, hence the same P, K, R~; with one class per
// node this is exactly the reference's backward loop):
//     R~ = I + sum_j B_j' P_j B_j,   S = sum_j B_j' P_j A_j,   R~ = C C' (Cholesky),   K = -R~^-1 S,
//     P  = I + K'K + sum_j (A_j + B_j K)' P_j (A_j + B_j K),    P_leaf = I.
// All matrices live in dynamic shared memory (nx, nu <= 64 -> <= 176 KB of the 227 KB a CTA may use on sm_100a).
// offline_projection_kernel (cache.py:235-242) has no device counterpart: the AVaR kernel projector is closed-form
// (node_ops.cuh kernel_projection).
#include "kernels.cuh"

namespace rb {

constexpr int kOffThreads = 256;

size_t offline_smem_bytes(int nx, int nu) {
    // PB (nx*nu) | Rt (nu*nu) | S/K (nu*nx) | Rinv (nu*nu) | Abar (nx*nx) | T (nx*nx) | Pacc (nx*nx)
    return sizeof(double) * ((size_t)nx * nu + 2 * (size_t)nu * nu + (size_t)nu * nx + 3 * (size_t)nx * nx);
}

__global__ void __launch_bounds__(kOffThreads) k_offline_level(const __grid_constant__ Params P, ClassView cv,
                                                              int level_begin, int level_count,
                                                              double *__restrict__ Ptab, double *__restrict__ Ktab,
                                                              double *__restrict__ KRcatT, int *__restrict__ status) {
    extern __shared__ double smem[];
    const int nx = P.L.nx, nu = P.L.nu;
    double *PB = smem;                 // nx x nu
    double *Rt = PB + nx * nu;         // nu x nu  (Cholesky factor in the lower triangle afterwards)
    double *Km = Rt + nu * nu;         // nu x nx  (S, then K)
    double *Ri = Km + nu * nx;         // nu x nu  (R~^-1)
    double *Ab = Ri + nu * nu;         // nx x nx
    double *T = Ab + nx * nx;          // nx x nx
    double *Pa = T + nx * nx;          // nx x nx
    const int tid = threadIdx.x, nt = blockDim.x;
    if ((int)blockIdx.x >= level_count) return;
    const int c = cv.level_list[level_begin + blockIdx.x];
    const int k0 = cv.child_ptr[c], k1 = cv.child_ptr[c + 1];

    for (int i = tid; i < nu * nu; i += nt) Rt[i] = (i / nu == i % nu) ? 1.0 : 0.0;
    for (int i = tid; i < nu * nx; i += nt) Km[i] = 0.0;
    __syncthreads();
    for (int kk = k0; kk < k1; ++kk) {
        const double *A = P.m.A + (long long)cv.child_dyn[kk] * nx * nx;
        const double *B = P.m.B + (long long)cv.child_dyn[kk] * nx * nu;
        const int cj = cv.child_cls[kk];
        const double *Pj = cj >= 0 ? Ptab + (long long)cj * nx * nx : nullptr;
        // PB = P_j B
        for (int i = tid; i < nx * nu; i += nt) {
            const int r = i / nu, a = i % nu;
            double acc;
            if (Pj) {
                acc = 0.0;
                for (int l = 0; l < nx; ++l) acc = fma(Pj[r * nx + l], B[l * nu + a], acc);
            } else {
                acc = B[r * nu + a];
            }
            PB[i] = acc;
        }
        __syncthreads();
        // Rt += B' (P B);  S += (P B)' A   [ = B' P A since P is symmetric ]
        for (int i = tid; i < nu * nu; i += nt) {
            const int a = i / nu, b = i % nu;
            double acc = 0.0;
            for (int l = 0; l < nx; ++l) acc = fma(B[l * nu + a], PB[l * nu + b], acc);
            Rt[i] += acc;
        }
        for (int i = tid; i < nu * nx; i += nt) {
            const int a = i / nx, k = i % nx;
            double acc = 0.0;
            for (int l = 0; l < nx; ++l) acc = fma(PB[l * nu + a], A[l * nx + k], acc);
            Km[i] += acc;
        }
        __syncthreads();
    }
    // keep a copy of R~ in Ri's place? not needed: Cholesky in place, lower triangle
    for (int col = 0; col < nu; ++col) {
        if (tid == 0) {
            const double piv = Rt[col * nu + col];
            if (!(piv > 0.0)) atomicOr(status, 4);
            Rt[col * nu + col] = sqrt(piv);
        }
        __syncthreads();
        const double d = Rt[col * nu + col];
        for (int r = col + 1 + tid; r < nu; r += nt) Rt[r * nu + col] /= d;
        __syncthreads();
        for (int i = tid; i < (nu - col - 1) * (nu - col - 1); i += nt) {
            const int r = col + 1 + i / (nu - col - 1), cc = col + 1 + i % (nu - col - 1);
            if (cc <= r) Rt[r * nu + cc] -= Rt[r * nu + col] * Rt[cc * nu + col];
        }
        __syncthreads();
    }
    // K = -R~^-1 S and R~^-1: one thread per right-hand-side column, forward then backward substitution
    for (int col = tid; col < nx + nu; col += nt) {
        const bool is_k = col < nx;
        double *M = is_k ? Km : Ri;
        const int ld = is_k ? nx : nu, cidx = is_k ? col : col - nx;
        if (!is_k)
            for (int r = 0; r < nu; ++r) M[r * ld + cidx] = (r == cidx) ? 1.0 : 0.0;
        for (int r = 0; r < nu; ++r) {   // C z = rhs
            double acc = is_k ? -M[r * ld + cidx] : M[r * ld + cidx];
            for (int l = 0; l < r; ++l) acc -= Rt[r * nu + l] * M[l * ld + cidx];
            M[r * ld + cidx] = acc / Rt[r * nu + r];
        }
        for (int r = nu - 1; r >= 0; --r) {   // C' x = z
            double acc = M[r * ld + cidx];
            for (int l = r + 1; l < nu; ++l) acc -= Rt[l * nu + r] * M[l * ld + cidx];
            M[r * ld + cidx] = acc / Rt[r * nu + r];
        }
    }
    __syncthreads();
    // P = I + K'K + sum_j Abar_j' P_j Abar_j
    for (int i = tid; i < nx * nx; i += nt) {
        const int r = i / nx, k = i % nx;
        double acc = (r == k) ? 1.0 : 0.0;
        for (int a = 0; a < nu; ++a) acc = fma(Km[a * nx + r], Km[a * nx + k], acc);
        Pa[i] = acc;
    }
    __syncthreads();
    for (int kk = k0; kk < k1; ++kk) {
        const double *A = P.m.A + (long long)cv.child_dyn[kk] * nx * nx;
        const double *B = P.m.B + (long long)cv.child_dyn[kk] * nx * nu;
        const int cj = cv.child_cls[kk];
        const double *Pj = cj >= 0 ? Ptab + (long long)cj * nx * nx : nullptr;
        for (int i = tid; i < nx * nx; i += nt) {
            const int r = i / nx, k = i % nx;
            double acc = A[i];
            for (int a = 0; a < nu; ++a) acc = fma(B[r * nu + a], Km[a * nx + k], acc);
            Ab[i] = acc;
        }
        __syncthreads();
        for (int i = tid; i < nx * nx; i += nt) {
            const int r = i / nx, k = i % nx;
            double acc;
            if (Pj) {
                acc = 0.0;
                for (int l = 0; l < nx; ++l) acc = fma(Pj[r * nx + l], Ab[l * nx + k], acc);
            } else {
                acc = Ab[i];
            }
            T[i] = acc;
        }
        __syncthreads();
        for (int i = tid; i < nx * nx; i += nt) {
            const int r = i / nx, k = i % nx;
            double acc = 0.0;
            for (int l = 0; l < nx; ++l) acc = fma(Ab[l * nx + r], T[l * nx + k], acc);
            Pa[i] += acc;
        }
        __syncthreads();
    }
    for (int i = tid; i < nx * nx; i += nt) Ptab[(long long)c * nx * nx + i] = Pa[i];
    // K (for K' r) and the concatenation [K R~^-1] stored reduction-index-major (for u = K x + R~^-1 r)
    double *KR = KRcatT + (long long)c * (nx + nu) * nu;
    for (int i = tid; i < nu * nx; i += nt) {
        const int a = i / nx, k = i % nx;
        Ktab[(long long)c * nu * nx + i] = Km[i];
        KR[(long long)k * nu + a] = Km[i];
    }
    for (int i = tid; i < nu * nu; i += nt) {
        const int a = i / nu, b = i % nu;
        KR[(long long)(nx + b) * nu + a] = Ri[i];
    }
}

// ----------------------------------------------------------------------------------------------------------------
// lambda_max(L* L).  L* L is block diagonal: per nonleaf node the blocks
//   G_x = sum_j sqrtQ_j' sqrtQ_j (+ I with rectangles),  G_u = same with sqrtR,  [[I + b b', -b], [-b', 1]] for (y, s),
// 1/2 for tau_j and leaf s, and sqrtQf' sqrtQf (+ I) for leaf x.  Nodes whose children carry the same cost matrices
// share G_x / G_u, so the host groups them and one warp per group runs a cyclic Jacobi eigenvalue iteration.
// kind 0: G_x of group g (list = cost rows of the children), 1: G_u, 2: leaf block of leaf-cost row g.
// work: [num_groups][dim*dim] scratch.
__global__ void k_gram_eig(const __grid_constant__ Params P, const int *__restrict__ grp_ptr,
                           const int *__restrict__ grp_idx, int kind, int num_groups, double *__restrict__ work,
                           double *__restrict__ out_max, int *__restrict__ status) {
    const int g = blockIdx.x;
    if (g >= num_groups) return;
    const int lane = threadIdx.x;
    const int dim = kind == 1 ? P.L.nu : P.L.nx;
    double *G = work + (long long)g * dim * dim;
    const double *tab = kind == 0 ? P.m.sqT : (kind == 1 ? P.m.srT : P.m.sqfT);
    const bool add_eye = kind == 2 ? P.L.has_leaf_rect : P.L.has_nl_rect;
    const int k0 = kind == 2 ? g : grp_ptr[g], k1 = kind == 2 ? g + 1 : grp_ptr[g + 1];
    for (int i = lane; i < dim * dim; i += 32) {
        const int r = i / dim, c = i % dim;
        double acc = (add_eye && r == c) ? 1.0 : 0.0;
        for (int kk = k0; kk < k1; ++kk) {
            const double *MT = tab + (long long)(kind == 2 ? g : grp_idx[kk]) * dim * dim;  // MT[l][k] = M[k][l]
            for (int l = 0; l < dim; ++l) acc = fma(MT[r * dim + l], MT[c * dim + l], acc);  // (M'M)[r][c]
        }
        G[i] = acc;
    }
    __syncwarp();
    // cyclic Jacobi (eigenvalues only): rotate rows/columns p, q until the off-diagonal mass vanishes
    bool converged = false;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = lane; i < dim * dim; i += 32) {
            if (i / dim != i % dim) off = fma(G[i], G[i], off);
            else diag = fma(G[i], G[i], diag);
        }
        off = warp_sum(off);
        diag = warp_sum(diag);
        if (off <= 1e-34 * diag) {   // off-diagonal mass below (1e-17)^2 of the diagonal: converged
            converged = true;
            break;
        }
        if (!(off == off) || !(diag == diag)) break;   // NaN in the cost matrices: reported below
        for (int p = 0; p < dim - 1; ++p) {
            for (int q = p + 1; q < dim; ++q) {
                const double apq = G[p * dim + q];
                if (apq != 0.0) {
                    const double app = G[p * dim + p], aqq = G[q * dim + q];
                    const double theta = (aqq - app) / (2.0 * apq);
                    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                    const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
                    __syncwarp();
                    for (int k = lane; k < dim; k += 32) {   // columns p, q
                        const double gkp = G[k * dim + p], gkq = G[k * dim + q];
                        G[k * dim + p] = cs * gkp - sn * gkq;
                        G[k * dim + q] = sn * gkp + cs * gkq;
                    }
                    __syncwarp();
                    for (int k = lane; k < dim; k += 32) {   // rows p, q
                        const double gpk = G[p * dim + k], gqk = G[q * dim + k];
                        G[p * dim + k] = cs * gpk - sn * gqk;
                        G[q * dim + k] = sn * gpk + cs * gqk;
                    }
                    __syncwarp();
                }
            }
        }
    }
    if (!converged) {   // the 30-sweep cap was hit (or NaN): one more look at the off-diagonal mass, with a usable tolerance
        double off = 0.0, diag = 0.0;
        for (int i = lane; i < dim * dim; i += 32) {
            if (i / dim != i % dim) off = fma(G[i], G[i], off);
            else diag = fma(G[i], G[i], diag);
        }
        off = warp_sum(off);
        diag = warp_sum(diag);
        // a diagonal maximum taken before convergence UNDER-estimates lambda_max -> alpha too large -> divergence: refuse
        if (!(off <= 1e-26 * diag) && lane == 0) atomicOr(status, 8);
    }
    double best = 0.0;
    for (int k = lane; k < dim; k += 32) best = fmax(best, G[k * dim + k]);
    best = warp_max(best);
    if (lane == 0) atomic_max_nonneg(out_max, best);
}

// (y_i, s_i) block [[I + b b', -b], [-b', 1]], b = [pi; 0; 1]: eigenvalues 1 and the roots of
// l^2 - (2 + |b|^2) l + 1 = 0, so l_max = (2 + |b|^2 + sqrt(|b|^4 + 4 |b|^2)) / 2.  Also covers the 1/2 blocks.
__global__ void k_risk_block_eig(const __grid_constant__ Params P, double *__restrict__ out_max) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int node = blockIdx.x * (blockDim.x >> 5) + warp;
    if (node >= P.L.m) return;
    const int c0 = P.t.child_first[node], cc = P.t.child_count[node];
    double bb = 0.0;
    for (int e = lane; e < cc; e += 32) bb = fma(P.t.cond_prob[c0 + e], P.t.cond_prob[c0 + e], bb);
    bb = warp_sum(bb) + 1.0;
    if (lane == 0) atomic_max_nonneg(out_max, fmax(0.5, (2.0 + bb + sqrt(bb * bb + 4.0 * bb)) / 2.0));
}

}  // namespace rb
