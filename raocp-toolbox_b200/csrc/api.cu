// api.cu -- host side of the C-ABI (include/raocp_b200.h): handle, device memory, launches, CUDA graphs.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/raocp_b200.h"
#include "kernels.cuh"

using namespace rb;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

// control block in device memory: kernels.cuh Ctrl + the parameters of the stopping test
struct HostCtrl {
    Ctrl c;
};

}  // namespace

struct rb_solver {
    Params P{};
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    std::string err;
    std::vector<void *> allocs;
    // host copies of the topology
    std::vector<int> stage_off, parent, child_first, child_count, dyn_idx, cost_idx, cls;
    int num_cls = 0, num_dyn = 0, num_cost = 0, num_leafcost = 0;
    int64_t np = 0, nd = 0;  // compact sizes
    // compact segment tables: offset in compact vector, offset in padded buffer, length
    struct Seg {
        int64_t compact, padded, len;
    };
    std::vector<Seg> pseg, dseg;
    // iterates: index 0 = current, 1 = old (API semantics)
    double *prim[2] = {nullptr, nullptr};
    double *dual[2] = {nullptr, nullptr};
    double *q = nullptr, *r = nullptr, *x0 = nullptr;
    Ctrl *ctrl = nullptr;
    double *slots = nullptr, *last = nullptr, *hist = nullptr;
    int hist_capacity = 0;
    int *status = nullptr;
    bool have_x0 = false, have_offline = false;
    // offline
    std::vector<int> cls_child_ptr, cls_child_dyn, cls_child_cls, level_list, level_ptr;
    int *d_cls_child_ptr = nullptr, *d_cls_child_dyn = nullptr, *d_cls_child_cls = nullptr, *d_level_list = nullptr;
    double *Ptab = nullptr, *Ktab = nullptr, *KTtab = nullptr, *RinvTtab = nullptr;
    // residual temporaries (allocated on first use)
    double *tp[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double *td[4] = {nullptr, nullptr, nullptr, nullptr};
    double *staging_p = nullptr, *staging_d = nullptr;  // device staging for host-vector operator calls
    // fused loop
    cudaGraphExec_t graph[2] = {nullptr, nullptr};
    double graph_alpha = 0.0;
    int64_t launches = 0;
    int kernels_per_iter = 0;
};

namespace {

int fail(rb_solver *s, int code, const std::string &msg) {
    if (s) s->err = msg;
    else g_create_error = msg;
    return code;
}

#define RB_CUDA(s, call)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail((s), e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? RB_ERR_NO_DEVICE : RB_ERR_CUDA, \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                          \
    } while (0)

template <typename T>
int upload(rb_solver *s, const T *host, size_t count, T **dev) {
    *dev = nullptr;
    if (count == 0) count = 1;
    RB_CUDA(s, cudaMalloc((void **)dev, count * sizeof(T)));
    s->allocs.push_back(*dev);
    if (host) RB_CUDA(s, cudaMemcpy(*dev, host, count * sizeof(T), cudaMemcpyHostToDevice));
    else RB_CUDA(s, cudaMemset(*dev, 0, count * sizeof(T)));
    return RB_OK;
}

template <typename T>
int dev_zero(rb_solver *s, size_t count, T **dev) {
    return upload<T>(s, nullptr, count, dev);
}

std::vector<double> transpose_tab(const double *tab, int count, int rows, int cols) {
    std::vector<double> out((size_t)count * rows * cols);
    for (int t = 0; t < count; ++t)
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c)
                out[(size_t)t * rows * cols + (size_t)c * rows + r] = tab[(size_t)t * rows * cols + (size_t)r * cols + c];
    return out;
}

int is_diag(const double *tab, int count, int dim) {
    for (int t = 0; t < count; ++t)
        for (int r = 0; r < dim; ++r)
            for (int c = 0; c < dim; ++c)
                if (r != c && tab[(size_t)t * dim * dim + (size_t)r * dim + c] != 0.0) return 0;
    return 1;
}

inline int64_t round16(int64_t v) { return (v + 15) / 16 * 16; }

dim3 node_grid(const rb_solver *s, int nodes) {
    return dim3((unsigned)((nodes + kWarpsPerBlock - 1) / kWarpsPerBlock), (unsigned)s->P.L.batch);
}

int check_status(rb_solver *s) {
    int st = 0;
    RB_CUDA(s, cudaMemcpyAsync(&st, s->status, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    RB_CUDA(s, cudaStreamSynchronize(s->stream));
    if (st) {
        RB_CUDA(s, cudaMemsetAsync(s->status, 0, sizeof(int), s->stream));
        if (st & 1) return fail(s, RB_ERR_NUMERIC, "Rectangle constraint - 'nan' value cannot be constrained");
        if (st & 4) return fail(s, RB_ERR_NUMERIC, "offline factorisation: R~ is not positive definite");
        return fail(s, RB_ERR_NUMERIC, "non-finite value in the residuals");
    }
    return RB_OK;
}

// compact host vector <-> padded device buffer, all instances (2-D copies: one per segment)
int copy_segments(rb_solver *s, const std::vector<rb_solver::Seg> &segs, int64_t compact_stride, int64_t padded_stride,
                  double *dev, const double *host_in, double *host_out) {
    for (const auto &g : segs) {
        if (g.len == 0) continue;
        if (host_in)
            RB_CUDA(s, cudaMemcpy2DAsync(dev + g.padded, padded_stride * sizeof(double), host_in + g.compact,
                                         compact_stride * sizeof(double), g.len * sizeof(double), s->P.L.batch,
                                         cudaMemcpyHostToDevice, s->stream));
        else
            RB_CUDA(s, cudaMemcpy2DAsync(host_out + g.compact, compact_stride * sizeof(double), dev + g.padded,
                                         padded_stride * sizeof(double), g.len * sizeof(double), s->P.L.batch,
                                         cudaMemcpyDeviceToHost, s->stream));
    }
    RB_CUDA(s, cudaStreamSynchronize(s->stream));
    return RB_OK;
}

int launch_ok(rb_solver *s, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(s, RB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return RB_OK;
}

#define RB_LAUNCHED(s, what)                       \
    do {                                           \
        ++(s)->launches;                           \
        int rc_ = launch_ok((s), (what));          \
        if (rc_ != RB_OK) return rc_;              \
    } while (0)

int need_offline(rb_solver *s) {
    if (!s->have_offline) return fail(s, RB_ERR_STATE, "rb_offline() has not been run");
    return RB_OK;
}

}  // namespace

// ====================================================================================================================
extern "C" {

const char *rb_last_error(const rb_solver *s) { return s ? s->err.c_str() : g_create_error.c_str(); }

int rb_create(const rb_problem *pb, rb_solver **out) {
    if (!pb || !out) return fail(nullptr, RB_ERR_INVALID, "null argument");
    *out = nullptr;
    if (pb->n < 2 || pb->m < 1 || pb->m >= pb->n || pb->nx < 1 || pb->nu < 1 || pb->batch < 1)
        return fail(nullptr, RB_ERR_INVALID, "bad sizes");
    if (pb->nx > kMaxDim || pb->nu > kMaxDim)
        return fail(nullptr, RB_ERR_INVALID, "nx and nu must be <= 64 in this build (kMaxDim)");
    if ((pb->num_nl_rect > 0 && !pb->nl_rect_idx) || (pb->num_leaf_rect > 0 && !pb->leaf_rect_idx))
        return fail(nullptr, RB_ERR_INVALID, "rectangle tables without index arrays");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, RB_ERR_NO_DEVICE,
                    std::string("no CUDA device (raocp_b200 has no CPU fallback): ") + cudaGetErrorString(e));
    if (pb->device < 0 || pb->device >= ndev) return fail(nullptr, RB_ERR_INVALID, "bad device ordinal");

    rb_solver *s = new rb_solver();
    s->device = pb->device;
    auto bail = [&](int rc) {
        g_create_error = s->err;
        rb_destroy(s);
        return rc;
    };
#define TRY(x)                         \
    do {                               \
        int rc_ = (x);                 \
        if (rc_ != RB_OK) return bail(rc_); \
    } while (0)
#define TRYC(call)                                                                                   \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            s->err = std::string(#call) + ": " + cudaGetErrorString(e_);                             \
            return bail(RB_ERR_CUDA);                                                                \
        }                                                                                            \
    } while (0)

    TRYC(cudaSetDevice(s->device));
    TRYC(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    const int n = pb->n, m = pb->m, nx = pb->nx, nu = pb->nu, nl = n - m;
    s->stage_off.assign(pb->stage_off, pb->stage_off + pb->num_stages + 1);
    s->parent.assign(pb->parent, pb->parent + n);
    s->child_first.assign(pb->child_first, pb->child_first + m);
    s->child_count.assign(pb->child_count, pb->child_count + m);
    s->dyn_idx.assign(pb->dyn_idx, pb->dyn_idx + n);
    s->cost_idx.assign(pb->cost_idx, pb->cost_idx + n);
    s->cls.assign(pb->cls, pb->cls + m);
    s->num_cls = pb->num_cls;
    s->num_dyn = pb->num_dyn;
    s->num_cost = pb->num_cost;
    s->num_leafcost = pb->num_leafcost;
    // ---- validate the topology the kernels rely on
    if (s->stage_off[0] != 0 || s->stage_off[pb->num_stages] != n || s->stage_off[pb->num_stages - 1] != m) {
        s->err = "stage_off must partition 0..n with the leaves as the last stage";
        return bail(RB_ERR_INVALID);
    }
    {
        int expect = 1;
        for (int i = 0; i < m; ++i) {
            if (s->child_count[i] < 1 || s->child_first[i] != expect) {
                s->err = "children of the nonleaf nodes must be consecutive ranges in node order (renumber the tree "
                         "breadth first)";
                return bail(RB_ERR_INVALID);
            }
            for (int j = expect; j < expect + s->child_count[i]; ++j)
                if (j >= n || s->parent[j] != i) {
                    s->err = "parent[] and child ranges disagree";
                    return bail(RB_ERR_INVALID);
                }
            expect += s->child_count[i];
        }
        if (expect != n) {
            s->err = "child ranges do not cover all nodes";
            return bail(RB_ERR_INVALID);
        }
        for (int i = 0; i < m; ++i) {
            if (s->cls[i] < 0 || s->cls[i] >= s->num_cls) {
                s->err = "cls[] out of range";
                return bail(RB_ERR_INVALID);
            }
            for (int j = s->child_first[i]; j < s->child_first[i] + s->child_count[i]; ++j)
                if (j < m && s->cls[j] <= s->cls[i]) {
                    s->err = "class of a child must be larger than the class of its parent";
                    return bail(RB_ERR_INVALID);
                }
        }
    }
    // ---- layout
    Layout &L = s->P.L;
    L.n = n; L.m = m; L.nleaf = nl; L.nx = nx; L.nu = nu; L.nxu = nx + nu;
    L.num_stages = pb->num_stages; L.batch = pb->batch;
    L.has_nl_rect = pb->num_nl_rect > 0; L.has_leaf_rect = pb->num_leaf_rect > 0;
    std::vector<int> yoff(m + 1, 0);
    for (int i = 0; i < m; ++i) yoff[i + 1] = yoff[i] + 2 * s->child_count[i] + 1;
    L.ysz = yoff[m];
    {
        int64_t pc = 0, pp = 0;  // compact / padded cursors
        auto addp = [&](long long &field, int64_t len) {
            field = pp;
            s->pseg.push_back({pc, pp, len});
            pc += len;
            pp = round16(pp + len);
        };
        addp(L.px, (int64_t)n * nx); addp(L.pu, (int64_t)m * nu); addp(L.py, L.ysz); addp(L.ptau, n); addp(L.ps, n);
        s->np = pc; L.np_pad = pp;
        pc = 0; pp = 0;
        auto addd = [&](long long &field, int64_t len) {
            field = pp;
            s->dseg.push_back({pc, pp, len});
            pc += len;
            pp = round16(pp + len);
        };
        addd(L.d1, L.ysz); addd(L.d2, m); addd(L.d3, (int64_t)(n - 1) * nx); addd(L.d4, (int64_t)(n - 1) * nu);
        addd(L.d5, n - 1); addd(L.d6, n - 1); addd(L.d7, L.has_nl_rect ? (int64_t)m * (nx + nu) : 0);
        addd(L.d11, (int64_t)nl * nx); addd(L.d12, nl); addd(L.d13, nl); addd(L.d14, L.has_leaf_rect ? (int64_t)nl * nx : 0);
        s->nd = pc; L.nd_pad = pp;
    }
    // ---- topology and tables to the device
    Topo &T = s->P.t;
    int *tmp_i = nullptr;
    double *tmp_d = nullptr;
    TRY(upload(s, pb->parent, n, &tmp_i)); T.parent = tmp_i;
    TRY(upload(s, pb->child_first, m, &tmp_i)); T.child_first = tmp_i;
    TRY(upload(s, pb->child_count, m, &tmp_i)); T.child_count = tmp_i;
    TRY(upload(s, yoff.data(), m + 1, &tmp_i)); T.yoff = tmp_i;
    TRY(upload(s, pb->dyn_idx, n, &tmp_i)); T.dyn_idx = tmp_i;
    TRY(upload(s, pb->cost_idx, n, &tmp_i)); T.cost_idx = tmp_i;
    TRY(upload(s, pb->leafcost_idx, nl, &tmp_i)); T.leafcost_idx = tmp_i;
    TRY(upload(s, L.has_nl_rect ? pb->nl_rect_idx : nullptr, m, &tmp_i)); T.nl_rect_idx = tmp_i;
    TRY(upload(s, L.has_leaf_rect ? pb->leaf_rect_idx : nullptr, nl, &tmp_i)); T.leaf_rect_idx = tmp_i;
    TRY(upload(s, pb->cls, m, &tmp_i)); T.cls = tmp_i;
    TRY(upload(s, pb->cond_prob, n, &tmp_d)); T.cond_prob = tmp_d;
    TRY(upload(s, pb->risk_alpha, m, &tmp_d)); T.risk_alpha = tmp_d;
    Tabs &M = s->P.m;
    TRY(upload(s, pb->A, (size_t)pb->num_dyn * nx * nx, &tmp_d)); M.A = tmp_d;
    TRY(upload(s, pb->B, (size_t)pb->num_dyn * nx * nu, &tmp_d)); M.B = tmp_d;
    {
        auto at = transpose_tab(pb->A, pb->num_dyn, nx, nx);
        TRY(upload(s, at.data(), at.size(), &tmp_d)); M.AT = tmp_d;
        auto bt = transpose_tab(pb->B, pb->num_dyn, nx, nu);
        TRY(upload(s, bt.data(), bt.size(), &tmp_d)); M.BT = tmp_d;
        auto sq = transpose_tab(pb->sqrtQ, pb->num_cost, nx, nx);
        TRY(upload(s, sq.data(), sq.size(), &tmp_d)); M.sqT = tmp_d;
        auto sr = transpose_tab(pb->sqrtR, pb->num_cost, nu, nu);
        TRY(upload(s, sr.data(), sr.size(), &tmp_d)); M.srT = tmp_d;
        auto sf = transpose_tab(pb->sqrtQf, pb->num_leafcost, nx, nx);
        TRY(upload(s, sf.data(), sf.size(), &tmp_d)); M.sqfT = tmp_d;
        M.sq_diag = is_diag(pb->sqrtQ, pb->num_cost, nx);
        M.sr_diag = is_diag(pb->sqrtR, pb->num_cost, nu);
        M.sqf_diag = is_diag(pb->sqrtQf, pb->num_leafcost, nx);
    }
    TRY(upload(s, L.has_nl_rect ? pb->nl_lo : nullptr, (size_t)pb->num_nl_rect * (nx + nu), &tmp_d)); M.nl_lo = tmp_d;
    TRY(upload(s, L.has_nl_rect ? pb->nl_hi : nullptr, (size_t)pb->num_nl_rect * (nx + nu), &tmp_d)); M.nl_hi = tmp_d;
    TRY(upload(s, L.has_leaf_rect ? pb->leaf_lo : nullptr, (size_t)pb->num_leaf_rect * nx, &tmp_d)); M.leaf_lo = tmp_d;
    TRY(upload(s, L.has_leaf_rect ? pb->leaf_hi : nullptr, (size_t)pb->num_leaf_rect * nx, &tmp_d)); M.leaf_hi = tmp_d;
    // ---- class structure for the offline factorisation: representative = first node of the class
    {
        std::vector<int> rep(s->num_cls, -1);
        for (int i = 0; i < m; ++i)
            if (rep[s->cls[i]] < 0) rep[s->cls[i]] = i;
        std::vector<int> level(s->num_cls, 0);
        s->cls_child_ptr.assign(1, 0);
        for (int c = 0; c < s->num_cls; ++c) {
            if (rep[c] < 0) {
                s->err = "class without nodes";
                return bail(RB_ERR_INVALID);
            }
            const int i = rep[c];
            for (int j = s->child_first[i]; j < s->child_first[i] + s->child_count[i]; ++j) {
                s->cls_child_dyn.push_back(s->dyn_idx[j]);
                s->cls_child_cls.push_back(j < m ? s->cls[j] : -1);
            }
            s->cls_child_ptr.push_back((int)s->cls_child_dyn.size());
        }
        int max_level = 0;
        for (int c = s->num_cls - 1; c >= 0; --c) {  // children have larger class ids: already final when visited
            int lv = 0;
            for (int k = s->cls_child_ptr[c]; k < s->cls_child_ptr[c + 1]; ++k)
                if (s->cls_child_cls[k] >= 0) lv = std::max(lv, level[s->cls_child_cls[k]] + 1);
            level[c] = lv;
            max_level = std::max(max_level, lv);
        }
        s->level_ptr.assign(max_level + 2, 0);
        for (int c = 0; c < s->num_cls; ++c) ++s->level_ptr[level[c] + 1];
        for (int l = 0; l <= max_level; ++l) s->level_ptr[l + 1] += s->level_ptr[l];
        s->level_list.resize(s->num_cls);
        std::vector<int> cur(s->level_ptr.begin(), s->level_ptr.end() - 1);
        for (int c = 0; c < s->num_cls; ++c) s->level_list[cur[level[c]]++] = c;
        TRY(upload(s, s->cls_child_ptr.data(), s->cls_child_ptr.size(), &s->d_cls_child_ptr));
        TRY(upload(s, s->cls_child_dyn.data(), s->cls_child_dyn.size(), &s->d_cls_child_dyn));
        TRY(upload(s, s->cls_child_cls.data(), s->cls_child_cls.size(), &s->d_cls_child_cls));
        TRY(upload(s, s->level_list.data(), s->level_list.size(), &s->d_level_list));
    }
    TRY(dev_zero(s, (size_t)s->num_cls * nx * nx, &s->Ptab));
    TRY(dev_zero(s, (size_t)s->num_cls * nu * nx, &s->Ktab));
    TRY(dev_zero(s, (size_t)s->num_cls * nu * nx, &s->KTtab));
    TRY(dev_zero(s, (size_t)s->num_cls * nu * nu, &s->RinvTtab));
    M.K = s->Ktab; M.KT = s->KTtab; M.RinvT = s->RinvTtab;
    // ---- iterates and scratch
    const size_t B = (size_t)L.batch;
    for (int w = 0; w < 2; ++w) {
        TRY(dev_zero(s, B * L.np_pad, &s->prim[w]));
        TRY(dev_zero(s, B * L.nd_pad, &s->dual[w]));
    }
    TRY(dev_zero(s, B * n * nx, &s->q));
    TRY(dev_zero(s, B * m * nu, &s->r));
    TRY(dev_zero(s, B * nx, &s->x0));
    TRY(dev_zero(s, 1, &s->ctrl));
    TRY(dev_zero(s, B * 6, &s->slots));
    TRY(dev_zero(s, B * 6, &s->last));
    TRY(dev_zero(s, 1, &s->status));
    s->kernels_per_iter = 1 + 2 * (L.num_stages - 1) + 1 + 1 + 1;
    *out = s;
    return RB_OK;
#undef TRY
#undef TRYC
}

void rb_destroy(rb_solver *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (int i = 0; i < 2; ++i)
        if (s->graph[i]) cudaGraphExecDestroy(s->graph[i]);
    for (void *p : s->allocs) cudaFree(p);
    if (s->hist) cudaFree(s->hist);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int rb_set_stream(rb_solver *s, void *cuda_stream) {
    if (!s) return RB_ERR_INVALID;
    RB_CUDA(s, cudaStreamSynchronize(s->stream));
    for (int i = 0; i < 2; ++i)
        if (s->graph[i]) {
            cudaGraphExecDestroy(s->graph[i]);
            s->graph[i] = nullptr;
        }
    if (cuda_stream) {
        if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
        s->stream = (cudaStream_t)cuda_stream;
        s->own_stream = false;
    } else if (!s->own_stream) {
        RB_CUDA(s, cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        s->own_stream = true;
    }
    return RB_OK;
}

int rb_sizes(const rb_solver *s, int64_t *np, int64_t *nd) {
    if (!s) return RB_ERR_INVALID;
    if (np) *np = s->np;
    if (nd) *nd = s->nd;
    return RB_OK;
}

int rb_synchronize(rb_solver *s) {
    if (!s) return RB_ERR_INVALID;
    RB_CUDA(s, cudaStreamSynchronize(s->stream));
    return RB_OK;
}

int rb_launch_count(const rb_solver *s, int64_t *k) {
    if (!s || !k) return RB_ERR_INVALID;
    *k = s->launches;
    return RB_OK;
}

// ---- offline -------------------------------------------------------------------------------------------------------
int rb_offline(rb_solver *s) {
    if (!s) return RB_ERR_INVALID;
    RB_CUDA(s, cudaSetDevice(s->device));
    const size_t smem = offline_smem_bytes(s->P.L.nx, s->P.L.nu);
    RB_CUDA(s, cudaFuncSetAttribute(k_offline_level, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ClassView cv{s->d_cls_child_ptr, s->d_cls_child_dyn, s->d_cls_child_cls, s->d_level_list};
    const int levels = (int)s->level_ptr.size() - 1;
    for (int l = 0; l < levels; ++l) {
        const int begin = s->level_ptr[l], count = s->level_ptr[l + 1] - begin;
        if (count == 0) continue;
        k_offline_level<<<count, 256, smem, s->stream>>>(s->P, cv, begin, count, s->Ptab, s->Ktab, s->KTtab, s->RinvTtab,
                                                        s->status);
        RB_LAUNCHED(s, "k_offline_level");
    }
    int rc = check_status(s);
    if (rc != RB_OK) return rc;
    s->have_offline = true;
    return RB_OK;
}

int rb_get_offline(rb_solver *s, double *Pm, double *K, double *Rinv) {
    if (!s) return RB_ERR_INVALID;
    int rc = need_offline(s);
    if (rc != RB_OK) return rc;
    const int nx = s->P.L.nx, nu = s->P.L.nu;
    RB_CUDA(s, cudaStreamSynchronize(s->stream));
    if (Pm) RB_CUDA(s, cudaMemcpy(Pm, s->Ptab, (size_t)s->num_cls * nx * nx * sizeof(double), cudaMemcpyDeviceToHost));
    if (K) RB_CUDA(s, cudaMemcpy(K, s->Ktab, (size_t)s->num_cls * nu * nx * sizeof(double), cudaMemcpyDeviceToHost));
    if (Rinv) {
        std::vector<double> t((size_t)s->num_cls * nu * nu);
        RB_CUDA(s, cudaMemcpy(t.data(), s->RinvTtab, t.size() * sizeof(double), cudaMemcpyDeviceToHost));
        auto tt = transpose_tab(t.data(), s->num_cls, nu, nu);
        std::memcpy(Rinv, tt.data(), tt.size() * sizeof(double));
    }
    return RB_OK;
}

// ---- iterate access -------------------------------------------------------------------------------------------------
int rb_set_primal(rb_solver *s, int which, const double *compact) {
    if (!s || !compact || which < 0 || which > 1) return RB_ERR_INVALID;
    return copy_segments(s, s->pseg, s->np, s->P.L.np_pad, s->prim[which], compact, nullptr);
}
int rb_get_primal(rb_solver *s, int which, double *compact) {
    if (!s || !compact || which < 0 || which > 1) return RB_ERR_INVALID;
    return copy_segments(s, s->pseg, s->np, s->P.L.np_pad, s->prim[which], nullptr, compact);
}
int rb_set_dual(rb_solver *s, int which, const double *compact) {
    if (!s || !compact || which < 0 || which > 1) return RB_ERR_INVALID;
    return copy_segments(s, s->dseg, s->nd, s->P.L.nd_pad, s->dual[which], compact, nullptr);
}
int rb_get_dual(rb_solver *s, int which, double *compact) {
    if (!s || !compact || which < 0 || which > 1) return RB_ERR_INVALID;
    return copy_segments(s, s->dseg, s->nd, s->P.L.nd_pad, s->dual[which], nullptr, compact);
}

int rb_set_initial_state(rb_solver *s, const double *x0) {
    if (!s || !x0) return RB_ERR_INVALID;
    const Layout &L = s->P.L;
    RB_CUDA(s, cudaMemcpyAsync(s->x0, x0, (size_t)L.batch * L.nx * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    // old_primal[0] = state (cache.py:81)
    RB_CUDA(s, cudaMemcpy2DAsync(s->prim[1] + L.px, L.np_pad * sizeof(double), x0, L.nx * sizeof(double),
                                 L.nx * sizeof(double), L.batch, cudaMemcpyHostToDevice, s->stream));
    RB_CUDA(s, cudaStreamSynchronize(s->stream));
    s->have_x0 = true;
    return RB_OK;
}

int rb_update_cache(rb_solver *s) {
    if (!s) return RB_ERR_INVALID;
    const Layout &L = s->P.L;
    RB_CUDA(s, cudaMemcpyAsync(s->prim[1], s->prim[0], (size_t)L.batch * L.np_pad * sizeof(double), cudaMemcpyDeviceToDevice,
                               s->stream));
    RB_CUDA(s, cudaMemcpyAsync(s->dual[1], s->dual[0], (size_t)L.batch * L.nd_pad * sizeof(double), cudaMemcpyDeviceToDevice,
                               s->stream));
    return RB_OK;
}

// ---- linear operator on host vectors --------------------------------------------------------------------------------
static int ensure_staging(rb_solver *s) {
    const Layout &L = s->P.L;
    if (!s->staging_p) {
        int rc = dev_zero(s, (size_t)L.batch * L.np_pad, &s->staging_p);
        if (rc != RB_OK) return rc;
        rc = dev_zero(s, (size_t)L.batch * L.nd_pad, &s->staging_d);
        if (rc != RB_OK) return rc;
    }
    return RB_OK;
}

int rb_apply_L(rb_solver *s, const double *primal_in, double *dual_out) {
    if (!s || !primal_in || !dual_out) return RB_ERR_INVALID;
    int rc = ensure_staging(s);
    if (rc != RB_OK) return rc;
    rc = copy_segments(s, s->pseg, s->np, s->P.L.np_pad, s->staging_p, primal_in, nullptr);
    if (rc != RB_OK) return rc;
    k_l_axpby<<<node_grid(s, s->P.L.n), kThreads, 0, s->stream>>>(s->P, s->staging_p, nullptr, 1.0, 0.0, nullptr,
                                                                 s->staging_d, 0.0, 1.0);
    RB_LAUNCHED(s, "k_l_axpby");
    return copy_segments(s, s->dseg, s->nd, s->P.L.nd_pad, s->staging_d, nullptr, dual_out);
}

int rb_apply_Lt(rb_solver *s, const double *dual_in, double *primal_out) {
    if (!s || !dual_in || !primal_out) return RB_ERR_INVALID;
    int rc = ensure_staging(s);
    if (rc != RB_OK) return rc;
    rc = copy_segments(s, s->dseg, s->nd, s->P.L.nd_pad, s->staging_d, dual_in, nullptr);
    if (rc != RB_OK) return rc;
    k_lt_axpby<<<node_grid(s, s->P.L.n), kThreads, 0, s->stream>>>(s->P, s->staging_d, nullptr, s->staging_p, 0.0, 1.0);
    RB_LAUNCHED(s, "k_lt_axpby");
    return copy_segments(s, s->pseg, s->np, s->P.L.np_pad, s->staging_p, nullptr, primal_out);
}

int rb_lambda_max(rb_solver *s, double *lambda_max) {
    if (!s || !lambda_max) return RB_ERR_INVALID;
    const Layout &L = s->P.L;
    // group the nonleaf nodes by the cost rows of their children
    std::map<std::vector<int>, int> groups;
    std::vector<int> grp_ptr(1, 0), grp_idx;
    for (int i = 0; i < L.m; ++i) {
        std::vector<int> key(s->cost_idx.begin() + s->child_first[i],
                             s->cost_idx.begin() + s->child_first[i] + s->child_count[i]);
        if (groups.emplace(key, (int)groups.size()).second) {
            grp_idx.insert(grp_idx.end(), key.begin(), key.end());
            grp_ptr.push_back((int)grp_idx.size());
        }
    }
    const int ng = (int)groups.size();
    int *d_ptr = nullptr, *d_idx = nullptr;
    double *d_work = nullptr, *d_out = nullptr;
    const int dim = std::max(L.nx, L.nu);
    RB_CUDA(s, cudaMalloc((void **)&d_ptr, grp_ptr.size() * sizeof(int)));
    RB_CUDA(s, cudaMalloc((void **)&d_idx, std::max<size_t>(grp_idx.size(), 1) * sizeof(int)));
    RB_CUDA(s, cudaMalloc((void **)&d_work, (size_t)std::max(ng, s->num_leafcost) * dim * dim * sizeof(double)));
    RB_CUDA(s, cudaMalloc((void **)&d_out, sizeof(double)));
    RB_CUDA(s, cudaMemcpyAsync(d_ptr, grp_ptr.data(), grp_ptr.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    RB_CUDA(s, cudaMemcpyAsync(d_idx, grp_idx.data(), grp_idx.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    RB_CUDA(s, cudaMemsetAsync(d_out, 0, sizeof(double), s->stream));
    for (int kind = 0; kind < 3; ++kind) {
        const int count = kind == 2 ? s->num_leafcost : ng;
        k_gram_eig<<<count, 32, 0, s->stream>>>(s->P, d_ptr, d_idx, kind, count, d_work, d_out);
        RB_LAUNCHED(s, "k_gram_eig");
    }
    k_risk_block_eig<<<(L.m + 3) / 4, 128, 0, s->stream>>>(s->P, d_out);
    RB_LAUNCHED(s, "k_risk_block_eig");
    RB_CUDA(s, cudaMemcpyAsync(lambda_max, d_out, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    RB_CUDA(s, cudaStreamSynchronize(s->stream));
    cudaFree(d_ptr); cudaFree(d_idx); cudaFree(d_work); cudaFree(d_out);
    return RB_OK;
}

// ---- half steps -------------------------------------------------------------------------------------------------------
int rb_primal_half(rb_solver *s, double alpha) {
    if (!s) return RB_ERR_INVALID;
    k_lt_axpby<<<node_grid(s, s->P.L.n), kThreads, 0, s->stream>>>(s->P, s->dual[1], s->prim[1], s->prim[0], 1.0, -alpha);
    RB_LAUNCHED(s, "k_lt_axpby");
    return RB_OK;
}

int rb_s0_shift(rb_solver *s, double alpha) {
    if (!s) return RB_ERR_INVALID;
    k_s0_shift<<<(s->P.L.batch + 127) / 128, 128, 0, s->stream>>>(s->P, s->prim[0], alpha);
    RB_LAUNCHED(s, "k_s0_shift");
    return RB_OK;
}

int rb_project_dynamics(rb_solver *s) {
    if (!s) return RB_ERR_INVALID;
    int rc = need_offline(s);
    if (rc != RB_OK) return rc;
    if (!s->have_x0) return fail(s, RB_ERR_STATE, "initial state not set (cache_initial_state)");
    const Layout &L = s->P.L;
    for (int t = L.num_stages - 1; t >= 0; --t) {
        const int lo = s->stage_off[t], hi = s->stage_off[t + 1];
        k_dyn_bwd<<<node_grid(s, hi - lo), kThreads, 0, s->stream>>>(s->P, s->prim[0], s->q, s->r, lo, hi);
        RB_LAUNCHED(s, "k_dyn_bwd");
    }
    k_set_root<<<std::min(L.batch, 1024), 64, 0, s->stream>>>(s->P, s->prim[0], s->x0);
    RB_LAUNCHED(s, "k_set_root");
    for (int t = 0; t < L.num_stages - 1; ++t) {
        const int lo = s->stage_off[t], hi = s->stage_off[t + 1];
        k_dyn_fwd<<<node_grid(s, hi - lo), kThreads, 0, s->stream>>>(s->P, s->prim[0], s->r, lo, hi);
        RB_LAUNCHED(s, "k_dyn_fwd");
    }
    return RB_OK;
}

int rb_project_kernel(rb_solver *s) {
    if (!s) return RB_ERR_INVALID;
    k_kernel_proj<<<node_grid(s, s->P.L.m), kThreads, 0, s->stream>>>(s->P, s->prim[0]);
    RB_LAUNCHED(s, "k_kernel_proj");
    return RB_OK;
}

int rb_prox_f(rb_solver *s, double alpha) {
    int rc = rb_s0_shift(s, alpha);
    if (rc != RB_OK) return rc;
    rc = rb_project_dynamics(s);
    if (rc != RB_OK) return rc;
    return rb_project_kernel(s);
}

int rb_dual_half(rb_solver *s, double alpha) {
    if (!s) return RB_ERR_INVALID;
    k_l_axpby<<<node_grid(s, s->P.L.n), kThreads, 0, s->stream>>>(s->P, s->prim[0], s->prim[1], 2.0, -1.0, s->dual[1],
                                                                 s->dual[0], 1.0, alpha);
    RB_LAUNCHED(s, "k_l_axpby");
    return RB_OK;
}

static int prox_g_mode(rb_solver *s, double alpha, int mode) {
    if (!s) return RB_ERR_INVALID;
    k_prox_g<<<node_grid(s, s->P.L.n), kThreads, 0, s->stream>>>(s->P, s->dual[0], alpha, mode, s->status);
    RB_LAUNCHED(s, "k_prox_g");
    return (mode & 12) ? check_status(s) : RB_OK;
}
int rb_prox_g_conj(rb_solver *s, double alpha) { return prox_g_mode(s, alpha, 31); }
int rb_modify_dual(rb_solver *s, double alpha) { return prox_g_mode(s, alpha, 1); }
int rb_add_halves(rb_solver *s) { return prox_g_mode(s, 1.0, 2); }
int rb_project_nonleaf(rb_solver *s) { return prox_g_mode(s, 1.0, 4); }
int rb_project_leaf(rb_solver *s) { return prox_g_mode(s, 1.0, 8); }

int rb_modify_projection(rb_solver *s, double alpha, const double *modified_dual) {
    if (!s || !modified_dual) return RB_ERR_INVALID;
    int rc = ensure_staging(s);
    if (rc != RB_OK) return rc;
    rc = copy_segments(s, s->dseg, s->nd, s->P.L.nd_pad, s->staging_d, modified_dual, nullptr);
    if (rc != RB_OK) return rc;
    const long long count = (long long)s->P.L.batch * s->P.L.nd_pad;
    // d = alpha * (w - d)   (cache.py:392-393)
    k_axpby<<<1184, 256, 0, s->stream>>>(s->dual[0], alpha, s->staging_d, -alpha, s->dual[0], count);
    RB_LAUNCHED(s, "k_axpby");
    return RB_OK;
}

// ---- residuals, reference formulas with stand-alone kernels ------------------------------------------------------------
int rb_residuals(rb_solver *s, double alpha, double *norms, double *vectors) {
    if (!s || !norms) return RB_ERR_INVALID;
    const Layout &L = s->P.L;
    const long long pc = (long long)L.batch * L.np_pad, dc = (long long)L.batch * L.nd_pad;
    for (int i = 0; i < 6; ++i)
        if (!s->tp[i]) {
            int rc = dev_zero(s, (size_t)pc, &s->tp[i]);
            if (rc != RB_OK) return rc;
        }
    for (int i = 0; i < 4; ++i)
        if (!s->td[i]) {
            int rc = dev_zero(s, (size_t)dc, &s->td[i]);
            if (rc != RB_OK) return rc;
        }
    double *dp = s->tp[0], *lt = s->tp[1], *xi1 = s->tp[2], *pn = s->tp[3], *xi0 = s->tp[4], *delta0 = s->tp[5];
    double *dd = s->td[0], *lp = s->td[1], *xi2 = s->td[2], *delta2 = s->td[3];
    const dim3 g = node_grid(s, L.n);
    cudaStream_t st = s->stream;
    k_axpby<<<1184, 256, 0, st>>>(dp, 1.0, s->prim[1], -1.0, s->prim[0], pc);            // p - p_new
    k_axpby<<<1184, 256, 0, st>>>(dd, 1.0, s->dual[1], -1.0, s->dual[0], dc);            // d - d_new
    k_lt_axpby<<<g, kThreads, 0, st>>>(s->P, dd, nullptr, lt, 0.0, 1.0);                 // L*(d - d_new)
    k_div_add<<<1184, 256, 0, st>>>(xi1, dp, alpha, -1.0, lt, pc);                       // xi1
    k_axpby<<<1184, 256, 0, st>>>(pn, 1.0, s->prim[0], -1.0, s->prim[1], pc);            // p_new - p = delta1
    k_l_axpby<<<g, kThreads, 0, st>>>(s->P, pn, nullptr, 1.0, 0.0, nullptr, lp, 0.0, 1.0);  // L(p_new - p)
    k_div_add<<<1184, 256, 0, st>>>(xi2, dd, alpha, 1.0, lp, dc);                        // xi2
    k_lt_axpby<<<g, kThreads, 0, st>>>(s->P, xi2, nullptr, lt, 0.0, 1.0);                // L* xi2
    k_axpby<<<1184, 256, 0, st>>>(xi0, 1.0, xi1, 1.0, lt, pc);                           // xi0
    k_axpby<<<1184, 256, 0, st>>>(delta2, 1.0, s->dual[0], -1.0, s->dual[1], dc);        // delta2
    k_lt_axpby<<<g, kThreads, 0, st>>>(s->P, delta2, nullptr, lt, 0.0, 1.0);             // L* delta2
    k_axpby<<<1184, 256, 0, st>>>(delta0, 1.0, pn, -1.0, lt, pc);                        // delta0
    s->launches += 12;
    int rc = launch_ok(s, "residual kernels");
    if (rc != RB_OK) return rc;
    RB_CUDA(s, cudaMemsetAsync(s->slots, 0, (size_t)L.batch * 6 * sizeof(double), st));
    const double *vecs[6] = {xi0, xi1, xi2, delta0, pn, delta2};
    for (int i = 0; i < 6; ++i) {
        const bool is_p = (i != 2 && i != 5);
        k_absmax<<<dim3(148, L.batch), 256, 0, st>>>(vecs[i], is_p ? L.np_pad : L.nd_pad, s->slots + i, 6, s->status);
        RB_LAUNCHED(s, "k_absmax");
    }
    RB_CUDA(s, cudaMemcpyAsync(norms, s->slots, (size_t)L.batch * 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB_CUDA(s, cudaMemsetAsync(s->slots, 0, (size_t)L.batch * 6 * sizeof(double), st));
    RB_CUDA(s, cudaStreamSynchronize(st));
    if (vectors) {
        // per instance: [xi0 | xi1 | xi2 | delta0 | delta1 | delta2], compact
        const int64_t per = 4 * s->np + 2 * s->nd;
        int64_t off = 0;
        for (int i = 0; i < 6; ++i) {
            const bool is_p = (i != 2 && i != 5);
            std::vector<rb_solver::Seg> segs = is_p ? s->pseg : s->dseg;
            for (auto &sg : segs) sg.compact += off;
            rc = copy_segments(s, segs, per, is_p ? L.np_pad : L.nd_pad, const_cast<double *>(vecs[i]), nullptr, vectors);
            if (rc != RB_OK) return rc;
            off += is_p ? s->np : s->nd;
        }
    }
    return RB_OK;
}

// ---- fused loop -------------------------------------------------------------------------------------------------------
namespace {

// enqueue the kernels of one iteration reading (prim[src], dual[src]) and writing (prim[dst], dual[dst])
int enqueue_iteration(rb_solver *s, double alpha, int src, int dst, int max_iters, double tol) {
    const Layout &L = s->P.L;
    cudaStream_t st = s->stream;
    k_fused_primal<<<node_grid(s, L.n), kThreads, 0, st>>>(s->P, s->ctrl, s->prim[src], s->dual[src], s->prim[dst], alpha);
    for (int t = L.num_stages - 1; t >= 0; --t) {
        const int lo = s->stage_off[t], hi = s->stage_off[t + 1];
        k_fused_bwd<<<node_grid(s, hi - lo), kThreads, 0, st>>>(s->P, s->ctrl, s->prim[dst], s->q, s->r, lo, hi);
    }
    for (int t = 0; t < L.num_stages - 1; ++t) {
        const int lo = s->stage_off[t], hi = s->stage_off[t + 1];
        k_fused_fwd<<<node_grid(s, hi - lo), kThreads, 0, st>>>(s->P, s->ctrl, s->prim[dst], s->r, s->x0, lo, hi);
    }
    k_fused_dual<<<node_grid(s, L.n), kThreads, 0, st>>>(s->P, s->ctrl, s->prim[src], s->prim[dst], s->dual[src],
                                                        s->dual[dst], alpha, s->slots);
    k_check<<<1, 32, 0, st>>>(s->P, s->ctrl, s->slots, s->last, s->hist, s->hist_capacity, max_iters, tol);
    return launch_ok(s, "fused iteration");
}

int run_fused(rb_solver *s, double alpha, int max_iters, double tol, int fixed_iters, double *xi_hist, double *delta_hist,
              int hist_capacity, int *iters_out, int *status_out) {
    int rc = need_offline(s);
    if (rc != RB_OK) return rc;
    if (!s->have_x0) return fail(s, RB_ERR_STATE, "initial state not set (cache_initial_state)");
    const Layout &L = s->P.L;
    cudaStream_t st = s->stream;
    const bool want_hist = xi_hist || delta_hist;
    const int total = fixed_iters > 0 ? fixed_iters : max_iters + 1;
    if (s->hist) {
        RB_CUDA(s, cudaStreamSynchronize(st));
        cudaFree(s->hist);
        s->hist = nullptr;
    }
    s->hist_capacity = 0;
    if (want_hist) {
        s->hist_capacity = std::min(total, hist_capacity);
        RB_CUDA(s, cudaMalloc((void **)&s->hist, (size_t)std::max(1, s->hist_capacity) * L.batch * 6 * sizeof(double)));
    }
    RB_CUDA(s, cudaMemsetAsync(s->ctrl, 0, sizeof(Ctrl), st));
    RB_CUDA(s, cudaMemsetAsync(s->slots, 0, (size_t)L.batch * 6 * sizeof(double), st));
    const int eff_max = fixed_iters > 0 ? fixed_iters - 1 : max_iters;
    const double eff_tol = fixed_iters > 0 ? -1.0 : tol;
    // iteration k reads buffer (k even ? old : current) and writes the other one
    int launched = 0;
    Ctrl hc{};
    const int chunk = 64;
    while (launched < total) {
        const int todo = std::min(chunk, total - launched);
        for (int k = 0; k < todo; ++k) {
            const int it = launched + k;
            const int src = (it % 2 == 0) ? 1 : 0, dst = 1 - src;
            rc = enqueue_iteration(s, alpha, src, dst, eff_max, eff_tol);
            if (rc != RB_OK) return rc;
        }
        launched += todo;
        s->launches += (int64_t)todo * s->kernels_per_iter;
        if (fixed_iters > 0) continue;  // benchmark mode: never synchronise inside
        RB_CUDA(s, cudaMemcpyAsync(&hc, s->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
        RB_CUDA(s, cudaStreamSynchronize(st));
        if (hc.done) break;
    }
    RB_CUDA(s, cudaMemcpyAsync(&hc, s->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
    RB_CUDA(s, cudaStreamSynchronize(st));
    const int iters = hc.iters;
    // the last executed iteration (index iters-1) wrote buffer dst; make current == old == final iterate
    const int final_buf = ((iters - 1) % 2 == 0) ? 0 : 1;
    const int other = 1 - final_buf;
    RB_CUDA(s, cudaMemcpyAsync(s->prim[other], s->prim[final_buf], (size_t)L.batch * L.np_pad * sizeof(double),
                               cudaMemcpyDeviceToDevice, st));
    RB_CUDA(s, cudaMemcpyAsync(s->dual[other], s->dual[final_buf], (size_t)L.batch * L.nd_pad * sizeof(double),
                               cudaMemcpyDeviceToDevice, st));
    if (want_hist) {
        const int rows = std::min(iters, s->hist_capacity);
        std::vector<double> h((size_t)rows * L.batch * 6);
        RB_CUDA(s, cudaMemcpyAsync(h.data(), s->hist, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
        RB_CUDA(s, cudaStreamSynchronize(st));
        for (size_t i = 0; i < (size_t)rows * L.batch; ++i)
            for (int k = 0; k < 3; ++k) {
                if (xi_hist) xi_hist[i * 3 + k] = h[i * 6 + k];
                if (delta_hist) delta_hist[i * 3 + k] = h[i * 6 + 3 + k];
            }
    }
    RB_CUDA(s, cudaStreamSynchronize(st));
    if (iters_out) *iters_out = iters;
    // reference status (solver.py:166-169): 0 if the final iteration index is < max_iters
    if (status_out) *status_out = (iters - 1 < max_iters) ? 0 : 1;
    if (hc.status) {
        RB_CUDA(s, cudaMemsetAsync(s->ctrl, 0, sizeof(Ctrl), st));
        if (hc.status & 1) return fail(s, RB_ERR_NUMERIC, "Rectangle constraint - 'nan' value cannot be constrained");
        return fail(s, RB_ERR_NUMERIC, "non-finite value in the residuals");
    }
    return RB_OK;
}

}  // namespace

int rb_iterate(rb_solver *s, double alpha, int32_t max_iters, double tol, int32_t check_every, double *xi_hist,
               double *delta_hist, int32_t hist_capacity, int32_t *iters, int32_t *status) {
    if (!s || max_iters < 0) return RB_ERR_INVALID;
    (void)check_every;  // the stopping test runs on the device after every iteration
    return run_fused(s, alpha, max_iters, tol, 0, xi_hist, delta_hist, hist_capacity, iters, status);
}

int rb_iterate_fixed(rb_solver *s, double alpha, int32_t iters, double *norms) {
    if (!s || iters < 1) return RB_ERR_INVALID;
    int rc = run_fused(s, alpha, iters - 1, -1.0, iters, nullptr, nullptr, 0, nullptr, nullptr);
    if (rc != RB_OK) return rc;
    if (norms) {
        RB_CUDA(s, cudaMemcpy(norms, s->last, (size_t)s->P.L.batch * 6 * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return RB_OK;
}

// ---- stand-alone projections ---------------------------------------------------------------------------------------------
int rb_cone_project(int32_t cone, int32_t dim, const double *in, double *out) {
    if (cone < 0 || cone > 3 || dim < 1 || !in || !out) return fail(nullptr, RB_ERR_INVALID, "bad cone arguments");
    if (dim * sizeof(double) > 48 * 1024) return fail(nullptr, RB_ERR_INVALID, "cone dimension too large");
    double *d_in = nullptr, *d_out = nullptr;
    RB_CUDA(nullptr, cudaMalloc((void **)&d_in, dim * sizeof(double)));
    RB_CUDA(nullptr, cudaMalloc((void **)&d_out, dim * sizeof(double)));
    RB_CUDA(nullptr, cudaMemcpy(d_in, in, dim * sizeof(double), cudaMemcpyHostToDevice));
    k_cone<<<1, 32, dim * sizeof(double)>>>(cone, dim, d_in, d_out);
    RB_CUDA(nullptr, cudaGetLastError());
    RB_CUDA(nullptr, cudaMemcpy(out, d_out, dim * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d_in);
    cudaFree(d_out);
    return RB_OK;
}

int rb_box_project(int32_t dim, const double *in, const double *lo, const double *hi, double *out) {
    if (dim < 1 || !in || !lo || !hi || !out) return fail(nullptr, RB_ERR_INVALID, "bad box arguments");
    double *d = nullptr;
    int *d_status = nullptr;
    RB_CUDA(nullptr, cudaMalloc((void **)&d, 4 * (size_t)dim * sizeof(double)));
    RB_CUDA(nullptr, cudaMalloc((void **)&d_status, sizeof(int)));
    RB_CUDA(nullptr, cudaMemset(d_status, 0, sizeof(int)));
    RB_CUDA(nullptr, cudaMemcpy(d, in, dim * sizeof(double), cudaMemcpyHostToDevice));
    RB_CUDA(nullptr, cudaMemcpy(d + dim, lo, dim * sizeof(double), cudaMemcpyHostToDevice));
    RB_CUDA(nullptr, cudaMemcpy(d + 2 * dim, hi, dim * sizeof(double), cudaMemcpyHostToDevice));
    k_box<<<(dim + 255) / 256, 256>>>(dim, d, d + dim, d + 2 * dim, d + 3 * dim, d_status);
    RB_CUDA(nullptr, cudaGetLastError());
    int st = 0;
    RB_CUDA(nullptr, cudaMemcpy(out, d + 3 * dim, dim * sizeof(double), cudaMemcpyDeviceToHost));
    RB_CUDA(nullptr, cudaMemcpy(&st, d_status, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d);
    cudaFree(d_status);
    if (st) return fail(nullptr, RB_ERR_NUMERIC, "Rectangle constraint - 'nan' value cannot be constrained");
    return RB_OK;
}

}  // extern "C"
