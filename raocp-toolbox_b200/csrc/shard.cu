// shard.cu -- one scenario tree sharded over the GPUs of a box by SUBTREE (BASELINE.json north_star).
//
// The tree is cut at the first sweep level (the first stage with >= 64 nodes): rank r owns a contiguous block of the
// subtrees rooted there -- in breadth-first numbering that is one contiguous node range per stage -- and every rank
// replicates the few nodes above the cut ("top").  Per Chambolle-Pock iteration ONE small all-gather crosses NVLink:
// per cut node j its backward DP message q_j (nx doubles) and its dual d2_j (1 double, needed by the parent's primal
// pass), plus the six residual maxima of the previous iteration.  Everything else is local:
//     primal pass (owned) -> backward sweep (owned subtrees) -> pack | all-gather | unpack + stopping test
//     -> primal pass (top) -> top sweep (replicated) -> forward sweep (owned) -> dual pass (owned + top)
// NCCL is reached through dlopen (libnccl.so.2 is already in the process when torch is imported): the library keeps
// no link-time dependency on it and single-GPU users never load it.
#include <dlfcn.h>

#include "kernels.cuh"

namespace rb {

// message of one rank: [cut nodes of the rank][nx + 1] then 6 residual maxima; stride = cap * (nx + 1) + 6
__global__ void k_shard_pack(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl, ShardPlan sp,
                             const double *__restrict__ q, const double *__restrict__ dual_src,
                             const double *__restrict__ slots, double *__restrict__ send) {
    if (ctrl->done) return;
    const int nx = P.L.nx, w = nx + 1;
    const int count = sp.cut_hi - sp.cut_lo;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count * w; i += gridDim.x * blockDim.x) {
        const int c = i / w, k = i - c * w, node = sp.cut_first + sp.cut_lo + c;
        send[i] = k < nx ? q[(long long)node * nx + k] : dual_src[P.L.d2 + node];
    }
    if (blockIdx.x == 0 && threadIdx.x < 6) send[(long long)sp.cap * w + threadIdx.x] = slots[threadIdx.x];
}

__global__ void k_shard_unpack(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl, ShardPlan sp,
                               const double *__restrict__ recv, double *__restrict__ q, double *__restrict__ dual_src,
                               double *__restrict__ slots) {
    if (ctrl->done) return;
    const int nx = P.L.nx, w = nx + 1;
    const long long stride = (long long)sp.cap * w + 6;
    for (int r = 0; r < sp.world; ++r) {
        if (r == sp.rank) continue;
        const int lo = sp.cut_bounds[r], count = sp.cut_bounds[r + 1] - lo;
        const double *msg = recv + r * stride;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count * w; i += gridDim.x * blockDim.x) {
            const int c = i / w, k = i - c * w, node = sp.cut_first + lo + c;
            if (k < nx) q[(long long)node * nx + k] = msg[i];
            else dual_src[P.L.d2 + node] = msg[i];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 6) {   // global residual maxima (bit patterns of non-negative doubles)
        unsigned long long best = 0ull;
        for (int r = 0; r < sp.world; ++r) {
            const unsigned long long v =
                (unsigned long long)__double_as_longlong(recv[r * stride + (long long)sp.cap * w + threadIdx.x]);
            best = v > best ? v : best;
        }
        slots[threadIdx.x] = __longlong_as_double((long long)best);
    }
}

// ---- NCCL through dlopen ----------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
}  // namespace

const char *nccl_load() {
    if (g_nccl.lib) return nullptr;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return "libnccl.so.2 not found (import torch first, or add NCCL to the library path)";
    g_nccl.GetUniqueId = (int (*)(void *))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void **, int, NcclId, int))dlsym(lib, "ncclCommInitRank");
    g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(lib, "ncclAllGather");
    g_nccl.CommDestroy = (int (*)(void *))dlsym(lib, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) return "NCCL symbols missing";
    g_nccl.lib = lib;
    return nullptr;
}
int nccl_unique_id(NcclId *id) { return g_nccl.GetUniqueId(id); }
int nccl_comm_init(void **comm, int world, const NcclId &id, int rank) { return g_nccl.CommInitRank(comm, world, id, rank); }
int nccl_all_gather_f64(const double *send, double *recv, size_t count, void *comm, cudaStream_t st) {
    return g_nccl.AllGather(send, recv, count, 8 /* ncclFloat64 */, comm, st);
}
void nccl_comm_destroy(void *comm) {
    if (g_nccl.lib && comm) g_nccl.CommDestroy(comm);
}
const char *nccl_error(int code) { return g_nccl.GetErrorString ? g_nccl.GetErrorString(code) : "nccl error"; }

}  // namespace rb
