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
                             const double *__restrict__ q, const double *__restrict__ aux,
                             const double *__restrict__ slots, double *__restrict__ send) {
    // aux[node]: the scalar that travels with q_j -- d2_j (aux = dual + L.d2: the top's primal pass forms sbar_j from it) or, in
    // the pipelined loop, sbar_j itself (aux = pbar + L.ps, written by the owner's dual pass)
    if (ctrl->done) return;
    const int nx = P.L.nx, w = nx + 1;
    const int count = sp.cut_hi - sp.cut_lo;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count * w; i += gridDim.x * blockDim.x) {
        const int c = i / w, k = i - c * w, node = sp.cut_first + sp.cut_lo + c;
        send[i] = k < nx ? q[(long long)node * nx + k] : aux[node];
    }
    if (blockIdx.x == 0 && threadIdx.x < 6) send[(long long)sp.cap * w + threadIdx.x] = slots[threadIdx.x];
}

__global__ void k_shard_unpack(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl, ShardPlan sp,
                               const double *__restrict__ recv, double *__restrict__ q, double *__restrict__ aux,
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
            else aux[node] = msg[i];
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

// ---- device-initiated exchange over NVLink peer memory ("peer-mapped buffers + flags") ---------------------------------
// Every rank's receive buffer and flag array are mapped into every other rank (CUDA IPC).  push: CTA r of rank R writes R's
// message straight into rank r's receive slot [parity][R] (st.global over NVLink; r == R: the local copy), fences at system scope
// and releases flag [parity][R] = seq.  pull: waits until the flags of all ranks carry seq, then unpacks like k_shard_unpack.
// No host call, no NCCL kernel: the exchange is two small launches inside the iteration's CUDA graph.  seq (device counter,
// bumped by k_shard_seq once per exchange) also selects the parity, so a slot is rewritten two exchanges later -- after its
// reader has demonstrably passed the exchange in between.  A wait that exceeds ~2 s sets status bit 16 instead of hanging.
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(256) k_shard_push(const __grid_constant__ Params P, const Ctrl *__restrict__ ctrl, ShardPlan sp,
                                                    const double *__restrict__ q, const double *__restrict__ aux,
                                                    const double *__restrict__ slots, PeerXchg px) {
    if (ctrl->done) return;
    const int nx = P.L.nx, w = nx + 1, peer = blockIdx.x;
    const unsigned long long seq = *px.seq;
    const int parity = (int)(seq & 1ull);
    const long long stride = (long long)sp.cap * w + 6;
    double *dst = px.recv[peer] + ((long long)parity * sp.world + sp.rank) * stride;
    const int count = sp.cut_hi - sp.cut_lo;
    for (int i = threadIdx.x; i < count * w; i += blockDim.x) {
        const int c = i / w, k = i - c * w, node = sp.cut_first + sp.cut_lo + c;
        dst[i] = k < nx ? q[(long long)node * nx + k] : aux[node];
    }
    if (threadIdx.x < 6) dst[(long long)sp.cap * w + threadIdx.x] = slots[threadIdx.x];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(px.flag[peer] + parity * sp.world + sp.rank, seq);
}

__global__ void __launch_bounds__(256) k_shard_pull(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl, ShardPlan sp,
                                                    double *__restrict__ q, double *__restrict__ aux, double *__restrict__ slots,
                                                    PeerXchg px) {
    if (ctrl->done) return;
    const int nx = P.L.nx, w = nx + 1;
    const unsigned long long seq = *px.seq;
    const int parity = (int)(seq & 1ull);
    const long long stride = (long long)sp.cap * w + 6;
    const unsigned long long *flags = px.flag[sp.rank] + parity * sp.world;
    if (threadIdx.x < sp.world) {
        const long long t0 = clock64();
        while (ld_acquire_sys(flags + threadIdx.x) != seq) {
            if (clock64() - t0 > 4000000000LL) {   // ~2 s: a peer is gone -- report instead of hanging the GPU
                atomicOr(&ctrl->status, 16);
                break;
            }
        }
    }
    __syncthreads();
    const double *recv = px.recv[sp.rank] + (long long)parity * sp.world * stride;
    for (int r = 0; r < sp.world; ++r) {
        if (r == sp.rank) continue;
        const int lo = sp.cut_bounds[r], count = sp.cut_bounds[r + 1] - lo;
        const double *msg = recv + r * stride;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count * w; i += gridDim.x * blockDim.x) {
            const int c = i / w, k = i - c * w, node = sp.cut_first + lo + c;
            const double v = __ldcg(msg + i);
            if (k < nx) q[(long long)node * nx + k] = v;
            else aux[node] = v;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 6) {   // global residual maxima (bit patterns of non-negative doubles)
        unsigned long long best = 0ull;
        for (int r = 0; r < sp.world; ++r) {
            const unsigned long long v = (unsigned long long)__double_as_longlong(__ldcg(recv + r * stride + (long long)sp.cap * w + threadIdx.x));
            best = v > best ? v : best;
        }
        slots[threadIdx.x] = __longlong_as_double((long long)best);
    }
}

__global__ void k_shard_seq(const Ctrl *__restrict__ ctrl, unsigned long long *seq) {
    if (!ctrl->done) *seq = *seq + 1ull;
}

// push, pull and the stopping test of the previous iteration in ONE launch (batch 1; the three-launch form above stays for batches
// and as the reference the tests compare with).  CTA b pushes this rank's message to rank b and releases its flag there, then waits
// for rank b's message here and unpacks it; the maxima of the peers are folded into `slots` with atomicMax (own maxima are already
// there; a message that was packed after a neighbour's maxima had been folded in is still <= the global maximum, so every rank ends
// with the same six numbers).  The last CTA to arrive runs k_check's body and bumps the sequence number.  Nothing waits before it
// has pushed, so the CTAs of the W ranks cannot wait on each other in a cycle.
__global__ void __launch_bounds__(256) k_shard_xchg(const __grid_constant__ Params P, Ctrl *__restrict__ ctrl, ShardPlan sp,
                                                    double *__restrict__ q, double *__restrict__ aux, double *__restrict__ slots,
                                                    PeerXchg px, double *__restrict__ last, double *__restrict__ host_last,
                                                    int do_check) {
    __shared__ int is_last;
    if (ctrl->done) return;
    const int nx = P.L.nx, w = nx + 1, peer = blockIdx.x;
    const unsigned long long seq = *px.seq;
    const int parity = (int)(seq & 1ull);
    const long long stride = (long long)sp.cap * w + 6;
    if (peer != sp.rank) {
        double *dst = px.recv[peer] + ((long long)parity * sp.world + sp.rank) * stride;
        const int count = sp.cut_hi - sp.cut_lo;
        for (int i = threadIdx.x; i < count * w; i += blockDim.x) {
            const int c = i / w, k = i - c * w, node = sp.cut_first + sp.cut_lo + c;
            dst[i] = k < nx ? q[(long long)node * nx + k] : aux[node];
        }
        if (threadIdx.x < 6) dst[(long long)sp.cap * w + threadIdx.x] = __ldcg(slots + threadIdx.x);
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            st_release_sys(px.flag[peer] + parity * sp.world + sp.rank, seq);
            const unsigned long long *flag = px.flag[sp.rank] + parity * sp.world + peer;
            const long long t0 = clock64();
            while (ld_acquire_sys(flag) != seq) {
                if (clock64() - t0 > 4000000000LL) {   // ~2 s: a peer is gone -- report instead of hanging the GPU
                    atomicOr(&ctrl->status, 16);
                    break;
                }
            }
        }
        __syncthreads();
        const double *msg = px.recv[sp.rank] + ((long long)parity * sp.world + peer) * stride;
        const int lo = sp.cut_bounds[peer], rcount = sp.cut_bounds[peer + 1] - lo;
        for (int i = threadIdx.x; i < rcount * w; i += blockDim.x) {
            const int c = i / w, k = i - c * w, node = sp.cut_first + lo + c;
            const double v = __ldcg(msg + i);
            if (k < nx) q[(long long)node * nx + k] = v;
            else aux[node] = v;
        }
        if (threadIdx.x < 6)   // bit patterns of non-negative doubles; NaN sorts above +inf and sticks
            atomicMax(reinterpret_cast<unsigned long long *>(slots) + threadIdx.x,
                      (unsigned long long)__double_as_longlong(__ldcg(msg + (long long)sp.cap * w + threadIdx.x)));
    }
    __threadfence();
    __syncthreads();
    unsigned int *arrived = reinterpret_cast<unsigned int *>(px.seq + 1);
    if (threadIdx.x == 0) is_last = atomicAdd(arrived, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last || threadIdx.x >= 32) return;
    __threadfence();
    const int lane = threadIdx.x;
    const Ctrl c = *ctrl;
    if (do_check && c.pending) {   // k_check (fused.cu), batch 1
        const double mine = lane < 6 ? __ldcg(slots + lane) : 0.0;
        const bool nan = mine != mine, bad = lane < 3 && !(mine <= c.tol);
        if (lane < 6) {
            if (c.hist && c.iters < c.hist_capacity) c.hist[(long long)c.iters * 6 + lane] = mine;
            last[lane] = mine;
            if (host_last && c.mirror) host_last[lane] = mine;
            slots[lane] = 0.0;
        }
        const bool all_ok = !__any_sync(0xffffffffu, bad), any_nan = __any_sync(0xffffffffu, nan);
        if (lane == 0) {
            if (any_nan) ctrl->status |= 2;
            ctrl->iters = c.iters + 1;
            ctrl->pending = 0;
            if (c.iters >= c.max_iters || all_ok) ctrl->done = 1;
        }
    }
    if (lane == 0) {
        *arrived = 0u;
        *px.seq = seq + 1ull;
    }
}

void launch_shard_push(cudaStream_t st, const Params &P, const Ctrl *ctrl, const ShardPlan &sp, const double *q, const double *aux,
                       const double *slots, const PeerXchg &px) {
    k_shard_push<<<sp.world, 256, 0, st>>>(P, ctrl, sp, q, aux, slots, px);
}
void launch_shard_pull(cudaStream_t st, const Params &P, Ctrl *ctrl, const ShardPlan &sp, double *q, double *aux, double *slots,
                       const PeerXchg &px) {
    k_shard_pull<<<4, 256, 0, st>>>(P, ctrl, sp, q, aux, slots, px);
    k_shard_seq<<<1, 1, 0, st>>>(ctrl, px.seq);
}

void launch_shard_xchg(cudaStream_t st, const Params &P, Ctrl *ctrl, const ShardPlan &sp, double *q, double *aux, double *slots,
                       const PeerXchg &px, double *last, double *host_last, bool check) {
    k_shard_xchg<<<sp.world, 256, 0, st>>>(P, ctrl, sp, q, aux, slots, px, last, host_last, check ? 1 : 0);
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
