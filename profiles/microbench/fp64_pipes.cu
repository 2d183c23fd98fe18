// Micro-benchmark: FP64 FMA pipe vs FP64 tensor-core mma.sync (m8n8k4) on sm_100a -- throughput and dependent latency.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k_dmma(double *out, int iters, long long *cycles) {
    double c0[ILP], c1[ILP];
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; ++i) c0[i] = c1[i] = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma(c0[i], c1[i], a, b);
    }
    long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int ILP>
__global__ void k_dfma(double *out, int iters, long long *cycles) {
    double c[ILP];
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
    }
    long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <typename F>
void run(const char *name, F launch, double fma_per_thread_iter_total, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-34s %8.3f ms  %8.2f TFLOP/s\n", name, ms, 2.0 * fma_per_thread_iter_total / (ms * 1e-3) / 1e12);
}

int main() {
    double *out;
    long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * 148 * 16 * 1024);
    cudaMalloc(&cyc, 8);
    const int iters = 20000;
    // throughput: 148*4 CTAs x 256 threads
    const int grid = 148 * 4, threads = 256;
    const double warps = (double)grid * threads / 32;
    run("DFMA ILP8 (grid 592x256)", [&] { k_dfma<8><<<grid, threads>>>(out, iters, cyc); }, warps * 32.0 * 8 * iters, iters);
    run("DMMA m8n8k4 ILP4 (grid 592x256)", [&] { k_dmma<4><<<grid, threads>>>(out, iters, cyc); }, warps * 256.0 * 4 * iters, iters);
    run("DMMA m8n8k4 ILP8 (grid 592x256)", [&] { k_dmma<8><<<grid, threads>>>(out, iters, cyc); }, warps * 256.0 * 8 * iters, iters);
    // one warp per SM sub-partition: issue rate of a single warp
    run("DMMA ILP4, 1 warp/SMSP (148x128)", [&] { k_dmma<4><<<148, 128>>>(out, iters, cyc); }, 148.0 * 4 * 256.0 * 4 * iters, iters);
    run("DFMA ILP8, 1 warp/SMSP (148x128)", [&] { k_dfma<8><<<148, 128>>>(out, iters, cyc); }, 148.0 * 4 * 32.0 * 8 * iters, iters);
    // latency: one warp, dependent chain
    k_dmma<1><<<1, 32>>>(out, iters, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DMMA dependent latency  %.1f cycles\n", (double)h / iters);
    k_dfma<1><<<1, 32>>>(out, iters, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent latency  %.1f cycles\n", (double)h / iters);
    k_dmma<4><<<1, 32>>>(out, iters, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DMMA 4 independent, one warp: %.1f cycles per mma\n", (double)h / iters / 4);
    k_dfma<8><<<1, 32>>>(out, iters, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA 8 independent, one warp: %.1f cycles per fma\n", (double)h / iters / 8);
    return 0;
}
