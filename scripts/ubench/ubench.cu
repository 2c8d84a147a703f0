// Micro-benchmarks that size the design (run on the B200 box): L2 / HBM read bandwidth, grid-barrier latency,
// kernel-to-kernel gap inside a CUDA graph, fp32 red.global throughput (scalar / v4), plain-store partials.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void k_read(const float4* __restrict__ p, size_t n4, int reps, float* out) {
    float acc = 0.f;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            float4 v = __ldcg(p + i); acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 123.456f) out[0] = acc;
}
__global__ void k_gridbar(int iters, unsigned* ctr, long long* cycles) {
    cg::grid_group g = cg::this_grid();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) g.sync();
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = (t1 - t0) / iters;
}
// hand-rolled barrier: one atomic per CTA on a monotonically increasing counter, spin with ld.acquire
__global__ void k_mybar(int iters, unsigned* ctr, long long* cycles) {
    long long t0 = clock64();
    unsigned target = 0;
    for (int i = 0; i < iters; ++i) {
        target += gridDim.x;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(ctr, 1u);
            unsigned v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < target);
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = (t1 - t0) / iters;
}
__global__ void k_empty(float* p) { if (p && threadIdx.x == 9999) p[0] = 1.f; }
__global__ void k_red1(float* dst, int n_per_cta, int ksplit) {   // every CTA adds n_per_cta floats; ksplit CTAs share a tile
    float* d = dst + (size_t)(blockIdx.x / ksplit) * n_per_cta;
    for (int i = threadIdx.x; i < n_per_cta; i += blockDim.x) atomicAdd(d + i, 1.0f);
}
__global__ void k_red4(float* dst, int n_per_cta, int ksplit) {
    float* d = dst + (size_t)(blockIdx.x / ksplit) * n_per_cta;
    for (int i = threadIdx.x * 4; i < n_per_cta; i += blockDim.x * 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + i), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
}
__global__ void k_st4(float* dst, int n_per_cta) {
    float4* d = (float4*)(dst + (size_t)blockIdx.x * n_per_cta);
    for (int i = threadIdx.x; i < n_per_cta / 4; i += blockDim.x) d[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
static float time_ms(cudaStream_t st, cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main() {
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int dev = 0; cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
    printf("device %s SMs %d L2 %d MB clock %d kHz\n", pr.name, pr.multiProcessorCount, pr.l2CacheSize >> 20, pr.clockRate);
    float* out; CK(cudaMalloc(&out, 1024));
    // ---- read bandwidth vs working set
    size_t sizes_mb[] = {8, 32, 64, 1024};
    for (size_t mb : sizes_mb) {
        size_t bytes = mb << 20; float4* p; CK(cudaMalloc(&p, bytes)); CK(cudaMemset(p, 0, bytes));
        int reps = mb >= 1024 ? 4 : 64;
        for (int grid : {148, 296, 592, 1184}) {
            k_read<<<grid, 512, 0, st>>>(p, bytes / 16, 2, out);
            CK(cudaEventRecord(e0, st));
            k_read<<<grid, 512, 0, st>>>(p, bytes / 16, reps, out);
            CK(cudaEventRecord(e1, st));
            float ms = time_ms(st, e0, e1);
            printf("read %5zu MB grid %4d x512: %.1f GB/s\n", mb, grid, (double)bytes * reps / ms / 1e6);
        }
        CK(cudaFree(p));
    }
    // ---- grid barrier
    unsigned* ctr; long long* cyc; CK(cudaMalloc(&ctr, 4)); CK(cudaMalloc(&cyc, 8));
    for (int grid : {64, 128, 148}) for (int threads : {128, 256}) {
        int iters = 200; void* args[] = {&iters, &ctr, &cyc};
        CK(cudaMemset(ctr, 0, 4));
        CK(cudaLaunchCooperativeKernel((void*)k_gridbar, dim3(grid), dim3(threads), args, 0, st));
        CK(cudaEventRecord(e0, st));
        CK(cudaLaunchCooperativeKernel((void*)k_gridbar, dim3(grid), dim3(threads), args, 0, st));
        CK(cudaEventRecord(e1, st));
        float ms = time_ms(st, e0, e1); long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
        printf("cg grid.sync grid %3d x%3d: %.2f us/barrier (%lld cycles)\n", grid, threads, ms * 1e3 / iters, c);
        CK(cudaMemset(ctr, 0, 4));
        CK(cudaLaunchCooperativeKernel((void*)k_mybar, dim3(grid), dim3(threads), args, 0, st));
        CK(cudaMemset(ctr, 0, 4));
        CK(cudaEventRecord(e0, st));
        CK(cudaLaunchCooperativeKernel((void*)k_mybar, dim3(grid), dim3(threads), args, 0, st));
        CK(cudaEventRecord(e1, st));
        ms = time_ms(st, e0, e1); CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
        printf("atomic barrier grid %3d x%3d: %.2f us/barrier (%lld cycles)\n", grid, threads, ms * 1e3 / iters, c);
    }
    // ---- launch gap: 1000 dependent empty kernels, stream vs graph
    {
        const int N = 1000;
        for (int i = 0; i < 10; ++i) k_empty<<<148, 128, 0, st>>>(nullptr);
        CK(cudaEventRecord(e0, st));
        for (int i = 0; i < N; ++i) k_empty<<<148, 128, 0, st>>>(nullptr);
        CK(cudaEventRecord(e1, st));
        printf("stream launches: %.2f us per dependent empty kernel\n", time_ms(st, e0, e1) * 1e3 / N);
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < N; ++i) k_empty<<<148, 128, 0, st>>>(nullptr);
        CK(cudaStreamEndCapture(st, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
        CK(cudaGraphLaunch(ge, st));
        CK(cudaEventRecord(e0, st)); CK(cudaGraphLaunch(ge, st)); CK(cudaEventRecord(e1, st));
        printf("graph launches:  %.2f us per dependent empty kernel\n", time_ms(st, e0, e1) * 1e3 / N);
    }
    // ---- split-K reductions: 128 CTAs, 16K floats each
    {
        float* dst; CK(cudaMalloc(&dst, 128 * 16384 * 4)); CK(cudaMemset(dst, 0, 128 * 16384 * 4));
        for (int ks : {1, 4, 8, 16}) {
            k_red1<<<128, 256, 0, st>>>(dst, 16384, ks);
            CK(cudaEventRecord(e0, st)); for (int i = 0; i < 20; ++i) k_red1<<<128, 256, 0, st>>>(dst, 16384, ks); CK(cudaEventRecord(e1, st));
            float a = time_ms(st, e0, e1) * 1e3 / 20;
            k_red4<<<128, 256, 0, st>>>(dst, 16384, ks);
            CK(cudaEventRecord(e0, st)); for (int i = 0; i < 20; ++i) k_red4<<<128, 256, 0, st>>>(dst, 16384, ks); CK(cudaEventRecord(e1, st));
            float b = time_ms(st, e0, e1) * 1e3 / 20;
            printf("128 CTAs x 16K fp32 red, %2d CTAs per tile: scalar %.2f us, v4 %.2f us (incl ~launch)\n", ks, a, b);
        }
        k_st4<<<128, 256, 0, st>>>(dst, 16384);
        CK(cudaEventRecord(e0, st)); for (int i = 0; i < 20; ++i) k_st4<<<128, 256, 0, st>>>(dst, 16384); CK(cudaEventRecord(e1, st));
        printf("128 CTAs x 16K fp32 plain st.v4: %.2f us\n", time_ms(st, e0, e1) * 1e3 / 20);
    }
    return 0;
}
