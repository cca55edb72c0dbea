// Microbenchmark: issue rate of the warp-level mma.sync forms that a GF(2) matrix product could use on sm_100a
// (b1 and.popc m16n8k256, s8 m16n8k32, bf16 m16n8k16).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rates mma_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(256) k_rate(int iters, int *sink, uint32_t seed)
{
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u, b0 = a0 * 11u, b1 = a0 * 13u;
    int c[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 2)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.xor.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 0x7fffffff) sink[0] = s;
}

template <int KIND>
void run(const char *name, double macs_per_mma)
{
    int *sink; cudaMalloc(&sink, 4);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int grid = prop.multiProcessorCount * 4, iters = 20000;
    k_rate<KIND><<<grid, 256>>>(100, sink, 1);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_rate<KIND><<<grid, 256>>>(iters, sink, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    const double mmas = double(grid) * 8 * iters * 8;
    printf("{\"form\": \"%s\", \"ms\": %.3f, \"mma_per_s\": %.3e, \"Tmac_per_s\": %.1f, \"err\": \"%s\"}\n", name, ms, mmas / ms * 1e3,
           mmas * macs_per_mma / ms * 1e3 / 1e12, cudaGetErrorString(e));
    cudaFree(sink);
}

int main()
{
    run<0>("b1.and.popc m16n8k256", 16.0 * 8 * 256);
    run<3>("b1.xor.popc m16n8k256", 16.0 * 8 * 256);
    run<1>("s8 m16n8k32", 16.0 * 8 * 32);
    run<2>("bf16 m16n8k16", 16.0 * 8 * 16);
    return 0;
}
