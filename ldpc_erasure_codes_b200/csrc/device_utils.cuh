// device_utils.cuh -- sm_100a PTX helpers (mbarrier, TMA / bulk async copies) and Threefry.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace ldpc {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier -----------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a barrier that never completes (a byte-count mismatch, a lost arrive) traps the
// kernel after a few seconds instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); spins++) {
        if (spins > (1u << 26)) {
            printf("libldpc_cuda: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

// ---- TMA (cp.async.bulk.tensor) and 1-D bulk copies ---------------------------------
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2,
                                            uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2),
          "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *smem_src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, int c3,
                                            uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
          "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, const void *smem_src, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (TMA store source)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- Threefry4x32-20 ----------------------------------------------------------------
// Same function as OpenCL/device/threefry.h:299-745 (Random123), written as a loop;
// pinned by the Random123 known-answer vectors through tests/test_parity_gpu.py.
__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, unsigned r) { return (x << r) | (x >> (32u - r)); }

__host__ __device__ __forceinline__ void threefry4x32_20(const uint32_t c[4], const uint32_t k[4], uint32_t out[4])
{
    const uint32_t ks[5] = {k[0], k[1], k[2], k[3], 0x1BD11BDAu ^ k[0] ^ k[1] ^ k[2] ^ k[3]};
    uint32_t x0 = c[0] + ks[0], x1 = c[1] + ks[1], x2 = c[2] + ks[2], x3 = c[3] + ks[3];
#define TF_EVEN(a, b) x0 += x1; x1 = rotl32(x1, a); x1 ^= x0; x2 += x3; x3 = rotl32(x3, b); x3 ^= x2;
#define TF_ODD(a, b)  x0 += x3; x3 = rotl32(x3, a); x3 ^= x0; x2 += x1; x1 = rotl32(x1, b); x1 ^= x2;
#define TF_KEY(s) x0 += ks[(s) % 5]; x1 += ks[((s) + 1) % 5]; x2 += ks[((s) + 2) % 5]; x3 += ks[((s) + 3) % 5] + (s);
    TF_EVEN(10, 26) TF_ODD(11, 21) TF_EVEN(13, 27) TF_ODD(23, 5) TF_KEY(1)
    TF_EVEN(6, 20) TF_ODD(17, 11) TF_EVEN(25, 10) TF_ODD(18, 20) TF_KEY(2)
    TF_EVEN(10, 26) TF_ODD(11, 21) TF_EVEN(13, 27) TF_ODD(23, 5) TF_KEY(3)
    TF_EVEN(6, 20) TF_ODD(17, 11) TF_EVEN(25, 10) TF_ODD(18, 20) TF_KEY(4)
    TF_EVEN(10, 26) TF_ODD(11, 21) TF_EVEN(13, 27) TF_ODD(23, 5) TF_KEY(5)
#undef TF_EVEN
#undef TF_ODD
#undef TF_KEY
    out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
}

}  // namespace ldpc
