// host_inplace.cuh -- the device side of the in-place host entry points (ldpc_decode_host_inplace).
//
// The reference's run() (main.cpp:555-659) reads the whole decoder output back.  Of those k*S bytes per codeword the
// host already holds every symbol that was received; only the symbols the decoder RECOVERED are news.  With the
// caller's buffer page-locked (and therefore addressable from the device) they are written straight into it:
// one S-byte store burst per erased systematic symbol, ~p*k*S instead of k*S bytes per codeword over PCIe.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ldpc {

struct WritebackParams {
    const uint8_t *out;     // [B][k][S] decoder output (device)
    const uint32_t *mask;   // [B][NW] erasure mask (device)
    uint8_t *dst;           // [B][n][S] the caller's codewords, page-locked host memory seen from the device
    long long B;
    int k, n, S, NW;
};

// One warp per codeword (grid-stride).  S/16 lanes copy one symbol, 16 bytes each; a warp-iteration takes as many
// erased symbols of one mask word as fit.  Symbols wider than 512 bytes are copied by the whole warp in turns.
__global__ void __launch_bounds__(256) writeback_erased_kernel(const WritebackParams p)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const int V = p.S >> 4;                               // 16-byte pieces per symbol
    const int spw = V >= 32 ? 1 : 32 / V;                 // symbols per warp-iteration
    const int slot = V >= 32 ? 0 : lane / V, part0 = V >= 32 ? lane : lane % V;
    const int kw = (p.k + 31) >> 5;
    for (long long cw = warp0; cw < p.B; cw += nwarps) {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.out + size_t(cw) * p.k * p.S);
        uint4 *dst = reinterpret_cast<uint4 *>(p.dst + size_t(cw) * p.n * p.S);
        for (int w = 0; w < kw; w++) {
            uint32_t bits = p.mask[cw * p.NW + w];
            if ((w << 5) + 32 > p.k) bits &= 0xFFFFFFFFu >> ((w << 5) + 32 - p.k);
            const int cnt = __popc(bits);
            for (int base = 0; base < cnt; base += spw) {
                const int idx = base + slot;
                if (idx < cnt && slot < spw) {
                    const int j = (w << 5) + int(__fns(bits, 0, idx + 1));
                    for (int q = part0; q < V; q += 32) dst[size_t(j) * V + q] = src[size_t(j) * V + q];
                }
            }
        }
    }
}

// The way up for the same entry point: the device fetches the RECEIVED symbols from the caller's page-locked buffer itself
// (erased ones are not read, their rows in device memory are zeroed), so that ~(1-p)*n*S instead of n*S bytes per codeword
// cross PCIe.  Thread t of a codeword owns the 16-byte pieces t, t + T, ...: a warp reads 512 contiguous bytes, minus the
// erased symbols in them.  Four independent loads per thread keep enough reads in flight to cover the PCIe round trip.
struct GatherParams {
    const uint8_t *src;     // [B][n][S] the caller's codewords (page-locked host memory seen from the device)
    const uint32_t *mask;   // [B][NW] erasure mask (device)
    uint8_t *dst;           // [B][n][S] device copy
    long long B;
    int n, S, NW;
};

__global__ void __launch_bounds__(256) gather_received_kernel(const GatherParams p)
{
    const int V = p.S >> 4;
    const long long per_cw = (long long)p.n * V;                       // 16-byte pieces per codeword
    const long long total = per_cw * p.B;
    const long long T = (long long)gridDim.x * blockDim.x;
    const uint4 *src = reinterpret_cast<const uint4 *>(p.src);
    uint4 *dst = reinterpret_cast<uint4 *>(p.dst);
    for (long long q0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; q0 < total; q0 += 4 * T) {
        uint4 v[4];
        bool live[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long q = q0 + u * T;
            live[u] = false;
            v[u] = make_uint4(0u, 0u, 0u, 0u);
            if (q < total) {
                const long long cw = q / per_cw;
                const int j = int((q - cw * per_cw) / V);
                live[u] = !((p.mask[cw * p.NW + (j >> 5)] >> (j & 31)) & 1u);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (live[u]) v[u] = __ldcs(src + q0 + u * T);
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (q0 + u * T < total) dst[q0 + u * T] = v[u];
    }
}

}  // namespace ldpc
