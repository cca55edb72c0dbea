// fec_packets.cuh -- FEC packet front-ends: codewords <-> packet stream.
//
// Reference: the streaming single-work-item kernels
//   OpenCL/device/ldpc_erasure_encoder_VITA_in_UDP_out.cl:100-104,170-175   (sender: header word, then the symbol)
//   OpenCL/device/ldpc_erasure_decoder_with_reordering_logic.cl:59-131      (receiver: parse, place, clear the flag)
// A packet is one 64-bit FEC header word -- the 32-bit value [class:8 | block:8 | symbol:16] repeated in both
// halves, class code 1 -- followed by the S-byte symbol.  The FPGA handles one packet at a time into two block
// buffers {current, next}; here a whole window of B <= 256 blocks (block numbers are modulo 256) is assembled at
// once: every packet is independent, so it is a scatter -- one sub-warp per packet, 8-byte accesses because
// packets are 8 + S bytes apart.  Both kernels are HBM bound (copy with a header): 2 * (8 + S) bytes per packet.
// Variable payload length (sender :162 `num_longs_used = ceil((packetLen - 2) / 2)`, :186-197; receiver :94-111 reads that
// many 8-byte words): a packet still occupies an 8 + S byte slot, `len8[packet]` says how many payload words are
// valid; the sender zero-fills the rest of the slot, the receiver leaves the rest of the symbol zero.
#pragma once
#include <cstdint>

namespace ldpc {

constexpr uint32_t kFecClass = 1u;

__host__ __device__ inline unsigned long long fec_header(uint32_t block, uint32_t symbol)
{
    const unsigned long long d = ((unsigned long long)(kFecClass & 0xffu) << 24) | ((unsigned long long)(block & 0xffu) << 16) | (symbol & 0xffffu);
    return (d << 32) | d;
}

// cw [B][n][S] -> packets [B*n][8+S]; LPP lanes per packet (a power of two <= 32)
template <int LPP>
__global__ void __launch_bounds__(256) packetize_kernel(const unsigned long long *__restrict__ cw, unsigned long long *__restrict__ packets,
                                                        long long n_packets, int n, int words /* S/8 */, uint32_t block0,
                                                        const uint16_t *__restrict__ len8 /* or nullptr */)
{
    const int sub = threadIdx.x % LPP;
    for (long long pk = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / LPP; pk < n_packets;
         pk += (long long)gridDim.x * blockDim.x / LPP) {
        const unsigned long long *src = cw + pk * words;
        unsigned long long *dst = packets + pk * (words + 1);
        const int used = len8 ? min(int(len8[pk]), words) : words;
        if (sub == 0) dst[0] = fec_header(block0 + uint32_t(pk / n), uint32_t(pk % n));
        for (int w = sub; w < words; w += LPP) dst[1 + w] = w < used ? src[w] : 0ull;
    }
}

// mask words of B blocks: every symbol erased until its packet arrives (receiver :59-68)
__global__ void __launch_bounds__(256) fec_mask_init_kernel(uint32_t *mask, long long B, int n, int NW, uint32_t *counts)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < B * NW; i += (long long)gridDim.x * blockDim.x) {
        const int w = int(i % NW);
        uint32_t x = 0xFFFFFFFFu;
        if (w == NW - 1 && (n & 31)) x >>= 32 - (n & 31);
        mask[i] = x;
    }
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i <= B; i += (long long)gridDim.x * blockDim.x) counts[i] = 0u;
}

// packets [N][8+S] in arrival order -> cw [B][n][S] (zeroed by the caller), mask bits cleared, counts[rel]++ ;
// LPP lanes per packet (a power of two <= 32)
template <int LPP>
__global__ void __launch_bounds__(256) depacketize_kernel(const unsigned long long *__restrict__ packets, long long n_packets,
                                                          unsigned long long *__restrict__ cw, uint32_t *mask, uint32_t *counts,
                                                          long long B, int n, int NW, int words, uint32_t block0,
                                                          const uint16_t *__restrict__ len8 /* or nullptr */)
{
    __shared__ unsigned int hist[257];            // per-block packet counts of this CTA ([256] = dropped)
    for (int i = threadIdx.x; i < 257; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const int sub = threadIdx.x % LPP;
    for (long long pk = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / LPP; pk < n_packets;
         pk += (long long)gridDim.x * blockDim.x / LPP) {
        const unsigned long long *src = packets + pk * (words + 1);
        const unsigned long long h = src[0];
        const uint32_t lo = uint32_t(h), hi = uint32_t(h >> 32);
        const uint32_t rel = (((lo >> 16) & 0xffu) - block0) & 0xffu, sym = lo & 0xffffu;
        const bool ok = lo == hi && ((lo >> 24) & 0xffu) == kFecClass && (long long)rel < B && int(sym) < n;
        if (!ok) {
            if (sub == 0) atomicAdd(&hist[256], 1u);
            continue;
        }
        unsigned long long *dst = cw + ((long long)rel * n + sym) * words;
        const int used = len8 ? min(int(len8[pk]), words) : words;
        for (int w = sub; w < used; w += LPP) dst[w] = src[1 + w];
        if (sub == 0) {
            atomicAnd(&mask[(long long)rel * NW + (sym >> 5)], ~(1u << (sym & 31)));
            atomicAdd(&hist[rel], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 257; i += blockDim.x)
        if (hist[i]) atomicAdd(&counts[i == 256 ? B : i], hist[i]);
}

}  // namespace ldpc
