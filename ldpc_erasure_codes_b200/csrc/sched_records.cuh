// sched_records.cuh -- completes the schedule blobs of the peeling decoder with the records the
// payload executor walks (payload_exec.cuh, "WALK").
//
// An entry (check c, symbol v) of level >= 2 is applied by the executor as
//     row[v] = (XOR of c's RECEIVED members) ^ (XOR of c's PRODUCED members),
// the first part in a dependency-free bulk pass, the second along the level chain.  The produced
// members of an entry are the members of its check that were erased on arrival, the target
// excepted: when the check fires (OpenCL/device/ldpc_erasure_decoder.cl:76-90, cnt == 1) every
// one of them has been recovered by an earlier entry.  They depend on the erasure mask alone, and
// listing them is embarrassingly parallel -- one thread per entry tests the <= RW members of the
// check against the mask -- so it runs here, at full occupancy, rather than inside the serial
// replay of peel_schedule_kernel (whose 8 lanes per codeword and 18 warps per SM would pay for it
// in lockstep) or inside the executor (once per byte slice of a codeword instead of once).
//
// Record (8 bytes, hmat.hpp): 5 x 12-bit symbol indices, padding = the executor's zero row;
// bit 63 = the check has more than five produced members (no member listed: full-row form).
// The kernel also cuts the walk into passes (<= epw entries of one level each: first entry | (count - 1) << 11, a
// warp scan over the level sizes) and finalises hdr[1] = levels | records << 16, hdr[2] |= passes << 16 and the blob
// length.  One warp per codeword.
#pragma once
#include "device_utils.cuh"

namespace ldpc {

struct RecParams {
    const uint32_t *mask;       // [B][NW]
    uint8_t *sched;             // [B][stride] blobs from peel_schedule_kernel
    uint32_t *sched_len;        // [B] blob bytes (rewritten: records included)
    const uint16_t *cidx;       // [m][RW]
    long long B;
    int n, m, RW, NW, stride;
    int blob_cap;               // bytes of a blob the executor can stage in shared memory
    int zrow;                   // record padding: the executor's zero row
    int epw;                    // entries per pass of the executor's walk (32 / lanes per entry)
};

constexpr int kRecWarps = 8;

template <int RWQ>   // uint4 chunks of a padded check row
__global__ void __launch_bounds__(kRecWarps * 32) sched_records_kernel(const RecParams p)
{
    extern __shared__ __align__(16) uint32_t rsh[];
    const int cidx_words = p.m * RWQ * 4;
    const int mw = (p.NW + 1 + 3) & ~3;                 // mask words per warp: NW + the always-zero word row padding points at
    uint32_t *msk = rsh + cidx_words + (threadIdx.x >> 5) * mw;
    const uint32_t padidx = uint32_t(p.NW) * 32u;
    {
        const uint32_t *s32 = reinterpret_cast<const uint32_t *>(p.cidx);
        for (int i = threadIdx.x; i < cidx_words; i += blockDim.x) {
            uint32_t w = s32[i];
            if ((w & 0xFFFFu) == 0xFFFFu) w = (w & 0xFFFF0000u) | padidx;
            if ((w >> 16) == 0xFFFFu) w = (w & 0xFFFFu) | (padidx << 16);
            rsh[i] = w;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned long long z = (unsigned long long)uint32_t(p.zrow);
    for (long long cw = (long long)blockIdx.x * kRecWarps + (threadIdx.x >> 5); cw < p.B; cw += (long long)gridDim.x * kRecWarps) {
        for (int w = lane; w <= p.NW; w += 32) {
            uint32_t x = w < p.NW ? p.mask[cw * p.NW + w] : 0u;
            if (w == p.NW - 1 && (p.n & 31)) x &= 0xFFFFFFFFu >> (32 - (p.n & 31));
            msk[w] = x;
        }
        uint8_t *blob = p.sched + cw * (long long)p.stride;
        uint32_t *hdr = reinterpret_cast<uint32_t *>(blob);
        const uint32_t ne = hdr[0], nl = hdr[1] & 0xFFFFu;
        const uint32_t *ent = hdr + 4;
        const uint16_t *lvo = reinterpret_cast<const uint16_t *>(ent + ne);
        const uint32_t n1 = nl >= 2u ? uint32_t(lvo[1]) : ne;
        // passes of the walk over the levels 2..nl (level index L = 1..nl-1 covers entries [lvo[L], lvo[L+1]))
        const uint32_t pt_off = 16u + 4u * ne + 2u * (nl + 1u);
        uint16_t *g_pt = reinterpret_cast<uint16_t *>(blob + pt_off);
        uint32_t npass = 0;
        for (uint32_t L0 = 1; L0 < nl; L0 += 32) {
            const uint32_t L = L0 + lane;
            const uint32_t s0 = L < nl ? uint32_t(lvo[L]) : 0u, s1 = L < nl ? uint32_t(lvo[L + 1]) : 0u;
            const uint32_t np = (s1 - s0 + uint32_t(p.epw) - 1u) / uint32_t(p.epw);
            uint32_t inc = np;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) inc += t;
            }
            uint32_t at = npass + inc - np;
            for (uint32_t pos = s0; pos < s1; pos += uint32_t(p.epw))
                g_pt[at++] = uint16_t(pos | ((min(uint32_t(p.epw), s1 - pos) - 1u) << 11));
            npass += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
        const uint32_t rec_off = (pt_off + 2u * npass + 7u) & ~7u;
        uint32_t nrec = 0;
        if (n1 < ne && rec_off + 8u <= uint32_t(p.blob_cap)) nrec = min(ne - n1, (uint32_t(p.blob_cap) - rec_off) / 8u);
        unsigned long long *g_rec = reinterpret_cast<unsigned long long *>(blob + rec_off);
        __syncwarp();
        for (uint32_t i = lane; i < nrec; i += 32) {
            const uint32_t e = ent[n1 + i];
            const uint32_t v = e & 0xFFFFu;
            const uint4 *row = reinterpret_cast<const uint4 *>(rsh) + size_t(e >> 16) * RWQ;
            unsigned long long rec = 0ull;
            int nd = 0;
#pragma unroll
            for (int q = 0; q < RWQ; q++) {
                const uint4 r4 = row[q];
                const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const uint32_t u = (t & 1) ? (rr[t >> 1] >> 16) : (rr[t >> 1] & 0xFFFFu);
                    const bool hit = ((msk[u >> 5] >> (u & 31)) & 1u) && u != v;
                    if (hit) { rec = (rec << 12) | u; nd++; }
                }
            }
#pragma unroll
            for (int j = 0; j < 5; j++)
                if (j >= nd) rec = (rec << 12) | z;
            if (nd > 5) rec = z | (z << 12) | (z << 24) | (z << 36) | (z << 48) | (1ull << 63);
            g_rec[i] = rec;
        }
        __syncwarp();
        if (lane == 0) {
            hdr[1] = nl | (nrec << 16);
            hdr[2] = (hdr[2] & 0xFFFFu) | (npass << 16);
            p.sched_len[cw] = (rec_off + 8u * nrec + 15u) & ~15u;
        }
        __syncwarp();
    }
}

}  // namespace ldpc
