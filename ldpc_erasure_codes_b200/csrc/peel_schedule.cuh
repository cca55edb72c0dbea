// peel_schedule.cuh -- "pattern phase" of the peeling decoder.
//
// The reference decoder (OpenCL/device/ldpc_erasure_decoder.cl:49-93) sweeps the checks in
// order, num_iter times, and a check fires when exactly one of its members is erased.  Which
// check recovers which symbol, and in which sweep, depends only on the erasure pattern -- not
// on the payload.  This kernel replays that serial schedule EXACTLY on the erasure mask alone
// (one group of G lanes per codeword) and emits, per codeword, the list of (check, symbol)
// recoveries grouped into dependency levels.  The payload executor (payload_exec.cuh) then
// applies the list to the symbol payload with wide XORs, level by level.
//
// Exactness: a check is "pending" once its erased-member count is 1.  Pending checks are kept
// in two bitmaps -- `cur` (still to be visited in this sweep: index above the check being
// processed) and `nxt` (will be visited in the next sweep).  Popping the lowest set bit of
// `cur` (ballot + ffs) visits exactly the checks that fire in the reference's sweep, in the
// reference's order; the sweep counter reproduces num_iter.  The early stop of
// ldpc_erasure_decoder_old.pro:116-123 is output-neutral and implicit (nothing pending).
//
// Per-check state word: [cnt:5 | level:11 | xor of erased member indices:16].  With cnt == 1
// the xor field IS the erased member.  `level` = deepest recovery this check has seen among
// its members; the recovery it performs itself gets level + 1.
#pragma once
#include <cooperative_groups.h>

#include "device_utils.cuh"

namespace ldpc {
namespace cg = cooperative_groups;

struct PeelParams {
    const uint32_t *mask;       // [B][NW]
    uint8_t *sched;             // [B][stride] schedule blobs
    uint32_t *sched_len;        // [B] bytes of each blob (multiple of 16)
    uint8_t *fail;              // [B] or nullptr
    uint32_t *resid;            // [B] or nullptr: erasures left after peeling (hybrid stage input)
    unsigned long long *stats;  // [8] frames, ldpc_errors, rs_errors, ...
    const uint16_t *cidx;       // [m][RW]
    const uint16_t *vadj;       // [n][VW]
    long long B;
    int n, k, m, RW, VW, NW, MW, stride, max_iter, rs_n, rs_k, groups_per_block, count_stats;
};

__host__ __device__ inline int peel_group_words(int m, int MW, int NW) { return 3 * m + 2 + 2 * MW + NW; }

__device__ __forceinline__ int popc_range(const uint32_t *w, int lo, int hi)  // bits [lo, hi)
{
    int cnt = 0;
    for (int i = lo >> 5; i <= (hi - 1) >> 5; i++) {
        uint32_t x = w[i];
        const int b0 = i << 5;
        if (lo > b0) x &= 0xFFFFFFFFu << (lo - b0);
        if (hi < b0 + 32) x &= 0xFFFFFFFFu >> (b0 + 32 - hi);
        cnt += __popc(x);
    }
    return cnt;
}

template <int G>
__global__ void __launch_bounds__(1024) peel_schedule_kernel(const PeelParams p)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ unsigned int s_stat[4];
    uint16_t *vadj_s = reinterpret_cast<uint16_t *>(smem_raw);
    uint16_t *cidx_s = vadj_s + size_t(p.n) * p.VW;
    uint32_t *grp_base = reinterpret_cast<uint32_t *>(cidx_s + size_t(p.m) * p.RW);

    {   // stage the adjacency tables (uint4 copies; both tables are multiples of 16 bytes)
        const uint4 *src = reinterpret_cast<const uint4 *>(p.vadj);
        uint4 *dst = reinterpret_cast<uint4 *>(vadj_s);
        const int nv = (p.n * p.VW * 2) / 16;
        for (int i = threadIdx.x; i < nv; i += blockDim.x) dst[i] = src[i];
        src = reinterpret_cast<const uint4 *>(p.cidx);
        dst = reinterpret_cast<uint4 *>(cidx_s);
        const int nc = (p.m * p.RW * 2) / 16;
        for (int i = threadIdx.x; i < nc; i += blockDim.x) dst[i] = src[i];
        if (threadIdx.x < 4) s_stat[threadIdx.x] = 0;
    }
    __syncthreads();

    auto tile = cg::tiled_partition<G>(cg::this_thread_block());
    const int lane = tile.thread_rank();
    const int grp = threadIdx.x / G;
    const int m = p.m, MW = p.MW, NW = p.NW, RW = p.RW, VW = p.VW;

    uint32_t *state = grp_base + size_t(grp) * peel_group_words(m, MW, NW);
    uint32_t *ent = state + m;
    uint32_t *lvlcnt = ent + m;          // m + 2 words
    uint32_t *bm_a = lvlcnt + m + 2;
    uint32_t *bm_b = bm_a + MW;
    uint32_t *msk = bm_b + MW;

    unsigned int my_fail = 0, my_rs = 0, my_frames = 0;

    for (long long cw = (long long)blockIdx.x * p.groups_per_block + grp; cw < p.B;
         cw += (long long)gridDim.x * p.groups_per_block) {
        // ---- 1. erasure mask -> shared, erasure counts ---------------------------------
        const uint32_t *gm = p.mask + cw * NW;
        int n_er = 0;
        for (int w = lane; w < NW; w += G) {
            uint32_t x = gm[w];
            if (w == NW - 1 && (p.n & 31)) x &= 0xFFFFFFFFu >> (32 - (p.n & 31));
            msk[w] = x;
            n_er += __popc(x);
        }
        for (int w = lane; w < MW; w += G) { bm_a[w] = 0; bm_b[w] = 0; }
        tile.sync();
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) n_er += tile.shfl_xor(n_er, o);
        int rem_sys = 0;
        if (lane == 0) rem_sys = popc_range(msk, 0, p.k);
        rem_sys = tile.shfl(rem_sys, 0);
        if (p.count_stats && p.rs_n > 0) {  // RS-equivalent MDS counting, perf_tests.cl:70-80
            for (int b = lane; b < p.n / p.rs_n; b += G)
                if (popc_range(msk, b * p.rs_n, (b + 1) * p.rs_n) > p.rs_n - p.rs_k) my_rs++;
        }

        // ---- 2. per-check state from the mask ------------------------------------------
        for (int c = lane; c < m; c += G) {
            const uint4 *row = reinterpret_cast<const uint4 *>(cidx_s + size_t(c) * RW);
            uint32_t cnt = 0, x = 0;
            for (int q = 0; q < RW / 8; q++) {
                const uint4 r4 = row[q];
                const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const uint32_t u = (rr[t >> 1] >> ((t & 1) * 16)) & 0xFFFFu;
                    if (u != 0xFFFFu) {
                        const uint32_t bit = (msk[u >> 5] >> (u & 31)) & 1u;
                        cnt += bit;
                        x ^= bit ? u : 0u;
                    }
                }
            }
            state[c] = (cnt << 27) | x;
            if (cnt == 1) atomicOr(&bm_a[c >> 5], 1u << (c & 31));
        }
        tile.sync();

        // ---- 3. replay of the serial sweeps --------------------------------------------
        uint32_t *cur = bm_a, *nxt = bm_b;
        int sweep = 1, w0 = 0;
        uint32_t ne = 0;
        while (n_er > 0 && p.max_iter > 0) {
            int c = -1;
            while (w0 < MW) {
                const int w = w0 + lane;
                const uint32_t x = (w < MW) ? cur[w] : 0u;
                const unsigned b = tile.ballot(x != 0u);
                if (b) {
                    const int src = __ffs(b) - 1;
                    const uint32_t xw = tile.shfl(x, src);
                    c = ((w0 + src) << 5) + (__ffs(xw) - 1);
                    if (lane == src) cur[w] = x & (x - 1u);
                    w0 += src;
                    break;
                }
                w0 += G;
            }
            tile.sync();
            if (c < 0) {  // sweep over: anything queued for the next one?
                sweep++;
                if (sweep > p.max_iter) break;
                uint32_t *t = cur; cur = nxt; nxt = t;
                w0 = 0;
                bool any = false;
                for (int w = lane; w < MW; w += G) any |= (cur[w] != 0u);
                if (!tile.any(any)) break;
                continue;
            }
            const uint32_t st = state[c];
            if ((st >> 27) != 1u) continue;  // its symbol was recovered by an earlier check
            const uint32_t v = st & 0xFFFFu;
            const uint32_t lvl = ((st >> 16) & 0x7FFu) + 1u;
            if (lane == 0) ent[ne] = v | (uint32_t(c) << 16);
            ne++;
            n_er--;
            if (int(v) < p.k) rem_sys--;
            for (int j = lane; j < VW; j += G) {
                const uint32_t c2 = vadj_s[size_t(v) * VW + j];
                if (c2 != 0xFFFFu) {
                    const uint32_t s2 = state[c2];
                    const uint32_t cnt2 = (s2 >> 27) - 1u;
                    const uint32_t l2 = max((s2 >> 16) & 0x7FFu, lvl);
                    state[c2] = (cnt2 << 27) | (l2 << 16) | ((s2 & 0xFFFFu) ^ v);
                    if (cnt2 == 1u) {
                        if (int(c2) > c) atomicOr(&cur[c2 >> 5], 1u << (c2 & 31));
                        else atomicOr(&nxt[c2 >> 5], 1u << (c2 & 31));
                    }
                }
            }
            tile.sync();
        }

        // ---- 4. counting sort of the recoveries by level, blob to global ---------------
        for (int i = lane; i < m + 2; i += G) lvlcnt[i] = 0;
        tile.sync();
        uint32_t nl = 0;
        for (uint32_t e = lane; e < ne; e += G) {
            const uint32_t L = (state[ent[e] >> 16] >> 16) & 0x7FFu;
            atomicAdd(&lvlcnt[L], 1u);
            nl = max(nl, L);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) nl = max(nl, tile.shfl_xor(nl, o));
        tile.sync();
        uint8_t *blob = p.sched + cw * (long long)p.stride;
        uint32_t *g_ent = reinterpret_cast<uint32_t *>(blob) + 4;
        uint16_t *g_off = reinterpret_cast<uint16_t *>(g_ent + ne);
        uint32_t run = 0;  // exclusive scan over levels 1..nl, G at a time
        for (uint32_t base = 1; base <= nl; base += G) {
            const uint32_t L = base + lane;
            const uint32_t cnt = (L <= nl) ? lvlcnt[L] : 0u;
            uint32_t inc = cnt;
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                const uint32_t t = tile.shfl_up(inc, o);
                if (lane >= o) inc += t;
            }
            const uint32_t excl = run + inc - cnt;
            if (L <= nl) { lvlcnt[L] = excl; g_off[L - 1] = uint16_t(excl); }
            run += tile.shfl(inc, G - 1);
        }
        if (lane == 0) {
            g_off[nl] = uint16_t(ne);
            uint32_t *hdr = reinterpret_cast<uint32_t *>(blob);
            hdr[0] = ne; hdr[1] = nl; hdr[2] = uint32_t(n_er); hdr[3] = uint32_t(rem_sys);
            p.sched_len[cw] = (16u + 4u * ne + 2u * (nl + 1u) + 15u) & ~15u;
            if (p.fail) p.fail[cw] = rem_sys > 0 ? 1 : 0;
            if (p.resid) p.resid[cw] = uint32_t(n_er);
            my_frames++;
            if (rem_sys > 0) my_fail++;
        }
        tile.sync();
        for (uint32_t e = lane; e < ne; e += G) {
            const uint32_t w = ent[e];
            const uint32_t L = (state[w >> 16] >> 16) & 0x7FFu;
            g_ent[atomicAdd(&lvlcnt[L], 1u)] = w;
        }
        tile.sync();
    }

    if (p.count_stats) {
        if (my_frames) atomicAdd(&s_stat[0], my_frames);
        if (my_fail) atomicAdd(&s_stat[1], my_fail);
        if (my_rs) atomicAdd(&s_stat[2], my_rs);
        __syncthreads();
        if (threadIdx.x < 3 && s_stat[threadIdx.x])
            atomicAdd(&p.stats[threadIdx.x], (unsigned long long)s_stat[threadIdx.x]);
    }
}

}  // namespace ldpc
