// peel_schedule.cuh -- "pattern phase" of the peeling decoder.
//
// The reference decoder (OpenCL/device/ldpc_erasure_decoder.cl:49-93) sweeps the checks in
// order, num_iter times, and a check fires when exactly one of its members is erased.  Which
// check recovers which symbol, and in which sweep, depends only on the erasure pattern -- not
// on the payload.  This kernel replays that serial schedule EXACTLY on the erasure mask alone
// and emits, per codeword, the list of (check, symbol) recoveries grouped into dependency
// levels.  The payload executor (payload_exec.cuh) then applies the list to the symbol payload
// with wide XORs, level by level.
//
// Exactness: a check is "pending" once its erased-member count is 1.  Pending checks are kept
// in two bitmaps -- `cur` (still to be visited in this sweep: index above the check being
// processed) and `nxt` (will be visited in the next sweep).  Popping the lowest set bit of
// `cur` (ballot + ffs) visits exactly the checks that fire in the reference's sweep, in the
// reference's order; the sweep counter reproduces num_iter.  The early stop of
// ldpc_erasure_decoder_old.pro:116-123 is output-neutral and implicit (nothing pending).
//
// Mapping: G = 8 lanes per codeword, four codewords per warp, and the four groups of a warp
// run in LOCKSTEP (one loop, warp-uniform trip count, full-mask ballots/shuffles): diverged
// sub-warp groups would be serialised by the SIMT scheduler.  Per-codeword shared state is
// 2.4 KB for (2040,1530), so ~72 codewords are in flight per SM.  All shared accesses go
// through one base pointer and 32-bit word offsets (the cur/nxt swap is an offset swap).
//
// Per-check state word: [cnt:5 | level:11 | xor of erased member indices:16].  With cnt == 1
// the xor field IS the erased member.  `level` = deepest recovery seen among the check's
// members; the recovery the check performs itself gets level + 1.  A check that has fired is
// marked cnt = 31 with (level, symbol) of its recovery kept in the word, so the recovery list
// is read back from the state array (no separate entry list).
#pragma once
#include "device_utils.cuh"

namespace ldpc {

struct PeelParams {
    const uint32_t *mask;       // [B][NW]
    unsigned int *work_ctr;     // zeroed before the launch: next unclaimed codeword
    uint8_t *sched;             // [B][stride] schedule blobs
    uint32_t *sched_len;        // [B] bytes of each blob (multiple of 16)
    uint8_t *fail;              // [B] or nullptr: a systematic symbol is still erased (perf_tests.cl:215-228)
    uint8_t *fail_any;          // [B] or nullptr: any of the n symbols is still erased (LDPCErasureCodes_MessagePassingAlgSim.m:229-236)
    uint32_t *resid;            // [B] or nullptr: erasures left after peeling (hybrid stage input)
    unsigned long long *stats;  // [8] frames, ldpc_errors, rs_errors, (hybrid: [3..5]), [6] frames with any symbol left erased
    const uint16_t *cidx;       // [m][RW]
    const uint16_t *vadj;       // [n][VW]
    long long B;
    int n, k, m, RW, VW, NW, MW, stride, max_iter, rs_n, rs_k, groups_per_block, count_stats;
    unsigned int *ge_list;      // hybrid mode: codewords that still have erasures are appended here ...
    unsigned int *ge_count;     // ... (nullptr otherwise)
};

constexpr int kPeelG = 8;          // lanes per codeword (4 codewords per warp, in lockstep); 4 was measured slower
constexpr uint32_t kFired = 31u;

// per-codeword shared words: state[m] | cur[BMW] | nxt[BMW] | msk[NW + 1] + 3 scratch words, each 4-word aligned
__host__ __device__ inline int peel_bitmap_words(int MW) { return MW <= 16 ? 16 : (MW <= 32 ? 32 : 64); }
__host__ __device__ inline int peel_group_words(int m, int MW, int NW)
{
    return ((m + 3) & ~3) + 2 * peel_bitmap_words(MW) + ((NW + 4 + 3) & ~3);
}

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
__device__ __forceinline__ int group_sum(int v)  // over the G lanes of a group (xor < G stays inside it)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
template <int G>
__device__ __forceinline__ uint32_t group_max(uint32_t v)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}

template <int BMW, int VW>   // BMW = bitmap words per codeword (16, 32, 64); VW = padded variable degree (4, 8)
__global__ void __launch_bounds__(1024) peel_schedule_kernel(const PeelParams p)
{
    constexpr int G = kPeelG;
    constexpr int WPL = BMW / G;              // bitmap words scanned per lane (a multiple of 4)
    constexpr int NPL = (VW + G - 1) / G;     // neighbour checks handled per lane
    constexpr unsigned GMASK = (1u << G) - 1u;
    static_assert(WPL == 2 || WPL % 4 == 0, "bitmap words per lane are read with 64- or 128-bit loads");
    constexpr unsigned FULL = 0xFFFFFFFFu;
    extern __shared__ __align__(16) uint32_t sh[];
    __shared__ unsigned int s_stat[4];
    const int m = p.m, NW = p.NW, RW = p.RW;
    const int m4 = (m + 3) & ~3;
    const int vadj_words = ((p.n * VW + 7) & ~7) / 2;    // u16 table sizes in 32-bit words (the host pads the table to 16 bytes)
    const int cidx_words = (m * RW) / 2;
    uint16_t *cidx_s = reinterpret_cast<uint16_t *>(sh + vadj_words);
    const int grp_base = vadj_words + cidx_words;
    const uint32_t padidx = uint32_t(NW) * 32u;   // row padding points at the always-zero word msk[NW]

    {   // stage the adjacency tables; check-row padding is redirected to the zero mask bit
        const uint4 *src = reinterpret_cast<const uint4 *>(p.vadj);
        uint4 *dst = reinterpret_cast<uint4 *>(sh);
        for (int i = threadIdx.x; i < vadj_words / 4; i += blockDim.x) dst[i] = src[i];
        const uint32_t *s32 = reinterpret_cast<const uint32_t *>(p.cidx);
        for (int i = threadIdx.x; i < cidx_words; i += blockDim.x) {
            uint32_t w = s32[i];
            if ((w & 0xFFFFu) == 0xFFFFu) w = (w & 0xFFFF0000u) | padidx;
            if ((w >> 16) == 0xFFFFu) w = (w & 0xFFFFu) | (padidx << 16);
            sh[vadj_words + i] = w;
        }
        if (threadIdx.x < 4) s_stat[threadIdx.x] = 0;
    }
    __syncthreads();

    const int wl = threadIdx.x & 31;          // lane in warp
    const int lane = wl % G;                  // lane in group
    const int gshift = wl - lane;             // first warp lane of this group
    const int grp = threadIdx.x / G;
    const int state_off = grp_base + grp * peel_group_words(m, p.MW, NW);
    const int bma_off = state_off + m4;
    const int bmb_off = bma_off + BMW;
    const int msk_off = bmb_off + BMW;
    const int LC = 2 * BMW + NW;              // level counters that fit in the dead bitmap/mask area
    const int lc_off = bma_off;
    const uint32_t sh_a = smem_u32(sh);
    const int dummy_off = msk_off + NW + 1 + lane % 3;     // (distinct lanes may share a scratch word: values are never read back)   // scratch words behind the mask 

    unsigned int my_fail = 0, my_rs = 0, my_frames = 0, my_any = 0;
    // Codewords are claimed a warp-load (32/G) at a time from a global counter: replay lengths vary by
    // codeword, and a static split leaves most of the GPU idle in the last round.
    auto claim = [&]() -> long long {
        unsigned int base = 0;
        if (wl == 0) base = atomicAdd(p.work_ctr, (unsigned int)(32 / G));
        base = __shfl_sync(FULL, base, 0);
        return (long long)base + wl / G;
    };

    for (long long cw = claim(); __any_sync(FULL, cw < p.B); cw = claim()) {
        const bool valid = cw < p.B;
        // ---- 1. erasure mask -> shared, erasure counts ---------------------------------
        int n_er = 0, rem_sys = 0;
        for (int w = lane; w <= NW; w += G) {
            uint32_t x = (valid && w < NW) ? p.mask[cw * NW + w] : 0u;
            if (w == NW - 1 && (p.n & 31)) x &= FULL >> (32 - (p.n & 31));
            sh[msk_off + w] = x;
            n_er += __popc(x);
            const int b0 = w << 5;   // systematic part: bits below k
            uint32_t xs = x;
            if (b0 >= p.k) xs = 0u;
            else if (b0 + 32 > p.k) xs &= FULL >> (b0 + 32 - p.k);
            rem_sys += __popc(xs);
        }
#pragma unroll
        for (int i = 0; i < 2 * WPL; i++) sh[bma_off + lane + i * G] = 0u;
        n_er = group_sum<G>(n_er);
        rem_sys = group_sum<G>(rem_sys);
        __syncwarp();
        if (p.count_stats && p.rs_n > 0 && valid) {  // RS-equivalent MDS counting, perf_tests.cl:70-80
            for (int b = lane; b < p.n / p.rs_n; b += G)
                if (popc_range(sh + msk_off, b * p.rs_n, (b + 1) * p.rs_n) > p.rs_n - p.rs_k) my_rs++;
        }

        // ---- 2. per-check state from the mask ------------------------------------------
        for (int c = lane; c < m; c += G) {
            const uint4 *row = reinterpret_cast<const uint4 *>(cidx_s + c * RW);
            uint32_t cnt = 0, x = 0;
            for (int q = 0; q < RW / 8; q++) {
                const uint4 r4 = row[q];
                const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const uint32_t u = (t & 1) ? (rr[t >> 1] >> 16) : (rr[t >> 1] & 0xFFFFu);
                    const uint32_t bit = (sh[msk_off + (u >> 5)] >> (u & 31)) & 1u;
                    cnt += bit;
                    x ^= (0u - bit) & u;
                }
            }
            sh[state_off + c] = (cnt << 27) | x;
            if (cnt == 1) atomicOr(&sh[bma_off + (c >> 5)], 1u << (c & 31));
        }
        __syncwarp();

        // ---- 3. replay of the serial sweeps, four codewords per warp in lockstep -------
        // (explicit shared-space addresses and PTX accessors keep this loop, the kernel's hot spot, to
        //  ~90 instructions per step; cur/nxt swap by swapping two address registers)
        uint32_t cur_a = sh_a + 4u * bma_off, nxt_a = sh_a + 4u * bmb_off;
        const uint32_t st_a = sh_a + 4u * state_off;
        const uint32_t dummy_a = sh_a + 4u * dummy_off;
        const uint32_t vadj_a = sh_a;                       // the variable->check table sits first
        int sweep = 1;
        uint32_t ne = 0;
        bool active = valid && n_er > 0 && p.max_iter > 0;
        bool prev_empty = false;
        while (__any_sync(FULL, active)) {
            // lowest pending check of this sweep: lane L owns bitmap words [L*WPL, (L+1)*WPL)
            uint32_t xs[WPL];
            if (WPL == 2) {
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(xs[0]), "=r"(xs[WPL - 1]) : "r"(cur_a + lane * 8u) : "memory");
            } else {
#pragma unroll
                for (int q = 0; q < WPL / 4; q++)
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(xs[q * 4 + 0]), "=r"(xs[(q * 4 + 1) % WPL]), "=r"(xs[(q * 4 + 2) % WPL]), "=r"(xs[(q * 4 + 3) % WPL])
                                 : "r"(cur_a + lane * (WPL * 4u) + q * 16u) : "memory");
            }
            int myc = -1;
#pragma unroll
            for (int i = WPL - 1; i >= 0; i--)
                if (xs[i]) myc = ((lane * WPL + i) << 5) + __ffs(xs[i]) - 1;
            const unsigned gb = (__ballot_sync(FULL, myc >= 0) >> gshift) & GMASK;
            const int src = __ffs(gb) - 1;                                   // -1: nothing pending
            const int c = __shfl_sync(FULL, myc, gshift + (src & (G - 1)));  // (src = -1 reads a lane whose myc is -1)
            // Straight-line, select-based body: the four codewords of the warp take the same
            // instruction path whatever they do this step (pop / stale pop / end of sweep / idle).
            const bool pop = active && c >= 0;
            const bool swp = active && c < 0;
            if (pop && lane == src)
                asm volatile("red.shared.and.b32 [%0], %1;" ::"r"(cur_a + 4u * (uint32_t(c) >> 5)), "r"(~(1u << (c & 31))) : "memory");
            uint32_t st;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(st) : "r"(st_a + (pop ? 4u * uint32_t(c) : 0u)) : "memory");
            const bool fire = pop && (st >> 27) == 1u;   // else: its symbol was recovered by an earlier check
            const uint32_t v = fire ? (st & 0xFFFFu) : 0u;
            const uint32_t lvl = ((st >> 16) & 0x7FFu) + 1u;
            ne += fire ? 1u : 0u;
            n_er -= fire ? 1 : 0;
            rem_sys -= (fire && int(v) < p.k) ? 1 : 0;
#pragma unroll
            for (int t = 0; t < NPL; t++) {
                const int j = lane * NPL + t;
                uint32_t c2 = 0xFFFFu;
                if (j < VW) {
                    unsigned short c2h;
                    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(c2h) : "r"(vadj_a + 2u * (v * VW + j)) : "memory");
                    c2 = c2h;
                }
                const bool upd = fire && c2 != 0xFFFFu;
                const uint32_t a2 = upd ? st_a + 4u * c2 : dummy_a;        // idle lanes hit a scratch word
                uint32_t s2;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(s2) : "r"(a2) : "memory");
                const uint32_t cnt2 = (s2 >> 27) - 1u;
                const uint32_t l2 = max((s2 >> 16) & 0x7FFu, lvl);
                const bool self = int(c2) == c;
                const uint32_t nv = self ? ((kFired << 27) | (lvl << 16) | v) : ((cnt2 << 27) | (l2 << 16) | ((s2 & 0xFFFFu) ^ v));
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(a2), "r"(nv) : "memory");
                if (upd && !self && cnt2 == 1u)
                    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"((int(c2) > c ? cur_a : nxt_a) + 4u * (c2 >> 5)), "r"(1u << (c2 & 31)) : "memory");
                if (upd && !self && cnt2 == 0u) {
                    // the check was pending and its last erased member has just been recovered by another check: it would be
                    // popped later and do nothing (13 % of the pops at the threshold); its bit is in one of the two bitmaps
                    asm volatile("red.shared.and.b32 [%0], %1;" ::"r"(cur_a + 4u * (c2 >> 5)), "r"(~(1u << (c2 & 31))) : "memory");
                    asm volatile("red.shared.and.b32 [%0], %1;" ::"r"(nxt_a + 4u * (c2 >> 5)), "r"(~(1u << (c2 & 31))) : "memory");
                }
            }
            // end of sweep: swap the bitmaps; stop at the cap or at the fixed point (two empty sweeps)
            sweep += swp ? 1 : 0;
            const uint32_t t_a = cur_a;
            cur_a = swp ? nxt_a : cur_a;
            nxt_a = swp ? t_a : nxt_a;
            active = active && n_er > 0 && !(swp && (sweep > p.max_iter || prev_empty));
            prev_empty = swp ? true : (pop ? false : prev_empty);
            __syncwarp();
        }

        // ---- 4. counting sort of the recoveries by level, blob to global ---------------
        for (int i = lane; i < LC; i += G) sh[lc_off + i] = 0u;
        __syncwarp();
        uint32_t nl = 0;
        for (int c = lane; c < m; c += G) {
            const uint32_t st = sh[state_off + c];
            if (valid && (st >> 27) == kFired) {
                const uint32_t L = (st >> 16) & 0x7FFu;
                if (int(L) < LC) atomicAdd(&sh[lc_off + L], 1u);
                nl = max(nl, L);
            }
        }
        nl = group_max<G>(nl);
        __syncwarp();
        uint8_t *blob = p.sched + (valid ? cw : 0) * (long long)p.stride;
        uint32_t *g_ent = reinterpret_cast<uint32_t *>(blob) + 4;
        uint16_t *g_off = reinterpret_cast<uint16_t *>(g_ent + ne);
        const unsigned gm = GMASK << gshift;
        if (int(nl) < LC) {
            uint32_t run = 0;  // exclusive scan over levels 1..nl, G at a time
            for (uint32_t base = 1; base <= nl; base += G) {
                const uint32_t L = base + lane;
                const uint32_t cnt = (L <= nl) ? sh[lc_off + L] : 0u;
                uint32_t inc = cnt;
#pragma unroll
                for (int o = 1; o < G; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(gm, inc, o, G);
                    if (lane >= o) inc += t;
                }
                const uint32_t excl = run + inc - cnt;
                if (L <= nl) { sh[lc_off + L] = excl; if (valid) g_off[L - 1] = uint16_t(excl); }
                run += __shfl_sync(gm, inc, G - 1, G);
            }
            __syncwarp(gm);
            for (int c = lane; c < m; c += G) {
                const uint32_t st = sh[state_off + c];
                if (valid && (st >> 27) == kFired) {
                    const uint32_t L = (st >> 16) & 0x7FFu;
                    g_ent[atomicAdd(&sh[lc_off + L], 1u)] = (st & 0xFFFFu) | (uint32_t(c) << 16);
                }
            }
        } else if (valid) {
            // more levels than counters (never seen on the committed codes): one pass per level
            uint32_t run = 0;
            for (uint32_t L = 1; L <= nl; L++) {
                if (lane == 0) g_off[L - 1] = uint16_t(run);
                for (int c0 = 0; c0 < m; c0 += G) {
                    const int c = c0 + lane;
                    const uint32_t st = (c < m) ? sh[state_off + c] : 0u;
                    const bool hit = (st >> 27) == kFired && ((st >> 16) & 0x7FFu) == L;
                    const unsigned hb = (__ballot_sync(gm, hit) >> gshift) & GMASK;
                    if (hit) g_ent[run + __popc(hb & ((1u << lane) - 1u))] = (st & 0xFFFFu) | (uint32_t(c) << 16);
                    run += __popc(hb);
                }
            }
        }
        if (valid && lane == 0) {
            g_off[nl] = uint16_t(ne);
            uint32_t *hdr = reinterpret_cast<uint32_t *>(blob);
            hdr[0] = ne; hdr[1] = nl; hdr[2] = uint32_t(n_er); hdr[3] = uint32_t(rem_sys);
            p.sched_len[cw] = (16u + 4u * ne + 2u * (nl + 1u) + 15u) & ~15u;
            if (p.fail) p.fail[cw] = rem_sys > 0 ? 1 : 0;
            if (p.fail_any) p.fail_any[cw] = n_er > 0 ? 1 : 0;
            if (p.resid) p.resid[cw] = uint32_t(n_er);
            if (p.ge_list && n_er > 0) p.ge_list[atomicAdd(p.ge_count, 1u)] = (unsigned int)cw;
            my_frames++;
            if (rem_sys > 0) my_fail++;
            if (n_er > 0) my_any++;
        }
        __syncwarp();
    }

    if (p.count_stats) {
        if (my_frames) atomicAdd(&s_stat[0], my_frames);
        if (my_fail) atomicAdd(&s_stat[1], my_fail);
        if (my_rs) atomicAdd(&s_stat[2], my_rs);
        if (my_any) atomicAdd(&s_stat[3], my_any);
        __syncthreads();
        if (threadIdx.x < 4 && s_stat[threadIdx.x])   // [6] = frames with any symbol unknown after peeling
            atomicAdd(&p.stats[threadIdx.x < 3 ? threadIdx.x : 6], (unsigned long long)s_stat[threadIdx.x]);
    }
}

}  // namespace ldpc
