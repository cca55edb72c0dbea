// erasure_gen.cuh -- GPU-side synthetic erasure channel and payload filler.
//
// i.i.d. model = the reference's decoder-side `data_in` kernel
// (OpenCL/device/ldpc_erasure_decoder_top.cl:57-120): Threefry4x32-20, key {tid = 1, seed},
// counter word 0 = 1 + frame * n + symbol (pre-incremented from 0, never reset between frames
// :75,:96), erased iff ((int)out.v[0] & 0x3F) < PER_numerator_div_64 (:105).
// IID32 (extension) compares the whole 32-bit word against a threshold.
// Bursty model = Matlab/Bursty_Error_Channel_Model_Generator.m:12-47 with u1 = v[0]/2^32,
// u2 = v[1]/2^32 of the same counter (extension; the CPU checker defines the draws the same way).  The chain state at
// the start of a codeword is recovered exactly and independently of batch sharding by walking
// BACKWARDS over the transition uniforms to the nearest symbol whose state map is constant
// (both states lead to the same next state) or to the beginning of the stream (state 0).
#pragma once
#include "device_utils.cuh"

namespace ldpc {

struct GenParams {
    uint32_t *mask;         // [B][NW]
    long long B;
    unsigned long long frame0;
    uint32_t seed;
    int n, NW;
    int model;              // 0 iid64, 1 iid32, 2 bursty
    uint32_t thr;           // iid: P (0..64) or 32-bit threshold
    // bursty: integer thresholds, v <= t  <=>  v / 2^32 <= prob ; flag = prob >= 1 (always)
    uint32_t t_alpha, t_beta, t_p01, t_p10;
    int a_alpha, a_beta, a_p01, a_p10;   // "always" flags (prob >= 1); prob < 0 -> never (thr = 0 and never_* set)
    int n_alpha, n_beta, n_p01, n_p10;   // "never" flags (prob < 0)
};

__device__ __forceinline__ bool le_prob(uint32_t v, uint32_t t, int always, int never)
{
    return always ? true : (never ? false : v <= t);
}

// next state of the two-state chain for transition word v2
__device__ __forceinline__ int bursty_next(const GenParams &p, int state, uint32_t v2)
{
    if (state == 0) return le_prob(v2, p.t_p01, p.a_p01, p.n_p01) ? 1 : 0;
    return le_prob(v2, p.t_p10, p.a_p10, p.n_p10) ? 0 : 1;
}

// one warp per mask word: lane = bit
__global__ void gen_erasures_iid_kernel(const GenParams p)
{
    const long long nwords = p.B * p.NW;
    const int lane = threadIdx.x & 31;
    const uint32_t key[4] = {1u, p.seed, 0u, 0u};
    for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < nwords;
         w += (long long)gridDim.x * (blockDim.x >> 5)) {
        const long long b = w / p.NW;
        const int sym = int(w % p.NW) * 32 + lane;
        bool er = false;
        if (sym < p.n) {
            const uint32_t ctr[4] = {uint32_t(1ull + (p.frame0 + (unsigned long long)b) * (unsigned long long)p.n + sym), 0u, 0u, 0u};
            uint32_t o[4];
            threefry4x32_20(ctr, key, o);
            er = (p.model == 0) ? ((o[0] & 0x3Fu) < p.thr) : (o[0] < p.thr);
        }
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, er);
        if (lane == 0) p.mask[w] = bal;
    }
}

// bursty: one warp per codeword.  Pass 1 finds the state at the codeword's first symbol,
// pass 2 composes the per-symbol state maps with a warp scan, 32 symbols at a time.
__global__ void gen_erasures_bursty_kernel(const GenParams p)
{
    const int lane = threadIdx.x & 31;
    const uint32_t key[4] = {1u, p.seed, 0u, 0u};
    for (long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < p.B;
         b += (long long)gridDim.x * (blockDim.x >> 5)) {
        const unsigned long long pos0 = (p.frame0 + (unsigned long long)b) * (unsigned long long)p.n;  // global symbol index
        // ---- state entering symbol pos0: find the last constant map before it, then replay
        int state = 0;
        unsigned long long start = 0;  // first symbol whose map still has to be applied
        for (unsigned long long hi = pos0; hi > 0;) {
            const unsigned long long lo = hi >= 32 ? hi - 32 : 0;
            const unsigned long long s = lo + lane;
            int f0 = 0, f1 = 1;
            if (s < hi) {
                const uint32_t ctr[4] = {uint32_t(1ull + s), 0u, 0u, 0u};
                uint32_t o[4];
                threefry4x32_20(ctr, key, o);
                f0 = bursty_next(p, 0, o[1]);
                f1 = bursty_next(p, 1, o[1]);
            }
            const unsigned cm = __ballot_sync(0xFFFFFFFFu, (s < hi) && (f0 == f1));
            if (cm) {
                const int L = 31 - __clz(cm);
                state = __shfl_sync(0xFFFFFFFFu, f0, L);
                start = lo + L + 1;
                break;
            }
            hi = lo;
        }
        for (unsigned long long blk = start; blk < pos0; blk += 32) {
            const unsigned long long s2 = blk + lane;
            int g0 = 0, g1 = 1;
            if (s2 < pos0) {
                const uint32_t ctr[4] = {uint32_t(1ull + s2), 0u, 0u, 0u};
                uint32_t o[4];
                threefry4x32_20(ctr, key, o);
                g0 = bursty_next(p, 0, o[1]);
                g1 = bursty_next(p, 1, o[1]);
            }
            const int cnt = (pos0 - blk) < 32 ? int(pos0 - blk) : 32;
            for (int l = 0; l < cnt; l++) {
                const int m0 = __shfl_sync(0xFFFFFFFFu, g0, l);
                const int m1 = __shfl_sync(0xFFFFFFFFu, g1, l);
                state = state ? m1 : m0;
            }
        }
        // ---- the codeword itself
        for (int base = 0; base < p.n; base += 32) {
            const int sym = base + lane;
            uint32_t o[4] = {0, 0, 0, 0};
            int f0 = 0, f1 = 1;
            if (sym < p.n) {
                const uint32_t ctr[4] = {uint32_t(1ull + pos0 + sym), 0u, 0u, 0u};
                threefry4x32_20(ctr, key, o);
                f0 = bursty_next(p, 0, o[1]);
                f1 = bursty_next(p, 1, o[1]);
            }
            // inclusive scan of map composition: (g after f)(x) = g(f(x))
            int c0 = f0, c1 = f1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int p0 = __shfl_up_sync(0xFFFFFFFFu, c0, d);
                const int p1 = __shfl_up_sync(0xFFFFFFFFu, c1, d);
                if (lane >= d) {
                    const int n0 = p0 ? c1 : c0;   // apply earlier map first, then this one
                    const int n1 = p1 ? c1 : c0;
                    c0 = n0; c1 = n1;
                }
            }
            // state BEFORE this lane's symbol = composed map of lanes < lane applied to `state`
            int e0 = __shfl_up_sync(0xFFFFFFFFu, c0, 1);
            int e1 = __shfl_up_sync(0xFFFFFFFFu, c1, 1);
            const int st_in = (lane == 0) ? state : (state ? e1 : e0);
            bool er = false;
            if (sym < p.n)
                er = st_in == 0 ? le_prob(o[0], p.t_alpha, p.a_alpha, p.n_alpha)
                                : le_prob(o[0], p.t_beta, p.a_beta, p.n_beta);
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, er);
            if (lane == 0) p.mask[b * p.NW + (base >> 5)] = bal;
            const int l0 = __shfl_sync(0xFFFFFFFFu, c0, 31);
            const int l1 = __shfl_sync(0xFFFFFFFFu, c1, 31);
            state = state ? l1 : l0;
        }
    }
}

// zero the payload of erased symbols: [B][n][S], S % 16 == 0.  One warp per mask word.
__global__ void zero_erased_kernel(const uint32_t *mask, uint8_t *payload, long long B, int n, int NW, int S)
{
    const long long nwords = B * NW;
    const int lane = threadIdx.x & 31;
    const int vec = S / 16;  // uint4 per symbol
    for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < nwords;
         w += (long long)gridDim.x * (blockDim.x >> 5)) {
        uint32_t x = mask[w];
        const long long b = w / NW;
        const int sym0 = int(w % NW) * 32;
        while (x) {
            const int bit = __ffs(x) - 1;
            x &= x - 1;
            const int sym = sym0 + bit;
            if (sym < n) {
                uint4 *dst = reinterpret_cast<uint4 *>(payload + (size_t(b) * n + sym) * S);
                for (int i = lane; i < vec; i += 32) dst[i] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    }
}

// counter-based random bytes: 16-byte block i = Threefry4x32-20(key {2, seed}, ctr {lo(i), hi(i), 0, 0})
__global__ void fill_random_kernel(uint4 *dst, long long nblocks, uint32_t seed, unsigned long long block0)
{
    const uint32_t key[4] = {2u, seed, 0u, 0u};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nblocks;
         i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long g = block0 + (unsigned long long)i;
        const uint32_t ctr[4] = {uint32_t(g), uint32_t(g >> 32), 0u, 0u};
        uint32_t o[4];
        threefry4x32_20(ctr, key, o);
        dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

}  // namespace ldpc
