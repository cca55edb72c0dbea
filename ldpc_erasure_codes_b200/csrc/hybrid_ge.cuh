// hybrid_ge.cuh -- hybrid-ML stage: GF(2) elimination on the residual stopping set.
//
// Reference: Matlab/My_LDPC_HybridML_Erasure_Decoder.m:48-87.  After the capped peeling sweeps the
// MATLAB code solves  H(:,E) x = H(:,known) y(known)  for the still-erased set E by Gaussian
// elimination with row swaps (:57-75) and Jordan back-elimination (:77-86); it aborts when a column
// has no pivot (:59-62), i.e. exactly when rank(H(:,E)) < |E|.  On success the solution is unique,
// so ANY exact solver returns the same bytes; on abort the contract (SURVEY a-10) is "leave the
// peeling result, report ml_fail".  This kernel therefore uses its own GPU-friendly elimination:
//
//   one CTA per codeword that left the peeling stage with erasures (compacted list from the peel
//   kernel).  The m x e bit matrix A = H(:,E) is built bit-packed in shared memory next to an
//   m x m identity, [A | I], one row per check.  Gauss-Jordan: per column, the lowest unused row
//   with a 1 is the pivot (atomicMin over the CTA), the pivot row is broadcast through shared memory
//   and XORed -- 32 words per warp instruction -- into every other row that has the bit.  No pivot
//   => rank deficient => ml_fail.  At the end row pivot(j) of the I part lists which check
//   syndromes add up to unknown j.  Payload: the syndromes  s_r = XOR of the KNOWN members of check r
//   (received or peeled; the executor stored all n symbols of such codewords) are formed in shared
//   memory, 64 bytes of every symbol at a time, and every erased SYSTEMATIC symbol is written as the
//   XOR of its syndromes straight into the decoder output.  There is no dependency chain in the
//   payload part: all unknowns are independent combinations of the syndromes.
//
// Codes whose [A | I] does not fit in shared memory (m = 1000, 2000) keep the matrix in a per-CTA
// global workspace (it stays L2 resident); the code path is the same.
#pragma once
#include <string>

#include "../../include/ldpc_cuda.h"
#include "device_utils.cuh"
#include "hmat.hpp"

namespace ldpc {

constexpr int kGeThreads = 512;

struct GeParams {
    const uint32_t *mask;            // [B][NW] erasure masks as received
    const uint8_t *sched;            // schedule blobs: the symbols peeling recovered
    const unsigned int *list;        // codewords that still have erasures
    const unsigned int *list_count;
    const uint8_t *full;             // [B][n][S] every symbol after peeling (valid for listed codewords)
    uint8_t *out;                    // [B][k][S]
    uint8_t *fail;                   // [B]
    unsigned long long *stats;       // [3] ml_attempts, [4] ml_failures, [5] ml_recovered
    const uint16_t *cidx;            // [m][RW]
    uint32_t *gmat;                  // per-CTA global workspace for [A | I], or nullptr (shared memory)
    int n, k, m, RW, NW, MW, S, stride;
    int RSW;                         // words per matrix row: 2 * MW + 1 (odd: rows start in different banks)
};

__host__ __device__ inline size_t ge_small_bytes(int m, int NW)
{   // er[NW] pref[NW+1] prow[RSW<=257] varlist[m] pivrow[m] used[m] inv[m] + scalars, generously rounded
    return size_t(NW) * 4 + size_t(NW + 1) * 4 + 260 * 4 + size_t(m) * 2 * 2 + size_t(m) * 2 + 64 + 64;
}

template <bool GMAT>   // where [A | I] lives: per-CTA global workspace (true) or shared memory (false)
__global__ void __launch_bounds__(kGeThreads) hybrid_ge_kernel(const GeParams p)
{
    extern __shared__ __align__(16) uint8_t ge_smem[];
    __shared__ int s_piv;
    __shared__ int s_e;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = p.m, NW = p.NW, MW = p.MW, RSW = p.RSW, RW = p.RW;

    // shared layout: [synd m*64][er NW][pref NW+1][prow RSW][varlist m u16][pivrow m u16][used m u8][inv m u8][mat ...]
    uint8_t *synd = ge_smem;
    uint32_t *er = reinterpret_cast<uint32_t *>(synd + size_t(m) * 64);
    uint32_t *pref = er + NW;
    uint32_t *prow = pref + NW + 1;
    uint16_t *varlist = reinterpret_cast<uint16_t *>(prow + 260);
    uint16_t *pivrow = varlist + m;
    uint8_t *used = reinterpret_cast<uint8_t *>(pivrow + m);
    uint8_t *inv = used + m;
    uint32_t *mat_s = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(inv + m) + 15) & ~uintptr_t(15));
    uint32_t *mat = GMAT ? p.gmat + size_t(blockIdx.x) * m * RSW : mat_s;

    const unsigned int count = *p.list_count;
    for (unsigned int li = blockIdx.x; li < count; li += gridDim.x) {
        const long long cw = p.list[li];
        // ---- 1. residual erased set: received mask minus what peeling recovered ------------
        for (int w = tid; w < NW; w += kGeThreads) {
            uint32_t x = p.mask[cw * NW + w];
            if (w == NW - 1 && (p.n & 31)) x &= 0xFFFFFFFFu >> (32 - (p.n & 31));
            er[w] = x;
        }
        __syncthreads();
        const uint32_t *hdr = reinterpret_cast<const uint32_t *>(p.sched + cw * (long long)p.stride);
        const int ne = int(hdr[0]);
        for (int i = tid; i < ne; i += kGeThreads) {
            const uint32_t v = hdr[4 + i] & 0xFFFFu;
            atomicAnd(&er[v >> 5], ~(1u << (v & 31)));
        }
        __syncthreads();
        if (tid == 0) {   // column index of an erased symbol = its rank in the erased set
            uint32_t run = 0;
            for (int w = 0; w < NW; w++) { pref[w] = run; run += __popc(er[w]); }
            pref[NW] = run;
            s_e = int(run);
        }
        __syncthreads();
        const int e = s_e;
        bool ok = e <= m;   // more unknowns than checks cannot have full column rank
        if (ok) {
            for (int w = tid; w < NW; w += kGeThreads) {
                uint32_t x = er[w];
                uint32_t b = pref[w];
                while (x) { varlist[b++] = uint16_t(w * 32 + __ffs(x) - 1); x &= x - 1u; }
            }
            // ---- 2. [A | I], one row per check ----------------------------------------------
            for (int i = tid; i < m * RSW; i += kGeThreads) mat[i] = 0u;
            __syncthreads();
            for (int r = tid; r < m; r += kGeThreads) {
                uint32_t *row = mat + size_t(r) * RSW;
                bool any = false;
                for (int j = 0; j < RW; j++) {
                    const uint32_t u = p.cidx[r * RW + j];
                    if (u == 0xFFFFu) continue;
                    const uint32_t x = er[u >> 5];
                    if ((x >> (u & 31)) & 1u) {
                        const uint32_t col = pref[u >> 5] + __popc(x & ((1u << (u & 31)) - 1u));
                        row[col >> 5] |= 1u << (col & 31);
                        any = true;
                    }
                }
                row[MW + (r >> 5)] |= 1u << (r & 31);
                inv[r] = any ? 1 : 0;
                used[r] = 0;
            }
            __syncthreads();
            // ---- 3. Gauss-Jordan, pivot row broadcast through shared memory --------------------
            for (int col = 0; col < e; col++) {
                if (tid == 0) s_piv = 0x7FFFFFFF;
                __syncthreads();
                const int cwrd = col >> 5;
                const uint32_t cbit = 1u << (col & 31);
                for (int r = tid; r < m; r += kGeThreads)
                    if (inv[r] && !used[r] && (mat[size_t(r) * RSW + cwrd] & cbit)) atomicMin(&s_piv, r);
                __syncthreads();
                const int piv = s_piv;
                if (piv == 0x7FFFFFFF) { ok = false; break; }   // no pivot: rank deficient (HybridML.m:59-62)
                for (int i = tid; i < RSW; i += kGeThreads) prow[i] = mat[size_t(piv) * RSW + i];
                if (tid == 0) { used[piv] = 1; pivrow[col] = uint16_t(piv); }
                __syncthreads();
                // rows are tested 32 at a time (one per lane, ballot), then each row that has the bit is
                // updated by the whole warp, one matrix word per lane
                for (int r0 = warp * 32; r0 < m; r0 += kGeThreads) {
                    const int r = r0 + lane;
                    const bool hit = r < m && r != piv && inv[r] && (mat[size_t(r) * RSW + cwrd] & cbit);
                    unsigned todo = __ballot_sync(0xFFFFFFFFu, hit);
                    while (todo) {
                        const int rr = r0 + __ffs(todo) - 1;
                        todo &= todo - 1u;
                        uint32_t *row = mat + size_t(rr) * RSW;
                        for (int i = lane; i < 2 * MW; i += 32) row[i] ^= prow[i];   // (word 2*MW is padding)
                    }
                }
                __syncthreads();
            }
        }
        // ---- 4. payload: syndromes, then every erased systematic symbol -----------------------
        if (ok && p.full) {   // (full == nullptr: error-rate run, pattern only)
            const uint8_t *full = p.full + size_t(cw) * p.n * p.S;
            uint8_t *out = p.out + size_t(cw) * p.k * p.S;
            const int qd = tid & 3;             // 16-byte quarter of a 64-byte chunk
            for (int ch = 0; ch < p.S; ch += 64) {
                const int cb = min(64, p.S - ch);   // S is a multiple of 16
                for (int r = tid >> 2; r < m; r += kGeThreads / 4) {
                    if (!inv[r] || qd * 16 >= cb) continue;
                    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
                    for (int j = 0; j < RW; j++) {
                        const uint32_t u = p.cidx[r * RW + j];
                        if (u == 0xFFFFu || ((er[u >> 5] >> (u & 31)) & 1u)) continue;
                        const uint4 v = *reinterpret_cast<const uint4 *>(full + size_t(u) * p.S + ch + qd * 16);
                        acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
                    }
                    *reinterpret_cast<uint4 *>(synd + size_t(r) * 64 + qd * 16) = acc;
                }
                __syncthreads();
                for (int col = tid >> 2; col < e; col += kGeThreads / 4) {
                    const int u = varlist[col];
                    if (u >= p.k || qd * 16 >= cb) continue;       // only systematic symbols are output
                    const uint32_t *trow = mat + size_t(pivrow[col]) * RSW + MW;
                    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
                    for (int w = 0; w < MW; w++) {
                        uint32_t bits = trow[w];
                        while (bits) {
                            const int r = w * 32 + __ffs(bits) - 1;
                            bits &= bits - 1u;
                            const uint4 v = *reinterpret_cast<const uint4 *>(synd + size_t(r) * 64 + qd * 16);
                            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
                        }
                    }
                    *reinterpret_cast<uint4 *>(out + size_t(u) * p.S + ch + qd * 16) = acc;
                }
                __syncthreads();
            }
        }
        if (tid == 0) {
            atomicAdd(&p.stats[3], 1ull);
            if (ok) {
                if (p.fail) p.fail[cw] = 0;
                if (hdr[3] != 0u) atomicAdd(&p.stats[5], 1ull);   // a frame the peel kernel counted as an error
            } else {
                atomicAdd(&p.stats[4], 1ull);
            }
        }
        __syncthreads();
    }
}


// ------------------------------------------------------------------------------------------
// Fast path: one WARP per stalled codeword, direct elimination on the payload.
//
// The CTA-wide kernel above pays three block barriers per pivot and then a second pass that
// applies the m x m combination matrix to the syndromes.  Stopping sets are small next to H
// (n2040/k1530 at 13/64: ~255 unknowns touching ~295 of the 510 checks), so here
//   * only the checks that have a residual member become rows, and the right-hand side is the
//     payload itself: [A | b], b = 64 bytes of the check's syndrome.  Row operations act on A and b
//     together, so when A has become a permutation the pivot row of column j HOLDS unknown j --
//     there is no combination matrix and no second pass;
//   * lane L owns rows L, L+32, ...: it tests the pivot column in its rows, the pivot is the lowest
//     unused hit (redux.min), every lane then updates its own hit rows from the pivot row (all lanes
//     read the same pivot words: a broadcast).  Row pitches are odd (A) / 20 words (b), so the 32
//     rows touched by one instruction sit in 32 different banks.  One __syncwarp per pivot;
//   * symbols wider than 64 bytes are solved 64 bytes at a time (A is rebuilt, it is cheap).
// Several warps share a CTA, each with its own slot of shared memory; a codeword whose matrix does
// not fit the slot is appended to `list_out` for the next stage (bigger slots, finally the CTA kernel).
// ------------------------------------------------------------------------------------------
constexpr int kGeBPitch = 20;            // words per b row: 16 payload + 4 so that rows 0..7 cover all banks

struct GeWarpParams {
    GeParams g;
    unsigned int *list_out;              // codewords deferred to the next stage
    unsigned int *count_out;
    unsigned int *work_ctr;              // next unclaimed list position
    int slot_words;                      // shared memory per warp, 32-bit words
};

__host__ __device__ inline int ge_warp_fixed_words(int m, int NW)
{   // er[NW+1] pref[NW+1] rowmap[m] varlist[m] pivrow[m] (u16 each), rounded to 16 bytes
    return ((2 * (NW + 1) + 3 * ((m + 1) / 2)) + 3) & ~3;
}
__host__ __device__ inline long long ge_warp_matrix_words(int rows, int e, bool payload)
{
    return (long long)rows * ((((e + 31) / 32) | 1) + (payload ? kGeBPitch : 0));
}

__global__ void __launch_bounds__(512) hybrid_ge_warp_kernel(const GeWarpParams q)
{
    extern __shared__ __align__(16) uint32_t gw_smem[];
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const GeParams &p = q.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = p.m, NW = p.NW, RW = p.RW;
    uint32_t *er = gw_smem + size_t(warp) * q.slot_words;
    uint32_t *pref = er + NW + 1;
    uint16_t *rowmap = reinterpret_cast<uint16_t *>(pref + NW + 1);
    uint16_t *varlist = rowmap + m + (m & 1);
    uint16_t *pivrow = varlist + m + (m & 1);
    uint32_t *area = er + ge_warp_fixed_words(m, NW);
    const long long area_words = q.slot_words - ge_warp_fixed_words(m, NW);
    const bool payload = p.full != nullptr;
    const unsigned int count = *p.list_count;

    for (;;) {
        unsigned int li = 0;
        if (lane == 0) li = atomicAdd(q.work_ctr, 1u);
        li = __shfl_sync(FULL, li, 0);
        if (li >= count) break;
        const long long cw = p.list[li];

        // ---- 1. residual erased set, column numbering ------------------------------------------
        for (int w = lane; w < NW; w += 32) {
            uint32_t x = p.mask[cw * NW + w];
            if (w == NW - 1 && (p.n & 31)) x &= FULL >> (32 - (p.n & 31));
            er[w] = x;
        }
        __syncwarp();
        const uint32_t *hdr = reinterpret_cast<const uint32_t *>(p.sched + cw * (long long)p.stride);
        const int ne = int(hdr[0]);
        for (int i = lane; i < ne; i += 32) {
            const uint32_t v = hdr[4 + i] & 0xFFFFu;
            atomicAnd(&er[v >> 5], ~(1u << (v & 31)));
        }
        __syncwarp();
        int e = 0;
        for (int w0 = 0; w0 < NW; w0 += 32) {       // exclusive prefix of the popcounts
            const int w = w0 + lane;
            const int c = w < NW ? __popc(er[w]) : 0;
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, inc, o);
                if (lane >= o) inc += t;
            }
            if (w < NW) pref[w] = uint32_t(e + inc - c);
            e += __shfl_sync(FULL, inc, 31);
        }
        __syncwarp();
        bool ok = e <= m;                            // more unknowns than checks cannot have full column rank
        int R = 0;
        if (ok) {
            for (int w = lane; w < NW; w += 32) {
                uint32_t x = er[w];
                uint32_t b = pref[w];
                while (x) { varlist[b++] = uint16_t(w * 32 + __ffs(x) - 1); x &= x - 1u; }
            }
            // ---- 2. the checks with a residual member become the rows ----------------------------
            for (int r0 = 0; r0 < m; r0 += 32) {
                const int r = r0 + lane;
                bool any = false;
                if (r < m)
                    for (int j = 0; j < RW; j++) {
                        const uint32_t u = __ldg(p.cidx + r * RW + j);
                        if (u != 0xFFFFu && ((er[u >> 5] >> (u & 31)) & 1u)) { any = true; break; }
                    }
                const unsigned bal = __ballot_sync(FULL, any);
                if (any) rowmap[R + __popc(bal & ((1u << lane) - 1u))] = uint16_t(r);
                R += __popc(bal);
            }
            __syncwarp();
        }
        if (ok && ge_warp_matrix_words(R, e, payload) > area_words) {      // does not fit this stage's slot
            if (lane == 0) q.list_out[atomicAdd(q.count_out, 1u)] = (unsigned int)cw;
            __syncwarp();
            continue;
        }
        const int EW = (e + 31) / 32;
        const int pa = EW | 1;                       // odd pitch of an A row
        uint32_t *bmat = area;                       // [R][kGeBPitch], 16-byte aligned
        uint32_t *amat = area + (payload ? size_t(R) * kGeBPitch : 0);
        const int RPL = (R + 31) / 32;               // rows per lane (<= 64)

        const int nchunk = payload ? (p.S + 63) / 64 : 1;
        for (int chn = 0; ok && chn < nchunk; chn++) {
            const int ch = chn * 64;
            const int cb = payload ? min(64, p.S - ch) : 0;     // S is a multiple of 16
            // ---- 3. [A | b] ----------------------------------------------------------------------
            for (int i = 0; i < RPL; i++) {
                const int ri = lane + 32 * i;
                if (ri >= R) break;
                uint32_t *row = amat + size_t(ri) * pa;
                for (int w = 0; w < pa; w++) row[w] = 0u;
                const int r = rowmap[ri];
                for (int j = 0; j < RW; j++) {
                    const uint32_t u = __ldg(p.cidx + r * RW + j);
                    if (u == 0xFFFFu) continue;
                    const uint32_t x = er[u >> 5];
                    if ((x >> (u & 31)) & 1u) {
                        const uint32_t col = pref[u >> 5] + __popc(x & ((1u << (u & 31)) - 1u));
                        row[col >> 5] |= 1u << (col & 31);
                    }
                }
            }
            if (payload) {   // syndromes: XOR of the KNOWN members, 8 rows x 4 quarters per pass
                const uint8_t *full = p.full + size_t(cw) * p.n * p.S + ch;
                const int qd = lane & 3;
                for (int ri = lane >> 2; ri < R; ri += 8) {
                    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
                    if (qd * 16 < cb) {
                        const int r = rowmap[ri];
                        for (int j0 = 0; j0 < RW; j0 += 8) {        // RW is a multiple of 8; eight gathers in flight
                            const uint4 r4 = __ldg(reinterpret_cast<const uint4 *>(p.cidx + r * RW + j0));
                            const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
                            uint4 v[8];
                            bool use[8];
#pragma unroll
                            for (int t = 0; t < 8; t++) {
                                const uint32_t u = (t & 1) ? (rr[t >> 1] >> 16) : (rr[t >> 1] & 0xFFFFu);
                                use[t] = u != 0xFFFFu && !((er[(u == 0xFFFFu ? 0u : u) >> 5] >> (u & 31)) & 1u);
                                v[t] = *reinterpret_cast<const uint4 *>(full + size_t(use[t] ? u : 0u) * p.S + qd * 16);
                            }
#pragma unroll
                            for (int t = 0; t < 8; t++)
                                if (use[t]) { acc.x ^= v[t].x; acc.y ^= v[t].y; acc.z ^= v[t].z; acc.w ^= v[t].w; }
                        }
                    }
                    *reinterpret_cast<uint4 *>(bmat + size_t(ri) * kGeBPitch + qd * 4) = acc;
                }
            }
            __syncwarp();
            // ---- 4. Gauss-Jordan, lane L owns rows L, L+32, ... -----------------------------------
            // (loads are issued in independent batches and the pivot row is held in registers: the
            //  compiler cannot reorder shared-memory loads across the stores of a read-modify-write loop)
            unsigned long long used = 0ull;
            const int nq = cb / 16;
            for (int col = 0; col < e; col++) {
                const int wj = col >> 5;
                const uint32_t bj = 1u << (col & 31);
                unsigned long long hits = 0ull;
                for (int i0 = 0; i0 < RPL; i0 += 8) {
                    uint32_t wv[8];
#pragma unroll
                    for (int t = 0; t < 8; t++) wv[t] = amat[size_t(min(lane + 32 * (i0 + t), R - 1)) * pa + wj];
#pragma unroll
                    for (int t = 0; t < 8; t++)
                        if (lane + 32 * (i0 + t) < R && (wv[t] & bj)) hits |= 1ull << (i0 + t);
                }
                const unsigned long long cand = hits & ~used;
                const unsigned mine = cand ? unsigned((__ffsll((long long)cand) - 1) * 32 + lane) : 0xFFFFFFFFu;
                const unsigned piv = __reduce_min_sync(FULL, mine);
                if (piv == 0xFFFFFFFFu) { ok = false; break; }   // no pivot: rank deficient (HybridML.m:59-62)
                if (int(piv & 31u) == lane) {
                    used |= 1ull << (piv >> 5);
                    hits &= ~(1ull << (piv >> 5));
                    pivrow[col] = uint16_t(piv);
                }
                const uint32_t *ap = amat + size_t(piv) * pa;
                uint4 pb[4];
#pragma unroll
                for (int t = 0; t < 4; t++)
                    pb[t] = t < nq ? reinterpret_cast<const uint4 *>(bmat + size_t(piv) * kGeBPitch)[t] : make_uint4(0u, 0u, 0u, 0u);
                while (hits) {
                    const int ri = lane + 32 * (__ffsll((long long)hits) - 1);
                    hits &= hits - 1ull;
                    uint32_t *ar = amat + size_t(ri) * pa;
                    uint4 *br = reinterpret_cast<uint4 *>(bmat + size_t(ri) * kGeBPitch);
                    uint4 rb[4];
#pragma unroll
                    for (int t = 0; t < 4; t++) if (t < nq) rb[t] = br[t];
                    for (int w0 = wj; w0 < EW; w0 += 4) {       // (pitch pa >= EW; words past EW are never read back)
                        uint32_t x[4], y[4];
#pragma unroll
                        for (int t = 0; t < 4; t++) { const int w = min(w0 + t, EW - 1); x[t] = ar[w]; y[t] = ap[w]; }
#pragma unroll
                        for (int t = 0; t < 4; t++) if (w0 + t < EW) ar[w0 + t] = x[t] ^ y[t];
                    }
#pragma unroll
                    for (int t = 0; t < 4; t++)
                        if (t < nq) {
                            rb[t].x ^= pb[t].x; rb[t].y ^= pb[t].y; rb[t].z ^= pb[t].z; rb[t].w ^= pb[t].w;
                            br[t] = rb[t];
                        }
                }
                __syncwarp();
            }
            // ---- 5. the pivot row of column j now holds unknown j ----------------------------------
            if (ok && payload) {
                uint8_t *out = p.out + size_t(cw) * p.k * p.S + ch;
                const int qd = lane & 3;
                for (int col = lane >> 2; col < e; col += 8) {
                    const int u = varlist[col];
                    if (u >= p.k || qd * 16 >= cb) continue;       // only systematic symbols are output
                    *reinterpret_cast<uint4 *>(out + size_t(u) * p.S + qd * 16) =
                        *reinterpret_cast<const uint4 *>(bmat + size_t(pivrow[col]) * kGeBPitch + qd * 4);
                }
            }
            __syncwarp();
        }
        if (lane == 0) {
            atomicAdd(&p.stats[3], 1ull);
            if (ok) {
                if (p.fail) p.fail[cw] = 0;
                if (hdr[3] != 0u) atomicAdd(&p.stats[5], 1ull);   // a frame the peel kernel counted as an error
            } else {
                atomicAdd(&p.stats[4], 1ull);
            }
        }
        __syncwarp();
    }
}

// ---- host side -------------------------------------------------------------------------------
struct HybridScratch {
    uint8_t *d_full = nullptr;          // [max_batch][n][S]
    unsigned int *d_list = nullptr;     // [max_batch] stalled codewords, written by the peel kernel
    unsigned int *d_list2 = nullptr, *d_list3 = nullptr;   // deferred by warp stage 1 / stage 2
    unsigned int *d_count = nullptr;    // [8]: list counts [0..2], work counters [3..4]
    int wpc[2] = {0, 0}, slot_words[2] = {0, 0};            // the two warp stages: warps per CTA, words per slot
    uint32_t *d_gmat = nullptr;         // per-CTA matrices when they do not fit in shared memory
    int grid = 0, smem = 0, RSW = 0;
    bool ready = false;
};

inline void hybrid_free(HybridScratch &h)
{
    cudaFree(h.d_full); cudaFree(h.d_list); cudaFree(h.d_list2); cudaFree(h.d_list3); cudaFree(h.d_count); cudaFree(h.d_gmat);
    h = HybridScratch();
}

// Allocates the stage's scratch on first use (the full-codeword buffer is large: max_batch * n * S).
inline int hybrid_prepare(HybridScratch &h, const HostCode &code, int S, int NW, int MW, int num_sms, int smem_optin,
                          long long max_batch, std::string &err)
{
    if (h.ready) return LDPC_OK;
    auto bad = [&](const char *what, cudaError_t e) {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? LDPC_ERR_NOMEM : LDPC_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaMalloc(&h.d_full, size_t(max_batch) * code.n * S)) != cudaSuccess) return bad("hybrid: full-codeword scratch", e);
    if ((e = cudaMalloc(&h.d_list, size_t(max_batch) * 4)) != cudaSuccess) return bad("hybrid: list", e);
    if ((e = cudaMalloc(&h.d_list2, size_t(max_batch) * 4)) != cudaSuccess) return bad("hybrid: list", e);
    if ((e = cudaMalloc(&h.d_list3, size_t(max_batch) * 4)) != cudaSuccess) return bad("hybrid: list", e);
    if ((e = cudaMalloc(&h.d_count, 8 * 4)) != cudaSuccess) return bad("hybrid: count", e);
    {   // warp stages: stage 1 sized for a typical stopping set (3/4 of the checks involved), stage 2 for the worst case
        const long long budget_w = (long long)(smem_optin - 2048) / 4;
        const int fixed = ge_warp_fixed_words(code.m, NW);
        const int typ = 3 * code.m / 4;
        long long s1 = fixed + ge_warp_matrix_words(typ, typ, true);
        long long s2 = fixed + ge_warp_matrix_words(code.m, code.m, true);
        int w1 = int(std::max<long long>(1, std::min<long long>(16, budget_w / s1)));
        int w2 = int(std::max<long long>(1, std::min<long long>(16, budget_w / s2)));
        if (const char *ev = getenv("LDPC_CUDA_GE_WPC")) w1 = std::max(1, std::min(16, atoi(ev)));
        w2 = std::min(w2, w1);
        h.wpc[0] = w1; h.slot_words[0] = int((budget_w / w1) & ~3ll);
        h.wpc[1] = w2; h.slot_words[1] = int((budget_w / w2) & ~3ll);
        if (h.slot_words[0] <= fixed + 64 || h.slot_words[1] <= fixed + 64) { h.wpc[0] = h.wpc[1] = 0; }   // CTA kernel only
        if ((e = cudaFuncSetAttribute(hybrid_ge_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 1024)) != cudaSuccess)
            return bad("hybrid: cudaFuncSetAttribute", e);
    }
    h.RSW = 2 * MW + 1;
    const size_t base = size_t(code.m) * 64 + ge_small_bytes(code.m, NW) + 16;
    const size_t mat = size_t(code.m) * h.RSW * 4;
    const size_t budget = size_t(smem_optin) - 1024;
    if (base > budget) { err = "hybrid: code too large for the elimination kernel"; return LDPC_ERR_UNSUPPORTED; }
    if (base + mat <= budget) {
        h.smem = int(base + mat);
        h.grid = num_sms * int(std::max<size_t>(1, std::min<size_t>(2, budget / (base + mat))));
    } else {
        h.smem = int(base);
        h.grid = num_sms;
        if ((e = cudaMalloc(&h.d_gmat, size_t(h.grid) * mat)) != cudaSuccess) return bad("hybrid: matrix workspace", e);
    }
    if ((e = cudaFuncSetAttribute(hybrid_ge_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(budget))) != cudaSuccess)
        return bad("hybrid: cudaFuncSetAttribute", e);
    if ((e = cudaFuncSetAttribute(hybrid_ge_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(budget))) != cudaSuccess)
        return bad("hybrid: cudaFuncSetAttribute", e);
    h.ready = true;
    return LDPC_OK;
}

}  // namespace ldpc
