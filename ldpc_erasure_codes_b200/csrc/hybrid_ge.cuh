// hybrid_ge.cuh -- hybrid-ML stage: GF(2) elimination on the residual stopping set.
//
// Reference: Matlab/My_LDPC_HybridML_Erasure_Decoder.m:48-87.  After the capped peeling sweeps the
// MATLAB code solves  H(:,E) x = H(:,known) y(known)  for the still-erased set E by Gaussian
// elimination with row swaps (:57-75) and Jordan back-elimination (:77-86); it aborts when a column
// has no pivot (:59-62), i.e. exactly when rank(H(:,E)) < |E|.  On success the solution is unique,
// so ANY exact solver returns the same bytes; on abort the contract (SURVEY a-10) is "leave the
// peeling result, report ml_fail".  Three solvers run back to back, each taking what the previous one
// deferred: inactivation decoding per warp (hybrid_inact_kernel + hybrid_apply_kernel, below),
// plain Gauss-Jordan per warp (hybrid_ge_warp_kernel), and the CTA-wide kernel described first:
//
//   one CTA per codeword that left the peeling stage with erasures (compacted list from the peel
//   kernel).  The m x e bit matrix A = H(:,E) is built bit-packed in shared memory next to an
//   m x m identity, [A | I], one row per check.  Gauss-Jordan: per column, the lowest unused row
//   with a 1 is the pivot (atomicMin over the CTA), the pivot row is broadcast through shared memory
//   and XORed -- 32 words per warp instruction -- into every other row that has the bit.  No pivot
//   => rank deficient => ml_fail.  At the end row pivot(j) of the I part lists which check
//   syndromes add up to unknown j.  Payload: the syndromes  s_r = XOR of the KNOWN members of check r
//   (received or peeled; the executor formed them while the codeword was in its shared memory) are
//   read 64 bytes of every symbol at a time, and every erased SYSTEMATIC symbol is written as the
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
    const uint8_t *synd;             // [B][m][S] per check: XOR of the members known after peeling (from the executor;
                                     // valid for listed codewords), or nullptr = error-rate run, pattern only
    uint8_t *out;                    // [B][k][S]
    uint8_t *fail;                   // [B]
    uint8_t *fail_any;               // [B] or nullptr: cleared when the elimination succeeds (all n symbols are then known)
    unsigned long long *stats;       // [3] ml_attempts, [4] ml_failures, [5] ml_recovered
    const uint16_t *cidx;            // [m][RW]
    const uint16_t *vadj;            // [n][VW] variable -> checks
    unsigned long long *phase_cycles;   // optional [8]: inactivation stage phase timers
    int VW;
    uint32_t *gmat;                  // per-CTA global workspace for [A | I], or nullptr (shared memory)
    int n, k, m, RW, NW, MW, S, stride;
    int RSW;                         // words per matrix row: 2 * MW + 1 (odd: rows start in different banks)
};

__host__ __device__ inline size_t ge_small_bytes(int m, int NW)
{   // er[NW] pref[NW+1] prow[RSW<=257] varlist[m] pivrow[m] used[m] inv[m] + scalars, generously rounded
    return size_t(NW) * 4 + size_t(NW + 1) * 4 + 260 * 4 + size_t(m) * 2 * 2 + size_t(m) * 2 + 64 + 64;
}


// 16 bytes (quarter qd of the 64-byte chunk at ch) of check r's right-hand side
__device__ __forceinline__ uint4 ge_rhs_quarter(const GeParams &p, long long cw, int r, int ch, int qd)
{
    return *reinterpret_cast<const uint4 *>(p.synd + (size_t(cw) * p.m + r) * p.S + ch + qd * 16);
}

// A warp loads the right-hand sides of its R rows: 8 rows x 4 quarters per instruction, four instructions in flight.
__device__ __forceinline__ void ge_load_rhs(const GeParams &p, long long cw, const uint16_t *rowmap, int R, int ch, int nq,
                                            uint32_t *bmat, int pitch, int lane)
{
    const int qd = lane & 3, hs = lane >> 2;
    for (int r0 = hs; r0 < R; r0 += 32) {
        uint4 v[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int ri = r0 + 8 * t;
            v[t] = (ri < R && qd < nq) ? ge_rhs_quarter(p, cw, rowmap[ri], ch, qd) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int ri = r0 + 8 * t;
            if (ri < R) *reinterpret_cast<uint4 *>(bmat + size_t(ri) * pitch + qd * 4) = v[t];
        }
    }
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
        if (ok && p.synd) {   // (synd == nullptr: error-rate run, pattern only)
            uint8_t *out = p.out + size_t(cw) * p.k * p.S;
            const int qd = tid & 3;             // 16-byte quarter of a 64-byte chunk
            for (int ch = 0; ch < p.S; ch += 64) {
                const int cb = min(64, p.S - ch);   // S is a multiple of 16
                for (int r = tid >> 2; r < m; r += kGeThreads / 4) {
                    if (!inv[r] || qd * 16 >= cb) continue;
                    *reinterpret_cast<uint4 *>(synd + size_t(r) * 64 + qd * 16) = ge_rhs_quarter(p, cw, r, ch, qd);
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
                if (p.fail_any) p.fail_any[cw] = 0;
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
    // inactivation stage, split form: the pattern kernel records a plan per solvable codeword, the apply
    // kernel replays it on the payload
    uint32_t *plan;                      // [slots][plan_words] or nullptr (solve in place / pattern only)
    unsigned int *plan_count;            // slots written
    int plan_words;                      // stride, 32-bit words
    int apply_slot_words;                // the apply kernel's shared memory per warp: what must fit to be planned
};

// plan of one codeword (32-bit words): header {cw, e, R, npeel, ninact, rank deficient, 0, 0}, then at fixed offsets
// er[NW] | rowmap[m] varlist[m] pl_col[m] pl_row[m] (u16) | icol[64] (u16) | ipart[m] (u64) | nbr[m*VW] (u16)
struct GePlanLayout {
    int er, rowmap, varlist, pl_col, pl_row, icol, ipart, nbr, words;
};
__host__ __device__ inline GePlanLayout ge_plan_layout(int m, int NW, int VW)
{
    GePlanLayout L;
    const int h = (m + 1) / 2;
    L.er = 8; L.rowmap = L.er + NW; L.varlist = L.rowmap + h; L.pl_col = L.varlist + h; L.pl_row = L.pl_col + h;
    L.icol = L.pl_row + h; L.ipart = (L.icol + 32 + 1) & ~1; L.nbr = L.ipart + 2 * m;
    L.words = (L.nbr + (m * VW + 1) / 2 + 3) & ~3;
    return L;
}

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
    const bool payload = p.synd != nullptr;
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
            if (payload) ge_load_rhs(p, cw, rowmap, R, ch, cb / 16, bmat, kGeBPitch, lane);   // the executor's syndromes
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
                if (p.fail_any) p.fail_any[cw] = 0;
                if (hdr[3] != 0u) atomicAdd(&p.stats[5], 1ull);   // a frame the peel kernel counted as an error
            } else {
                atomicAdd(&p.stats[4], 1ull);
            }
        }
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------
// First stage: inactivation decoding, one warp per stalled codeword.
//
// A stopping set is sparse: it stalls peeling only because no check has exactly ONE unknown left.
// Declaring a few unknowns "inactive" (to be solved later) restarts peeling; on n2040/k1530 at 13/64
// about 6 inactivations let peeling finish the other ~250 unknowns.  This is Gauss-Jordan elimination
// with a good pivot order, so its verdict (full column rank or not) and its solution are those of the
// reference's elimination -- but a peeled pivot touches only the <= VW-1 other checks of its variable
// instead of ~35 filled-in rows, and the dense part is |inactive| x |inactive|.
//
//   per involved check (row): state = [used:1 | deg:15 | xor of its ACTIVE unknowns:16], ipart = which
//   inactive unknowns its equation contains (<= 64), b = 64 bytes of right-hand side.
//   phase 1  while unknowns are active: pop a row with deg 1 -> its unknown u is "peeled" with this row as
//            pivot; every other check of u (static adjacency, H's column) gets  row ^= pivot row  (ipart, b)
//            and deg-1.  No such row: take the unused row of smallest degree and inactivate one of its
//            unknowns (bit q of ipart in all its checks, deg-1).
//   phase 2  the unused rows are now equations in the inactive unknowns only: Gauss-Jordan on
//            [ipart | b] (<= 64 columns); no pivot => rank deficient => ml_fail.
//   phase 3  unknown of pivot row r  =  b[r] ^ XOR of the inactive unknowns in ipart[r]; systematic
//            ones are written to the decoder output.  (All of phase 3 is independent work.)
// More than 64 inactivations, or rows that do not fit the slot: deferred to the next stage.
// ------------------------------------------------------------------------------------------
constexpr int kInactMax = 64;

__host__ __device__ inline int ge_inact_fixed_words(int m, int NW, int MW)
{   // er[NW+1] pref[NW+1] d1[MW] act[MW] icol/ipiv[64] | u16: inv_rowmap rowmap pl_col pl_row varlist, [m] each
    return ((2 * (NW + 1) + 2 * MW + kInactMax) + 5 * ((m + 1) / 2) + 3) & ~3;
}
__host__ __device__ inline long long ge_inact_area_words(int rows, int e, int VW, bool payload)
{   // state[R] | ipart[R] u64 | nbr[e][VW] u16 | b[R][16]
    return ((3ll * rows + 3) & ~3ll) + ((((long long)e * VW + 1) / 2 + 3) & ~3ll) + (payload ? 16ll * rows : 0);
}

__global__ void __launch_bounds__(512) hybrid_inact_kernel(const GeWarpParams q)
{
    extern __shared__ __align__(16) uint32_t gw_smem[];
    constexpr unsigned FULL = 0xFFFFFFFFu;
    constexpr uint32_t USED = 0x80000000u, TAKEN = 0x40000000u;
    const GeParams &p = q.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = p.m, NW = p.NW, MW = p.MW, RW = p.RW, VW = p.VW;
    uint32_t *er = gw_smem + size_t(warp) * q.slot_words;     // residual erased set (the unknowns)
    uint32_t *pref = er + NW + 1;                               // column number of the first unknown of each mask word
    uint32_t *d1 = pref + NW + 1;                               // rows whose degree became 1
    uint32_t *act = d1 + MW;                                    // columns still active
    uint16_t *icol = reinterpret_cast<uint16_t *>(act + MW);    // [64] inactive columns, [64] their pivot rows
    uint16_t *ipiv = icol + kInactMax;
    const int mp = m + (m & 1);
    uint16_t *inv_rowmap = ipiv + kInactMax;
    uint16_t *rowmap = inv_rowmap + mp;
    uint16_t *pl_col = rowmap + mp;
    uint16_t *pl_row = pl_col + mp;
    uint16_t *varlist = pl_row + mp;
    uint32_t *area = er + ge_inact_fixed_words(m, NW, MW);
    const long long area_words = q.slot_words - ge_inact_fixed_words(m, NW, MW);
    const bool planning = q.plan != nullptr;                    // record the solution order, leave the payload to the apply kernel
    const bool payload = p.synd != nullptr && !planning;
    const unsigned int count = *p.list_count;
    const int qd = lane & 3, hs = lane >> 2;                    // payload lanes: 16-byte quarter, hit slot

    for (;;) {
        unsigned int li = 0;
        if (lane == 0) li = atomicAdd(q.work_ctr, 1u);
        li = __shfl_sync(FULL, li, 0);
        if (li >= count) break;
        const long long cw = p.list[li];
        const uint32_t *hdr = reinterpret_cast<const uint32_t *>(p.sched + cw * (long long)p.stride);
        const int ne = int(hdr[0]);
        bool ok = true, defer = false;
        long long tmark = p.phase_cycles ? clock64() : 0;
        auto lap = [&](int ph) {
            if (p.phase_cycles) {
                const long long t = clock64();
                if (lane == 0) atomicAdd(&p.phase_cycles[ph], (unsigned long long)(t - tmark));
                tmark = t;
            }
        };
        const int nchunk = payload ? (p.S + 63) / 64 : 1;
        for (int chn = 0; ok && !defer && chn < nchunk; chn++) {
            const int ch = chn * 64;
            const int cb = payload ? min(64, p.S - ch) : 0;
            // ---- residual erased set, column numbering -----------------------------------------------
            for (int w = lane; w < NW; w += 32) {
                uint32_t x = p.mask[cw * NW + w];
                if (w == NW - 1 && (p.n & 31)) x &= FULL >> (32 - (p.n & 31));
                er[w] = x;
            }
            for (int w = lane; w < MW; w += 32) d1[w] = 0u;
            __syncwarp();
            for (int i = lane; i < ne; i += 32) {
                const uint32_t v = hdr[4 + i] & 0xFFFFu;
                atomicAnd(&er[v >> 5], ~(1u << (v & 31)));
            }
            __syncwarp();
            int e = 0;
            for (int w0 = 0; w0 < NW; w0 += 32) {
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
            if (e > m) { ok = false; break; }                    // more unknowns than checks: no full column rank
            for (int w = lane; w < NW; w += 32) {
                uint32_t x = er[w];
                uint32_t b = pref[w];
                while (x) { varlist[b++] = uint16_t(w * 32 + __ffs(x) - 1); x &= x - 1u; }
            }
            lap(0);
            for (int w = lane; w < MW; w += 32) act[w] = (w * 32 + 32 <= e) ? FULL : (w * 32 < e ? (FULL >> (32 - (e - w * 32))) : 0u);
            // ---- rows: the checks with a residual member; state = [deg | xor of the active columns] ------
            int R = 0;
            for (int r0 = 0; r0 < m; r0 += 32) {
                const int r = r0 + lane;
                uint32_t cnt = 0, xr = 0;
                if (r < m)
                    for (int j0 = 0; j0 < RW; j0 += 8) {
                        const uint4 r4 = __ldg(reinterpret_cast<const uint4 *>(p.cidx + r * RW + j0));
                        const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                        for (int t = 0; t < 8; t++) {
                            const uint32_t u = (t & 1) ? (rr[t >> 1] >> 16) : (rr[t >> 1] & 0xFFFFu);
                            const uint32_t us = u == 0xFFFFu ? 0u : u;
                            const uint32_t x = er[us >> 5];
                            const uint32_t bit = u == 0xFFFFu ? 0u : (x >> (us & 31)) & 1u;
                            const uint32_t col = pref[us >> 5] + __popc(x & ((1u << (us & 31)) - 1u));
                            cnt += bit;
                            xr ^= (0u - bit) & col;
                        }
                    }
                const unsigned bal = __ballot_sync(FULL, cnt > 0);
                const int ri = R + __popc(bal & ((1u << lane) - 1u));
                if (r < m) inv_rowmap[r] = cnt > 0 ? uint16_t(ri) : uint16_t(0xFFFFu);
                if (cnt > 0) rowmap[ri] = uint16_t(r);
                R += __popc(bal);
                if (cnt > 0 && ri < area_words) area[ri] = (cnt << 16) | xr;     // state (the fit is checked below)
            }
            __syncwarp();
            lap(1);
            if (ge_inact_area_words(R, e, VW, payload) > area_words) { defer = true; break; }
            if (planning && ge_inact_area_words(R, e, VW, true) > q.apply_slot_words - ge_inact_fixed_words(m, NW, MW)) { defer = true; break; }
            uint32_t *state = area;                                              // [R]
            unsigned long long *ipart = reinterpret_cast<unsigned long long *>(area + ((R + 1) & ~1));   // [R]
            uint16_t *nbr = reinterpret_cast<uint16_t *>(area + ((3 * R + 3) & ~3));   // [e][VW] rows of each unknown
            uint32_t *bmat = area + ((3 * R + 3) & ~3) + (((e * VW + 1) / 2 + 3) & ~3);   // [R][16], 16-byte aligned
            for (int ri = lane; ri < R; ri += 32) {
                ipart[ri] = 0ull;
                if ((state[ri] >> 16) == 1u) atomicOr(&d1[ri >> 5], 1u << (ri & 31));
            }
            for (int i = lane; i < e * VW; i += 32) {            // H's columns for the unknowns, as compact row numbers
                const uint32_t chk = __ldg(p.vadj + size_t(varlist[i / VW]) * VW + (i % VW));
                nbr[i] = chk == 0xFFFFu ? uint16_t(0xFFFFu) : inv_rowmap[chk];
            }
            if (payload) ge_load_rhs(p, cw, rowmap, R, ch, cb / 16, bmat, 16, lane);          // the executor's syndromes
            __syncwarp();

            lap(2);
            // ---- phase 1: peel, inactivating when stuck --------------------------------------------------
            const int MWc = (R + 31) >> 5;
            const int RPL = (R + 31) >> 5;
            int remaining = e, npeel = 0, ninact = 0;
            while (remaining > 0) {
                int ri = -1;                                     // lowest row marked degree-1
                for (int w0 = 0; w0 < MWc && ri < 0; w0 += 32) {
                    const uint32_t x = (w0 + lane < MWc) ? d1[w0 + lane] : 0u;
                    const unsigned bal = __ballot_sync(FULL, x != 0u);
                    if (bal) {
                        const int fl = __ffs(bal) - 1;
                        const uint32_t xf = __shfl_sync(FULL, x, fl);
                        ri = (w0 + fl) * 32 + __ffs(xf) - 1;
                        if (lane == fl) d1[w0 + fl] = x & (x - 1u);
                    }
                }
                uint32_t col;
                const bool peel = ri >= 0;
                if (peel) {
                    const uint32_t st = state[ri];
                    if ((st >> 16) != 1u) { __syncwarp(); continue; }     // stale mark (degree dropped to 0)
                    col = st & 0xFFFFu;
                } else {
                    // stuck: unused row of smallest degree >= 2; inactivate its first active unknown
                    unsigned best = 0xFFFFFFFFu;
                    for (int i0 = 0; i0 < RPL; i0 += 8) {
                        uint32_t sv[8];
#pragma unroll
                        for (int t = 0; t < 8; t++) sv[t] = state[min(lane + 32 * (i0 + t), R - 1)];
#pragma unroll
                        for (int t = 0; t < 8; t++) {
                            const int rr_ = lane + 32 * (i0 + t);
                            const uint32_t dg = sv[t] >> 16;          // (used rows carry the flag in this field)
                            if (rr_ < R && dg >= 2u && dg < 0x4000u) best = min(best, (dg << 16) | uint32_t(rr_));
                        }
                    }
                    best = __reduce_min_sync(FULL, best);
                    if (best == 0xFFFFFFFFu) { ok = false; break; }       // an unknown that no equation constrains
                    if (ninact == kInactMax) { defer = true; break; }
                    ri = int(best & 0xFFFFu);
                    const int r = rowmap[ri];
                    uint32_t mc = 0xFFFFu;
                    bool a = false;
                    if (lane < RW) {
                        const uint32_t mu = __ldg(p.cidx + r * RW + lane);
                        if (mu != 0xFFFFu) {
                            const uint32_t x = er[mu >> 5];
                            if ((x >> (mu & 31)) & 1u) {
                                mc = pref[mu >> 5] + __popc(x & ((1u << (mu & 31)) - 1u));
                                a = (act[mc >> 5] >> (mc & 31)) & 1u;
                            }
                        }
                    }
                    const unsigned bal = __ballot_sync(FULL, a);          // (!= 0: the row's degree counts active members)
                    col = __shfl_sync(FULL, mc, __ffs(bal) - 1);
                }
                // the rows of this unknown (static: H's column)
                uint32_t rj = 0xFFFFu;
                if (lane < VW) rj = nbr[col * VW + lane];
                const bool nb = rj != 0xFFFFu;
                const bool other = nb && !(peel && int(rj) == ri);
                const unsigned long long pI = peel ? ipart[ri] : (1ull << ninact);
                uint4 pb = make_uint4(0u, 0u, 0u, 0u);
                if (payload && peel) pb = *reinterpret_cast<const uint4 *>(bmat + size_t(ri) * 16 + qd * 4);
                if (other) {
                    const uint32_t s2 = state[rj];
                    const uint32_t n2 = (((s2 >> 16) - 1u) << 16) | ((s2 & 0xFFFFu) ^ col);
                    state[rj] = n2;
                    ipart[rj] ^= pI;                                     // (inactivation: a new bit, so ^ sets it)
                    if ((n2 >> 16) == 1u) atomicOr(&d1[rj >> 5], 1u << (rj & 31));
                }
                if (peel && nb && int(rj) == ri) state[ri] = USED;
                if (lane == 0) {
                    act[col >> 5] &= ~(1u << (col & 31));
                    if (peel) { pl_col[npeel] = uint16_t(col); pl_row[npeel] = uint16_t(ri); }
                    else icol[ninact] = uint16_t(col);
                }
                if (payload && peel) {                                   // b[row] ^= b[pivot], eight rows per instruction
                    const unsigned hb = __ballot_sync(FULL, other);
                    const int nh = __popc(hb);
                    for (int h0 = 0; h0 < nh; h0 += 8) {
                        const int src = __fns(hb, 0, h0 + hs + 1);       // lane holding the (h0+hs)-th hit, -1 if none
                        const int r2 = __shfl_sync(FULL, int(rj), src < 0 ? 0 : src);
                        if (src >= 0 && src < 32) {
                            uint4 *br = reinterpret_cast<uint4 *>(bmat + size_t(r2) * 16 + qd * 4);
                            uint4 a4 = *br;
                            a4.x ^= pb.x; a4.y ^= pb.y; a4.z ^= pb.z; a4.w ^= pb.w;
                            *br = a4;
                        }
                    }
                }
                if (peel) npeel++; else ninact++;
                remaining--;
                __syncwarp();
            }
            lap(3);
            if (!ok || defer) break;
            unsigned int plan_slot = 0;
            if (planning) {   // snapshot for the apply kernel: ipart as it is BEFORE the dense solve
                if (lane == 0) plan_slot = atomicAdd(q.plan_count, 1u);
                plan_slot = __shfl_sync(FULL, plan_slot, 0);
                const GePlanLayout L = ge_plan_layout(m, NW, VW);
                uint32_t *pw = q.plan + size_t(plan_slot) * q.plan_words;
                if (lane == 0) { pw[0] = uint32_t(cw); pw[1] = uint32_t(e); pw[2] = uint32_t(R); pw[3] = uint32_t(npeel); pw[4] = uint32_t(ninact); pw[5] = 0u; }
                for (int i = lane; i < NW; i += 32) pw[L.er + i] = er[i];
                const uint32_t *s_rowmap = reinterpret_cast<const uint32_t *>(rowmap), *s_var = reinterpret_cast<const uint32_t *>(varlist);
                const uint32_t *s_pc = reinterpret_cast<const uint32_t *>(pl_col), *s_pr = reinterpret_cast<const uint32_t *>(pl_row);
                for (int i = lane; i < (R + 1) / 2; i += 32) pw[L.rowmap + i] = s_rowmap[i];
                for (int i = lane; i < (e + 1) / 2; i += 32) pw[L.varlist + i] = s_var[i];
                for (int i = lane; i < (npeel + 1) / 2; i += 32) { pw[L.pl_col + i] = s_pc[i]; pw[L.pl_row + i] = s_pr[i]; }
                pw[L.icol + lane] = reinterpret_cast<const uint32_t *>(icol)[lane];
                const uint32_t *s_ip = reinterpret_cast<const uint32_t *>(ipart);
                for (int i = lane; i < 2 * R; i += 32) pw[L.ipart + i] = s_ip[i];
                const uint32_t *s_nb = reinterpret_cast<const uint32_t *>(nbr);
                for (int i = lane; i < (e * VW + 1) / 2; i += 32) pw[L.nbr + i] = s_nb[i];
            }

            // ---- phase 2: the unused rows are equations in the inactive unknowns only ----------------------
            for (int c = 0; c < ninact; c++) {
                unsigned best = 0xFFFFFFFFu;
                for (int r0 = lane; r0 < R; r0 += 32)
                    if (!(state[r0] & (USED | TAKEN)) && ((ipart[r0] >> c) & 1u)) { best = uint32_t(r0); break; }
                best = __reduce_min_sync(FULL, best);
                if (best == 0xFFFFFFFFu) { ok = false; break; }           // no pivot: rank deficient (HybridML.m:59-62)
                const int pr = int(best);
                const unsigned long long pI = ipart[pr];
                uint32_t pb = 0u;
                if (payload && lane < 16) pb = bmat[size_t(pr) * 16 + lane];
                __syncwarp();
                if (lane == 0) { state[pr] |= TAKEN; ipiv[c] = uint16_t(pr); }
                for (int r0 = 0; r0 < R; r0 += 32) {
                    const int r2 = r0 + lane;
                    const bool hit = r2 < R && r2 != pr && !(state[r2] & USED) && ((ipart[r2] >> c) & 1u);
                    if (hit) ipart[r2] ^= pI;
                    if (payload) {
                        unsigned hb = __ballot_sync(FULL, hit);
                        while (hb) {
                            const int hl = __ffs(hb) - 1;
                            hb &= hb - 1u;
                            if (lane < 16) bmat[size_t(r0 + hl) * 16 + lane] ^= pb;
                        }
                    }
                }
                __syncwarp();
            }
            lap(4);
            if (planning && !ok && lane == 0) q.plan[size_t(plan_slot) * q.plan_words + 5] = 1u;   // rank deficient: nothing to apply
            if (!ok) break;

            // ---- phase 3: read the unknowns off their pivot rows ----------------------------------------------
            if (payload) {
                uint8_t *out = p.out + size_t(cw) * p.k * p.S + ch;
                for (int i = hs; i < npeel + ninact; i += 8) {
                    const bool pe = i < npeel;
                    const int u = varlist[pe ? pl_col[i] : icol[i - npeel]];
                    const int r = pe ? pl_row[i] : ipiv[i - npeel];
                    if (u >= p.k || qd * 16 >= cb) continue;              // only systematic symbols are output
                    uint4 acc = *reinterpret_cast<const uint4 *>(bmat + size_t(r) * 16 + qd * 4);
                    unsigned long long bits = pe ? ipart[r] : 0ull;
                    while (bits) {
                        const int c = __ffsll((long long)bits) - 1;
                        bits &= bits - 1ull;
                        const uint4 v = *reinterpret_cast<const uint4 *>(bmat + size_t(ipiv[c]) * 16 + qd * 4);
                        acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
                    }
                    *reinterpret_cast<uint4 *>(out + size_t(u) * p.S + qd * 16) = acc;
                }
            }
            __syncwarp();
            lap(5);
        }
        if (p.phase_cycles && lane == 0) atomicAdd(&p.phase_cycles[6], 1ull);
        if (defer) {
            if (lane == 0) q.list_out[atomicAdd(q.count_out, 1u)] = (unsigned int)cw;
            __syncwarp();
            continue;
        }
        if (lane == 0) {
            atomicAdd(&p.stats[3], 1ull);
            if (ok) {
                if (p.fail) p.fail[cw] = 0;
                if (p.fail_any) p.fail_any[cw] = 0;
                if (hdr[3] != 0u) atomicAdd(&p.stats[5], 1ull);   // a frame the peel kernel counted as an error
            } else {
                atomicAdd(&p.stats[4], 1ull);
            }
        }
        __syncwarp();
    }
}


// Payload part of the inactivation stage: replays a recorded plan.  One warp per codeword; the pattern
// kernel above (which needs 3 words per row instead of 19) found the pivots at full occupancy.
__global__ void __launch_bounds__(512) hybrid_apply_kernel(const GeWarpParams q)
{
    extern __shared__ __align__(16) uint32_t gw_smem[];
    constexpr unsigned FULL = 0xFFFFFFFFu;
    constexpr uint32_t USED = 0x80000000u, TAKEN = 0x40000000u;
    const GeParams &p = q.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = p.m, NW = p.NW, MW = p.MW, VW = p.VW;
    uint32_t *er = gw_smem + size_t(warp) * q.slot_words;     // same fixed layout as the pattern kernel
    uint32_t *pref = er + NW + 1;
    uint32_t *d1 = pref + NW + 1;
    uint32_t *act = d1 + MW;
    uint16_t *icol = reinterpret_cast<uint16_t *>(act + MW);
    uint16_t *ipiv = icol + kInactMax;
    const int mp = m + (m & 1);
    uint16_t *inv_rowmap = ipiv + kInactMax;
    uint16_t *rowmap = inv_rowmap + mp;
    uint16_t *pl_col = rowmap + mp;
    uint16_t *pl_row = pl_col + mp;
    uint16_t *varlist = pl_row + mp;
    uint32_t *area = er + ge_inact_fixed_words(m, NW, MW);
    const GePlanLayout L = ge_plan_layout(m, NW, VW);
    const unsigned int count = *q.plan_count;
    const int qd = lane & 3, hs = lane >> 2;

    for (;;) {
        unsigned int li = 0;
        if (lane == 0) li = atomicAdd(q.work_ctr, 1u);
        li = __shfl_sync(FULL, li, 0);
        if (li >= count) break;
        const uint32_t *pw = q.plan + size_t(li) * q.plan_words;
        long long tmark = p.phase_cycles ? clock64() : 0;
        auto lap = [&](int ph) {
            if (p.phase_cycles) {
                const long long t = clock64();
                if (lane == 0) atomicAdd(&p.phase_cycles[ph], (unsigned long long)(t - tmark));
                tmark = t;
            }
        };
        if (pw[5] != 0u) continue;                               // rank deficient: the peeling result stays
        const long long cw = pw[0];
        const int e = int(pw[1]), R = int(pw[2]), npeel = int(pw[3]), ninact = int(pw[4]);
        uint32_t *state = area;                                              // [R] flags only
        unsigned long long *ipart = reinterpret_cast<unsigned long long *>(area + ((R + 1) & ~1));
        uint16_t *nbr = reinterpret_cast<uint16_t *>(area + ((3 * R + 3) & ~3));
        uint32_t *bmat = area + ((3 * R + 3) & ~3) + (((e * VW + 1) / 2 + 3) & ~3);
        // ---- the plan -> shared memory ----------------------------------------------------------
        for (int i = lane; i < NW; i += 32) er[i] = pw[L.er + i];
        for (int i = lane; i < (R + 1) / 2; i += 32) reinterpret_cast<uint32_t *>(rowmap)[i] = pw[L.rowmap + i];
        for (int i = lane; i < (e + 1) / 2; i += 32) reinterpret_cast<uint32_t *>(varlist)[i] = pw[L.varlist + i];
        for (int i = lane; i < (npeel + 1) / 2; i += 32) {
            reinterpret_cast<uint32_t *>(pl_col)[i] = pw[L.pl_col + i];
            reinterpret_cast<uint32_t *>(pl_row)[i] = pw[L.pl_row + i];
        }
        reinterpret_cast<uint32_t *>(icol)[lane] = pw[L.icol + lane];
        for (int i = lane; i < (e * VW + 1) / 2; i += 32) reinterpret_cast<uint32_t *>(nbr)[i] = pw[L.nbr + i];
        __syncwarp();
        lap(0);

        for (int ch = 0; ch < p.S; ch += 64) {
            const int cb = min(64, p.S - ch);
            const int nq = cb / 16;
            for (int i = lane; i < 2 * R; i += 32) reinterpret_cast<uint32_t *>(ipart)[i] = pw[L.ipart + i];
            for (int i = lane; i < R; i += 32) state[i] = 0u;
            __syncwarp();
            for (int i = lane; i < npeel; i += 32) state[pl_row[i]] = USED;
            ge_load_rhs(p, cw, rowmap, R, ch, nq, bmat, 16, lane);     // right-hand sides: the executor's syndromes
            __syncwarp();
            lap(1);
            // ---- phase 1 replay: b[row] ^= b[pivot] for the other rows of each peeled unknown, in order ----
            // 32 steps at a time: lane t fetches step t's pivot and the rows of its unknown (one 8- or 16-byte
            // load); the steps then run one after the other with their operands coming from shuffles, so that
            // the only memory round trip in a step is the row update itself.
            for (int i0 = 0; i0 < npeel; i0 += 32) {
                const bool mine = i0 + lane < npeel;
                const int mycol = mine ? pl_col[i0 + lane] : 0;
                const int myrow = mine ? pl_row[i0 + lane] : 0;
                uint32_t nw[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};   // up to 8 rows, two per word
                if (mine) {
                    const uint32_t *nr = reinterpret_cast<const uint32_t *>(nbr + size_t(mycol) * VW);
#pragma unroll
                    for (int w = 0; w < 4; w++) if (2 * w < VW) nw[w] = nr[w];
                }
                const int cnt = min(32, npeel - i0);
                for (int t = 0; t < cnt; t++) {
                    const int ri = __shfl_sync(FULL, myrow, t);
                    const uint32_t w0 = __shfl_sync(FULL, nw[0], t), w1 = __shfl_sync(FULL, nw[1], t);
                    const uint32_t w2 = __shfl_sync(FULL, nw[2], t), w3 = __shfl_sync(FULL, nw[3], t);
                    const uint32_t wsel = (hs >> 1) == 0 ? w0 : ((hs >> 1) == 1 ? w1 : ((hs >> 1) == 2 ? w2 : w3));
                    const uint32_t r2 = (hs & 1) ? (wsel >> 16) : (wsel & 0xFFFFu);       // hit slot hs of this step
                    if (r2 != 0xFFFFu && int(r2) != ri && qd < nq) {
                        const uint4 pb = *reinterpret_cast<const uint4 *>(bmat + size_t(ri) * 16 + qd * 4);
                        uint4 *br = reinterpret_cast<uint4 *>(bmat + size_t(r2) * 16 + qd * 4);
                        uint4 a4 = *br;
                        a4.x ^= pb.x; a4.y ^= pb.y; a4.z ^= pb.z; a4.w ^= pb.w;
                        *br = a4;
                    }
                    __syncwarp();
                }
            }
            lap(2);
            // ---- phase 2: dense solve on [ipart | b] of the unused rows (the plan is known to have full rank) ----
            // Only the rows no peeled unknown used as its pivot take part: R - npeel of them, listed once.
            uint16_t *ulist = inv_rowmap;                 // (the inverse row map is not needed in this kernel)
            int nU = 0;
            for (int r0 = 0; r0 < R; r0 += 32) {
                const int r = r0 + lane;
                const bool un = r < R && !(state[r] & USED);
                const unsigned bal = __ballot_sync(FULL, un);
                if (un) ulist[nU + __popc(bal & ((1u << lane) - 1u))] = uint16_t(r);
                nU += __popc(bal);
            }
            __syncwarp();
            for (int c = 0; c < ninact; c++) {
                unsigned best = 0xFFFFFFFFu;
                for (int i = lane; i < nU; i += 32) {
                    const int r = ulist[i];
                    if (!(state[r] & TAKEN) && ((ipart[r] >> c) & 1u)) { best = uint32_t(r); break; }
                }
                best = __reduce_min_sync(FULL, best);
                if (best == 0xFFFFFFFFu) break;                           // (cannot happen: the pattern kernel found the pivots)
                const int pr = int(best);
                const unsigned long long pI = ipart[pr];
                uint4 pb = make_uint4(0u, 0u, 0u, 0u);
                if (qd < nq) pb = *reinterpret_cast<const uint4 *>(bmat + size_t(pr) * 16 + qd * 4);
                __syncwarp();
                if (lane == 0) { state[pr] |= TAKEN; ipiv[c] = uint16_t(pr); }
                for (int i0 = 0; i0 < nU; i0 += 8) {      // four lanes per row: test, then one quarter of b each
                    const int r2 = ulist[min(i0 + hs, nU - 1)];
                    const unsigned long long ip = ipart[r2];
                    const bool hit = i0 + hs < nU && r2 != pr && ((ip >> c) & 1u);
                    __syncwarp();                          // (all four lanes have read ipart[r2])
                    if (hit) {
                        if (qd == 0) ipart[r2] = ip ^ pI;
                        if (qd < nq) {
                            uint4 *br = reinterpret_cast<uint4 *>(bmat + size_t(r2) * 16 + qd * 4);
                            uint4 a4 = *br;
                            a4.x ^= pb.x; a4.y ^= pb.y; a4.z ^= pb.z; a4.w ^= pb.w;
                            *br = a4;
                        }
                    }
                }
                __syncwarp();
            }
            lap(3);
            // ---- phase 3: read the unknowns off their pivot rows -----------------------------------------
            uint8_t *out = p.out + size_t(cw) * p.k * p.S + ch;
            for (int i = hs; i < npeel + ninact; i += 8) {
                const bool pe = i < npeel;
                const int u = varlist[pe ? pl_col[i] : icol[i - npeel]];
                const int r = pe ? pl_row[i] : ipiv[i - npeel];
                if (u >= p.k || qd >= nq) continue;                       // only systematic symbols are output
                uint4 acc = *reinterpret_cast<const uint4 *>(bmat + size_t(r) * 16 + qd * 4);
                unsigned long long bits = pe ? ipart[r] : 0ull;
                while (bits) {
                    const int c = __ffsll((long long)bits) - 1;
                    bits &= bits - 1ull;
                    const uint4 v = *reinterpret_cast<const uint4 *>(bmat + size_t(ipiv[c]) * 16 + qd * 4);
                    acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
                }
                *reinterpret_cast<uint4 *>(out + size_t(u) * p.S + qd * 16) = acc;
            }
            __syncwarp();
            lap(4);
        }
        if (p.phase_cycles && lane == 0) atomicAdd(&p.phase_cycles[6], 1ull);
    }
}

// ---- host side -------------------------------------------------------------------------------
struct HybridScratch {
    uint8_t *d_synd = nullptr;          // [max_batch][m][S] check syndromes written by the executor
    unsigned int *d_list = nullptr;     // [max_batch] stalled codewords, written by the peel kernel
    unsigned int *d_list2 = nullptr, *d_list3 = nullptr, *d_list4 = nullptr;   // deferred by warp stages 0 / 1 / 2
    unsigned int *d_count = nullptr;    // [16]: list counts [0..3], work counters [4..6], plan counts [8..9], apply work counters [10..11]
    // warp stages: [0] inactivation, typical slots; [1] inactivation, worst-case slots; [2] per-warp Gauss-Jordan
    int wpc[3] = {0, 0, 0}, slot_words[3] = {0, 0, 0};
    uint32_t *d_plan = nullptr;          // [max_batch][plan_words] recorded solutions of the inactivation stage
    int plan_words = 0;
    int wpc_pat[2] = {0, 0}, slot_words_pat[2] = {0, 0};     // the pattern kernel of stages 0 / 1 (no payload: smaller slots, more warps)
    uint32_t *d_gmat = nullptr;         // per-CTA matrices when they do not fit in shared memory
    int grid = 0, smem = 0, RSW = 0;
    bool ready = false;
};

inline void hybrid_free(HybridScratch &h)
{
    cudaFree(h.d_synd); cudaFree(h.d_list); cudaFree(h.d_list2); cudaFree(h.d_list3); cudaFree(h.d_list4); cudaFree(h.d_plan); cudaFree(h.d_count); cudaFree(h.d_gmat);
    h = HybridScratch();
}

// Allocates the stage's scratch on first use (syndromes max_batch * m * S, plans max_batch * ~16 KB).
inline int hybrid_prepare(HybridScratch &h, const HostCode &code, int S, int NW, int MW, int num_sms, int smem_optin,
                          long long max_batch, std::string &err)
{
    if (h.ready) return LDPC_OK;
    auto bad = [&](const char *what, cudaError_t e) {   // a failed set-up leaves nothing allocated: a retry starts clean
        err = std::string(what) + ": " + cudaGetErrorString(e);
        hybrid_free(h);
        return e == cudaErrorMemoryAllocation ? LDPC_ERR_NOMEM : LDPC_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaMalloc(&h.d_synd, size_t(max_batch) * code.m * S)) != cudaSuccess) return bad("hybrid: syndrome scratch", e);
    if ((e = cudaMalloc(&h.d_list, size_t(max_batch) * 4)) != cudaSuccess) return bad("hybrid: list", e);
    if ((e = cudaMalloc(&h.d_list2, size_t(max_batch) * 4)) != cudaSuccess) return bad("hybrid: list", e);
    if ((e = cudaMalloc(&h.d_list3, size_t(max_batch) * 4)) != cudaSuccess) return bad("hybrid: list", e);
    if ((e = cudaMalloc(&h.d_list4, size_t(max_batch) * 4)) != cudaSuccess) return bad("hybrid: list", e);
    if ((e = cudaMalloc(&h.d_count, 16 * 4)) != cudaSuccess) return bad("hybrid: count", e);
    {   // warp stages: [0] inactivation decoding, slots for a typical stopping set (2/3 of the checks involved);
        //              [1] the same with worst-case slots, for what [0] defers for size;
        //              [2] plain Gauss-Jordan per warp, worst-case slots: more than kInactMax inactivations
        const long long budget_w = (long long)(smem_optin - 2048) / 4;
        // rows a typical slot holds: 2/3 of m (LDPC_CUDA_GE_TYP, in %).  Measured on (2040,1530), 10 sweeps first: 85 % is 10 % faster at
        // 12/64 (3.16 against 3.51 ms per 32 768 codewords: few frames stall and almost all of them then fit the first stage) and equal at
        // 13/64 with 32 768 codewords per chunk, but 4 % slower at 13/64 with the bench's 65 536 (184.8 against 177.8 ms per 1 Mi codewords);
        // 75, 80, 90, 100 % are slower at 13/64 either way (one warp less per CTA without fewer deferrals)
        int typ = 2 * code.m / 3;
        if (const char *et = getenv("LDPC_CUDA_GE_TYP")) typ = std::max(16, std::min(code.m, code.m * atoi(et) / 100));
        const int fx = ge_inact_fixed_words(code.m, NW, MW);
        const long long need[3] = {fx + ge_inact_area_words(typ, typ, code.VW, true),
                                   fx + ge_inact_area_words(code.m, code.m, code.VW, true),
                                   ge_warp_fixed_words(code.m, NW) + ge_warp_matrix_words(code.m, code.m, true)};
        for (int i = 0; i < 3; i++) {
            int w = int(std::max<long long>(1, std::min<long long>(16, budget_w / need[i])));
            if (i == 0)
                if (const char *ev = getenv("LDPC_CUDA_GE_WPC")) w = std::max(1, std::min(16, atoi(ev)));
            h.wpc[i] = w;
            h.slot_words[i] = int((budget_w / w) & ~3ll);
            if (i != 0 && h.slot_words[i] < need[i]) h.wpc[i] = 0;                  // worst case does not fit: stage skipped
        }
        if (h.slot_words[0] <= fx + 64) h.wpc[0] = 0;
        if (h.wpc[1] >= h.wpc[0] && h.wpc[0] > 0) { h.wpc[0] = h.wpc[1]; h.slot_words[0] = h.slot_words[1]; h.wpc[1] = 0; }   // no need for two sizes
        for (int i = 0; i < 2; i++) {   // pattern slots: typical for stage 0 when a worst-case stage follows, else worst case
            const int rows = (i == 0 && h.wpc[1] > 0) ? typ : code.m;
            const long long sp = fx + ge_inact_area_words(rows, rows, code.VW, false);
            h.wpc_pat[i] = int(std::max<long long>(1, std::min<long long>(16, budget_w / sp)));
            h.slot_words_pat[i] = int((budget_w / h.wpc_pat[i]) & ~3ll);
            if (h.wpc[i] == 0 || h.slot_words_pat[i] < sp) h.wpc_pat[i] = 0;
        }
        if ((e = cudaFuncSetAttribute(hybrid_ge_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 1024)) != cudaSuccess)
            return bad("hybrid: cudaFuncSetAttribute", e);
        if ((e = cudaFuncSetAttribute(hybrid_inact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 1024)) != cudaSuccess)
            return bad("hybrid: cudaFuncSetAttribute", e);
        if ((e = cudaFuncSetAttribute(hybrid_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 1024)) != cudaSuccess)
            return bad("hybrid: cudaFuncSetAttribute", e);
        h.plan_words = ge_plan_layout(code.m, NW, code.VW).words;
        if (h.wpc[0] > 0 && h.wpc_pat[0] > 0)
            if ((e = cudaMalloc(&h.d_plan, size_t(max_batch) * h.plan_words * 4)) != cudaSuccess) return bad("hybrid: plan buffer", e);
    }
    h.RSW = 2 * MW + 1;
    const size_t base = size_t(code.m) * 64 + ge_small_bytes(code.m, NW) + 16;
    const size_t mat = size_t(code.m) * h.RSW * 4;
    const size_t budget = size_t(smem_optin) - 1024;
    if (base > budget) { err = "hybrid: code too large for the elimination kernel"; hybrid_free(h); return LDPC_ERR_UNSUPPORTED; }
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
