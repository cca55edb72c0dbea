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

// ---- host side -------------------------------------------------------------------------------
struct HybridScratch {
    uint8_t *d_full = nullptr;          // [max_batch][n][S]
    unsigned int *d_list = nullptr;     // [max_batch]
    unsigned int *d_count = nullptr;    // [1]
    uint32_t *d_gmat = nullptr;         // per-CTA matrices when they do not fit in shared memory
    int grid = 0, smem = 0, RSW = 0;
    bool ready = false;
};

inline void hybrid_free(HybridScratch &h)
{
    cudaFree(h.d_full); cudaFree(h.d_list); cudaFree(h.d_count); cudaFree(h.d_gmat);
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
    if ((e = cudaMalloc(&h.d_count, 4)) != cudaSuccess) return bad("hybrid: count", e);
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
