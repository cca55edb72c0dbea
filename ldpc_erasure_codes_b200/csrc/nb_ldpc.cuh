// nb_ldpc.cuh -- non-binary GF(2^8) LDPC erasure code.
//
// The reference's MATLAB study of LDPC codes over GF(256) (Matlab/ErasureCodes_NonBinaryLDPCSim.m,
// Matlab/My_LDPC_HybridML_NonBinary_Erasure_Decoder.m): the binary parity-check matrix keeps its structure and
// every 1 becomes a nonzero field element (sim :51-58); a check reads  sum_u h[c][u] * y[u] = 0  over GF(2^8),
// field polynomial 0x171 (sim :69, the field of Matlab/GF_256_add_mult_inv_tables.mat).
//   encoder  (sim :176-182): parity p = inv(h_diag) * sum_{others} h * c, row after row;
//   decoder  (decoder :19-55): serial sweeps, a check with ONE erased member recovers it as
//            inv(h) * sum_{others} h * y; then (:57-125) Gauss-Jordan over GF(256) on the residual set.
// A packet symbol is S bytes; a check's coefficient multiplies every byte of the symbol.
//
// WHICH check recovers WHICH symbol, and in which sweep, depends on the erasure mask alone and is the same as for
// the binary code of the same structure: the pattern phase is peel_schedule_kernel, unchanged.  What is new is the
// payload arithmetic:
//   nb_exec_kernel<W>  applies a schedule (per-codeword from the peel kernel, or the encoder's static one) to a W-byte
//     slice of a codeword held in shared memory, level by level: entry (c, v) is
//     row[v] = sum_{u != v} g_u * row[u] with g_u = h[c][u] / h[c][v].  Multiplication of four packed bytes by a constant
//     is bit-sliced over the CONSTANT: x * g = XOR_j bit_j(g) * (x * 2^j).  The thread accumulates the members into eight
//     bit planes, acc_j ^= x_u & mask_j(g_u) (one LOP3 per plane and word; the masks of every constant sit in shared
//     memory), and folds the planes once per entry by Horner's rule, r = (...(acc_7 * 2 ^ acc_6) * 2 ...) ^ acc_0
//     (seven doublings of five ALU operations) -- 8 operations per member and word instead of ~48.
//   nb_ge_kernel  one CTA per codeword that still has erasures: A = H(:, E) (dense bytes: in shared memory when the
//     residual set fits, else in an L2-resident per-CTA workspace), b = sum_{known} h * y, Gauss-Jordan with the lowest
//     unused row as pivot.  The MATLAB code aborts
//     iff a column has no pivot, i.e. iff rank(A) < |E|; otherwise the solution is unique and any exact solver
//     returns its bytes.  On abort the state after the sweeps is kept and the frame reported, as for the binary
//     hybrid decoder.
#pragma once
#include "device_utils.cuh"

namespace ldpc {

constexpr int kNbThreads = 256;

__device__ __forceinline__ uint32_t nb_xtime4(uint32_t x)   // four packed field elements times alpha (poly 0x171)
{
    return ((x & 0x7F7F7F7Fu) << 1) ^ (((x >> 7) & 0x01010101u) * 0x71u);
}

struct NbExecParams {
    const uint8_t *in;          // [B][rows_in][S]
    uint8_t *out;               // [B][rows_out][S]
    const uint8_t *sched;       // per-codeword blobs (stride sched_stride) or the static blob (stride 0)
    const uint16_t *cidx;       // [m][RW] members, pad 0xFFFF
    const uint8_t *coef;        // [m][RW] coefficients, pad 0
    const uint8_t *tab;         // log[256] | alog[512]
    const uint32_t *m8;         // [256][8] bit-plane masks of every constant
    long long B;
    int sched_stride, sched_max;
    int n, m, RW, S, rows_in, rows_out, slices;
};

// shared: slot [n][W] | cidx [m][RW] u16 | coef [m][RW] u8 | m8 [256][8] u32 | log/alog 768 | blob
template <int W>
__global__ void __launch_bounds__(kNbThreads) nb_exec_kernel(const NbExecParams p)
{
    constexpr int WPE = W / 4;                 // threads (32-bit words) per entry
    constexpr int EPP = kNbThreads / WPE;      // entries per pass of the CTA
    extern __shared__ __align__(16) uint8_t nsm[];
    uint32_t *slot = reinterpret_cast<uint32_t *>(nsm);
    uint16_t *cidx_s = reinterpret_cast<uint16_t *>(nsm + size_t(p.n) * W);
    uint8_t *coef_s = reinterpret_cast<uint8_t *>(cidx_s + size_t(p.m) * p.RW);
    uint32_t *m8 = reinterpret_cast<uint32_t *>(coef_s + ((size_t(p.m) * p.RW + 15) & ~size_t(15)));
    uint8_t *lg = reinterpret_cast<uint8_t *>(m8 + 256 * 8);
    uint8_t *al = lg + 256;
    uint8_t *blob = al + 512;
    const int tid = threadIdx.x;
    for (int i = tid; i < p.m * p.RW; i += kNbThreads) { cidx_s[i] = p.cidx[i]; coef_s[i] = p.coef[i]; }
    for (int i = tid; i < 256 * 8; i += kNbThreads) m8[i] = p.m8[i];
    for (int i = tid; i < 768; i += kNbThreads) lg[i] = p.tab[i];
    const bool dynamic = p.sched_stride != 0;
    if (!dynamic)
        for (int i = tid; i < p.sched_max / 4; i += kNbThreads) reinterpret_cast<uint32_t *>(blob)[i] = reinterpret_cast<const uint32_t *>(p.sched)[i];
    __syncthreads();

    const int wi = tid % WPE, es = tid / WPE;
    for (long long unit = blockIdx.x; unit < p.B * p.slices; unit += gridDim.x) {
        const long long b = unit / p.slices;
        const int sl = int(unit % p.slices);
        // ---- load the slice (and the codeword's schedule) ------------------------------------------------
        const uint8_t *src = p.in + size_t(b) * p.rows_in * p.S + size_t(sl) * W;
        for (int r = es; r < p.rows_in; r += EPP) slot[r * WPE + wi] = *reinterpret_cast<const uint32_t *>(src + size_t(r) * p.S + wi * 4);
        if (dynamic) {
            const uint32_t *g = reinterpret_cast<const uint32_t *>(p.sched + size_t(b) * p.sched_stride);
            const int words = (16 + 4 * int(g[0]) + 2 * (int(g[1] & 0xFFFFu) + 1) + 3) / 4;
            for (int i = tid; i < words; i += kNbThreads) reinterpret_cast<uint32_t *>(blob)[i] = g[i];
        }
        __syncthreads();
        const uint32_t *hdr = reinterpret_cast<const uint32_t *>(blob);
        const int ne = int(hdr[0]), nl = int(hdr[1] & 0xFFFFu);
        const uint32_t *ent = hdr + 4;
        const uint16_t *lvo = reinterpret_cast<const uint16_t *>(ent + ne);
        // ---- the schedule, level by level ------------------------------------------------------------------
        for (int L = 0; L < nl; L++) {
            const int s0 = lvo[L], s1 = lvo[L + 1];
            for (int i = s0 + es; i < s1; i += EPP) {
                const uint32_t e = ent[i];
                const uint32_t v = e & 0xFFFFu, c = e >> 16;
                const uint16_t *row = cidx_s + size_t(c) * p.RW;
                const uint8_t *hc = coef_s + size_t(c) * p.RW;
                int lhv = 0;                                   // log of the target's coefficient
                for (int t = 0; t < p.RW; t++)
                    if (row[t] == v) lhv = lg[hc[t]];
                uint32_t acc[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                for (int t = 0; t < p.RW; t++) {
                    const uint32_t u = row[t];
                    if (u == 0xFFFFu || u == v) continue;
                    const uint32_t g = al[int(lg[hc[t]]) + 255 - lhv];     // h[c][u] / h[c][v]
                    const uint32_t x = slot[u * WPE + wi];
                    const uint4 ma = *reinterpret_cast<const uint4 *>(m8 + g * 8), mb = *reinterpret_cast<const uint4 *>(m8 + g * 8 + 4);
                    acc[0] ^= x & ma.x; acc[1] ^= x & ma.y; acc[2] ^= x & ma.z; acc[3] ^= x & ma.w;
                    acc[4] ^= x & mb.x; acc[5] ^= x & mb.y; acc[6] ^= x & mb.z; acc[7] ^= x & mb.w;
                }
                uint32_t r = acc[7];
#pragma unroll
                for (int j = 6; j >= 0; j--) r = nb_xtime4(r) ^ acc[j];
                slot[v * WPE + wi] = r;
            }
            __syncthreads();
        }
        // ---- store ---------------------------------------------------------------------------------------------
        uint8_t *dst = p.out + size_t(b) * p.rows_out * p.S + size_t(sl) * W;
        for (int r = es; r < p.rows_out; r += EPP) *reinterpret_cast<uint32_t *>(dst + size_t(r) * p.S + wi * 4) = slot[r * WPE + wi];
        __syncthreads();
    }
}

struct NbGeParams {
    uint8_t *work;              // [B][n][S] codewords after the sweeps (unknown symbols: anything)
    const uint32_t *mask;       // [B][NW]
    const uint8_t *sched;       // blobs: which symbols the sweeps recovered
    const unsigned int *list;   // codewords that still have erasures
    const unsigned int *list_count;
    uint8_t *fail;              // [B] systematic-failure flags (cleared on success)
    unsigned long long *stats;  // [3] attempts, [4] aborts, [5] frames recovered that peeling had counted as errors
    const uint16_t *cidx;       // [m][RW]
    const uint8_t *coef;        // [m][RW]
    const uint8_t *tab;         // log | alog
    uint8_t *gA;                // [grid][m][m] workspace: A, for residual sets too large for shared memory
    uint8_t *gB;                // [grid][m][S] workspace: right-hand sides, likewise
    int n, k, m, RW, NW, S, stride;
    int work_bytes;             // shared bytes behind the small arrays: A (and b) live there when they fit
};

// shared: log/alog 768 | unknown bitmap [NW] u32 | prefix [NW] u16 | used [m] u8 | piv [m] u16 | rows list [m] u16 | E [m] u16 |
//         work area: A [m][lda] and b [m][S] when they fit (a residual set of ~250 symbols: 510 x 256 + 510 x 64 bytes)
__global__ void __launch_bounds__(kNbThreads) nb_ge_kernel(const NbGeParams p)
{
    extern __shared__ __align__(16) uint8_t gsm[];
    uint8_t *lg = gsm;
    uint8_t *al = lg + 256;
    uint32_t *unk = reinterpret_cast<uint32_t *>(gsm + 768);
    uint16_t *pre = reinterpret_cast<uint16_t *>(unk + p.NW);
    uint8_t *used = reinterpret_cast<uint8_t *>(pre + ((p.NW + 1) & ~1));
    uint16_t *piv = reinterpret_cast<uint16_t *>(used + ((p.m + 3) & ~3));
    uint16_t *rlist = piv + p.m;
    uint16_t *Evar = rlist + p.m;
    uint8_t *work_s = reinterpret_cast<uint8_t *>(Evar + p.m) + ((16 - (reinterpret_cast<uintptr_t>(Evar + p.m) & 15)) & 15);
    __shared__ int s_e, s_piv, s_nrows, s_abort;
    const int tid = threadIdx.x, m = p.m, S = p.S;
    for (int i = tid; i < 768; i += kNbThreads) lg[i] = p.tab[i];
    auto mul = [&](uint32_t a, uint32_t b) -> uint32_t { return (a && b) ? al[int(lg[a]) + int(lg[b])] : 0u; };
    __syncthreads();

    const unsigned int count = *p.list_count;
    for (unsigned int li = blockIdx.x; li < count; li += gridDim.x) {
        const long long cw = p.list[li];
        uint8_t *y = p.work + size_t(cw) * p.n * S;
        const uint32_t *hdr = reinterpret_cast<const uint32_t *>(p.sched + size_t(cw) * p.stride);
        // ---- the residual set E: erased on arrival and not recovered by the sweeps ---------------------
        for (int w = tid; w < p.NW; w += kNbThreads) {
            uint32_t x = p.mask[cw * p.NW + w];
            if (w == p.NW - 1 && (p.n & 31)) x &= 0xFFFFFFFFu >> (32 - (p.n & 31));
            unk[w] = x;
        }
        if (tid == 0) s_abort = 0;
        __syncthreads();
        const int ne = int(hdr[0]);
        for (int i = tid; i < ne; i += kNbThreads) {
            const uint32_t v = hdr[4 + i] & 0xFFFFu;
            atomicAnd(&unk[v >> 5], ~(1u << (v & 31)));
        }
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int w = 0; w < p.NW; w++) { pre[w] = uint16_t(run); run += __popc(unk[w]); }
            s_e = run;
        }
        __syncthreads();
        const int e = s_e;
        if (tid == 0) atomicAdd(&p.stats[3], 1ull);
        if (e > m) {                       // more unknowns than equations (the MATLAB code would index past its matrix)
            if (tid == 0) atomicAdd(&p.stats[4], 1ull);
            __syncthreads();
            continue;
        }
        // A and b: in shared memory when this residual set fits, else in the CTA's global workspace (row stride lda)
        const int lda_s = (e + 15) & ~15;
        const bool a_sh = m * lda_s <= p.work_bytes - 16;
        const bool b_sh = a_sh && m * lda_s + m * S <= p.work_bytes - 16;
        const int lda = a_sh ? lda_s : m;
        uint8_t *A = a_sh ? work_s : p.gA + size_t(blockIdx.x) * m * m;
        uint8_t *Bv = b_sh ? work_s + size_t(m) * lda_s : p.gB + size_t(blockIdx.x) * m * S;
        for (int w = tid; w < p.NW; w += kNbThreads) {
            uint32_t x = unk[w];
            int at = pre[w];
            while (x) { const int bit = __ffs(x) - 1; x &= x - 1; Evar[at++] = uint16_t(w * 32 + bit); }
        }
        // ---- A = H(:, E), b = sum over the known members h * y ------------------------------------------
        for (int i = tid; i < m * e; i += kNbThreads) A[size_t(i / e) * lda + (i % e)] = 0;
        for (int i = tid; i < m; i += kNbThreads) used[i] = 0;
        __syncthreads();
        for (int r = tid / 32; r < m; r += kNbThreads / 32) {             // a warp per check row
            const uint16_t *row = p.cidx + size_t(r) * p.RW;
            const uint8_t *hc = p.coef + size_t(r) * p.RW;
            for (int b0 = (tid & 31) * 4; b0 < S; b0 += 128) {            // four bytes of the right-hand side per lane
                uint32_t acc = 0;
                for (int t = 0; t < p.RW; t++) {
                    const uint32_t u = row[t];
                    if (u == 0xFFFFu) continue;
                    if ((unk[u >> 5] >> (u & 31)) & 1u) continue;
                    const uint32_t h = hc[t];
                    const uint32_t x = *reinterpret_cast<const uint32_t *>(y + size_t(u) * S + b0);
                    acc ^= mul(h, x & 0xFFu) | (mul(h, (x >> 8) & 0xFFu) << 8) | (mul(h, (x >> 16) & 0xFFu) << 16) | (mul(h, x >> 24) << 24);
                }
                *reinterpret_cast<uint32_t *>(Bv + size_t(r) * S + b0) = acc;
            }
            if ((tid & 31) < p.RW) {
                const uint32_t u = row[tid & 31];
                if (u != 0xFFFFu && ((unk[u >> 5] >> (u & 31)) & 1u))
                    A[size_t(r) * lda + pre[u >> 5] + __popc(unk[u >> 5] & ((1u << (u & 31)) - 1u))] = hc[tid & 31];
            }
        }
        __syncthreads();
        // ---- Gauss-Jordan: column by column, the lowest unused row with a nonzero entry is the pivot -----
        for (int col = 0; col < e; col++) {
            if (tid == 0) { s_piv = m; s_nrows = 0; }
            __syncthreads();
            for (int r = tid; r < m; r += kNbThreads)
                if (!used[r] && A[size_t(r) * lda + col]) atomicMin(&s_piv, r);
            __syncthreads();
            const int pr = s_piv;
            if (pr >= m) { if (tid == 0) s_abort = 1; break; }            // rank(A) < e (decoder :82-85)
            const uint32_t inv = al[255 - int(lg[A[size_t(pr) * lda + col]])];
            // the other rows with a nonzero entry in this column
            for (int r = tid; r < m; r += kNbThreads)
                if (r != pr && A[size_t(r) * lda + col]) rlist[atomicAdd(&s_nrows, 1)] = uint16_t(r);
            __syncthreads();
            // scale the pivot row (entries left of `col` in an unused row are zero)
            for (int j = col + tid; j < e; j += kNbThreads) A[size_t(pr) * lda + j] = uint8_t(mul(inv, A[size_t(pr) * lda + j]));
            for (int b0 = tid; b0 < S; b0 += kNbThreads) Bv[size_t(pr) * S + b0] = uint8_t(mul(inv, Bv[size_t(pr) * S + b0]));
            __syncthreads();
            const int nrows = s_nrows;
            for (int q = tid / 32; q < nrows; q += kNbThreads / 32) {      // a warp per row to clear
                const int r = rlist[q];
                const uint32_t f = A[size_t(r) * lda + col];
                __syncwarp();
                for (int j = col + (tid & 31); j < e; j += 32) A[size_t(r) * lda + j] ^= uint8_t(mul(f, A[size_t(pr) * lda + j]));
                for (int b0 = tid & 31; b0 < S; b0 += 32) Bv[size_t(r) * S + b0] ^= uint8_t(mul(f, Bv[size_t(pr) * S + b0]));
            }
            if (tid == 0) { used[pr] = 1; piv[col] = uint16_t(pr); }
            __syncthreads();
        }
        __syncthreads();
        if (s_abort) {
            if (tid == 0) atomicAdd(&p.stats[4], 1ull);
        } else {
            for (int i = tid; i < e * S; i += kNbThreads) y[size_t(Evar[i / S]) * S + (i % S)] = Bv[size_t(piv[i / S]) * S + (i % S)];
            if (tid == 0) {
                if (p.fail) p.fail[cw] = 0;
                if (hdr[3] != 0u) atomicAdd(&p.stats[5], 1ull);
            }
        }
        __syncthreads();
    }
}

}  // namespace ldpc
