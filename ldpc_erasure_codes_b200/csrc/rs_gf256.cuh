// rs_gf256.cuh -- Reed-Solomon GF(2^8) erasure codec (the reference's equal-rate comparison code).
//
// Field and code as in the reference's MATLAB:
//   * GF(2^8) with primitive polynomial 0x171, alpha = 2 (Matlab/Build_GF256_Lookup_Tables.m:11-24;
//     the committed table file GF_256_add_mult_inv_tables.mat is this field);
//   * generator G[i][j] = alpha^(i*j), i = 1..k, j = 1..n, systematised G_sys = G(:,1:k)^-1 G = [I | P]
//     (Matlab/Test_My_RS_Decode.m:30-37);
//   * decoding uses the FIRST k received symbols (Matlab/ReedSolomonErasureCodes.m:80-85) and solves
//     for the erased systematic symbols (Matlab/My_RS_Decode_Optimize_With_GFTables.m).  The code is
//     MDS, so the solution is unique and any exact solver returns the MATLAB decoder's bytes.
//
// GPU formulation: one CTA per codeword.
//   pattern part (per codeword, byte arithmetic with log/antilog tables in shared memory):
//     E = erased systematic symbols (t of them), R = the first t received repair symbols;
//     M[a][b] = G_sys[E_b][R_a] is inverted by Gauss-Jordan (t <= n-k), and the decode matrix
//     D (t x k) over the k used received symbols is formed:  u_E = M^-1 (c_R + P_R^T u_known).
//   payload part (what the time goes to, ~ t*k*S multiply-accumulates): a thread owns one 16-byte column (four 32-bit
//     words) of all symbols and every fourth output.  Per input symbol x the CTA builds, for each of its 64 columns, the
//     multiples v*x (v = 0..15) and (16v)*x from the eight doublings of the packed words (5 ALU ops per doubling) in shared
//     memory; an output with coefficient c then adds lo[c & 15] ^ hi[c >> 4]: two conflict-free 128-bit shared loads and four
//     3-input XORs per 16-byte multiply-accumulate (round 1: 32 LOP3, bit-sliced over the constant; the ALU pipe was the
//     limit at 60 % busy, profiles/r02_rs.md).  Tables are double-buffered: one CTA barrier per input symbol.
//   The encoder is the same payload routine with the static matrix D = P^T.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/ldpc_cuda.h"

struct rs_ctx {
    int n = 0, k = 0, S = 0, device = 0;
    long long max_batch = 0;
    int num_sms = 0;
    std::vector<uint8_t> gsys;   // k x n, host copy
    uint8_t *d_P = nullptr;      // [k][n-k] parity part of G_sys
    uint8_t *d_tab = nullptr;    // log[256] | alog[512]
    int smem = 0;
};

namespace ldpc {

constexpr int kRsThreads = 256;
constexpr int kRsMaxT = 128;     // n - k <= 128 (matrix dimensions)
constexpr int kRsOwn = 16;       // outputs a thread accumulates per pass over the inputs (4 * kRsOwn per CTA pass)
constexpr int kRsTabBytes = 32 * (kRsThreads / 4) * 16;   // multiples of one input symbol: 32 entries x 64 columns x 16 bytes
static_assert(kRsTabBytes <= kRsMaxT * 2 * kRsMaxT, "the first table sits in the pattern part's matrix");

struct RsParams {
    const uint8_t *in;      // decode: [B][n][S] received codewords; encode: [B][k][S] info
    uint8_t *out;           // decode: [B][k][S]; encode: [B][n][S]
    const uint32_t *mask;   // decode: [B][NW]
    uint8_t *fail;          // decode: [B] or nullptr
    const uint8_t *P;       // [k][n-k]
    const uint8_t *tab;     // log | alog
    long long B;
    int n, k, S, NW, encode;
};

__device__ __forceinline__ uint32_t gf_xtime4(uint32_t x)   // multiply four packed field elements by alpha (poly 0x171)
{
    return ((x & 0x7F7F7F7Fu) << 1) ^ (((x >> 7) & 0x01010101u) * 0x71u);
}

__global__ void __launch_bounds__(kRsThreads, 2) rs_codec_kernel(const RsParams p)
{
    extern __shared__ __align__(16) uint8_t rs_smem[];
    const int n = p.n, k = p.k, r_ = n - k, S = p.S;
    // shared layout: D[128][k8] | P[k][r_] | lg2[256] u16 | al2[1024] | aug[128][256] (payload part: multiples table 0) | multiples table 1 |
    //                rlist[256] elist[128] plist[128]
    // lg2 / al2: a product is al2[lg2[a] + lg2[b]] without a test for zero -- lg2[0] = 511, al2[i] = alpha^i for i <= 508 and 0 above
    const int k8 = (k + 7) & ~7;
    uint8_t *D = rs_smem;
    uint8_t *Ps = D + kRsMaxT * k8;
    uint16_t *lg2 = reinterpret_cast<uint16_t *>(Ps + ((k * r_ + 15) & ~15));
    uint8_t *al2 = reinterpret_cast<uint8_t *>(lg2 + 256);
    uint8_t *aug = al2 + 1024;
    uint8_t *tab1 = aug + kRsMaxT * 2 * kRsMaxT;
    uint8_t *rlist = tab1 + kRsTabBytes;
    uint8_t *elist = rlist + 256;
    uint8_t *plist = elist + kRsMaxT;
    __shared__ int s_t, s_nrecv, s_piv;
    const int tid = threadIdx.x;

    for (int i = tid; i < k * r_; i += kRsThreads) Ps[i] = p.P[i];
    for (int i = tid; i < 256; i += kRsThreads) lg2[i] = i ? uint16_t(p.tab[i]) : uint16_t(511);
    for (int i = tid; i < 1024; i += kRsThreads) al2[i] = i <= 508 ? p.tab[256 + i] : uint8_t(0);      // (the host table is alpha^i for i < 512)
    if (p.encode) {   // static matrix: parity b = sum_i P[i][b] u_i
        for (int i = tid; i < r_ * k; i += kRsThreads) D[(i / k) * k8 + (i % k)] = p.P[(i % k) * r_ + (i / k)];
        for (int i = tid; i < k; i += kRsThreads) rlist[i] = uint8_t(i);
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = kRsThreads / 32;
    constexpr int AS = 2 * kRsMaxT;              // row stride of the augmented matrix

    for (long long cw = blockIdx.x; cw < p.B; cw += gridDim.x) {
        int t = r_;
        bool ok = true;
        if (!p.encode) {
            // ---- which symbols are used: first k received; E = erased systematic; R = first t received repair ---
            if (warp == 0) {      // 32 symbols per step, places by ballot + popc
                const uint32_t *mk = p.mask + cw * p.NW;
                int nrecv = 0, te = 0;
                const unsigned lt = (1u << lane) - 1u;
                for (int j0 = 0; j0 < n; j0 += 32) {
                    const int j = j0 + lane;
                    const bool in_cw = j < n;
                    const bool er = in_cw && ((mk[j >> 5] >> (j & 31)) & 1u);
                    const unsigned be = __ballot_sync(0xFFFFFFFFu, er && j < k), br = __ballot_sync(0xFFFFFFFFu, in_cw && !er);
                    if (er && j < k) { const int at = te + __popc(be & lt); if (at < kRsMaxT) elist[at] = uint8_t(j); }
                    if (in_cw && !er) { const int at = nrecv + __popc(br & lt); if (at < k) rlist[at] = uint8_t(j); }
                    te += __popc(be);
                    nrecv = min(k, nrecv + __popc(br));
                }
                __syncwarp();
                if (nrecv == k) {   // the used repair symbols are the tail of rlist; exactly te of them
                    for (int a = lane; a < te; a += 32) plist[a] = rlist[k - te + a];
                } else {
                    te = min(te, kRsMaxT);
                }
                if (lane == 0) { s_nrecv = nrecv; s_t = te; }
            }
            __syncthreads();
            t = s_t;
            ok = (s_nrecv == k);
            if (ok && t > 0) {
                // ---- M[a][b] = G_sys[E_b][R_a], augmented with the identity; Gauss-Jordan.  Warp w owns rows w, w + 8, ...,
                //      a lane the columns lane, lane + 32, ...; pivot rows stay unscaled until the end (a row is divided by its
                //      diagonal once, afterwards): two barriers per column -----------------
                for (int a = warp; a < t; a += NWARP)
                    for (int c = lane; c < 2 * t; c += 32)
                        aug[a * AS + c] = (c < t) ? Ps[int(elist[c]) * r_ + (int(plist[a]) - k)] : uint8_t(c - t == a);
                __syncthreads();
                for (int col = 0; col < t; col++) {
                    if (warp == 0) {
                        int pv = -1;
                        for (int a0 = col; a0 < t && pv < 0; a0 += 32) {
                            const int a = a0 + lane;
                            const unsigned nz = __ballot_sync(0xFFFFFFFFu, a < t && aug[a * AS + col] != 0);
                            if (nz) pv = a0 + __ffs(nz) - 1;
                        }
                        if (lane == 0) s_piv = pv;   // an MDS code always has one
                    }
                    __syncthreads();
                    const int pv = s_piv;
                    if (pv < 0) { ok = false; break; }
                    if (pv != col) {
                        for (int c = tid; c < 2 * t; c += kRsThreads) {
                            const uint8_t x = aug[col * AS + c];
                            aug[col * AS + c] = aug[pv * AS + c];
                            aug[pv * AS + c] = x;
                        }
                        __syncthreads();
                    }
                    const uint32_t lpinv = (255u - lg2[aug[col * AS + col]]) % 255u;       // log of 1 / pivot
                    uint32_t lgp[AS / 32];                                                  // logs of my columns of the pivot row
#pragma unroll
                    for (int q = 0; q < AS / 32; q++) lgp[q] = (lane + 32 * q < 2 * t) ? lg2[aug[col * AS + lane + 32 * q]] : 511u;
                    for (int a = warp; a < t; a += NWARP) {
                        if (a == col) continue;
                        const uint32_t f = aug[a * AS + col];
                        __syncwarp();                                    // every lane has f before lane col % 32 clears it
                        if (f) {
                            const uint32_t lgf = (lg2[f] + lpinv) % 255u;                  // log of f / pivot
                            uint8_t cur[AS / 32], add[AS / 32];
#pragma unroll
                            for (int q = 0; q < AS / 32; q++)
                                if (32 * q < 2 * t) { cur[q] = aug[a * AS + lane + 32 * q]; add[q] = al2[lgf + lgp[q]]; }
#pragma unroll
                            for (int q = 0; q < AS / 32; q++)
                                if (32 * q < 2 * t && lane + 32 * q < 2 * t) aug[a * AS + lane + 32 * q] = cur[q] ^ add[q];
                        }
                    }
                    __syncthreads();
                }
                if (ok) {
                    // the inverse: right half of every row over its diagonal
                    for (int a = warp; a < t; a += NWARP) {
                        const uint32_t linv = (255u - lg2[aug[a * AS + a]]) % 255u;
                        for (int c = t + lane; c < 2 * t; c += 32) aug[a * AS + c] = al2[linv + lg2[aug[a * AS + c]]];
                    }
                    __syncthreads();
                    // ---- decode matrix over the used received symbols (rlist order).  First the logs of the entries of P it
                    //      needs, lgpt[ri][a] = log P[pos_ri][R_a] (the second multiples table is free until the payload part), then
                    //      warp = output row, lane = two inputs, four products per 8-byte load ----
                    uint16_t *lgpt = reinterpret_cast<uint16_t *>(tab1);
                    const int t4 = (t + 3) & ~3;
                    const int LS = t4 + 4;                                // (k - t <= 255 - 2t rows of t4 + 4 halfwords: at most 18 KB)
                    const int nsys = k - t;                               // received systematic symbols come first in rlist
                    for (int i = tid; i < nsys * (t4 / 4); i += kRsThreads) {
                        const int ri = i / (t4 / 4), a0 = (i % (t4 / 4)) * 4;
                        const uint8_t *prow = Ps + int(rlist[ri]) * r_;
                        uint16_t v[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) v[q] = (a0 + q < t) ? lg2[prow[int(plist[a0 + q]) - k]] : uint16_t(511);
                        *reinterpret_cast<uint2 *>(lgpt + ri * LS + a0) = make_uint2(uint32_t(v[0]) | (uint32_t(v[1]) << 16), uint32_t(v[2]) | (uint32_t(v[3]) << 16));
                    }
                    __syncthreads();
                    for (int b = warp; b < t; b += NWARP) {
                        uint32_t lgm[4];                                  // logs of row b of the inverse, spread over the lanes (t <= 128)
#pragma unroll
                        for (int q = 0; q < 4; q++) lgm[q] = (lane + 32 * q < t) ? lg2[aug[b * AS + t + lane + 32 * q]] : 511u;
                        for (int ri0 = 0; ri0 < nsys; ri0 += 64) {
                            const int ri_a = min(ri0 + lane, nsys - 1), ri_b = min(ri0 + 32 + lane, nsys - 1);
                            uint32_t d_a = 0, d_b = 0;
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                if (32 * q < t) {
                                    const int lim = min(32, t4 - 32 * q);
                                    for (int a = 0; a < lim; a += 4) {
                                        const uint2 pa = *reinterpret_cast<const uint2 *>(lgpt + ri_a * LS + 32 * q + a);
                                        const uint2 pb = *reinterpret_cast<const uint2 *>(lgpt + ri_b * LS + 32 * q + a);
                                        const uint32_t m0 = __shfl_sync(0xFFFFFFFFu, lgm[q], a), m1 = __shfl_sync(0xFFFFFFFFu, lgm[q], a + 1);
                                        const uint32_t m2 = __shfl_sync(0xFFFFFFFFu, lgm[q], a + 2), m3 = __shfl_sync(0xFFFFFFFFu, lgm[q], a + 3);
                                        d_a ^= uint32_t(al2[m0 + (pa.x & 0xFFFFu)]) ^ uint32_t(al2[m1 + (pa.x >> 16)]) ^ uint32_t(al2[m2 + (pa.y & 0xFFFFu)]) ^ uint32_t(al2[m3 + (pa.y >> 16)]);
                                        d_b ^= uint32_t(al2[m0 + (pb.x & 0xFFFFu)]) ^ uint32_t(al2[m1 + (pb.x >> 16)]) ^ uint32_t(al2[m2 + (pb.y & 0xFFFFu)]) ^ uint32_t(al2[m3 + (pb.y >> 16)]);
                                    }
                                }
                            }
                            if (ri0 + lane < nsys) D[b * k8 + ri0 + lane] = uint8_t(d_a);
                            if (ri0 + 32 + lane < nsys) D[b * k8 + ri0 + 32 + lane] = uint8_t(d_b);
                        }
                        for (int a = lane; a < t; a += 32) D[b * k8 + nsys + a] = aug[b * AS + t + a];     // the used repair symbols
                    }
                }
                __syncthreads();
            }
        }
        // ---- payload: thread (u, oq) owns the 16-byte column u of every symbol and the outputs b = oq, oq + 4, ...
        //      (kRsOwn of them per pass).  Per input symbol the CTA first builds, for each of its 64 columns, the 16 multiples
        //      v*x of the low nibble and the 16 multiples (16v)*x of the high one (thread (u, q) makes eight of the 32 entries
        //      from the doublings of x); every output then takes c*x = lo[c & 15] ^ hi[c >> 4] with two 128-bit shared loads
        //      instead of a bit-sliced product (32 LOP3 per 16 bytes).  The tables live where the pattern part kept its matrix. ---
        const int in_rows = p.encode ? k : n;
        const int out_rows = p.encode ? n : k;
        const uint4 *in = reinterpret_cast<const uint4 *>(p.in + size_t(cw) * in_rows * S);
        uint4 *out = reinterpret_cast<uint4 *>(p.out + size_t(cw) * out_rows * S);
        const int WS = S / 16;                       // 16-byte columns per symbol
        const int nt = (ok && t > 0) ? t : 0;
        constexpr int CPC = kRsThreads / 4;          // columns a CTA works on at a time
        const int oq = tid / CPC;                    // output quarter = table part this thread builds (warp uniform)
        const int ul = tid % CPC;
        uint4 *tabs0 = reinterpret_cast<uint4 *>(aug), *tabs1 = reinterpret_cast<uint4 *>(tab1);   // [32][CPC] each
        for (int u0 = 0; u0 < WS; u0 += CPC) {
            const int u = u0 + ul;
            const bool live = u < WS;
            if (!p.encode && !ok) {
                // undecodable: pass the received systematic symbols through, erased ones as zero
                const uint32_t *mk = p.mask + cw * p.NW;
                if (live)
                    for (int j = oq; j < k; j += 4)
                        out[size_t(j) * WS + u] = ((mk[j >> 5] >> (j & 31)) & 1u) ? make_uint4(0u, 0u, 0u, 0u) : in[size_t(j) * WS + u];
                continue;
            }
            for (int t0 = 0; t0 < (nt > 0 ? nt : 1); t0 += 4 * kRsOwn) {   // 4 * kRsOwn outputs per pass over the inputs
                uint4 acc[kRsOwn];
#pragma unroll
                for (int i = 0; i < kRsOwn; i++) acc[i] = make_uint4(0u, 0u, 0u, 0u);
                uint4 xn = make_uint4(0u, 0u, 0u, 0u);
                if (live) xn = in[size_t(rlist[0]) * WS + u];
                for (int ri = 0; ri < k; ri++) {
                    const int pos = rlist[ri];
                    const uint4 xv = xn;
                    if (live && ri + 1 < k) xn = in[size_t(rlist[ri + 1]) * WS + u];      // the next input, a step ahead
                    if (live && pos < k && t0 == 0 && oq == 0) out[size_t(pos) * WS + u] = xv;   // systematic symbols pass through
                    if (nt == 0) continue;
                    uint4 *tabs = (ri & 1) ? tabs1 : tabs0;
                    {   // my eight table entries: part oq = {lo 0-7, lo 8-15, hi 0-7, hi 8-15}
                        uint32_t b[4][4];            // base, 2*base, 4*base, 8*base (base = x or 16*x)
                        b[0][0] = xv.x; b[0][1] = xv.y; b[0][2] = xv.z; b[0][3] = xv.w;
                        if (oq >= 2) {
#pragma unroll
                            for (int d = 0; d < 4; d++)
#pragma unroll
                                for (int q = 0; q < 4; q++) b[0][q] = gf_xtime4(b[0][q]);
                        }
#pragma unroll
                        for (int j = 1; j < 4; j++)
#pragma unroll
                            for (int q = 0; q < 4; q++) b[j][q] = gf_xtime4(b[j - 1][q]);
                        const uint32_t hi_on = (oq & 1) ? 0xFFFFFFFFu : 0u;
                        uint4 *dstt = tabs + size_t((oq >> 1) * 16 + (oq & 1) * 8) * CPC + ul;
#pragma unroll
                        for (int v = 0; v < 8; v++) {
                            uint32_t e[4];
#pragma unroll
                            for (int q = 0; q < 4; q++)
                                e[q] = ((v & 1) ? b[0][q] : 0u) ^ ((v & 2) ? b[1][q] : 0u) ^ ((v & 4) ? b[2][q] : 0u) ^ (b[3][q] & hi_on);
                            dstt[size_t(v) * CPC] = make_uint4(e[0], e[1], e[2], e[3]);
                        }
                    }
                    __syncthreads();
                    const uint4 *lo = tabs + ul, *hi = tabs + size_t(16) * CPC + ul;
#pragma unroll
                    for (int g4 = 0; g4 < kRsOwn / 4; g4++) {
                        if (t0 + oq + 16 * g4 < nt) {
#pragma unroll
                            for (int ii = 0; ii < 4; ii++) {
                                const int i = g4 * 4 + ii;
                                const uint32_t c = D[(t0 + oq + 4 * i) * k8 + ri];   // rows >= nt are never read back
                                const uint4 a = lo[size_t(c & 15u) * CPC];
                                const uint4 h = hi[size_t(c >> 4) * CPC];
                                acc[i].x ^= a.x ^ h.x; acc[i].y ^= a.y ^ h.y; acc[i].z ^= a.z ^ h.z; acc[i].w ^= a.w ^ h.w;
                            }
                        }
                    }
                    // (no barrier here: the next symbol's multiples go to the other table; the one after that is built
                    //  behind the next barrier, which nobody passes before finishing these loads)
                }
                __syncthreads();                    // the tables are reused by the next pass / column block / the pattern part
                if (live) {
#pragma unroll
                    for (int i = 0; i < kRsOwn; i++) {
                        const int b = t0 + oq + 4 * i;
                        if (b < nt) {
                            const int row = p.encode ? (k + b) : int(elist[b]);
                            out[size_t(row) * WS + u] = acc[i];
                        }
                    }
                }
            }
        }
        if (!p.encode && tid == 0 && p.fail) p.fail[cw] = ok ? 0 : 1;
        __syncthreads();
    }
}

// ---- host side ---------------------------------------------------------------------------------
inline void rs_host_tables(uint8_t *lg, uint8_t *al)
{
    unsigned x = 1;
    for (int i = 0; i < 255; i++) {
        al[i] = uint8_t(x);
        lg[x] = uint8_t(i);
        x <<= 1;
        if (x & 0x100) x ^= 0x171;
    }
    for (int i = 255; i < 512; i++) al[i] = al[i - 255];
    lg[0] = 0;
}

inline int rs_create_impl(rs_ctx **out, int n, int k, int S, int device, int64_t max_batch, std::string &err)
{
    if (!out) { err = "out is NULL"; return LDPC_ERR_ARG; }
    *out = nullptr;
    if (n < 2 || n > 255 || k < 1 || k >= n) { err = "RS needs 1 <= k < n <= 255"; return LDPC_ERR_ARG; }
    if (n - k > kRsMaxT) { err = "RS: n - k > 128 is not supported"; return LDPC_ERR_UNSUPPORTED; }
    if (S <= 0 || S % 16) { err = "symbol_bytes must be a positive multiple of 16"; return LDPC_ERR_ARG; }
    if (max_batch <= 0) { err = "max_batch must be positive"; return LDPC_ERR_ARG; }
    uint8_t lg[256], al[512];
    rs_host_tables(lg, al);
    auto mul = [&](uint8_t a, uint8_t b) -> uint8_t { return (a && b) ? al[int(lg[a]) + int(lg[b])] : uint8_t(0); };
    // G[i][j] = alpha^((i+1)(j+1)), systematised by Gauss-Jordan on the leading k x k block
    std::vector<uint8_t> M(size_t(k) * n);
    for (int r = 0; r < k; r++)
        for (int c = 0; c < n; c++) M[size_t(r) * n + c] = al[((r + 1) * (c + 1)) % 255];
    for (int col = 0; col < k; col++) {
        int piv = -1;
        for (int r = col; r < k; r++) if (M[size_t(r) * n + col]) { piv = r; break; }
        if (piv < 0) { err = "RS generator: singular leading block"; return LDPC_ERR_ARG; }
        if (piv != col) for (int c = 0; c < n; c++) std::swap(M[size_t(col) * n + c], M[size_t(piv) * n + c]);
        const uint8_t iv = al[255 - lg[M[size_t(col) * n + col]]];
        for (int c = 0; c < n; c++) M[size_t(col) * n + c] = mul(iv, M[size_t(col) * n + c]);
        for (int r = 0; r < k; r++) {
            const uint8_t f = M[size_t(r) * n + col];
            if (r == col || !f) continue;
            for (int c = 0; c < n; c++) M[size_t(r) * n + c] ^= mul(f, M[size_t(col) * n + c]);
        }
    }
    rs_ctx *c = new (std::nothrow) rs_ctx();
    if (!c) { err = "out of host memory"; return LDPC_ERR_NOMEM; }
    c->n = n; c->k = k; c->S = S; c->device = device; c->max_batch = max_batch; c->gsys = M;
    auto bad = [&](const char *what, cudaError_t e) {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        cudaFree(c->d_P); cudaFree(c->d_tab);
        delete c;
        return e == cudaErrorMemoryAllocation ? LDPC_ERR_NOMEM : LDPC_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bad("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bad("cudaGetDeviceProperties", e);
    if (prop.major != 10) { err = "libldpc_cuda is built for sm_100a (B200) only"; delete c; return LDPC_ERR_UNSUPPORTED; }
    c->num_sms = prop.multiProcessorCount;
    const int r_ = n - k;
    std::vector<uint8_t> P(size_t(k) * r_);
    for (int i = 0; i < k; i++) for (int b = 0; b < r_; b++) P[size_t(i) * r_ + b] = M[size_t(i) * n + k + b];
    std::vector<uint8_t> tab(768);
    memcpy(tab.data(), lg, 256); memcpy(tab.data() + 256, al, 512);
    if ((e = cudaMalloc(&c->d_P, P.size())) != cudaSuccess) return bad("cudaMalloc", e);
    if ((e = cudaMalloc(&c->d_tab, tab.size())) != cudaSuccess) return bad("cudaMalloc", e);
    if ((e = cudaMemcpy(c->d_P, P.data(), P.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return bad("cudaMemcpy", e);
    if ((e = cudaMemcpy(c->d_tab, tab.data(), tab.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return bad("cudaMemcpy", e);
    const int k8 = (k + 7) & ~7;
    c->smem = kRsMaxT * k8 + ((k * r_ + 15) & ~15) + 1536 + kRsMaxT * 2 * kRsMaxT + kRsTabBytes + 256 + 2 * kRsMaxT + 64;
    // the attribute is per-function process state: always raise it to the device maximum, never to this context's size
    if ((e = cudaFuncSetAttribute(rs_codec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  int(prop.sharedMemPerBlockOptin) - 1024)) != cudaSuccess)
        return bad("cudaFuncSetAttribute", e);
    *out = c;
    return LDPC_OK;
}

inline int rs_destroy_impl(rs_ctx *c)
{
    if (!c) return LDPC_OK;
    cudaSetDevice(c->device);
    cudaFree(c->d_P); cudaFree(c->d_tab);
    delete c;
    return LDPC_OK;
}

inline int rs_get_generator_impl(const rs_ctx *c, uint8_t *gsys, std::string &err)
{
    if (!c || !gsys) { err = "NULL argument"; return LDPC_ERR_ARG; }
    memcpy(gsys, c->gsys.data(), c->gsys.size());
    return LDPC_OK;
}

inline int rs_launch(rs_ctx *c, const void *in, void *outp, const uint32_t *mask, uint8_t *failp, int64_t B, int encode,
                     cudaStream_t st, std::string &err)
{
    if (!c || B < 0) { err = "bad argument"; return LDPC_ERR_ARG; }
    if (B == 0) return LDPC_OK;
    if (!in || !outp || (!encode && !mask)) { err = "NULL buffer"; return LDPC_ERR_ARG; }
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) { err = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return LDPC_ERR_CUDA; }
    RsParams p;
    p.in = static_cast<const uint8_t *>(in); p.out = static_cast<uint8_t *>(outp); p.mask = mask; p.fail = failp;
    p.P = c->d_P; p.tab = c->d_tab; p.B = B; p.n = c->n; p.k = c->k; p.S = c->S; p.NW = (c->n + 31) / 32;
    p.encode = encode;
    const int per_sm = std::max(1, std::min(8, (227 * 1024) / (c->smem + 1024)));
    const int grid = int(std::min<long long>(B, (long long)c->num_sms * per_sm));
    rs_codec_kernel<<<grid, kRsThreads, c->smem, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("rs_codec_kernel: ") + cudaGetErrorString(e); return LDPC_ERR_CUDA; }
    return LDPC_OK;
}

inline int rs_encode_impl(rs_ctx *c, const void *d_info, void *d_cw, int64_t B, cudaStream_t st, std::string &err)
{
    return rs_launch(c, d_info, d_cw, nullptr, nullptr, B, 1, st, err);
}

inline int rs_decode_impl(rs_ctx *c, const void *d_cw, const uint32_t *d_mask, void *d_out, uint8_t *d_fail, int64_t B,
                          cudaStream_t st, std::string &err)
{
    return rs_launch(c, d_cw, d_out, d_mask, d_fail, B, 0, st, err);
}

}  // namespace ldpc
