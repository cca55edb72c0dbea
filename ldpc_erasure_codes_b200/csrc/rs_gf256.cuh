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
//   payload part (what the time goes to, ~ t*k*S multiply-accumulates): a thread owns one 16-byte
//     column (four 32-bit words) of all symbols and every fourth output.  Multiplication by a constant c is bit-sliced over the
//     CONSTANT: x*c = XOR_j bit_j(c) * (x * 2^j); the eight doublings x*2^j of a packed word cost 5 ALU
//     ops each and are shared by all t outputs, and bit_j(c) is applied as a precomputed 32-bit mask
//     (table of 8 masks per constant in shared memory, read with two broadcast 128-bit loads), so one
//     4-byte multiply-accumulate is 8 LOP3 -- no per-byte table lookups.
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
    uint32_t *d_m8 = nullptr;    // [256][8] bit masks of every constant
    int smem = 0;
};

namespace ldpc {

constexpr int kRsThreads = 256;
constexpr int kRsMaxT = 128;     // n - k <= 128 (matrix dimensions)
constexpr int kRsOwn = 16;       // outputs a thread accumulates per pass over the inputs (4 * kRsOwn per CTA pass)

struct RsParams {
    const uint8_t *in;      // decode: [B][n][S] received codewords; encode: [B][k][S] info
    uint8_t *out;           // decode: [B][k][S]; encode: [B][n][S]
    const uint32_t *mask;   // decode: [B][NW]
    uint8_t *fail;          // decode: [B] or nullptr
    const uint8_t *P;       // [k][n-k]
    const uint8_t *tab;     // log | alog
    const uint32_t *m8;     // [256][8]
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
    // shared layout: m8[256*8 u32] | D[128][k8] | P[k][r_] | log[256] alog[512] | aug[128][256] | rlist[256] elist[128] plist[128]
    const int k8 = (k + 7) & ~7;
    uint32_t *m8 = reinterpret_cast<uint32_t *>(rs_smem);
    uint8_t *D = reinterpret_cast<uint8_t *>(m8 + 256 * 8);
    uint8_t *Ps = D + kRsMaxT * k8;
    uint8_t *lg = Ps + ((k * r_ + 15) & ~15);
    uint8_t *al = lg + 256;
    uint8_t *aug = al + 512;
    uint8_t *rlist = aug + kRsMaxT * 2 * kRsMaxT;
    uint8_t *elist = rlist + 256;
    uint8_t *plist = elist + kRsMaxT;
    __shared__ int s_t, s_nrecv, s_piv;
    const int tid = threadIdx.x;

    for (int i = tid; i < 256 * 8; i += kRsThreads) m8[i] = p.m8[i];
    for (int i = tid; i < k * r_; i += kRsThreads) Ps[i] = p.P[i];
    for (int i = tid; i < 768; i += kRsThreads) lg[i] = p.tab[i];
    if (p.encode) {   // static matrix: parity b = sum_i P[i][b] u_i
        for (int i = tid; i < r_ * k; i += kRsThreads) D[(i / k) * k8 + (i % k)] = p.P[(i % k) * r_ + (i / k)];
        for (int i = tid; i < k; i += kRsThreads) rlist[i] = uint8_t(i);
    }
    __syncthreads();
    auto mul = [&](uint8_t a, uint8_t b) -> uint8_t { return (a && b) ? al[int(lg[a]) + int(lg[b])] : uint8_t(0); };

    for (long long cw = blockIdx.x; cw < p.B; cw += gridDim.x) {
        int t = r_;
        bool ok = true;
        if (!p.encode) {
            // ---- which symbols are used: first k received; E = erased systematic; R = first t received repair ---
            if (tid == 0) {
                const uint32_t *mk = p.mask + cw * p.NW;
                int nrecv = 0, te = 0;
                for (int j = 0; j < n; j++) {
                    const bool er = (mk[j >> 5] >> (j & 31)) & 1u;
                    if (er) { if (j < k) elist[te++] = uint8_t(j); }
                    else if (nrecv < k) { rlist[nrecv] = uint8_t(j); nrecv++; }
                }
                s_nrecv = nrecv;
                if (nrecv == k) {   // the used repair symbols are the tail of rlist; exactly te of them
                    for (int a = 0; a < te; a++) plist[a] = rlist[k - te + a];
                } else {
                    te = min(te, kRsMaxT);
                }
                s_t = te;
            }
            __syncthreads();
            t = s_t;
            ok = (s_nrecv == k);
            if (ok && t > 0) {
                // ---- M[a][b] = G_sys[E_b][R_a], augmented with the identity; Gauss-Jordan -----------------
                for (int i = tid; i < t * 2 * t; i += kRsThreads) {
                    const int a = i / (2 * t), c = i % (2 * t);
                    aug[a * 2 * kRsMaxT + c] = (c < t) ? Ps[int(elist[c]) * r_ + (int(plist[a]) - k)] : uint8_t(c - t == a);
                }
                __syncthreads();
                for (int col = 0; col < t; col++) {
                    if (tid == 0) {
                        int pv = -1;
                        for (int a = col; a < t && pv < 0; a++) if (aug[a * 2 * kRsMaxT + col]) pv = a;
                        s_piv = pv;   // an MDS code always has one
                    }
                    __syncthreads();
                    const int pv = s_piv;
                    if (pv < 0) { ok = false; break; }
                    if (pv != col) {
                        for (int c = tid; c < 2 * t; c += kRsThreads) {
                            const uint8_t x = aug[col * 2 * kRsMaxT + c];
                            aug[col * 2 * kRsMaxT + c] = aug[pv * 2 * kRsMaxT + c];
                            aug[pv * 2 * kRsMaxT + c] = x;
                        }
                        __syncthreads();
                    }
                    const uint8_t pinv = al[255 - int(lg[aug[col * 2 * kRsMaxT + col]])];
                    __syncthreads();
                    for (int c = tid; c < 2 * t; c += kRsThreads) aug[col * 2 * kRsMaxT + c] = mul(pinv, aug[col * 2 * kRsMaxT + c]);
                    __syncthreads();
                    for (int i = tid; i < t * 2 * t; i += kRsThreads) {
                        const int a = i / (2 * t), c = i % (2 * t);
                        if (a == col) continue;
                        const uint8_t f = aug[a * 2 * kRsMaxT + col];
                        // column `col` of row a is read by every thread of that row before anyone overwrites it:
                        // the element c == col is written last within this step by the barrier below
                        if (f && c != col) aug[a * 2 * kRsMaxT + c] ^= mul(f, aug[col * 2 * kRsMaxT + c]);
                    }
                    __syncthreads();
                    for (int a = tid; a < t; a += kRsThreads) if (a != col) aug[a * 2 * kRsMaxT + col] = 0;
                    __syncthreads();
                }
                // ---- decode matrix over the used received symbols (rlist order) -----------------------------
                if (ok) {
                    for (int i = tid; i < t * k; i += kRsThreads) {
                        const int b = i / k, ri = i % k;
                        const int pos = rlist[ri];
                        uint8_t d = 0;
                        if (pos < k) {
                            for (int a = 0; a < t; a++)
                                d ^= mul(aug[b * 2 * kRsMaxT + t + a], Ps[pos * r_ + (int(plist[a]) - k)]);
                        } else {
                            d = aug[b * 2 * kRsMaxT + t + (ri - (k - t))];
                        }
                        D[b * k8 + ri] = d;
                    }
                }
                __syncthreads();
            }
        }
        // ---- payload: thread (u, oq) owns the 16-byte column u of every symbol and the outputs
        //      b = oq, oq + 4, ... (kRsOwn of them per pass); the masks of a coefficient are shared by
        //      the four words of the column ------------------------------------------------------------
        const int in_rows = p.encode ? k : n;
        const int out_rows = p.encode ? n : k;
        const uint4 *in = reinterpret_cast<const uint4 *>(p.in + size_t(cw) * in_rows * S);
        uint4 *out = reinterpret_cast<uint4 *>(p.out + size_t(cw) * out_rows * S);
        const int WS = S / 16;                       // 16-byte columns per symbol
        const int nt = (ok && t > 0) ? t : 0;
        const int oq = tid / (kRsThreads / 4);       // output quarter (warp uniform)
        for (int u = tid % (kRsThreads / 4); u < WS; u += kRsThreads / 4) {
            if (!p.encode && !ok) {
                // undecodable: pass the received systematic symbols through, erased ones as zero
                const uint32_t *mk = p.mask + cw * p.NW;
                for (int j = oq; j < k; j += 4)
                    out[size_t(j) * WS + u] = ((mk[j >> 5] >> (j & 31)) & 1u) ? make_uint4(0u, 0u, 0u, 0u) : in[size_t(j) * WS + u];
                continue;
            }
            for (int t0 = 0; t0 < (nt > 0 ? nt : 1); t0 += 4 * kRsOwn) {   // 4 * kRsOwn outputs per pass over the inputs
                uint4 acc[kRsOwn];
#pragma unroll
                for (int i = 0; i < kRsOwn; i++) acc[i] = make_uint4(0u, 0u, 0u, 0u);
                for (int ri = 0; ri < k; ri++) {
                    const int pos = rlist[ri];
                    const uint4 xv = in[size_t(pos) * WS + u];
                    if (pos < k && t0 == 0 && oq == 0) out[size_t(pos) * WS + u] = xv;   // systematic symbols pass through
                    if (nt == 0) continue;
                    uint32_t x[8][4];
                    x[0][0] = xv.x; x[0][1] = xv.y; x[0][2] = xv.z; x[0][3] = xv.w;
#pragma unroll
                    for (int j = 1; j < 8; j++)
#pragma unroll
                        for (int q = 0; q < 4; q++) x[j][q] = gf_xtime4(x[j - 1][q]);
#pragma unroll
                    for (int g4 = 0; g4 < kRsOwn / 4; g4++) {
                        if (t0 + oq + 16 * g4 < nt) {
#pragma unroll
                            for (int ii = 0; ii < 4; ii++) {
                                const int i = g4 * 4 + ii;
                                const uint32_t c = D[(t0 + oq + 4 * i) * k8 + ri];   // rows >= nt are never read back
                                const uint4 ma = *reinterpret_cast<const uint4 *>(m8 + c * 8);
                                const uint4 mb = *reinterpret_cast<const uint4 *>(m8 + c * 8 + 4);
                                const uint32_t mm[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
#pragma unroll
                                for (int j = 0; j < 8; j++) {
                                    acc[i].x ^= mm[j] & x[j][0];
                                    acc[i].y ^= mm[j] & x[j][1];
                                    acc[i].z ^= mm[j] & x[j][2];
                                    acc[i].w ^= mm[j] & x[j][3];
                                }
                            }
                        }
                    }
                }
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
        cudaFree(c->d_P); cudaFree(c->d_tab); cudaFree(c->d_m8);
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
    std::vector<uint32_t> m8(256 * 8);
    for (int v = 0; v < 256; v++) for (int j = 0; j < 8; j++) m8[v * 8 + j] = ((v >> j) & 1) ? 0xFFFFFFFFu : 0u;
    if ((e = cudaMalloc(&c->d_P, P.size())) != cudaSuccess) return bad("cudaMalloc", e);
    if ((e = cudaMalloc(&c->d_tab, tab.size())) != cudaSuccess) return bad("cudaMalloc", e);
    if ((e = cudaMalloc(&c->d_m8, m8.size() * 4)) != cudaSuccess) return bad("cudaMalloc", e);
    if ((e = cudaMemcpy(c->d_P, P.data(), P.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return bad("cudaMemcpy", e);
    if ((e = cudaMemcpy(c->d_tab, tab.data(), tab.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return bad("cudaMemcpy", e);
    if ((e = cudaMemcpy(c->d_m8, m8.data(), m8.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return bad("cudaMemcpy", e);
    const int k8 = (k + 7) & ~7;
    c->smem = 256 * 8 * 4 + kRsMaxT * k8 + ((k * r_ + 15) & ~15) + 768 + kRsMaxT * 2 * kRsMaxT + 256 + 2 * kRsMaxT + 64;
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
    cudaFree(c->d_P); cudaFree(c->d_tab); cudaFree(c->d_m8);
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
    p.P = c->d_P; p.tab = c->d_tab; p.m8 = c->d_m8; p.B = B; p.n = c->n; p.k = c->k; p.S = c->S; p.NW = (c->n + 31) / 32;
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
