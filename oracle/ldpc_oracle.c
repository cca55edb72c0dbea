/*
 * ldpc_oracle.c -- CPU restatement of the reference's erasure-codec algorithms.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA
 * product path (libldpc_cuda).  It may be called from tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
 * and from nowhere else.  The product never routes through it.
 *
 * The reference (chadac8j/LDPC_Erasure_Codes) cannot be compiled here: its
 * device code is Intel-FPGA OpenCL (cl_intel_channels, AOCLUtils) and the rest
 * is MATLAB.  So each function below restates one reference loop in plain C
 * and cites the file:line it follows.  What pins it (see tests/test_oracle.py):
 *   - Threefry4x32-20: the public Random123 known-answer vectors
 *   - GF(256): Matlab/GF_256_add_mult_inv_tables.mat (full mul table + inverses)
 *   - H: the three committed .mat codes (re-exported under codes/)
 * No input->output vectors for encode / decode / hybrid / RS are committed in
 * the reference, so for payload bytes of those paths parity is pinned by this
 * restatement plus algebraic uniqueness ("parity unpinned" by reference
 * fixtures for: encoder parity bytes, hybrid-ML outputs, RS parity bytes,
 * bursty patterns) -- DESIGN.md says the same.
 *
 * Conventions: H is passed as CSR (row_ptr[m+1], col_idx[nnz]), 0-based,
 * column indices ascending inside a row -- i.e. the reference's "Vlist" row
 * {w, c_1..c_w} (1-based) with the weight moved into row_ptr.
 * Payload is [n][S] bytes, erasure flags are one byte per symbol.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* Threefry4x32, 20 rounds.  Follows OpenCL/device/threefry.h:299-745        */
/* (macro _threefry4x_tpl): key schedule ks[4]=parity^k0^k1^k2^k3 (:175,     */
/* :304-320), rotation table R_32x4 (:104-117), key injection after every    */
/* 4th round with the round-group number added to word 3 (:343-347).         */
/* ------------------------------------------------------------------------- */
static inline uint32_t rotl32(uint32_t x, unsigned r) { return (x << r) | (x >> (32 - r)); }

void orc_threefry4x32_20(const uint32_t ctr[4], const uint32_t key[4], uint32_t out[4])
{
    static const unsigned char R[8][2] = {
        {10, 26}, {11, 21}, {13, 27}, {23, 5}, {6, 20}, {17, 11}, {25, 10}, {18, 20}};
    uint32_t ks[5];
    uint32_t X[4];
    ks[4] = 0x1BD11BDAu;
    for (int i = 0; i < 4; i++) {
        ks[i] = key[i];
        X[i] = ctr[i] + key[i];
        ks[4] ^= key[i];
    }
    for (unsigned r = 0; r < 20; r++) {
        if ((r & 1) == 0) {
            X[0] += X[1]; X[1] = rotl32(X[1], R[r & 7][0]); X[1] ^= X[0];
            X[2] += X[3]; X[3] = rotl32(X[3], R[r & 7][1]); X[3] ^= X[2];
        } else {
            X[0] += X[3]; X[3] = rotl32(X[3], R[r & 7][0]); X[3] ^= X[0];
            X[2] += X[1]; X[1] = rotl32(X[1], R[r & 7][1]); X[1] ^= X[2];
        }
        if ((r & 3) == 3) {
            unsigned s = (r + 1) >> 2;
            for (int i = 0; i < 4; i++) X[i] += ks[(s + i) % 5];
            X[3] += s;
        }
    }
    for (int i = 0; i < 4; i++) out[i] = X[i];
}

/* ------------------------------------------------------------------------- */
/* i.i.d. erasure generator.  Follows `data_in` in                            */
/* OpenCL/device/ldpc_erasure_decoder_top.cl:68-117: key = {tid=1, seed},     */
/* counter word 0 pre-incremented from 0 and never reset between frames       */
/* (:75,:96) => i = 1 + frame*n + sym (uint32, wraps); erased iff             */
/* ((int)out.v[0] & 0x3F) < PER_numerator_div_64 (:105).                      */
/* mode 0 = that rule (threshold = P in 0..64);                               */
/* mode 1 = EXTENSION for rates that are not a multiple of 1/64:              */
/*          erased iff out.v[0] < threshold (threshold = floor(p*2^32)).      */
/* ------------------------------------------------------------------------- */
void orc_gen_erasures_iid(int n, uint32_t seed, int mode, uint32_t threshold,
                          uint64_t frame0, int64_t nframes, uint8_t *flags)
{
    const uint32_t key[4] = {1u, seed, 0u, 0u};
    for (int64_t f = 0; f < nframes; f++) {
        for (int s = 0; s < n; s++) {
            uint32_t ctr[4] = {(uint32_t)(1u + (uint32_t)((frame0 + (uint64_t)f) * (uint64_t)n) + (uint32_t)s), 0, 0, 0};
            uint32_t o[4];
            orc_threefry4x32_20(ctr, key, o);
            int e;
            if (mode == 0) e = ((int)(o[0] & 0x3Fu) < (int)threshold);
            else e = (o[0] < threshold);
            flags[f * (int64_t)n + s] = (uint8_t)e;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Bursty two-state channel.  Follows                                         */
/* Matlab/Bursty_Error_Channel_Model_Generator.m:12-47: per symbol two        */
/* uniforms in this order (:25-26); state 0: erase iff u1<=alpha (:28), go to */
/* 1 iff u2<=0.1/bias (:16,:32); state 1: erase iff u1<=beta (:38), go to 0   */
/* iff u2<=0.1 (:17,:42).  State persists across symbols and codewords        */
/* (Matlab/ErasureCodes_NonBinaryLDPCSim.m:163,192), initial state 0.         */
/* MATLAB's rand stream is not reproducible, so (EXTENSION, shared with the   */
/* CUDA path) u1 = v[0]/2^32, u2 = v[1]/2^32 of the same Threefry counter     */
/* scheme as the i.i.d. generator.  `state_io` carries the state in and out.  */
/* ------------------------------------------------------------------------- */
void orc_gen_erasures_bursty(int n, uint32_t seed, double alpha, double beta, double bias,
                             uint64_t frame0, int64_t nframes, int *state_io, uint8_t *flags)
{
    const uint32_t key[4] = {1u, seed, 0u, 0u};
    const double transition = 0.1;
    const double p01 = transition / bias;
    const double p10 = transition;
    int state = *state_io;
    for (int64_t f = 0; f < nframes; f++) {
        for (int s = 0; s < n; s++) {
            uint32_t ctr[4] = {(uint32_t)(1u + (uint32_t)((frame0 + (uint64_t)f) * (uint64_t)n) + (uint32_t)s), 0, 0, 0};
            uint32_t o[4];
            orc_threefry4x32_20(ctr, key, o);
            double u1 = (double)o[0] / 4294967296.0;
            double u2 = (double)o[1] / 4294967296.0;
            int err = 0;
            if (state == 0) {
                if (u1 <= alpha) err = 1;
                if (u2 <= p01) state = 1;
            } else {
                if (u1 <= beta) err = 1;
                if (u2 <= p10) state = 0;
            }
            flags[f * (int64_t)n + s] = (uint8_t)err;
        }
    }
    *state_io = state;
}

/* ------------------------------------------------------------------------- */
/* Systematic encoder.  Follows OpenCL/device/ldpc_erasure_encoder.cl:50-93   */
/* (bit-level twin Matlab/LDPCErasureCodes_MessagePassingAlgSim.m:164-168):   */
/* symbols 0..k-1 pass through (:62-71); parity row r = XOR of the first      */
/* w_r - 1 row members, i.e. every member but the last = the diagonal         */
/* (:72-83), rows in order because rows reference earlier parities.           */
/* ------------------------------------------------------------------------- */
void orc_ldpc_encode(int n, int k, const int32_t *row_ptr, const int32_t *col_idx, int S,
                     const uint8_t *info, uint8_t *cw)
{
    memcpy(cw, info, (size_t)k * S);
    for (int r = 0; r < n - k; r++) {
        uint8_t *acc = cw + (size_t)(k + r) * S;
        memset(acc, 0, (size_t)S);
        for (int j = row_ptr[r]; j < row_ptr[r + 1] - 1; j++) {
            const uint8_t *src = cw + (size_t)col_idx[j] * S;
            for (int l = 0; l < S; l++) acc[l] ^= src[l];
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Canonical peeling decoder.  Follows                                        */
/* OpenCL/device/ldpc_erasure_decoder.cl:49-93 (bit twin                      */
/* Matlab/My_LDPC_Erasure_Decoder.m:18-47): num_iter serial sweeps over the   */
/* checks in order; per check XOR the payload of ALL members (erased ones are */
/* zero by the stated assumption :17-20), count erasures, remember the last   */
/* erased index (:76-80); exactly one => write the accumulator there and      */
/* clear its flag, in place, visible to later rows of the same sweep          */
/* (:82-90).  early_stop != 0 adds the output-neutral stop of                 */
/* ldpc_erasure_decoder_old.pro:116-123 / My_LDPC_Erasure_Decoder.m:39-42     */
/* (leave the loop once no erasure is left).  Returns the number of sweeps    */
/* run.  The caller reads the first k payloads (:97-102) and derives          */
/* fail_sys = any(erased[0..k)) (ldpc_erasure_decoder_perf_tests.cl:215-228). */
/* ------------------------------------------------------------------------- */
int orc_ldpc_peel(int n, int k, const int32_t *row_ptr, const int32_t *col_idx, int S,
                  uint8_t *payload, uint8_t *erased, int max_iter, int early_stop)
{
    const int m = n - k;
    uint8_t *acc = (uint8_t *)malloc((size_t)S);
    int remaining = 0;
    for (int i = 0; i < n; i++) remaining += erased[i] ? 1 : 0;
    int it = 0;
    while (it < max_iter) {
        if (early_stop && remaining == 0) break;
        for (int c = 0; c < m; c++) {
            int num_erasures = 0;
            int erasure_ind = 0;
            memset(acc, 0, (size_t)S);
            for (int j = row_ptr[c]; j < row_ptr[c + 1]; j++) {
                const int u = col_idx[j];
                const uint8_t *src = payload + (size_t)u * S;
                for (int l = 0; l < S; l++) acc[l] ^= src[l];
                if (erased[u] == 1) {
                    num_erasures++;
                    erasure_ind = u;
                }
            }
            if (num_erasures == 1) {
                erased[erasure_ind] = 0;
                memcpy(payload + (size_t)erasure_ind * S, acc, (size_t)S);
                remaining--;
            }
        }
        it++;
    }
    free(acc);
    return it;
}

/* Same decoder, 64-bit XOR lanes, S % 8 == 0: the timed CPU baseline.         */
/* Identical results to orc_ldpc_peel (tests check it); it exists only so that */
/* the CPU number is not handicapped by a byte loop.                           */
int orc_ldpc_peel_u64(int n, int k, const int32_t *row_ptr, const int32_t *col_idx, int S,
                      uint8_t *payload, uint8_t *erased, int max_iter, int early_stop)
{
    const int m = n - k;
    const int W = S / 8;
    uint64_t *P = (uint64_t *)payload;
    uint64_t acc[W > 0 ? W : 1];
    int remaining = 0;
    for (int i = 0; i < n; i++) remaining += erased[i] ? 1 : 0;
    int it = 0;
    while (it < max_iter) {
        if (early_stop && remaining == 0) break;
        for (int c = 0; c < m; c++) {
            int num_erasures = 0;
            int erasure_ind = 0;
            for (int l = 0; l < W; l++) acc[l] = 0;
            for (int j = row_ptr[c]; j < row_ptr[c + 1]; j++) {
                const int u = col_idx[j];
                const uint64_t *src = P + (size_t)u * W;
                for (int l = 0; l < W; l++) acc[l] ^= src[l];
                if (erased[u] == 1) {
                    num_erasures++;
                    erasure_ind = u;
                }
            }
            if (num_erasures == 1) {
                erased[erasure_ind] = 0;
                uint64_t *dst = P + (size_t)erasure_ind * W;
                for (int l = 0; l < W; l++) dst[l] = acc[l];
                remaining--;
            }
        }
        it++;
    }
    return it;
}

/* ------------------------------------------------------------------------- */
/* Hybrid-ML decoder.  Follows Matlab/My_LDPC_HybridML_Erasure_Decoder.m:     */
/*  (1) up to `peel_iter` (=10, :9) serial sweeps with the "no erasure left"  */
/*      stop (:17-46);                                                        */
/*  (2) E = erased set ascending (:50), A = H(:,E) (m x e, :52),              */
/*      rhs = H(:,known) * y(known) per payload bit (:54);                    */
/*  (3) for col = 1..e: rows >= col with a 1 in this column (:58); none =>    */
/*      dont_do_jordan, break (:59-62); swap the first such row to `col`,     */
/*      A and rhs (:63-69); XOR the pivot row into the other listed rows      */
/*      (:71-74);                                                             */
/*  (4) if not aborted: col = e..2, clear the column above the diagonal       */
/*      (:77-86);                                                             */
/*  (5) y(E) = rhs(1:e) unconditionally (:87).                                */
/* `abort_writeback` = 1 reproduces (5) literally on abort (garbage, and for  */
/* e > m MATLAB would raise an index error -- here e > m is treated as an     */
/* abort before elimination).  0 = the contract the CUDA path implements      */
/* (SURVEY.md section 8 a-10): on abort leave the peeling state untouched.    */
/* Returns 0 = clean after peeling, 1 = GE ran and succeeded, 2 = GE aborted  */
/* (ml_fail).  *n_rowops receives the number of row XORs GE performed.        */
/* ------------------------------------------------------------------------- */
int orc_ldpc_hybrid(int n, int k, const int32_t *row_ptr, const int32_t *col_idx, int S,
                    uint8_t *payload, uint8_t *erased, int peel_iter, int abort_writeback,
                    int *n_rowops)
{
    const int m = n - k;
    if (n_rowops) *n_rowops = 0;
    orc_ldpc_peel(n, k, row_ptr, col_idx, S, payload, erased, peel_iter, 1);
    int e = 0;
    for (int i = 0; i < n; i++) e += erased[i] ? 1 : 0;
    if (e == 0) return 0;

    int *E = (int *)malloc(sizeof(int) * (size_t)e);
    int *pos = (int *)malloc(sizeof(int) * (size_t)n); /* column -> index in E, or -1 */
    for (int i = 0, j = 0; i < n; i++) {
        pos[i] = -1;
        if (erased[i]) { pos[i] = j; E[j++] = i; }
    }
    /* dense bit matrix A (m x e), one byte per bit: this is a restatement, not a fast path */
    uint8_t *A = (uint8_t *)calloc((size_t)m * (size_t)e, 1);
    uint8_t *rhs = (uint8_t *)calloc((size_t)m * (size_t)S, 1);
    for (int c = 0; c < m; c++) {
        for (int j = row_ptr[c]; j < row_ptr[c + 1]; j++) {
            const int u = col_idx[j];
            if (pos[u] >= 0) {
                A[(size_t)c * e + pos[u]] = 1;
            } else {
                const uint8_t *src = payload + (size_t)u * S;
                uint8_t *dst = rhs + (size_t)c * S;
                for (int l = 0; l < S; l++) dst[l] ^= src[l];
            }
        }
    }
    int rowops = 0;
    int dont_do_jordan = 0;
    uint8_t *tmp_row = (uint8_t *)malloc((size_t)(e > S ? e : S));
    if (e > m) {
        dont_do_jordan = 1;
    } else {
        for (int col = 0; col < e; col++) {
            int first = -1;
            for (int r = col; r < m; r++)
                if (A[(size_t)r * e + col]) { first = r; break; }
            if (first < 0) { dont_do_jordan = 1; break; }
            if (first != col) {
                memcpy(tmp_row, rhs + (size_t)col * S, (size_t)S);
                memcpy(rhs + (size_t)col * S, rhs + (size_t)first * S, (size_t)S);
                memcpy(rhs + (size_t)first * S, tmp_row, (size_t)S);
                memcpy(tmp_row, A + (size_t)col * e, (size_t)e);
                memcpy(A + (size_t)col * e, A + (size_t)first * e, (size_t)e);
                memcpy(A + (size_t)first * e, tmp_row, (size_t)e);
            }
            /* the rows listed after the first one (they kept their places: the swap moved
             * row `col`, which had a 0 or was itself `first`, into `first`) */
            for (int r = first + 1; r < m; r++) {
                if (r == col) continue;
                if (A[(size_t)r * e + col]) {
                    for (int j = 0; j < e; j++) A[(size_t)r * e + j] ^= A[(size_t)col * e + j];
                    for (int l = 0; l < S; l++) rhs[(size_t)r * S + l] ^= rhs[(size_t)col * S + l];
                    rowops++;
                }
            }
        }
        if (!dont_do_jordan) {
            for (int col = e - 1; col >= 1; col--) {
                for (int r = 0; r < col; r++) {
                    if (A[(size_t)r * e + col]) {
                        for (int j = 0; j < e; j++) A[(size_t)r * e + j] ^= A[(size_t)col * e + j];
                        for (int l = 0; l < S; l++) rhs[(size_t)r * S + l] ^= rhs[(size_t)col * S + l];
                        rowops++;
                    }
                }
            }
        }
    }
    int ret;
    if (!dont_do_jordan) {
        for (int j = 0; j < e; j++) {
            memcpy(payload + (size_t)E[j] * S, rhs + (size_t)j * S, (size_t)S);
            erased[E[j]] = 0;
        }
        ret = 1;
    } else {
        if (abort_writeback && e <= m) {
            for (int j = 0; j < e; j++)
                memcpy(payload + (size_t)E[j] * S, rhs + (size_t)j * S, (size_t)S);
        }
        ret = 2;
    }
    if (n_rowops) *n_rowops = rowops;
    free(tmp_row); free(rhs); free(A); free(pos); free(E);
    return ret;
}

/* ------------------------------------------------------------------------- */
/* RS-equivalent MDS counting.  Follows                                       */
/* OpenCL/device/ldpc_erasure_decoder_perf_tests.cl:48-50,70-80 and           */
/* Matlab/LDPCErasureCodes_MessagePassingAlgSim.m:199-205: split the frame    */
/* into n/RS_n consecutive blocks; a block fails iff #erasures > RS_n - RS_k. */
/* Returns the number of failed blocks of this frame.                         */
/* ------------------------------------------------------------------------- */
int orc_rs_mds_count(int n, int rs_n, int rs_k, const uint8_t *erased)
{
    int fails = 0;
    for (int b = 0; b < n / rs_n; b++) {
        int cnt = 0;
        for (int i = 0; i < rs_n; i++) cnt += erased[b * rs_n + i] ? 1 : 0;
        if (cnt > rs_n - rs_k) fails++;
    }
    return fails;
}

/* ------------------------------------------------------------------------- */
/* GF(2^8) tables.  Follows Matlab/Build_GF256_Lookup_Tables.m:7-67:          */
/* primitive-polynomial vector [1 0 1 1 1 0 0 0 1] read MSB-first (:11-14)    */
/* = 0x171, alpha = 2 (:24); antilog by repeated multiplication by alpha      */
/* (:21-32); mul(a,b) = antilog[(log a + log b) mod 255] (:48); inv(a) =      */
/* antilog[255 - log a] (:35-41); add = XOR (:61).                            */
/* ------------------------------------------------------------------------- */
static uint8_t gf_log_t[256];
static uint8_t gf_alog_t[512];
static int gf_ready = 0;

static void gf_init(void)
{
    if (gf_ready) return;
    unsigned x = 1;
    for (int i = 0; i < 255; i++) {
        gf_alog_t[i] = (uint8_t)x;
        gf_log_t[x] = (uint8_t)i;
        x <<= 1;
        if (x & 0x100) x ^= 0x171;
    }
    for (int i = 255; i < 512; i++) gf_alog_t[i] = gf_alog_t[i - 255];
    gf_log_t[0] = 0;
    gf_ready = 1;
}

static inline uint8_t gf_mul(uint8_t a, uint8_t b)
{
    if (a == 0 || b == 0) return 0;
    return gf_alog_t[gf_log_t[a] + gf_log_t[b]];
}

static inline uint8_t gf_inv(uint8_t a) { return gf_alog_t[255 - gf_log_t[a]]; }

void orc_gf256_tables(uint8_t *mul256x256, uint8_t *inv255, uint8_t *log256, uint8_t *alog255)
{
    gf_init();
    if (mul256x256)
        for (int a = 0; a < 256; a++)
            for (int b = 0; b < 256; b++) mul256x256[a * 256 + b] = gf_mul((uint8_t)a, (uint8_t)b);
    if (inv255)
        for (int a = 1; a < 256; a++) inv255[a - 1] = gf_inv((uint8_t)a);
    if (log256) memcpy(log256, gf_log_t, 256);
    if (alog255) memcpy(alog255, gf_alog_t, 255);
}

/* ------------------------------------------------------------------------- */
/* RS systematic generator.  Follows Matlab/Test_My_RS_Decode.m:30-37:        */
/* G(row,col) = alpha^(row*col), row = 1..k, col = 1..n; G_sys =              */
/* inv(G(:,1:k)) * G = [I | P].  Gsys is k x n row-major.  Returns 0, or -1   */
/* if the leading k x k block is singular (never for distinct alpha^j).       */
/* ------------------------------------------------------------------------- */
int orc_rs_gsys(int n, int k, uint8_t *Gsys)
{
    gf_init();
    uint8_t *M = (uint8_t *)malloc((size_t)k * (size_t)n);
    for (int r = 0; r < k; r++)
        for (int c = 0; c < n; c++) M[(size_t)r * n + c] = gf_alog_t[((r + 1) * (c + 1)) % 255];
    /* Gauss-Jordan on the leading block, row operations applied to all n columns */
    for (int col = 0; col < k; col++) {
        int piv = -1;
        for (int r = col; r < k; r++)
            if (M[(size_t)r * n + col]) { piv = r; break; }
        if (piv < 0) { free(M); return -1; }
        if (piv != col)
            for (int c = 0; c < n; c++) {
                uint8_t t = M[(size_t)col * n + c];
                M[(size_t)col * n + c] = M[(size_t)piv * n + c];
                M[(size_t)piv * n + c] = t;
            }
        uint8_t iv = gf_inv(M[(size_t)col * n + col]);
        for (int c = 0; c < n; c++) M[(size_t)col * n + c] = gf_mul(iv, M[(size_t)col * n + c]);
        for (int r = 0; r < k; r++) {
            if (r == col) continue;
            uint8_t f = M[(size_t)r * n + col];
            if (!f) continue;
            for (int c = 0; c < n; c++) M[(size_t)r * n + c] ^= gf_mul(f, M[(size_t)col * n + c]);
        }
    }
    memcpy(Gsys, M, (size_t)k * (size_t)n);
    free(M);
    return 0;
}

/* RS encode: c = u * G_sys per byte position (Matlab/ReedSolomonErasureCodes.m:53). */
/* info [k][S], cw [n][S].                                                          */
void orc_rs_encode(int n, int k, int S, const uint8_t *Gsys, const uint8_t *info, uint8_t *cw)
{
    gf_init();
    memset(cw, 0, (size_t)n * S);
    for (int c = 0; c < n; c++) {
        uint8_t *dst = cw + (size_t)c * S;
        for (int r = 0; r < k; r++) {
            uint8_t g = Gsys[(size_t)r * n + c];
            if (!g) continue;
            const uint8_t *src = info + (size_t)r * S;
            for (int l = 0; l < S; l++) dst[l] ^= gf_mul(g, src[l]);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* RS erasure decode.  Follows Matlab/My_RS_Decode_Optimize_With_GFTables.m:  */
/* 15-118 with the call site Matlab/ReedSolomonErasureCodes.m:80-85 (take the */
/* FIRST k received symbols, in order).  Steps restated one for one:          */
/*   GJ_mat(ii,:) = G(:,recv(ii)) (:19-23); move the unit entry of each       */
/*   received systematic row to the diagonal by column swaps, tracking        */
/*   bit_order_vec (:29-48); forward elimination of the repair rows with the  */
/*   accumulator carried alongside (:55-91, incl. the row swap from below     */
/*   when the diagonal is zero :80-90); Jordan back-substitution (:100-105);  */
/*   un-permute (:110-116).  The accumulator is an S-byte symbol here.        */
/* recv_idx: 0-based received positions (ascending), exactly k of them.       */
/* recv_val: [k][S].  out: [k][S].  Returns 0, or 1 if the matrix was found   */
/* rank deficient (:94-96; cannot happen for an MDS code).                    */
/* ------------------------------------------------------------------------- */
int orc_rs_decode(int n, int k, int S, const uint8_t *Gsys, const int32_t *recv_idx,
                  const uint8_t *recv_val, uint8_t *out)
{
    gf_init();
    uint8_t *GJ = (uint8_t *)malloc((size_t)k * (size_t)k);
    uint8_t *acc = (uint8_t *)malloc((size_t)k * (size_t)S);
    int *order = (int *)malloc(sizeof(int) * (size_t)k);
    uint8_t *tmp = (uint8_t *)malloc((size_t)(k > S ? k : S));
    for (int ii = 0; ii < k; ii++)
        for (int c = 0; c < k; c++) GJ[(size_t)ii * k + c] = Gsys[(size_t)c * n + recv_idx[ii]];
    int num_sys = 0;
    for (int ii = 0; ii < k; ii++)
        if (recv_idx[ii] < k) num_sys++;
    for (int i = 0; i < k; i++) order[i] = i;
    for (int ii = 0; ii < num_sys; ii++) {
        int col_ind = -1;
        for (int c = 0; c < k && col_ind < 0; c++)
            if (GJ[(size_t)ii * k + c] != 0) col_ind = c;
        for (int r = 0; r < k; r++) {
            uint8_t t = GJ[(size_t)r * k + ii];
            GJ[(size_t)r * k + ii] = GJ[(size_t)r * k + col_ind];
            GJ[(size_t)r * k + col_ind] = t;
        }
        /* literal :45-47 -- bit_order_vec(ii) receives the column NUMBER; that is a true   */
        /* swap because received systematic positions ascend, so column col_ind has not */
        /* been touched by an earlier iteration.                                        */
        int t = order[ii]; order[ii] = col_ind; order[col_ind] = t;
    }
    memcpy(acc, recv_val, (size_t)k * (size_t)S);
    int row = num_sys;
    int swap_ind = row + 1;
    int not_done = 1;
    while (row < k && not_done) {
        uint8_t *a_row = acc + (size_t)row * S;
        for (int jj = 0; jj < num_sys; jj++) {
            uint8_t g = GJ[(size_t)row * k + jj];
            if (g) {
                const uint8_t *src = acc + (size_t)jj * S;
                for (int l = 0; l < S; l++) a_row[l] ^= gf_mul(g, src[l]);
            }
            GJ[(size_t)row * k + jj] = 0;
        }
        for (int jj = num_sys; jj < row; jj++) {
            uint8_t g = GJ[(size_t)row * k + jj];
            if (g) {
                const uint8_t *src = acc + (size_t)jj * S;
                for (int l = 0; l < S; l++) a_row[l] ^= gf_mul(g, src[l]);
                for (int ll = jj; ll < k; ll++)
                    GJ[(size_t)row * k + ll] ^= gf_mul(g, GJ[(size_t)jj * k + ll]);
            }
        }
        if (GJ[(size_t)row * k + row] != 0) {
            uint8_t iv = gf_inv(GJ[(size_t)row * k + row]);
            for (int ll = row; ll < k; ll++) GJ[(size_t)row * k + ll] = gf_mul(iv, GJ[(size_t)row * k + ll]);
            for (int l = 0; l < S; l++) a_row[l] = gf_mul(iv, a_row[l]);
            row++;
            swap_ind = row + 1;
        } else {
            if (swap_ind >= k) {
                not_done = 0;
            } else {
                memcpy(tmp, GJ + (size_t)row * k, (size_t)k);
                memcpy(GJ + (size_t)row * k, GJ + (size_t)swap_ind * k, (size_t)k);
                memcpy(GJ + (size_t)swap_ind * k, tmp, (size_t)k);
                memcpy(tmp, acc + (size_t)row * S, (size_t)S);
                memcpy(acc + (size_t)row * S, acc + (size_t)swap_ind * S, (size_t)S);
                memcpy(acc + (size_t)swap_ind * S, tmp, (size_t)S);
                swap_ind++;
            }
        }
    }
    int rank_def = (row < k);
    for (int ii = k - 2; ii >= num_sys; ii--) {
        uint8_t *a_row = acc + (size_t)ii * S;
        for (int jj = ii + 1; jj < k; jj++) {
            uint8_t g = GJ[(size_t)ii * k + jj];
            if (g) {
                const uint8_t *src = acc + (size_t)jj * S;
                for (int l = 0; l < S; l++) a_row[l] ^= gf_mul(g, src[l]);
            }
            GJ[(size_t)ii * k + jj] = 0;
        }
    }
    for (int ii = 0; ii < num_sys; ii++)
        memcpy(out + (size_t)order[ii] * S, recv_val + (size_t)ii * S, (size_t)S);
    for (int ii = num_sys; ii < k; ii++)
        memcpy(out + (size_t)order[ii] * S, acc + (size_t)ii * S, (size_t)S);
    free(tmp); free(order); free(acc); free(GJ);
    return rank_def;
}

/* ------------------------------------------------------------------------- */
/* Batch drivers (OpenMP over codewords; codewords are independent, each      */
/* `while(1)` iteration of the reference kernels touches only its own         */
/* codeword[], ldpc_erasure_decoder.cl:27-104).  Used by the tests for bulk   */
/* comparisons and by bench.py as the timed CPU baseline.                     */
/* payload [B][n][S] in place, erased [B][n] in place, out [B][k][S] or NULL, */
/* fail_sys [B] or NULL, iters [B] or NULL.  mode 0 = peel, 1 = hybrid        */
/* (contract abort), status [B] (hybrid return code) or NULL.                 */
/* ------------------------------------------------------------------------- */
int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_ldpc_decode_batch(int n, int k, const int32_t *row_ptr, const int32_t *col_idx, int S,
                           int64_t B, uint8_t *payload, uint8_t *erased, uint8_t *out,
                           uint8_t *fail_sys, int32_t *iters, int32_t *status,
                           int max_iter, int early_stop, int mode, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t b = 0; b < B; b++) {
        uint8_t *p = payload + (size_t)b * n * S;
        uint8_t *e = erased + (size_t)b * n;
        int it = 0, st = 0;
        if (mode == 0) {
            if (S % 8 == 0 && (((uintptr_t)p) & 7) == 0)
                it = orc_ldpc_peel_u64(n, k, row_ptr, col_idx, S, p, e, max_iter, early_stop);
            else
                it = orc_ldpc_peel(n, k, row_ptr, col_idx, S, p, e, max_iter, early_stop);
        } else {
            st = orc_ldpc_hybrid(n, k, row_ptr, col_idx, S, p, e, max_iter, 0, NULL);
        }
        if (out) memcpy(out + (size_t)b * k * S, p, (size_t)k * S);
        if (fail_sys) {
            int f = 0;
            for (int i = 0; i < k; i++) f |= e[i];
            fail_sys[b] = (uint8_t)f;
        }
        if (iters) iters[b] = it;
        if (status) status[b] = st;
    }
}

void orc_ldpc_encode_batch(int n, int k, const int32_t *row_ptr, const int32_t *col_idx, int S,
                           int64_t B, const uint8_t *info, uint8_t *cw, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++)
        orc_ldpc_encode(n, k, row_ptr, col_idx, S, info + (size_t)b * k * S, cw + (size_t)b * n * S);
}

/* ------------------------------------------------------------------------- */
/* FEC packet front-ends.                                */
/* Packet = one 64-bit FEC header word + the S-byte symbol.  Header: the      */
/* 32-bit value [class:8 | block:8 | symbol:16] repeated in both halves,      */
/* OpenCL/device/ldpc_erasure_encoder_VITA_in_UDP_out.cl:100-104 (repair       */
/* symbols) and :170-175 (source symbols); class code 1 (:57).  Receiver:      */
/* OpenCL/device/ldpc_erasure_decoder_with_reordering_logic.cl:77-84 parses    */
/* (hdr >> 24) & 0xff, (hdr >> 16) & 0xff, hdr & 0xffff; a block buffer starts */
/* all-erased and all-zero (:59-68); a received symbol is stored at its        */
/* symbol number and its flag cleared (:94-131); packets of blocks outside     */
/* the window are dropped (:105,124).  The reference's window is {current,     */
/* next}; here it is [block0, block0 + B) modulo 256.  A packet counts for     */
/* its block every time it arrives (cur_block_num_cnt, :117), duplicates too.  */
/* ------------------------------------------------------------------------- */
static uint64_t orc_fec_header(uint32_t cls, uint32_t block, uint32_t symbol)
{
    const uint64_t d = 0xffffffffull & (((uint64_t)(cls & 0xffu) << 24) | ((uint64_t)(block & 0xffu) << 16) | (symbol & 0xffffu));
    return ((d << 32) & 0xffffffff00000000ull) | d;
}

void orc_packetize(int n, int S, uint32_t block0, int64_t B, const uint8_t *cw, uint8_t *packets)
{
    const size_t ps = 8 + (size_t)S;
    for (int64_t b = 0; b < B; b++)
        for (int s = 0; s < n; s++) {
            uint8_t *p = packets + ((size_t)b * n + s) * ps;
            const uint64_t h = orc_fec_header(1u, block0 + (uint32_t)b, (uint32_t)s);
            memcpy(p, &h, 8);
            memcpy(p + 8, cw + ((size_t)b * n + s) * S, (size_t)S);
        }
}

/* cw [B][n][S] (zeroed, then filled), flags [B][n] (1 = erased), counts [B+1] ([B] = dropped packets) */
void orc_depacketize(int n, int S, uint32_t block0, int64_t B, const uint8_t *packets, int64_t npackets,
                     uint8_t *cw, uint8_t *flags, uint32_t *counts)
{
    const size_t ps = 8 + (size_t)S;
    memset(cw, 0, (size_t)B * n * S);
    memset(flags, 1, (size_t)B * n);
    memset(counts, 0, (size_t)(B + 1) * 4);
    for (int64_t i = 0; i < npackets; i++) {
        const uint8_t *p = packets + (size_t)i * ps;
        uint64_t h;
        memcpy(&h, p, 8);
        const uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
        const uint32_t cls = (lo >> 24) & 0xffu, blk = (lo >> 16) & 0xffu, sym = lo & 0xffffu;
        const uint32_t rel = (blk - block0) & 0xffu;
        if (lo != hi || cls != 1u || (int64_t)rel >= B || (int)sym >= n) { counts[B]++; continue; }
        memcpy(cw + ((size_t)rel * n + sym) * S, p + 8, (size_t)S);
        flags[(size_t)rel * n + sym] = 0;
        counts[rel]++;
    }
}

/* hand-off rule of the receiver, ldpc_erasure_decoder_with_reordering_logic.cl:54-55,139 */
int orc_ready_to_decode(int n, int k, int cur_cnt, int next_cnt)
{
    const int m = n - k;
    const int desired = (int)((double)m * 0.8 + 0.5), minimum = (int)((double)m * 0.2 + 0.5);   /* round() */
    return (cur_cnt == n) || ((cur_cnt > k + desired) && (next_cnt > 10)) || ((cur_cnt > k + minimum) && (next_cnt > 100));
}

/* ------------------------------------------------------------------------- */
/* Non-binary GF(256) LDPC code.                         */
/*   coefficients: Matlab/ErasureCodes_NonBinaryLDPCSim.m:51-58 -- the binary  */
/*     H keeps its structure, every 1 becomes floor((GF_SIZE-1)*rand)+1, i.e.  */
/*     a nonzero field element.  MATLAB's rand stream is not reproducible, so  */
/*     the draw is Threefry4x32-20, key {3, seed, 0, 0}, counter {e, 0, 0, 0}  */
/*     for the e-th nonzero of H in row-major (CSR) order: 1 + (v[0] mod 255)  */
/*     (an extension in the spirit of the erasure generator; the CUDA path     */
/*     uses the same draw, or any table the caller supplies);                  */
/*   encoder: :176-182 -- parity p = inv(h_diag) * sum_{others} h * c;         */
/*   decoder: Matlab/My_LDPC_HybridML_NonBinary_Erasure_Decoder.m -- <= itenum */
/*     serial sweeps (:19-55: a check with ONE erased member recovers it as    */
/*     inv(h) * sum of h * y over the others), then Gauss-Jordan over GF(256)  */
/*     on the residual set (:57-125), field polynomial 0x171 (:69 of the sim). */
/* A symbol is S bytes; the check's coefficient multiplies every byte.         */
/* ------------------------------------------------------------------------- */
void orc_nb_coefficients(int64_t nnz, uint32_t seed, uint8_t *coef)
{
    for (int64_t e = 0; e < nnz; e++) {
        const uint32_t ctr[4] = {(uint32_t)e, (uint32_t)(e >> 32), 0u, 0u}, key[4] = {3u, seed, 0u, 0u};
        uint32_t out[4];
        orc_threefry4x32_20(ctr, key, out);
        coef[e] = (uint8_t)(1u + out[0] % 255u);
    }
}

static void nb_axpy(uint8_t *dst, const uint8_t *src, uint8_t a, int S)   /* dst += a * src */
{
    if (a == 0) return;
    const int la = gf_log_t[a];
    for (int l = 0; l < S; l++)
        if (src[l]) dst[l] ^= gf_alog_t[la + gf_log_t[src[l]]];
}

static void nb_scale(uint8_t *dst, uint8_t a, int S)                      /* dst *= a */
{
    const int la = gf_log_t[a];
    for (int l = 0; l < S; l++) dst[l] = (dst[l] && a) ? gf_alog_t[la + gf_log_t[dst[l]]] : 0;
}

void orc_nb_encode(int n, int k, const int32_t *row_ptr, const int32_t *col_idx, const uint8_t *coef, int S,
                   const uint8_t *info, uint8_t *cw)
{
    gf_init();
    memcpy(cw, info, (size_t)k * S);
    for (int r = 0; r < n - k; r++) {
        uint8_t *p = cw + (size_t)(k + r) * S;
        memset(p, 0, (size_t)S);
        const int last = row_ptr[r + 1] - 1;                   /* the diagonal: not used in the sum (sim :178) */
        for (int j = row_ptr[r]; j < last; j++) nb_axpy(p, cw + (size_t)col_idx[j] * S, coef[j], S);
        nb_scale(p, gf_inv(coef[last]), S);                    /* :181 */
    }
}

/* One codeword in place.  Returns 0 = clean after the sweeps, 1 = elimination ran and succeeded, 2 = it met a
 * column without pivot (ml_fail: the state after the sweeps is left, as for the binary hybrid decoder);
 * do_ml = 0 stops after the sweeps (returns 0 / 3 = erasures left).  *iters = sweeps made.                       */
int orc_nb_hybrid(int n, int k, const int32_t *row_ptr, const int32_t *col_idx, const uint8_t *coef, int S,
                  uint8_t *payload, uint8_t *erased, int peel_iter, int do_ml, int *iters)
{
    gf_init();
    const int m = n - k;
    int it = 0, left = 0;
    for (int i = 0; i < n; i++) left += erased[i] ? 1 : 0;
    uint8_t *acc = (uint8_t *)malloc((size_t)S);
    while (left > 0 && it < peel_iter) {                       /* decoder :19-55 */
        it++;
        for (int c = 0; c < m; c++) {
            int cnt = 0, at = -1;
            for (int j = row_ptr[c]; j < row_ptr[c + 1]; j++)
                if (erased[col_idx[j]]) { cnt++; at = j; }
            if (cnt != 1) continue;
            memset(acc, 0, (size_t)S);
            for (int j = row_ptr[c]; j < row_ptr[c + 1]; j++)
                if (j != at) nb_axpy(acc, payload + (size_t)col_idx[j] * S, coef[j], S);
            nb_scale(acc, gf_inv(coef[at]), S);
            memcpy(payload + (size_t)col_idx[at] * S, acc, (size_t)S);
            erased[col_idx[at]] = 0;
            left--;
        }
    }
    free(acc);
    if (iters) *iters = it;
    if (left == 0) return 0;
    if (!do_ml) return 3;

    const int e = left;
    int *E = (int *)malloc(sizeof(int) * (size_t)e);
    int *pos = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0, j = 0; i < n; i++) {
        pos[i] = -1;
        if (erased[i]) { pos[i] = j; E[j++] = i; }
    }
    uint8_t *A = (uint8_t *)calloc((size_t)m * (size_t)e, 1);      /* find_inv = H_sparse(:, erasure_ind) (:60) */
    uint8_t *rhs = (uint8_t *)calloc((size_t)m * (size_t)S, 1);    /* :68-77 */
    for (int c = 0; c < m; c++)
        for (int j = row_ptr[c]; j < row_ptr[c + 1]; j++) {
            const int u = col_idx[j];
            if (pos[u] >= 0) A[(size_t)c * e + pos[u]] = coef[j];
            else nb_axpy(rhs + (size_t)c * S, payload + (size_t)u * S, coef[j], S);
        }
    int abort_ml = e > m;
    uint8_t *tmp = (uint8_t *)malloc((size_t)(e > S ? e : S));
    for (int col = 0; col < e && !abort_ml; col++) {               /* :80-110 */
        int first = -1;
        for (int r = col; r < m; r++)
            if (A[(size_t)r * e + col]) { first = r; break; }
        if (first < 0) { abort_ml = 1; break; }
        if (first != col) {
            memcpy(tmp, rhs + (size_t)col * S, (size_t)S);
            memcpy(rhs + (size_t)col * S, rhs + (size_t)first * S, (size_t)S);
            memcpy(rhs + (size_t)first * S, tmp, (size_t)S);
            memcpy(tmp, A + (size_t)col * e, (size_t)e);
            memcpy(A + (size_t)col * e, A + (size_t)first * e, (size_t)e);
            memcpy(A + (size_t)first * e, tmp, (size_t)e);
        }
        const uint8_t inv = gf_inv(A[(size_t)col * e + col]);      /* :92-98 make the diagonal 1 */
        nb_scale(A + (size_t)col * e, inv, e);
        nb_scale(rhs + (size_t)col * S, inv, S);
        for (int r = first + 1; r < m; r++) {                      /* :100-108 the other rows that had a nonzero here */
            if (r == col) continue;
            const uint8_t f = A[(size_t)r * e + col];
            if (!f) continue;
            nb_axpy(A + (size_t)r * e, A + (size_t)col * e, f, e);
            nb_axpy(rhs + (size_t)r * S, rhs + (size_t)col * S, f, S);
        }
    }
    if (!abort_ml)
        for (int col = e - 1; col >= 1; col--)                     /* :112-122 */
            for (int r = 0; r < col; r++) {
                const uint8_t f = A[(size_t)r * e + col];
                if (!f) continue;
                nb_axpy(rhs + (size_t)r * S, rhs + (size_t)col * S, f, S);
                A[(size_t)r * e + col] = 0;
            }
    if (!abort_ml)
        for (int j = 0; j < e; j++) {                              /* :124 */
            memcpy(payload + (size_t)E[j] * S, rhs + (size_t)j * S, (size_t)S);
            erased[E[j]] = 0;
        }
    free(tmp); free(rhs); free(A); free(pos); free(E);
    return abort_ml ? 2 : 1;
}

void orc_nb_encode_batch(int n, int k, const int32_t *row_ptr, const int32_t *col_idx, const uint8_t *coef, int S,
                         int64_t B, const uint8_t *info, uint8_t *cw, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    gf_init();
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++)
        orc_nb_encode(n, k, row_ptr, col_idx, coef, S, info + (size_t)b * k * S, cw + (size_t)b * n * S);
}

void orc_nb_decode_batch(int n, int k, const int32_t *row_ptr, const int32_t *col_idx, const uint8_t *coef, int S,
                         int64_t B, uint8_t *payload, uint8_t *erased, uint8_t *out, uint8_t *fail_sys,
                         int32_t *iters, int32_t *status, int max_iter, int mode, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    gf_init();
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t b = 0; b < B; b++) {
        uint8_t *p = payload + (size_t)b * n * S;
        uint8_t *e = erased + (size_t)b * n;
        int it = 0;
        const int st = orc_nb_hybrid(n, k, row_ptr, col_idx, coef, S, p, e, max_iter, mode == 1, &it);
        if (out) memcpy(out + (size_t)b * k * S, p, (size_t)k * S);
        if (fail_sys) {
            int f = 0;
            for (int i = 0; i < k; i++) f |= e[i];
            fail_sys[b] = (uint8_t)f;
        }
        if (iters) iters[b] = it;
        if (status) status[b] = st;
    }
}
