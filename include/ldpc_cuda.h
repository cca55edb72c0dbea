/*
 * ldpc_cuda.h -- C ABI of libldpc_cuda, the B200-native packet-erasure codec.
 *
 * This library stands where the reference's OpenCL host<->device contract
 * stands (chadac8j/LDPC_Erasure_Codes, OpenCL/host/src/main.cpp): the reference
 * has no plugin/FFI layer, its boundary is "create context + buffers, set kernel
 * args, enqueue the data_in / codec / data_out task trio, finish, read back".
 * Each entry point below names the reference interface it replaces.
 *
 * Conventions
 *   - plain C, no exceptions, no torch / CUDA types in the signatures: device
 *     buffers are `void*` device pointers OWNED BY THE CALLER, streams are the
 *     caller's cudaStream_t passed as `void*` (NULL = default stream);
 *   - every function returns 0 (LDPC_OK) or a negative LDPC_ERR_* code and
 *     never calls exit() (the reference's checkError() prints and exits,
 *     main.cpp:493); ldpc_last_error_string() describes the last failure of
 *     the calling thread;
 *   - work is enqueued on the caller's stream, no hidden synchronisation
 *     except in the *_host variants and ldpc_get_stats();
 *   - a context is bound to one GPU and is thread-compatible (one context per
 *     GPU per host thread), like the reference's single-threaded host.  Contexts
 *     on different GPUs may live in one process and be driven from one thread or
 *     from one thread each (ldpc_*_host_multi does the latter).  A context's
 *     scratch (schedules, syndromes) is shared by all its calls: keep ONE
 *     stream's worth of work in flight per context -- calls on one stream are
 *     ordered; calls on different streams need the caller's own event between
 *     them.
 *
 * Data layout (the GPU form of the reference's `symbol_type`,
 * OpenCL/device/ldpc_erasure_decoder_top.cl:38-44 = {ulong symbol[128]; uchar
 * is_erasure}): structure-of-arrays --
 *   payload  [B][n][S] bytes, S = symbol_bytes (reference SYM_LEN*8 = 1024),
 *   erasure mask [B][mask_words] uint32, bit (i & 31) of word (i >> 5) set iff
 *            symbol i of that codeword is erased (mask_words = ceil(n/32)).
 * As in the reference (ldpc_erasure_decoder.cl:17-20) an erased symbol's payload
 * is expected to be all-zero on input; the decoder never reads it, and writes
 * it only when it recovers the symbol.
 */
#ifndef LDPC_CUDA_H
#define LDPC_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDPC_CUDA_ABI_VERSION 2

enum {
    LDPC_OK = 0,
    LDPC_ERR_ARG = -1,         /* bad argument                                     */
    LDPC_ERR_IO = -2,          /* cannot open / read the H file                    */
    LDPC_ERR_FORMAT = -3,      /* not a MAT-v5 file with a usable H_sparse         */
    LDPC_ERR_CUDA = -4,        /* CUDA runtime / driver error                      */
    LDPC_ERR_NOMEM = -5,       /* host or device allocation failed                 */
    LDPC_ERR_UNSUPPORTED = -6, /* code / symbol size outside what the kernels take */
    LDPC_ERR_NOT_TRIANGULAR = -7 /* encode asked on an H without the staircase diagonal */
};

/* decode modes: the reference's OpenCL decoder (peeling only) and the MATLAB hybrid */
enum {
    LDPC_MODE_PEEL = 0,   /* OpenCL/device/ldpc_erasure_decoder.cl:24-105            */
    LDPC_MODE_HYBRID = 1  /* Matlab/My_LDPC_HybridML_Erasure_Decoder.m               */
};

/* erasure models of ldpc_gen_erasures */
enum {
    LDPC_ERASURE_IID64 = 0,  /* reference rule: (x0 & 63) < P   (decoder_top.cl:105) */
    LDPC_ERASURE_IID32 = 1,  /* extension: x0 < floor(p * 2^32), any rate            */
    LDPC_ERASURE_BURSTY = 2  /* Matlab/Bursty_Error_Channel_Model_Generator.m        */
};

typedef struct ldpc_ctx ldpc_ctx;
typedef struct rs_ctx rs_ctx;

typedef struct ldpc_code_info {
    int32_t n, k, m;          /* code length, information symbols, checks          */
    int32_t nnz;              /* ones in H                                          */
    int32_t symbol_bytes;     /* S                                                  */
    int32_t mask_words;       /* uint32 words of erasure mask per codeword          */
    int32_t rs_n, rs_k;       /* RS-equivalent block for MDS counting (ldpc_params) */
    int32_t max_row_weight, max_col_weight;
    int32_t encode_levels;    /* dependency depth of the back-substitution          */
    int32_t slice_bytes;      /* W: bytes of a symbol one executor unit carries     */
    int32_t exec_slots;       /* units resident in shared memory per SM             */
    int32_t device;
    int64_t max_batch;        /* codewords one internal chunk processes             */
} ldpc_code_info;

typedef struct ldpc_erasure_model {
    int32_t model;            /* LDPC_ERASURE_*                                     */
    int32_t per_numerator_div_64; /* IID64: P, erase iff (x0 & 63) < P  (host flag -p) */
    uint32_t threshold32;     /* IID32: erase iff x0 < threshold32                  */
    double alpha, beta, bias; /* BURSTY: PER in good / bad state, good-state bias   */
} ldpc_erasure_model;

/* cumulative counters = the reference's error_type pushed through ERROR_STAT
 * (decoder_top.cl:46-49, perf_tests.cl:233-236) and printed by data_out (:151-156) */
typedef struct ldpc_stats {
    int64_t frames;           /* codewords decoded since create / reset             */
    int64_t ldpc_errors;      /* frames with a systematic symbol still erased       */
    int64_t rs_errors;        /* RS-equivalent blocks beyond MDS capacity           */
    int64_t ml_attempts;      /* hybrid mode: frames that entered GF(2) elimination */
    int64_t ml_failures;      /* hybrid mode: rank-deficient eliminations           */
    int64_t ml_recovered;     /* hybrid mode: frames counted in ldpc_errors whose systematic
                                 symbols the elimination then recovered (hybrid frame errors =
                                 ldpc_errors - ml_recovered)                         */
    int64_t any_errors;       /* frames with ANY of the n symbols still unknown after decoding:
                                 the MATLAB harness's block-error criterion
                                 (LDPCErasureCodes_MessagePassingAlgSim.m:229-236)    */
} ldpc_stats;

/* ---- context ------------------------------------------------------------------------
 * Replaces init_opencl() + the ldpc_params table select (main.cpp:439-544, :258-259) and
 * the code tables baked into the .aocx (OpenCL/device/LDPC_Vlist_data.h).
 * h_mat_path: MAT-v5 file holding sparse `H_sparse` (the reference's Matlab/ *.mat), or
 *   NULL to take the built-in code `code_ind` (0 = (2000,1000), 1 = (2040,1530) as in
 *   host flag -c, 2 = (4000,2000)); built-ins are read from the `codes/` directory that
 *   ships next to the library (or $LDPC_CUDA_CODES_DIR).
 * symbol_bytes: S, a multiple of 16.   max_batch: scratch is sized for this many
 *   codewords; larger calls are processed in chunks of max_batch.                        */
int ldpc_ctx_create(ldpc_ctx **out, const char *h_mat_path, int code_ind, int symbol_bytes,
                    int device, int64_t max_batch);
/* Replaces cleanup() (main.cpp:668-691). */
int ldpc_ctx_destroy(ldpc_ctx *ctx);
int ldpc_ctx_info(const ldpc_ctx *ctx, ldpc_code_info *info);
/* Host-only: parse a MAT-v5 file holding sparse `H_sparse` (the reference's Matlab/ *.mat) without
 * touching a GPU.  dims receives {m, n, nnz, triangular(0/1)}; row_ptr / col_idx may be NULL to
 * query the sizes first, else they must hold m+1 and nnz entries.                                   */
int ldpc_read_h_file(const char *h_mat_path, int32_t dims[4], int32_t *row_ptr, int32_t *col_idx);
/* Host copy of H as CSR (row_ptr[m+1], col_idx[nnz], 0-based ascending) = the Vlist rows. */
int ldpc_ctx_get_csr(const ldpc_ctx *ctx, int32_t *row_ptr, int32_t *col_idx);
/* Tuning knob for experiments: force the executor's slice width W (16/32/64/...) and slot
 * count (0 = choose automatically).                                                      */
int ldpc_ctx_set_exec_geometry(ldpc_ctx *ctx, int slice_bytes, int slots);

/* ---- encoder ------------------------------------------------------------------------
 * Replaces the encoder task trio data_in -> ldpc_erasure_encoder -> data_out
 * (OpenCL/device/ldpc_erasure_encoder_top.cl:43-92, ldpc_erasure_encoder.cl:26-95).
 * d_info [B][k][S] -> d_cw [B][n][S] (systematic symbols first, then the n-k parities). */
int ldpc_encode(ldpc_ctx *ctx, const void *d_info, void *d_cw, int64_t B, void *stream);

/* ---- erasure channel ----------------------------------------------------------------
 * Replaces the decoder-side data_in kernel (ldpc_erasure_decoder_top.cl:57-120):
 * Threefry4x32-20, key {1, seed}, counter 1 + (frame0 + b) * n + symbol.  Writes the
 * erasure mask of B codewords and, if d_payload != NULL, zeroes the erased symbols of
 * d_payload [B][n][S] in place (the "erased = all zero" convention).
 * The counter is 32 bits wide and wraps, as the reference's `c.v[0]++` does: for one seed the
 * frame sequence repeats after 2^32 / gcd(2^32, n) frames (2^28 = 2.7e8 for n = 2000 and 4000,
 * 2^29 = 5.4e8 for n = 2040).  Longer runs must change the seed per period (tools/bler_deep.py
 * does); frame0 + B beyond one period replays frames.                                       */
int ldpc_gen_erasures(ldpc_ctx *ctx, const ldpc_erasure_model *model, uint32_t seed,
                      uint64_t frame0, int64_t B, uint32_t *d_mask, void *d_payload, void *stream);

/* ---- decoder ------------------------------------------------------------------------
 * Replaces the ldpc_erasure_decoder task (args num_iter, code_ind: main.cpp:593-596) plus
 * the payload side of data_out.  d_cw [B][n][S] and d_mask [B][mask_words] are read only;
 * d_out [B][k][S] receives the first k symbols after decoding (decoder.cl:97-102);
 * d_fail [B] (may be NULL) receives 1 where a systematic symbol is still erased
 * (perf_tests.cl:215-228).  max_iter = the reference's num_iter (sweeps over the checks,
 * host flag -i, default 50); the decoder reproduces the serial sweep order exactly, so a
 * binding cap gives the reference's partial result.  In LDPC_MODE_HYBRID max_iter is the
 * sweep cap before elimination (10 in the MATLAB file) and d_fail additionally covers
 * rank-deficient eliminations.  Erased symbols of d_cw are all-zero by the reference's
 * convention (ldpc_gen_erasures and ldpc_depacketize leave them so), but no decoder here reads
 * them: a symbol that stays unknown is passed through to d_out as it came (peel mode) or as
 * zeros (hybrid mode, frames that reached the elimination stage).  A batch larger than
 * max_batch is cut into max_batch chunks that alternate between two internal streams; the
 * caller's stream continues when all of them are done.                                    */
int ldpc_decode(ldpc_ctx *ctx, const void *d_cw, const uint32_t *d_mask, void *d_out,
                uint8_t *d_fail, int max_iter, int mode, int64_t B, void *stream);
/* The same with the second failure criterion the reference uses: d_fail_any [B] (may be NULL)
 * receives 1 where ANY of the n symbols is still unknown after decoding -- the MATLAB harness
 * compares all n symbols with the truth (LDPCErasureCodes_MessagePassingAlgSim.m:229-236), the
 * OpenCL kernel only the first k (d_fail).  d_fail_any[b] >= d_fail[b].                      */
int ldpc_decode_ex(ldpc_ctx *ctx, const void *d_cw, const uint32_t *d_mask, void *d_out,
                   uint8_t *d_fail, uint8_t *d_fail_any, int max_iter, int mode, int64_t B, void *stream);

/* ---- error-rate run ---------------------------------------------------------------
 * The reference's committed flow in one call (main.cpp:555-659 with decoder_top.cl:57-158 and
 * ldpc_erasure_decoder_perf_tests.cl): `frames` all-zero codewords get synthetic erasures
 * (data_in), are decoded (only the erasure pattern matters for an all-zero codeword) and only
 * the cumulative counters come back (ERROR_STAT / data_out).  Runs the pattern phase only --
 * no payload is moved -- and ADDS to the context's counters (read them with ldpc_get_stats). */
int ldpc_simulate_fer(ldpc_ctx *ctx, const ldpc_erasure_model *model, uint32_t seed, uint64_t frame0,
                      int64_t frames, int max_iter, int mode, void *stream);

/* ---- statistics ---------------------------------------------------------------------
 * Replaces the ERROR_STAT channel + data_out report (decoder_top.cl:123-158).
 * Synchronises the context's device.                                                     */
int ldpc_get_stats(ldpc_ctx *ctx, ldpc_stats *out);
int ldpc_reset_stats(ldpc_ctx *ctx);

/* ---- profiling ----------------------------------------------------------------------
 * Replaces CL_QUEUE_PROFILING_ENABLE + getStartEndTime(kernel_event) (main.cpp:515,:652):
 * when enabled, every kernel the context launches is bracketed by CUDA events on the
 * launching stream; ldpc_profile_read() synchronises and returns per-kernel device time.
 * Launch counters are always maintained.                                                 */
/* (LDPC_K_CHANNEL also covers the packet front-end kernels) */
enum { LDPC_K_PEEL = 0, LDPC_K_EXEC_DECODE = 1, LDPC_K_EXEC_ENCODE = 2,
       LDPC_K_HYBRID = 3,        /* elimination stage 1: inactivation decoding, one warp per codeword (pattern part) */
       LDPC_K_CHANNEL = 4,
       LDPC_K_HYBRID_WARP = 5,   /* stage 2: per-warp Gauss-Jordan on what stage 1 deferred (rare)          */
       LDPC_K_HYBRID_CTA = 6,    /* stage 3: CTA-per-codeword Gauss-Jordan on what stage 2 deferred         */
       LDPC_K_HYBRID_APPLY = 7,  /* stage 1, payload part: replay of the recorded pivots                    */
       LDPC_K_KINDS = 8 };
typedef struct ldpc_profile {
    double ms[LDPC_K_KINDS];        /* summed device time per kernel kind (profiling on)  */
    int64_t launches[LDPC_K_KINDS]; /* kernel launches per kind since create / last reset */
    /* tuning aid, filled only when the context was created with LDPC_CUDA_PHASE_TIMING=1 in the
     * environment: SM cycles the executor's group leaders spent per phase, summed over units:
     * [0] claim + TMA issue, [1] waiting for the load, [2] XOR, [3] store, [4] number of units   */
    uint64_t exec_phase_cycles[8];
    uint64_t ge_phase_cycles[8];    /* LDPC_CUDA_PHASE_TIMING=1: inactivation stage, warp cycles per phase: setup, rows,
                                       adjacency+syndromes, peel/inactivate, dense solve, output; [6] = codewords      */
    uint64_t apply_phase_cycles[8]; /* apply kernel: plan load, right-hand sides, replay, dense solve, output; [6] = codewords */
} ldpc_profile;
int ldpc_profile_enable(ldpc_ctx *ctx, int on);
int ldpc_profile_read(ldpc_ctx *ctx, ldpc_profile *out, int reset);

/* ---- host-buffer entry points -------------------------------------------------------
 * The reference's run() (main.cpp:555-659): blocking host->device copy, kernels, blocking
 * device->host copy.  Host pointers (pinned memory makes the copies asynchronous and
 * overlapped; pageable memory works too).  Chunked and pipelined over internal streams. */
int ldpc_encode_host(ldpc_ctx *ctx, const void *h_info, void *h_cw, int64_t B);
int ldpc_decode_host(ldpc_ctx *ctx, const void *h_cw, const uint32_t *h_mask, void *h_out,
                     uint8_t *h_fail, int max_iter, int mode, int64_t B);
int ldpc_decode_host_ex(ldpc_ctx *ctx, const void *h_cw, const uint32_t *h_mask, void *h_out,
                        uint8_t *h_fail, uint8_t *h_fail_any, int max_iter, int mode, int64_t B);
/* The whole box: ctxs[n_ctx] = one context per GPU, same code and symbol size.  The
 * batch is cut into contiguous frame ranges -- GPU g takes frames [g*B/G, (g+1)*B/G) -- and one host
 * thread per GPU drives that GPU's pipeline; there is no exchange between the GPUs (codewords are
 * independent, ldpc_erasure_decoder.cl:27-104).  Counters stay per context (sum them with
 * ldpc_get_stats).  Returns the first failing GPU's code; its message names the GPU.             */
int ldpc_encode_host_multi(ldpc_ctx *const *ctxs, int n_ctx, const void *h_info, void *h_cw, int64_t B);
int ldpc_decode_host_multi(ldpc_ctx *const *ctxs, int n_ctx, const void *h_cw, const uint32_t *h_mask,
                           void *h_out, uint8_t *h_fail, uint8_t *h_fail_any, int max_iter, int mode,
                           int64_t B);

/* In place (round 2; the reference has no equivalent -- its run() reads the whole output buffer back, main.cpp:638).
 * Of a decoded codeword the host already holds every received symbol: only the erased systematic symbols are
 * news.  h_cw [B][n][S] must be PAGE-LOCKED host memory (cudaHostAlloc / cudaHostRegister; LDPC_ERR_ARG otherwise):
 * the device fetches the RECEIVED symbols from it (erased ones are not read: they count as all-zero, the input
 * convention above), decodes, and writes every systematic symbol whose mask bit is set back into it -- the recovered
 * value, or zero for a symbol that stays erased.  Received symbols and the parity part are not touched.  About
 * (1-p)*n*S up and p*k*S down instead of n*S and k*S bytes per codeword cross PCIe.
 * ldpc_encode_host_inplace: h_cw [B][n][S] with the information symbols in rows 0..k-1 (pinned or pageable); only
 * those rows are uploaded and only the n-k parity rows are written back.                                          */
int ldpc_decode_host_inplace(ldpc_ctx *ctx, void *h_cw, const uint32_t *h_mask, uint8_t *h_fail,
                             uint8_t *h_fail_any, int max_iter, int mode, int64_t B);
int ldpc_decode_host_inplace_multi(ldpc_ctx *const *ctxs, int n_ctx, void *h_cw, const uint32_t *h_mask,
                                   uint8_t *h_fail, uint8_t *h_fail_any, int max_iter, int mode, int64_t B);
int ldpc_encode_host_inplace(ldpc_ctx *ctx, void *h_cw, int64_t B);

/* ---- synthetic payload --------------------------------------------------------------
 * Counter-based uniform bytes (Threefry key {2, seed}, counter = 16-byte block index +
 * block0): bench / test input generator, independent of batch sharding.                  */
int ldpc_fill_random(void *d_dst, int64_t nbytes, uint32_t seed, uint64_t block0, int device,
                     void *stream);

/* ---- FEC packet front-ends ----------------------------------------
 * The reference's sender and receiver wrap every symbol in a packet: one 64-bit FEC header word
 * -- the 32-bit value [class:8 | block:8 | symbol:16] repeated in both halves, class code 1 --
 * followed by the S-byte symbol (OpenCL/device/ldpc_erasure_encoder_VITA_in_UDP_out.cl:100-104,
 * 170-175; OpenCL/device/ldpc_erasure_decoder_with_reordering_logic.cl:77-84).
 *   ldpc_packetize    codewords [B][n][S] -> packets [B*n][8+S], block numbers block0 + b (mod 256).
 *   ldpc_depacketize  packets [n_packets][8+S] in ANY order, with losses and duplicates -> codewords
 *                     [B][n][S] (missing symbols all-zero) + erasure mask (bit set = not received):
 *                     what ldpc_decode takes.  A packet belongs to the window iff
 *                     (block - block0) mod 256 < B (B <= 256) and symbol < n, class 1 and both header
 *                     halves equal; others are dropped (receiver :105,124).  d_counts[b] = packets
 *                     placed for block b (duplicates counted, like cur_block_num_cnt :117),
 *                     d_counts[B] = dropped packets.
 *   ldpc_ready_to_decode  the receiver's hand-off rule (:54-55,139): all n symbols in, or more than
 *                     k + round(0.8 m) with > 10 packets of the next block seen, or more than
 *                     k + round(0.2 m) with > 100.                                                 */
#define LDPC_FEC_CLASS 1
int ldpc_packetize(ldpc_ctx *ctx, const void *d_cw, uint32_t block0, int64_t B, void *d_packets, void *stream);
int ldpc_depacketize(ldpc_ctx *ctx, const void *d_packets, int64_t n_packets, uint32_t block0, int64_t B,
                     void *d_cw, uint32_t *d_mask, uint32_t *d_counts, void *stream);
int ldpc_ready_to_decode(const ldpc_ctx *ctx, int cur_block_cnt, int next_block_cnt);
/* Variable payload length (sender :162,186-197 `num_longs_used`; receiver :94-111): a packet still occupies an
 * 8 + S byte slot; d_len8[packet] (<= S / 8) says how many 8-byte payload words are valid.  The sender zero-fills the
 * rest of the slot, the receiver leaves the rest of the symbol zero.  d_len8 of ldpc_packetize_var is [B * n].       */
int ldpc_packetize_var(ldpc_ctx *ctx, const void *d_cw, const uint16_t *d_len8, uint32_t block0, int64_t B, void *d_packets,
                       void *stream);
int ldpc_depacketize_var(ldpc_ctx *ctx, const void *d_packets, const uint16_t *d_len8, int64_t n_packets, uint32_t block0,
                         int64_t B, void *d_cw, uint32_t *d_mask, uint32_t *d_counts, void *stream);
/* The receiver's two-buffer state machine (ldpc_erasure_decoder_with_reordering_logic.cl:45-142): the first usable
 * packet names the current block, the next block number (modulo 256) is assembled beside it, packets of any other
 * block are dropped, and after EVERY packet the hand-off rule (ldpc_ready_to_decode) is evaluated on the two arrival
 * counters; when it fires the current block is decoded as it stands and emitted, `next` becomes `current` and the freed
 * buffer is cleared.  ldpc_rx_stream_push takes a batch of packets in arrival order ([n_packets][8 + S] on the device,
 * d_len8 may be NULL), reads the headers back, runs that per-packet control on the host and the payload through
 * ldpc_depacketize / ldpc_decode on the GPU; blocks that became ready are written to d_out [cap][k][S], d_fail [cap],
 * h_blocks [cap] (host: their block numbers), *n_decoded of them (LDPC_ERR_ARG if more than cap).  ldpc_rx_stream_flush
 * ends the stream: the current block, then the next if it holds packets, decoded as they stand.
 * ldpc_rx_stream_state: {current, next, packets counted for current, for next}.  Synchronises `stream`.              */
typedef struct ldpc_rx_stream ldpc_rx_stream;
int ldpc_rx_stream_create(ldpc_rx_stream **out, ldpc_ctx *ctx, int max_iter, int mode);
int ldpc_rx_stream_destroy(ldpc_rx_stream *s);
int ldpc_rx_stream_push(ldpc_rx_stream *s, const void *d_packets, const uint16_t *d_len8, int64_t n_packets, void *d_out,
                        uint8_t *d_fail, int32_t *h_blocks, int cap, int *n_decoded, void *stream);
int ldpc_rx_stream_flush(ldpc_rx_stream *s, void *d_out, uint8_t *d_fail, int32_t *h_blocks, int cap, int *n_decoded,
                         void *stream);
int ldpc_rx_stream_state(const ldpc_rx_stream *s, int32_t state[4]);

/* ---- Reed-Solomon GF(2^8) comparison code -------------------------------------------
 * Field polynomial 0x171, alpha = 2 (Matlab/Build_GF256_Lookup_Tables.m:11-24);
 * G[i][j] = alpha^(i*j) systematised (Matlab/Test_My_RS_Decode.m:30-37);
 * decode from the first k received symbols (Matlab/ReedSolomonErasureCodes.m:80-85,
 * My_RS_Decode_Optimize_With_GFTables.m).                                                */
int rs_ctx_create(rs_ctx **out, int n, int k, int symbol_bytes, int device, int64_t max_batch);
int rs_ctx_destroy(rs_ctx *ctx);
/* Host copy of the k x n systematic generator (row-major). */
int rs_ctx_get_generator(const rs_ctx *ctx, uint8_t *gsys);
/* d_info [B][k][S] -> d_cw [B][n][S]. */
int rs_encode(rs_ctx *ctx, const void *d_info, void *d_cw, int64_t B, void *stream);
/* d_cw [B][n][S], d_mask [B][ceil(n/32)] -> d_out [B][k][S]; d_fail[b] = 1 iff fewer than
 * k symbols of codeword b were received (then d_out holds the received systematic symbols,
 * erased ones zero).                                                                      */
int rs_decode(rs_ctx *ctx, const void *d_cw, const uint32_t *d_mask, void *d_out,
              uint8_t *d_fail, int64_t B, void *stream);

/* ---- non-binary GF(2^8) LDPC code -------------------------------------
 * Matlab/ErasureCodes_NonBinaryLDPCSim.m, Matlab/My_LDPC_HybridML_NonBinary_Erasure_Decoder.m:
 * the binary H of `base` keeps its structure, every edge carries a nonzero element of GF(2^8)
 * (field polynomial 0x171); a check reads sum_u h[c][u] * y[u] = 0, the coefficient multiplying
 * every byte of the S-byte symbol.  coef_csr: one coefficient per nonzero of H in row-major (CSR)
 * order, or NULL to draw them as the simulation does (floor(255 * rand) + 1, sim :55) from
 * Threefry4x32-20 with key {3, coef_seed}, counter = index of the nonzero.  The context borrows
 * `base` (its peel kernel and scratch: the peeling schedule does not depend on the coefficients);
 * `base` must outlive it and the two must not be used from two threads at once.
 * encode: sim :176-182 (parity = inv(h_diag) * sum of h * c over the other members).
 * decode: mode PEEL = the serial sweeps of decoder :19-55 (max_iter of them; reference: 10);
 * mode HYBRID = then Gauss-Jordan over GF(2^8) on the residual set (:57-125); a column without
 * pivot leaves the state after the sweeps and counts as ml_failure, as for the binary code.
 * Buffers, masks, failure flags and counters as for ldpc_encode / ldpc_decode.                  */
typedef struct ldpc_nb_ctx ldpc_nb_ctx;
int ldpc_nb_ctx_create(ldpc_nb_ctx **out, ldpc_ctx *base, const uint8_t *coef_csr, uint32_t coef_seed);
int ldpc_nb_ctx_destroy(ldpc_nb_ctx *ctx);
int ldpc_nb_get_coefficients(const ldpc_nb_ctx *ctx, uint8_t *coef_csr);
int ldpc_nb_encode(ldpc_nb_ctx *ctx, const void *d_info, void *d_cw, int64_t B, void *stream);
int ldpc_nb_decode(ldpc_nb_ctx *ctx, const void *d_cw, const uint32_t *d_mask, void *d_out, uint8_t *d_fail,
                   int max_iter, int mode, int64_t B, void *stream);

/* ---- code design (host only): girth-8 triangular-form H generator and short-cycle checker -----
 * Replaces Matlab/Hgen_irregularDegree_no6cycles_systematic_encoding.m:94-224 ("bit filling": rows
 * are filled with variables drawn with probability ~ (edges still needed)^3, a draw is kept iff it
 * closes no 4- or 6-cycle; row r ends with the diagonal edge (r, k + r)) and
 * Matlab/Cycle_Finder_length4_fromroot.m / Cycle_Finder_length6.m.  The draws come from a seeded
 * generator (MATLAB's rand stream is not reproducible): same seed, same matrix.
 * deg_c_prof / deg_v_prof: n_*_deg rows of (count, degree), degrees in DESCENDING order; the sums of
 * count * degree must agree.  Output: dims = {m, n, nnz, 1}; CSR with ascending columns when
 * row_ptr [m+1] / col_idx [col_cap >= nnz] are given (call once with NULLs to size them: nnz <=
 * sum(count * degree) + m).  LDPC_ERR_UNSUPPORTED if no matrix was found in max_tries attempts.  */
int ldpc_h_generate(const int32_t *deg_c_prof, int n_c_deg, const int32_t *deg_v_prof, int n_v_deg,
                    uint64_t seed, int max_tries, int32_t dims[4], int32_t *row_ptr, int32_t *col_idx,
                    int64_t col_cap, int32_t *tries_used);
/* Number of variable nodes for which the reference's rooted finders report a cycle: length 4
 * (Cycle_Finder_length4_fromroot.m) and length <= 6 (Cycle_Finder_length6.m, which counts a 4-cycle
 * too).  Both zero <=> girth >= 8.                                                              */
int ldpc_h_count_short_cycles(const int32_t *row_ptr, const int32_t *col_idx, int m, int n,
                              int64_t *vars_on_4cycles, int64_t *vars_on_6cycles);
const char *ldpc_h_last_error_string(void);

/* ---- errors ------------------------------------------------------------------------- */
const char *ldpc_last_error_string(void);
int ldpc_cuda_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_CUDA_H */
