/*
 * ref_tu.c -- one translation unit of oracle/_ref/libldpc_ref.so: the shim vocabulary, then ONE of the
 * reference's device sources included unmodified from the build's view directory (symlinks into
 * /root/reference/OpenCL/device, made by oracle/Makefile), then a harness that plays the host:
 * it feeds the kernel's input channel from plain arrays and drains its output channel into plain arrays.
 *
 * TEST INFRASTRUCTURE ONLY.  Compiled several times (see oracle/Makefile):
 *   REF_KIND 1  decoder datapath file behind shim-declared symbol_type / channels (SYM_LEN = REF_SYM_LEN,
 *               the compile-time constant the reference's top file sets to 128, decoder_top.cl:38)
 *   REF_KIND 2  encoder datapath file, same
 *   REF_KIND 3  ldpc_erasure_decoder_top.cl as committed (threefry.h, LDPC_Vlist_data.h, typedefs, channels,
 *               data_in, data_out, and whatever `ldpc_erasure_decoder.cl` is in the view: the canonical file,
 *               _old.pro or _perf_tests.cl -- the author swapped them the same way)
 *   REF_KIND 4  ldpc_erasure_encoder_top.cl as committed
 *   REF_PREFIX  name prefix of the exported harness functions
 *   REF_ERRSTAT the decoder variant reports through ERROR_STAT instead of LDPC_DEC_DOUT (_old.pro, _perf_tests.cl)
 *   REF_ARGS2   the decoder variant takes (num_iter, code_ind)                       (_perf_tests.cl:30)
 */
#include "ref_shim.h"
#include <pthread.h>
#include <stdint.h>

#if REF_KIND == 1 || REF_KIND == 2
#define SYM_LEN REF_SYM_LEN
/* = ldpc_erasure_decoder_top.cl:38-49 and ldpc_erasure_encoder_top.cl:31-36 with SYM_LEN as a build parameter;
 * the REF_KIND 3/4 units take these from the reference's own top files and the tests compare the two */
typedef struct {
    unsigned long symbol[SYM_LEN];
    unsigned char is_erasure;
} symbol_type;
typedef struct {
    int num_LDPC_errors;
    int num_RS_errors;
} error_type;
#include "LDPC_Vlist_data.h"
#endif

#if REF_KIND == 1
channel symbol_type LDPC_DEC_DIN;
channel symbol_type LDPC_DEC_DOUT;
channel error_type ERROR_STAT;
#include "ldpc_erasure_decoder.cl"
#elif REF_KIND == 2
channel symbol_type LDPC_ENC_DIN;
channel symbol_type LDPC_ENC_DOUT;
#include "ldpc_erasure_encoder.cl"
#elif REF_KIND == 3
#undef UINT64_C                  /* Random123's openclfeatures.h defines its own */
#include "ldpc_erasure_decoder_top.cl"
#elif REF_KIND == 4
#include "LDPC_Vlist_data.h"      /* the encoder's top file has no table include of its own; the stub needs the master table */
#include "ldpc_erasure_encoder_top.cl"
#else
#error "REF_KIND"
#endif

#include "n2000_k1000_no6cycle_ldpc_Vlist_device.h"   /* the view's stand-in (guarded): ref_code_ind for variants that do not include it */

#define REF_CAT2(a, b) a##b
#define REF_CAT(a, b) REF_CAT2(a, b)
#define REF_FN(name) REF_CAT(REF_PREFIX, name)
#define REF_STACK_BYTES (64ul << 20)    /* codeword[n_ldpc] of 1032-byte symbols lives on the kernel's stack (perf_tests: 2 x 4000) */

int REF_FN(sym_bytes)(void) { return SYM_LEN * 8; }
int REF_FN(sizeof_symbol_type)(void) { return (int)sizeof(symbol_type); }
int REF_FN(code_params)(int code_ind, int out[6])
{
    if (code_ind < 0 || code_ind > 1) return -1;
    for (int i = 0; i < 6; i++) out[i] = ldpc_params[code_ind][i];
    return 0;
}
/* row r of the selected code as the kernels see it: out[0] = weight, then 1-based columns */
int REF_FN(vlist_row)(int code_ind, int r, short out[20])
{
    if (code_ind < 0 || code_ind > 1) return -1;
    for (int i = 0; i < 20; i++) out[i] = parity_check_mat_Vlist_master[ldpc_params[code_ind][2] + r][i];
    return 0;
}

typedef struct {
    /* arrays of the whole call */
    const uint8_t *in_payload;   /* [frames][rows_in][S] or NULL (all zero) */
    const uint8_t *in_flags;     /* [frames][rows_in] or NULL (none erased) */
    uint8_t *out_payload;        /* [frames][rows_out][S] or NULL */
    uint8_t *out_flags;          /* [frames][rows_out] or NULL */
    int *errstat;                /* [frames][2] cumulative counters as written to ERROR_STAT, or NULL */
    int rows_in, rows_out, code_ind, num_iter;
    /* this thread's frame range and cursors */
    long f0, f1, rd, wr, er;     /* rd / wr count symbols, er counts frames */
    pthread_t thr;
} ref_io;

static int ref_src(void *ctx, void *dst)
{
    ref_io *io = (ref_io *)ctx;
    const long f = io->f0 + io->rd / io->rows_in, i = io->rd % io->rows_in;
    if (f >= io->f1) return 0;
    symbol_type *s = (symbol_type *)dst;
    memset(s, 0, sizeof(*s));
    if (io->in_payload) memcpy(s->symbol, io->in_payload + ((size_t)f * io->rows_in + i) * (SYM_LEN * 8), SYM_LEN * 8);
    if (io->in_flags) s->is_erasure = io->in_flags[(size_t)f * io->rows_in + i];
    io->rd++;
    return 1;
}

static void ref_snk(void *ctx, const void *src)
{
    ref_io *io = (ref_io *)ctx;
    const symbol_type *s = (const symbol_type *)src;
    const long f = io->f0 + io->wr / io->rows_out, i = io->wr % io->rows_out;
    if (f >= io->f1) abort();
    if (io->out_payload) memcpy(io->out_payload + ((size_t)f * io->rows_out + i) * (SYM_LEN * 8), s->symbol, SYM_LEN * 8);
    if (io->out_flags) io->out_flags[(size_t)f * io->rows_out + i] = s->is_erasure;
    io->wr++;
}

#if REF_KIND == 1 || REF_KIND == 3
static void ref_err_snk(void *ctx, const void *src)
{
    ref_io *io = (ref_io *)ctx;
    const error_type *e = (const error_type *)src;
    const long f = io->f0 + io->er;
    if (f >= io->f1) abort();
    if (io->errstat) { io->errstat[2 * f] = e->num_LDPC_errors; io->errstat[2 * f + 1] = e->num_RS_errors; }
    io->er++;
}

static void *ref_dec_thread(void *arg)
{
    ref_io *io = (ref_io *)arg;
    ref_code_ind = io->code_ind;
    ref_chan *cin = ref_chan_get(&LDPC_DEC_DIN, sizeof(symbol_type));
    cin->source = ref_src; cin->ctx = io;
    ref_chan *cout = ref_chan_get(&LDPC_DEC_DOUT, sizeof(symbol_type));
    cout->sink = ref_snk; cout->ctx = io;
    ref_chan *cerr = ref_chan_get(&ERROR_STAT, sizeof(error_type));
    cerr->sink = ref_err_snk; cerr->ctx = io;
    if (!setjmp(ref_finish)) {
#ifdef REF_ARGS2
        ldpc_erasure_decoder((short)io->num_iter, io->code_ind);
#else
        ldpc_erasure_decoder((short)io->num_iter);
#endif
    }
    ref_chan_reset_all();
    return NULL;
}

/* Decodes `frames` codewords of code `code_ind` (0 = (2000,1000), 1 = (2040,1530): ldpc_params) with the
 * reference kernel, `nthreads` independent kernel instances each taking a contiguous range of frames (the
 * cumulative ERROR_STAT counters restart per instance: errstat_chunk receives each frame's instance start). */
int REF_FN(decode)(int code_ind, int num_iter, long frames, const uint8_t *payload, const uint8_t *flags,
                   uint8_t *out, uint8_t *out_flags, int *errstat, long *errstat_chunk, int nthreads)
{
    if (code_ind < 0 || code_ind > 1 || frames < 0) return -1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > frames) nthreads = frames > 0 ? (int)frames : 1;
    ref_io *ios = (ref_io *)calloc((size_t)nthreads, sizeof(ref_io));
    pthread_attr_t at;
    pthread_attr_init(&at);
    pthread_attr_setstacksize(&at, REF_STACK_BYTES);
    int rc = 0;
    for (int t = 0; t < nthreads; t++) {
        ref_io *io = &ios[t];
        io->in_payload = payload; io->in_flags = flags; io->out_payload = out; io->out_flags = out_flags; io->errstat = errstat;
        io->rows_in = ldpc_params[code_ind][0]; io->rows_out = ldpc_params[code_ind][1];
        io->code_ind = code_ind; io->num_iter = num_iter;
        io->f0 = frames * t / nthreads; io->f1 = frames * (t + 1) / nthreads;
        if (errstat_chunk) for (long f = io->f0; f < io->f1; f++) errstat_chunk[f] = io->f0;
        if (pthread_create(&io->thr, &at, ref_dec_thread, io)) { rc = -2; nthreads = t; break; }
    }
    for (int t = 0; t < nthreads; t++) {
        pthread_join(ios[t].thr, NULL);
#ifdef REF_ERRSTAT
        if (ios[t].er != ios[t].f1 - ios[t].f0) rc = -3;
#else
        if (ios[t].wr != (ios[t].f1 - ios[t].f0) * ios[t].rows_out) rc = -3;
#endif
    }
    pthread_attr_destroy(&at);
    free(ios);
    return rc;
}
#endif

#if REF_KIND == 3
static void ref_flag_snk(void *ctx, const void *src)
{
    ref_io *io = (ref_io *)ctx;
    io->out_flags[io->wr++] = ((const symbol_type *)src)->is_erasure;
}

static void *ref_data_in_thread(void *arg)
{
    ref_io *io = (ref_io *)arg;
    ref_chan *cin = ref_chan_get(&LDPC_DEC_DIN, sizeof(symbol_type));
    cin->sink = ref_flag_snk; cin->ctx = io;
    /* args as the host sets them (main.cpp:578-589): buffer (ignored by the kernel), n, seed, P, code, frames */
    data_in(NULL, (unsigned short)ldpc_params[io->code_ind][0], io->num_iter /* seed */, io->rows_in /* P */, io->code_ind, io->f1);
    ref_chan_reset_all();
    return NULL;
}

/* The reference's erasure generator (decoder_top.cl:57-120) run for `frames` frames from its start:
 * flags [frames][n] = the is_erasure field of what it pushes into LDPC_DEC_DIN. */
int REF_FN(data_in)(int code_ind, int seed, int per_numerator_div_64, long frames, uint8_t *flags)
{
    if (code_ind < 0 || code_ind > 1 || frames < 0) return -1;
    ref_io io;
    memset(&io, 0, sizeof(io));
    io.code_ind = code_ind; io.num_iter = seed; io.rows_in = per_numerator_div_64; io.f1 = frames; io.out_flags = flags;
    pthread_attr_t at;
    pthread_attr_init(&at);
    pthread_attr_setstacksize(&at, REF_STACK_BYTES);
    if (pthread_create(&io.thr, &at, ref_data_in_thread, &io)) return -2;
    pthread_join(io.thr, NULL);
    pthread_attr_destroy(&at);
    return io.wr == frames * ldpc_params[code_ind][0] ? 0 : -3;
}
#endif

#if REF_KIND == 2 || REF_KIND == 4
static void *ref_enc_thread(void *arg)
{
    ref_io *io = (ref_io *)arg;
    ref_code_ind = io->code_ind;
    ref_chan *cin = ref_chan_get(&LDPC_ENC_DIN, sizeof(symbol_type));
    cin->source = ref_src; cin->ctx = io;
    ref_chan *cout = ref_chan_get(&LDPC_ENC_DOUT, sizeof(symbol_type));
    cout->sink = ref_snk; cout->ctx = io;
    if (!setjmp(ref_finish)) ldpc_erasure_encoder();
    ref_chan_reset_all();
    return NULL;
}

/* info [frames][k][S] -> cw [frames][n][S] through the reference encoder kernel */
int REF_FN(encode)(int code_ind, long frames, const uint8_t *info, uint8_t *cw, int nthreads)
{
    if (code_ind < 0 || code_ind > 1 || frames < 0) return -1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > frames) nthreads = frames > 0 ? (int)frames : 1;
    ref_io *ios = (ref_io *)calloc((size_t)nthreads, sizeof(ref_io));
    pthread_attr_t at;
    pthread_attr_init(&at);
    pthread_attr_setstacksize(&at, REF_STACK_BYTES);
    int rc = 0;
    for (int t = 0; t < nthreads; t++) {
        ref_io *io = &ios[t];
        io->in_payload = info; io->out_payload = cw;
        io->rows_in = ldpc_params[code_ind][1]; io->rows_out = ldpc_params[code_ind][0];
        io->code_ind = code_ind;
        io->f0 = frames * t / nthreads; io->f1 = frames * (t + 1) / nthreads;
        if (pthread_create(&io->thr, &at, ref_enc_thread, io)) { rc = -2; nthreads = t; break; }
    }
    for (int t = 0; t < nthreads; t++) {
        pthread_join(ios[t].thr, NULL);
        if (ios[t].wr != (ios[t].f1 - ios[t].f0) * ios[t].rows_out) rc = -3;
    }
    pthread_attr_destroy(&at);
    free(ios);
    return rc;
}
#endif
