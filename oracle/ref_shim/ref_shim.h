/*
 * ref_shim.h -- the 40 lines of OpenCL-for-Intel-FPGA vocabulary that let gcc compile the
 * reference's own device sources UNMODIFIED, from where they lie under /root/reference:
 *     OpenCL/device/ldpc_erasure_decoder.cl        (canonical peeling decoder)
 *     OpenCL/device/ldpc_erasure_decoder_old.pro   (same sweep + early stop + ERROR_STAT counters)
 *     OpenCL/device/ldpc_erasure_decoder_perf_tests.cl (the 2-way-split variant the host sets args for)
 *     OpenCL/device/ldpc_erasure_encoder.cl
 *     OpenCL/device/ldpc_erasure_decoder_top.cl / ldpc_erasure_encoder_top.cl (data_in / data_out, typedefs, channels)
 *     OpenCL/device/threefry.h (Random123, through its own openclfeatures.h: build with -D__OPENCL_VERSION__=120)
 *
 * TEST INFRASTRUCTURE ONLY (oracle/_ref): nothing under ldpc_erasure_codes_b200/ links or loads it.
 *
 * What the shim supplies:
 *   - address-space and kernel qualifiers, OpenCL scalar / vector type names;
 *   - Intel channels: `channel T NAME __attribute__((depth(..)))` becomes a thread-local object whose
 *     ADDRESS names a FIFO; read_channel_intel / write_channel_intel go to ref_chan_read / ref_chan_write.
 *     A channel can have a source hook (called when a read finds the FIFO empty) and a sink hook (called
 *     on every write) so that the harness streams frames from / to plain arrays -- the role of the
 *     reference's data_in / data_out kernels.  The datapath kernels are infinite `while(1)` loops; a read
 *     that finds its channel empty and its source exhausted longjmp()s back to the harness: that is the
 *     kernel being "finished" by the host (clFinish on the data_out queue, main.cpp:632).
 */
#ifndef REF_SHIM_H
#define REF_SHIM_H

#include <setjmp.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned long ulong;
typedef unsigned int uint;
typedef unsigned short ushort;
typedef unsigned char uchar;
typedef struct { int x, y, z, w; } int4;

#define __kernel static __attribute__((unused))
#define __global
#define global
#define restrict __restrict__
#define __constant static const
#define channel static __thread

/* the kernels' progress printf()s (data_in "Enter data_in", data_out's FER line) are not wanted in a test log */
#define printf(...) ((void)0)

#define REF_MAX_CHAN 8
typedef struct ref_chan {
    const void *key;                 /* address of the channel object */
    size_t elem;                     /* bytes per element */
    unsigned char *buf;              /* FIFO storage (elements), grows on demand */
    size_t head, count, cap;
    int (*source)(void *ctx, void *dst);          /* 1 = produced one element into dst, 0 = exhausted */
    void (*sink)(void *ctx, const void *src);
    void *ctx;
} ref_chan;

static __thread ref_chan ref_chans[REF_MAX_CHAN];
static __thread int ref_nchan;
static __thread jmp_buf ref_finish;

static ref_chan *ref_chan_get(const void *key, size_t elem)
{
    for (int i = 0; i < ref_nchan; i++)
        if (ref_chans[i].key == key) return &ref_chans[i];
    if (ref_nchan == REF_MAX_CHAN) abort();
    ref_chan *c = &ref_chans[ref_nchan++];
    memset(c, 0, sizeof(*c));
    c->key = key;
    c->elem = elem;
    return c;
}

static void ref_chan_reset_all(void)
{
    for (int i = 0; i < ref_nchan; i++) free(ref_chans[i].buf);
    ref_nchan = 0;
}

static void ref_chan_push(ref_chan *c, const void *src)
{
    if (c->count == c->cap) {
        const size_t ncap = c->cap ? 2 * c->cap : 64;
        unsigned char *nb = (unsigned char *)malloc(ncap * c->elem);
        if (!nb) abort();
        for (size_t i = 0; i < c->count; i++)
            memcpy(nb + i * c->elem, c->buf + ((c->head + i) % c->cap) * c->elem, c->elem);
        free(c->buf);
        c->buf = nb; c->cap = ncap; c->head = 0;
    }
    memcpy(c->buf + ((c->head + c->count) % c->cap) * c->elem, src, c->elem);
    c->count++;
}

static void ref_chan_write(const void *key, const void *src, size_t elem)
{
    ref_chan *c = ref_chan_get(key, elem);
    if (c->sink) c->sink(c->ctx, src);
    else ref_chan_push(c, src);
}

static void ref_chan_read(const void *key, void *dst, size_t elem)
{
    ref_chan *c = ref_chan_get(key, elem);
    if (c->count) {
        memcpy(dst, c->buf + c->head * c->elem, elem);
        c->head = (c->head + 1) % c->cap;
        c->count--;
        return;
    }
    if (c->source && c->source(c->ctx, dst)) return;
    longjmp(ref_finish, 1);          /* nothing will ever arrive: the host "finishes" the kernel */
}

#define read_channel_intel(ch) \
    (__extension__({ __typeof__(ch) _ref_v; ref_chan_read(&(ch), &_ref_v, sizeof(_ref_v)); _ref_v; }))
#define write_channel_intel(ch, v) \
    do { __typeof__(ch) _ref_w = (v); ref_chan_write(&(ch), &_ref_w, sizeof(_ref_w)); } while (0)

#endif /* REF_SHIM_H */
