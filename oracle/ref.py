"""ctypes/numpy front-end of oracle/_ref/libldpc_ref.so: the REFERENCE'S OWN device sources
(OpenCL/device/ldpc_erasure_decoder.cl, _old.pro, _perf_tests.cl, ldpc_erasure_encoder.cl, the data_in
generator of ldpc_erasure_decoder_top.cl with Random123 threefry.h) compiled unmodified by gcc behind
oracle/ref_shim/ (see oracle/Makefile).

TEST INFRASTRUCTURE ONLY: used by tests/ to pin the restatement (oracle/ldpc_oracle.c) and the CUDA path
to the reference's bytes, and by bench.py's --impl reference / cpu_baseline legs.  The library is built in
the authoring container (where /root/reference exists) and travels prebuilt to the GPU box.

Variants:  "canon" = ldpc_erasure_decoder.cl (payloads + flags out, no early stop, no counters),
           "old"   = ldpc_erasure_decoder_old.pro (early stop, cumulative ERROR_STAT counters, no payload out;
                     its RS block size is hard-wired to (250,125), :31-32),
           "perf"  = ldpc_erasure_decoder_perf_tests.cl (2-way split; SURVEY a-9: known-buggy early stop),
each at symbol sizes 16 / 64 / 1024 bytes behind shim-declared types ("s2" / "s8" / "s128") and, at the
reference's own 1024 bytes, through the committed top files ("top").  Codes: 0 = (2000,1000), 1 = (2040,1530)
-- the two rows of ldpc_params; the (4000,2000) code exists only as a .mat and has no OpenCL table.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libldpc_ref.so")
REFERENCE = "/root/reference"


def build() -> str | None:
    """Builds oracle/_ref when the reference tree is present; returns the library path or None."""
    if os.path.isdir(os.path.join(REFERENCE, "OpenCL", "device")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
    return _SO if os.path.exists(_SO) else None


def available() -> bool:
    return build() is not None


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = build()
        if so is None:
            raise RuntimeError("oracle/_ref/libldpc_ref.so is missing and there is no reference tree to build it from")
        _lib = C.CDLL(so)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _unit(variant: str, S: int, top: bool) -> str:
    if top:
        assert S == 1024, "the committed top files fix SYM_LEN = 128 (1024-byte symbols)"
        return f"ref_top_{variant}_"
    assert S in (16, 64, 1024), "oracle/_ref is built for 16-, 64- and 1024-byte symbols"
    return f"ref_{variant}_s{S // 8}_"


def code_params(code_ind: int):
    out = (C.c_int * 6)()
    assert lib().ref_top_canon_code_params(C.c_int(code_ind), out) == 0
    return list(out)


def vlist_rows(code_ind: int):
    """The code's check rows as the kernels index them: list of 0-based ascending column lists."""
    n, k = code_params(code_ind)[:2]
    rows = []
    buf = (C.c_short * 20)()
    for r in range(n - k):
        lib().ref_top_canon_vlist_row(C.c_int(code_ind), C.c_int(r), buf)
        rows.append([buf[1 + i] - 1 for i in range(buf[0])])
    return rows


def sizeof_symbol_type(top=True):
    return lib().ref_top_canon_sizeof_symbol_type() if top else lib().ref_canon_s128_sizeof_symbol_type()


def data_in(code_ind: int, seed: int, P: int, frames: int) -> np.ndarray:
    """flags [frames][n] from the reference's data_in kernel (decoder_top.cl:57-120), frames 0..frames-1."""
    n = code_params(code_ind)[0]
    flags = np.zeros((frames, n), dtype=np.uint8)
    rc = lib().ref_top_canon_data_in(C.c_int(code_ind), C.c_int(seed), C.c_int(P), C.c_long(frames), _p(flags))
    assert rc == 0, rc
    return flags


def encode(code_ind: int, info: np.ndarray, top=False, nthreads=1) -> np.ndarray:
    """info [B][k][S] -> codewords [B][n][S] through ldpc_erasure_encoder.cl."""
    info = np.ascontiguousarray(info, dtype=np.uint8)
    B, k, S = info.shape
    n, kk = code_params(code_ind)[:2]
    assert k == kk
    cw = np.zeros((B, n, S), dtype=np.uint8)
    fn = getattr(lib(), ("ref_top_enc_" if top else f"ref_enc_s{S // 8}_") + "encode")
    if top:
        assert S == 1024
    rc = fn(C.c_int(code_ind), C.c_long(B), _p(info), _p(cw), C.c_int(nthreads))
    assert rc == 0, rc
    return cw


def decode(code_ind: int, payload: np.ndarray, erased: np.ndarray, num_iter=50, variant="canon", top=False, nthreads=1):
    """payload [B][n][S], erased [B][n] u8.  canon: dict(out [B][k][S], out_flags [B][k], fail_sys [B]);
    old / perf: dict(fail_sys [B], rs_errors [B]) from the per-frame increments of the ERROR_STAT counters."""
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    erased = np.ascontiguousarray(erased, dtype=np.uint8)
    B, n, S = payload.shape
    nn, k = code_params(code_ind)[:2]
    assert n == nn and erased.shape == (B, n)
    fn = getattr(lib(), _unit(variant, S, top) + "decode")
    if variant == "canon":
        out = np.zeros((B, k, S), dtype=np.uint8)
        oflags = np.zeros((B, k), dtype=np.uint8)
        rc = fn(C.c_int(code_ind), C.c_int(num_iter), C.c_long(B), _p(payload), _p(erased), _p(out), _p(oflags), None, None,
                C.c_int(nthreads))
        assert rc == 0, rc
        return dict(out=out, out_flags=oflags, fail_sys=(oflags.max(axis=1) if B else np.zeros(0, np.uint8)))
    err = np.zeros((B, 2), dtype=np.int32)
    chunk = np.zeros(B, dtype=np.int64)
    rc = fn(C.c_int(code_ind), C.c_int(num_iter), C.c_long(B), _p(payload), _p(erased), None, None, _p(err), _p(chunk),
            C.c_int(nthreads))
    assert rc == 0, rc
    prev = np.zeros_like(err)
    prev[1:] = err[:-1]
    prev[chunk == np.arange(B)] = 0          # first frame of each kernel instance: counters start at zero
    d = err - prev
    return dict(fail_sys=d[:, 0].astype(np.uint8), rs_errors=d[:, 1].astype(np.int32), counters=err)
